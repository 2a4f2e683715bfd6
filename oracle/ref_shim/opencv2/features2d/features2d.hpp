/* ORACLE — TEST INFRASTRUCTURE ONLY: forwards to minicv.hpp (oracle/ref_shim), no OpenCV here. */
#include "../../minicv.hpp"
