/* Host side of the TMA staging: cuTensorMapEncodeTiled through the runtime's driver entry point (no link against
 * libcuda). */
#include "sdyn_internal.h"
#include "tma.h"
#include <cstring>

namespace sdyn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static cudaError_t encoder(EncodeTiledFn* fn, const char** why)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess) { *why = "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)"; return e; }
    if (q != cudaDriverEntryPointSuccess || !p) { *why = "cuTensorMapEncodeTiled is not exported by this driver"; return cudaErrorNotSupported; }
    *fn = reinterpret_cast<EncodeTiledFn>(p);
    return cudaSuccess;
}

/* map over one level of a pyramid buffer: x = byte in the padded row, y = bordered row, z = frame */
static bool encode_one(EncodeTiledFn fn, const Geom& g, const LevelGeom& L, const uint8_t* dBuffer, int maxBatch, int boxW, int boxH,
                       CUtensorMap* out)
{
    /* origin = first padded byte of the first bordered row of the level (256-byte aligned inside the frame block) */
    void* base = const_cast<uint8_t*>(dBuffer) + (L.off - (long long)kEdge * L.pitch - kLeftPad);
    const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(L.h + 2 * kEdge), (cuuint64_t)maxBatch};
    const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.frameBytes};
    const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t encode_level_maps(const Geom& g, const uint8_t* dBuffer, int maxBatch, int boxW, int boxH, LevelMaps* out,
                              const char** why)
{
    EncodeTiledFn fn = nullptr;
    cudaError_t e = encoder(&fn, why);
    if (e != cudaSuccess) return e;
    std::memset(out, 0, sizeof *out);
    for (int l = 0; l < g.nlevels; ++l)
        if (!encode_one(fn, g, g.L[l], dBuffer, maxBatch, boxW, boxH, &out->m[l])) {
            *why = "cuTensorMapEncodeTiled rejected a pyramid level";
            return cudaErrorInvalidValue;
        }
    return cudaSuccess;
}

cudaError_t encode_resize_maps(const Geom& g, const uint8_t* dBuffer, int maxBatch, LevelMaps* out, const char** why)
{
    EncodeTiledFn fn = nullptr;
    cudaError_t e = encoder(&fn, why);
    if (e != cudaSuccess) return e;
    std::memset(out, 0, sizeof *out);
    for (int l = 1; l < g.nlevels; ++l) {
        if (g.L[l].rsPitch > kResizePitch || g.L[l].rsRows > 256) continue;      /* k_resize<0>: no TMA box */
        if (!encode_one(fn, g, g.L[l - 1], dBuffer, maxBatch, kResizePitch, g.L[l].rsRows, &out->m[l])) {
            *why = "cuTensorMapEncodeTiled rejected a resize source level";
            return cudaErrorInvalidValue;
        }
    }
    return cudaSuccess;
}

}  // namespace sdyn
