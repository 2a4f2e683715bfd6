/* Drop-in replacement of ONE member function of the reference's Frame: Frame::firstSeparate (src/Frame.cc:555-604), as
 * declared in the reference's own include/Frame.h:228.  Delete that function from src/Frame.cc (or weaken its symbol) and add
 * this file: the constructors keep calling it unchanged.  The box test runs on the extractor's GPU context. */
#include "Frame.h"
#include "sdyn_adapters.hpp"

namespace ORB_SLAM2
{

bool Frame::firstSeparate(const cv::Mat& /*mask*/, vector<cv::Rect2d>& boxes, vector<vector<int>>& index, vector<bool>& hasKpts)
{
    return sdyn_host::FirstSeparate(mpORBextractorLeft->Context(), *this, boxes, index, hasKpts);
}

}  // namespace ORB_SLAM2
