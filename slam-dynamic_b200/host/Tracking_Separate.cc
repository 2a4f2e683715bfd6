/* Drop-in replacement of ONE member function of the reference's Tracking: Tracking::Separate (src/Tracking.cc:1093-1239), as
 * declared in the reference's own include/Tracking.h:149.  Delete that function from src/Tracking.cc and add this file.
 * classifyH / classifyF (:1241-1367) are subsumed: the device classifies every cross-checked match of every box in the same
 * call that matches them. */
#include "Tracking.h"
#include "sdyn_adapters.hpp"
#include "sdyn_context.h"

namespace ORB_SLAM2
{

int Tracking::Separate(cv::Mat HorF, int flag, vector<vector<int>>& dynStatus)
{
    return sdyn_host::Separate(sdyn_host::ThreadContext(), mCurrentFrame, *mRefFrame, mLastFrame, HorF, flag, dynStatus);
}

}  // namespace ORB_SLAM2
