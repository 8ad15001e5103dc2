// Frame_orbx.cc — drop-in replacement for Frame::ComputeStereoMatches (src/Frame.cc:955-1133).  The pyramids
// the reference reads through mpORBextractorLeft/Right->mvImagePyramid stay on the device: the two extractor
// instances that just produced mvKeys / mvKeysRight (src/Frame.cc:111-114) still hold them.
#include "orbx_shim_config.h"

namespace orbx_shim { orbx_ext* extractor_handle(const void* self); }

namespace ORB_SLAM3 {

void Frame::ComputeStereoMatches() {
  mvuRight = std::vector<float>(N, -1.0f);
  mvDepth = std::vector<float>(N, -1.0f);
  orbx_ext* L = orbx_shim::extractor_handle(mpORBextractorLeft);
  orbx_ext* R = orbx_shim::extractor_handle(mpORBextractorRight);
  if (!L || !R || N == 0) return;
  auto flat = [](const std::vector<cv::KeyPoint>& v) {
    std::vector<orbx_keypoint> o(v.size());
    for (size_t i = 0; i < v.size(); ++i)
      o[i] = orbx_keypoint{v[i].pt.x, v[i].pt.y, v[i].size, v[i].angle, v[i].response, v[i].octave};
    return o;
  };
  const std::vector<orbx_keypoint> kl = flat(mvKeys), kr = flat(mvKeysRight);
  orbx_shim::check("orbx_stereo_match",
                   orbx_stereo_match(orbx_shim::context(), L, 0, R, 0, kl.data(), mDescriptors.data, (int)kl.size(),
                                     kr.data(), mDescriptorsRight.data, (int)kr.size(), mbf, mb, mvuRight.data(),
                                     mvDepth.data()));
}

}  // namespace ORB_SLAM3
