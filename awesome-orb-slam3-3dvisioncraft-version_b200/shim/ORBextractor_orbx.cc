// ORBextractor_orbx.cc — drop-in replacement for the reference's src/ORBextractor.cc.
// Same class, same signatures (include/ORBextractor.h:43-109); the work happens in liborbx.so.
//
// The class declaration cannot grow a member (Tracking.cc / Frame.cc must compile against the unmodified
// header), so the device handle of an instance lives in a side table keyed by `this`.  Instances are created
// three times per System (Tracking.cc:226-233) and live for the process; the inline `~ORBextractor(){}` in the
// reference header cannot be hooked, so handles are released at process exit.
#include "orbx_shim_config.h"
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace orbx_shim {

orbx_ctx* context() {
  static orbx_ctx* ctx = [] {
    const char* dev = std::getenv("ORBX_DEVICE");
    orbx_ctx* c = orbx_create(dev ? std::atoi(dev) : 0);
    if (!c) {
      std::fprintf(stderr, "orbx: %s\n", orbx_last_error());
      std::abort();
    }
    return c;
  }();
  return ctx;
}

// The device-resident DBoW2 vocabulary, uploaded once: ORBX_VOCABULARY names the same ORBvoc.bin the
// reference's System constructor loads (Examples pass it as argv[1]).
orbx_voc* vocabulary() {
  static orbx_voc* voc = [] {
    const char* path = std::getenv("ORBX_VOCABULARY");
    orbx_voc* v = path ? orbx_vocabulary_load(context(), path) : nullptr;
    if (!v) {
      std::fprintf(stderr, "orbx: vocabulary (%s): %s\n", path ? path : "ORBX_VOCABULARY not set", orbx_last_error());
      std::abort();
    }
    return v;
  }();
  return voc;
}

void die(const char* where, int status) {
  std::fprintf(stderr, "orbx: %s failed (%d): %s\n", where, status, orbx_last_error());
  std::abort();
}

namespace {
struct ExtState {
  orbx_ext* ext = nullptr;
  int w = 0, h = 0;
  std::vector<orbx_keypoint> kps;
  std::vector<uint8_t> desc;
};
std::mutex g_mu;
std::unordered_map<const void*, ExtState> g_ext;
}  // namespace

// (re)create the device extractor when the image size changes (the reference accepts any size per call)
static ExtState& state_for(const void* self, int nfeatures, float scale, int nlevels, int ini, int min, int w, int h) {
  std::lock_guard<std::mutex> lk(g_mu);
  ExtState& s = g_ext[self];
  if (!s.ext || w > s.w || h > s.h) {
    if (s.ext) orbx_extractor_destroy(s.ext);
    s.ext = orbx_extractor_create(context(), nfeatures, scale, nlevels, ini, min, w, h, 1);
    if (!s.ext) die("orbx_extractor_create", ORBX_EINVAL);
    s.w = w;
    s.h = h;
    const int cap = orbx_extractor_max_keypoints(s.ext);
    s.kps.resize(cap);
    s.desc.resize((size_t)cap * 32);
  }
  return s;
}

orbx_ext* extractor_handle(const void* self) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ext.find(self);
  return it == g_ext.end() ? nullptr : it->second.ext;
}

}  // namespace orbx_shim

namespace ORB_SLAM3 {

// src/ORBextractor.cc:408-468 — the scale tables stay host-side members because Frame's constructors copy them
// (src/Frame.cc:99-105); they are computed with the reference's exact float/double expressions.
ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  mvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels);
  mvInvScaleFactor.resize(nlevels);
  mvInvLevelSigma2.resize(nlevels);
  mvScaleFactor[0] = 1.0f;
  mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor;
    mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
  }
  for (int i = 0; i < nlevels; i++) {
    mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
    mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
  }
  mvImagePyramid.resize(nlevels);
  mnFeaturesPerLevel.resize(nlevels);   // informational; the device recomputes the same split
}

// src/ORBextractor.cc:1074-1156
int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
  if (_image.empty()) return -1;
  cv::Mat image = _image.getMat();
  assert(image.type() == CV_8UC1);
  orbx_shim::ExtState& s = orbx_shim::state_for(this, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST,
                                                image.cols, image.rows);
  int n = 0, mono = 0;
  orbx_shim::check("orbx_extract",
                   orbx_extract(s.ext, image.data, image.cols, image.rows, (int)image.step, vLappingArea[0],
                                vLappingArea[1], s.kps.data(), s.desc.data(), (int)s.kps.size(), &n, &mono));
  _keypoints = std::vector<cv::KeyPoint>(n);
  if (n == 0) {
    _descriptors.release();
  } else {
    _descriptors.create(n, 32, CV_8U);
    cv::Mat d = _descriptors.getMat();
    for (int i = 0; i < n; ++i) {
      const orbx_keypoint& k = s.kps[i];
      _keypoints[i] = cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, -1);
      std::memcpy(d.ptr(i), &s.desc[(size_t)i * 32], 32);
    }
  }
  // mvImagePyramid is public and read by Frame::ComputeStereoMatches (src/Frame.cc:962,1052-1071).  When the
  // stereo matcher is also replaced (Frame_orbx.cc) nothing reads it and this copy can be compiled out.
#ifndef ORBX_SHIM_NO_PYRAMID_COPY
  for (int l = 0; l < nlevels; ++l) {
    int w = 0, h = 0;
    orbx_shim::check("orbx_pyramid_level", orbx_pyramid_level(s.ext, 0, l, nullptr, 0, &w, &h));
    // same geometry as the reference: a (w+38)x(h+38) buffer whose ROI is the level
    cv::Mat temp(h + 2 * 19, w + 2 * 19, CV_8UC1);
    mvImagePyramid[l] = temp(cv::Rect(19, 19, w, h));
    orbx_shim::check("orbx_pyramid_level",
                     orbx_pyramid_level(s.ext, 0, l, mvImagePyramid[l].data, (int)mvImagePyramid[l].step, &w, &h));
    cv::copyMakeBorder(mvImagePyramid[l], temp, 19, 19, 19, 19, cv::BORDER_REFLECT_101 + cv::BORDER_ISOLATED);
  }
#endif
  return mono;
}

}  // namespace ORB_SLAM3
