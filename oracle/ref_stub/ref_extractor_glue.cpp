// ref_extractor_glue.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's UNMODIFIED
// ORB_SLAM3::ORBextractor (compiled in place from /root/reference/src/ORBextractor.cc against the OpenCV stand-in
// in this directory) so that tests can run it from Python (ctypes) and compare it with the oracle restatement.
#include "ORBextractor.h"   // the reference's own header (-I /root/reference/include)
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <new>

// ---- allocator mode -------------------------------------------------------------------------------------------
// DistributeOctTree sorts (size, ExtractorNode*) pairs (src/ORBextractor.cc:682), i.e. it breaks ties by heap address.
// With REF_BUMP_ALLOC the library is built with a monotonic operator new (addresses grow in allocation order and are
// never reused during one extraction), which is the rule the oracle and the device follow ("later-created node =
// larger address").  Without it the default glibc allocator decides, as it would in a reference binary.
#ifdef REF_BUMP_ALLOC
namespace {
struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  bool active = false;
} g_arena;
}
void* operator new(size_t n) {
  if (g_arena.active) {
    const size_t a = (n + 15) & ~(size_t)15;
    if (g_arena.used + a <= g_arena.cap) {
      void* p = g_arena.base + g_arena.used;
      g_arena.used += a;
      return p;
    }
  }
  void* p = std::malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept {
  if (g_arena.base && (char*)p >= g_arena.base && (char*)p < g_arena.base + g_arena.cap) return;
  std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }
static void arena_begin() {
  if (!g_arena.base) {
    g_arena.cap = (size_t)1 << 30;
    g_arena.base = (char*)std::malloc(g_arena.cap);
  }
  g_arena.used = 0;
  g_arena.active = true;
}
static void arena_end() { g_arena.active = false; }
#else
static void arena_begin() {}
static void arena_end() {}
#endif

struct RefKeyPoint { float x, y, size, angle, response; int octave; };

extern "C" {

int ref_alloc_mode() {
#ifdef REF_BUMP_ALLOC
  return 1;
#else
  return 0;
#endif
}

void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  return new ORB_SLAM3::ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_extractor_destroy(void* e) { delete (ORB_SLAM3::ORBextractor*)e; }

// returns the number of keypoints (<= cap), -1 for an empty image; *mono = operator()'s return value
int ref_extract(void* e_, const uint8_t* img, int w, int h, int stride, int lap0, int lap1, RefKeyPoint* kps,
                uint8_t* desc, int cap, int* mono) {
  ORB_SLAM3::ORBextractor* e = (ORB_SLAM3::ORBextractor*)e_;
  std::vector<int> lap = {lap0, lap1};
  int n = 0;
  arena_begin();
  {
    cv::Mat image = (img && w > 0 && h > 0) ? cv::Mat(h, w, CV_8UC1, (void*)img, (size_t)stride) : cv::Mat();
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    const int m = (*e)(image, cv::Mat(), keys, descriptors, lap);
    if (mono) *mono = m;
    n = m < 0 ? -1 : (int)keys.size();
    for (int i = 0; i < n && i < cap; ++i) {
      kps[i] = {keys[i].pt.x, keys[i].pt.y, keys[i].size, keys[i].angle, keys[i].response, keys[i].octave};
      std::memcpy(desc + (size_t)32 * i, descriptors.ptr(i), 32);
    }
    e->mvImagePyramid.clear();
    e->mvImagePyramid.resize(e->GetLevels());
  }
  arena_end();
  return n;
}

int ref_extractor_tables(void* e_, float* scale, float* inv, float* s2, float* is2) {
  ORB_SLAM3::ORBextractor* e = (ORB_SLAM3::ORBextractor*)e_;
  const int n = e->GetLevels();
  std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                     d = e->GetInverseScaleSigmaSquares();
  for (int i = 0; i < n; ++i) { scale[i] = a[i]; inv[i] = b[i]; s2[i] = c[i]; is2[i] = d[i]; }
  return n;
}

// pyramid level l of the last ref_pyramid call (ComputePyramid is protected: a derived class exposes it)
struct RefPyr : ORB_SLAM3::ORBextractor {
  using ORB_SLAM3::ORBextractor::ORBextractor;
  void pyr(cv::Mat im) { ComputePyramid(im); }
  int feat(int l) { return mnFeaturesPerLevel[l]; }
  int um(int v) { return umax[v]; }
};
int ref_pyramid_level(int nlevels, float scaleFactor, const uint8_t* img, int w, int h, int stride, int level, uint8_t* out,
                      int out_stride, int* lw, int* lh) {
  RefPyr e(1000, scaleFactor, nlevels, 20, 7);
  cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)stride);
  e.pyr(image);
  const cv::Mat& L = e.mvImagePyramid[level];
  *lw = L.cols; *lh = L.rows;
  if (out) for (int y = 0; y < L.rows; ++y) std::memcpy(out + (size_t)y * out_stride, L.ptr(y), L.cols);
  return 0;
}
int ref_features_per_level(int nfeatures, float scaleFactor, int nlevels, int* n, int* umax16) {
  RefPyr e(nfeatures, scaleFactor, nlevels, 20, 7);
  for (int l = 0; l < nlevels; ++l) n[l] = e.feat(l);
  if (umax16) for (int v = 0; v < 16; ++v) umax16[v] = e.um(v);
  return nlevels;
}

}  // extern "C"
