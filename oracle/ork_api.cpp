// ork_api.cpp — ORACLE (test infrastructure): flat C entry points for ctypes (tests/, bench.py's
// cpu_baseline leg, __graft_entry__.smoke()).  Not part of the product.
#include "ork.h"
#include <cstring>
#include <algorithm>

using namespace ork;

extern "C" {

void* ork_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  if (nfeatures < 0 || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !(scaleFactor > 1.f)) return nullptr;
  return new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ork_extractor_destroy(void* e) { delete (Extractor*)e; }

int ork_extractor_tables(void* ev, float* scale, float* inv, float* s2, float* is2, int* nfeat) {
  Extractor* e = (Extractor*)ev;
  for (int l = 0; l < e->nlevels; ++l) {
    if (scale) scale[l] = e->scale[l];
    if (inv) inv[l] = e->invScale[l];
    if (s2) s2[l] = e->sigma2[l];
    if (is2) is2[l] = e->invSigma2[l];
    if (nfeat) nfeat[l] = e->featuresPerLevel[l];
  }
  return e->nlevels;
}

int ork_extractor_umax(void* ev, int* umax16) {
  Extractor* e = (Extractor*)ev;
  for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
  return 16;
}

int ork_extract(void* ev, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                orbx_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out) {
  Extractor* e = (Extractor*)ev;
  std::vector<orbx_keypoint> K;
  std::vector<uint8_t> D;
  int mono = 0;
  int rc = e->extract(img, w, h, stride, lap0, lap1, K, D, &mono);
  if (n_out) *n_out = 0;
  if (rc != ORBX_OK) return rc;
  if ((int)K.size() > cap) return ORBX_ECAP;
  if (!K.empty()) {
    std::memcpy(kps, K.data(), K.size() * sizeof(orbx_keypoint));
    std::memcpy(desc, D.data(), D.size());
  }
  if (n_out) *n_out = (int)K.size();
  if (mono_out) *mono_out = mono;
  return ORBX_OK;
}

int ork_pyramid_level(void* ev, int level, uint8_t* dst, int dst_stride, int* w, int* h) {
  Extractor* e = (Extractor*)ev;
  if (level < 0 || level >= e->nlevels) return ORBX_EINVAL;
  const Gray& g = e->pyramid[level];
  if (w) *w = g.w;
  if (h) *h = g.h;
  if (dst)
    for (int y = 0; y < g.h; ++y) std::memcpy(dst + (size_t)y * dst_stride, g.row(y), g.w);
  return ORBX_OK;
}

int ork_candidates(void* ev, int level, int16_t* xy, uint8_t* score, int cap, int* n_out) {
  Extractor* e = (Extractor*)ev;
  if (level < 0 || level >= e->nlevels) return ORBX_EINVAL;
  const auto& c = e->cand[level];
  if (n_out) *n_out = (int)c.size();
  if ((int)c.size() > cap) return ORBX_ECAP;
  for (size_t i = 0; i < c.size(); ++i) {
    xy[2 * i] = (int16_t)c[i].x;
    xy[2 * i + 1] = (int16_t)c[i].y;
    score[i] = (uint8_t)c[i].response;
  }
  return ORBX_OK;
}

void ork_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                          int dstride) {
  resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
}

// out: [cap][3] int32 (x, y, score); returns the number of corners (may exceed cap).
int ork_fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, int nms, int32_t* out,
                 int cap) {
  std::vector<FastPoint> v;
  fast9_16(img, w, h, stride, threshold, nms != 0, v);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) {
    out[3 * i] = v[i].x;
    out[3 * i + 1] = v[i].y;
    out[3 * i + 2] = v[i].score;
  }
  return (int)v.size();
}

void ork_gaussian_blur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  gaussian_blur7_s2(src, w, h, sstride, dst, dstride);
}

float ork_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void ork_fast_atan2_array(const float* y, const float* x, float* out, int n) {
  for (int i = 0; i < n; ++i) out[i] = fast_atan2(y[i], x[i]);
}

// distribute_octree on an explicit candidate list (x,y relative to minBorder; response)
int ork_distribute_octree(const float* x, const float* y, const float* resp, int n, int minX, int maxX,
                          int minY, int maxY, int N, float* ox, float* oy, float* oresp, int cap) {
  std::vector<orbx_keypoint> K(n);
  for (int i = 0; i < n; ++i) { K[i] = orbx_keypoint{x[i], y[i], 7.f, -1.f, resp[i], 0}; }
  auto R = distribute_octree(K, minX, maxX, minY, maxY, N);
  for (int i = 0; i < (int)R.size() && i < cap; ++i) { ox[i] = R[i].x; oy[i] = R[i].y; oresp[i] = R[i].response; }
  return (int)R.size();
}

}  // extern "C"
