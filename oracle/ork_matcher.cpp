// ork_matcher.cpp — ORACLE (test infrastructure): sequential CPU restatement of the reference's
// descriptor matchers on the hot path, on flat arrays (same argument layout as include/orbx.h so a
// test can hand identical buffers to both sides).
//
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   src/Frame.cc:444-478,852-862,755-850
//   Frame::ComputeStereoMatches                                   src/Frame.cc:955-1133
//   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&)    src/ORBmatcher.cc:59-255
//   ORBmatcher::SearchByProjection(Frame&, const Frame&)          src/ORBmatcher.cc:2244-2509
//   ORBmatcher::SearchForTriangulation                            src/ORBmatcher.cc:1138-1428
//   ORBmatcher::ComputeThreeMaxima / DescriptorDistance           src/ORBmatcher.cc:2654-2716
//   Pinhole::project / epipolarConstrain                          src/CameraModels/Pinhole.cpp:31-50,155-177
//
// Where the reference goes through cv::Mat float algebra (3x3 products, K inverse) the oracle fixes
// one evaluation order in fp32 without FMA; OpenCV's own order is not reproducible here, so those
// quantities are tolerance-level w.r.t. the true reference (decisions can flip only within ~1e-5 px
// of a gate) but bit-exact between oracle and device.
#include "ork.h"
#include <algorithm>
#include <climits>
#include <cstring>

namespace ork {

static const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;

// ORBmatcher::DescriptorDistance: 8 x 32-bit SWAR popcount
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

struct Grid {
  float minX, minY, wInv, hInv;
  std::vector<int> cell[ORBX_GRID_COLS][ORBX_GRID_ROWS];
  const orbx_frame_desc* F;
  explicit Grid(const orbx_frame_desc* f) : F(f) {
    minX = f->min_x;
    minY = f->min_y;
    wInv = (float)ORBX_GRID_COLS / (f->max_x - f->min_x);   // src/Frame.cc:151-152
    hInv = (float)ORBX_GRID_ROWS / (f->max_y - f->min_y);
    for (int i = 0; i < f->n; ++i) {
      int px = (int)std::round((f->kps[i].x - minX) * wInv);   // PosInGrid
      int py = (int)std::round((f->kps[i].y - minY) * hInv);
      if (px < 0 || px >= ORBX_GRID_COLS || py < 0 || py >= ORBX_GRID_ROWS) continue;
      cell[px][py].push_back(i);
    }
  }
  void query(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
    out.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * wInv));
    if (nMinCellX >= ORBX_GRID_COLS) return;
    const int nMaxCellX = std::min(ORBX_GRID_COLS - 1, (int)std::ceil((x - minX + r) * wInv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * hInv));
    if (nMinCellY >= ORBX_GRID_ROWS) return;
    const int nMaxCellY = std::min(ORBX_GRID_ROWS - 1, (int)std::ceil((y - minY + r) * hInv));
    if (nMaxCellY < 0) return;
    const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
      for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
        for (int j : cell[ix][iy]) {
          const orbx_keypoint& kp = F->kps[j];
          if (checkLevels) {
            if (kp.octave < minLevel) continue;
            if (maxLevel >= 0 && kp.octave > maxLevel) continue;
          }
          const float dx = kp.x - x, dy = kp.y - y;
          if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(j);
        }
  }
};

static void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  ind1 = ind2 = ind3 = -1;
  for (int i = 0; i < L; ++i) {
    const int s = histo[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

static inline int rot_bin(float angleA, float angleB) {
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = angleA - angleB;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

}  // namespace ork

using namespace ork;

extern "C" {

int ork_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// Persistent grid handle (used by oracle/ref_stub/ref_matcher_glue.cpp: the reference's ORBmatcher.cc compiled
// unmodified calls Frame::GetFeaturesInArea, which lives in src/Frame.cc; the stand-in Frame answers it with this grid).
void* ork_grid_create(const orbx_frame_desc* F) { return new Grid(F); }
void ork_grid_destroy(void* g) { delete (Grid*)g; }
int ork_grid_query(const void* g, float x, float y, float r, int minL, int maxL, int32_t* out, int cap) {
  std::vector<int> v;
  ((const Grid*)g)->query(x, y, r, minL, maxL, v);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}

int ork_features_in_area(const orbx_frame_desc* F, int nq, const float* x, const float* y, const float* r,
                         const int32_t* minL, const int32_t* maxL, int32_t* out_idx, int cap, int32_t* out_n) {
  Grid g(F);
  std::vector<int> v;
  for (int q = 0; q < nq; ++q) {
    g.query(x[q], y[q], r[q], minL[q], maxL[q], v);
    out_n[q] = (int)v.size();
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out_idx[(size_t)q * cap + i] = v[i];
  }
  return ORBX_OK;
}

// Frame::ComputeStereoMatches.  pyrL/pyrR: per-level un-bordered images (mvImagePyramid[l]).
int ork_stereo_match(const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh,
                     const orbx_keypoint* kpL, const uint8_t* descL, int nL, const orbx_keypoint* kpR,
                     const uint8_t* descR, int nR, const float* scaleFactors, const float* invScaleFactors, float bf,
                     float b, float* uright, float* depth) {
  for (int i = 0; i < nL; ++i) uright[i] = depth[i] = -1.0f;
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  const int nRows = lh[0];
  std::vector<std::vector<int>> rowIdx(nRows);
  for (int iR = 0; iR < nR; ++iR) {
    const float kpY = kpR[iR].y;
    const float r = 2.0f * scaleFactors[kpR[iR].octave];
    const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; ++yi)
      if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);   // (the reference indexes unchecked)
  }
  const float minZ = b, minD = 0, maxD = bf / minZ;
  std::vector<std::pair<int, int>> distIdx;
  for (int iL = 0; iL < nL; ++iL) {
    const orbx_keypoint& kl = kpL[iL];
    const int levelL = kl.octave;
    const float vL = kl.y, uL = kl.x;
    const int row = (int)vL;
    if (row < 0 || row >= nRows) continue;
    const std::vector<int>& cands = rowIdx[row];
    if (cands.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH;
    int bestIdxR = 0;
    for (int iR : cands) {
      const orbx_keypoint& kr = kpR[iR];
      if (kr.octave < levelL - 1 || kr.octave > levelL + 1) continue;
      const float uR = kr.x;
      if (uR >= minU && uR <= maxU) {
        const int dist = descriptor_distance(descL + 32 * (size_t)iL, descR + 32 * (size_t)iR);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (bestDist < thOrbDist) {
      const float uR0 = kpR[bestIdxR].x;
      const float sf = invScaleFactors[kl.octave];
      const float scaleduL = std::round(kl.x * sf), scaledvL = std::round(kl.y * sf), scaleduR0 = std::round(uR0 * sf);
      const int w = 5, L = 5;
      const int W = lw[kl.octave], H = lh[kl.octave];
      const uint8_t* IL = pyrL[kl.octave];
      const uint8_t* IR = pyrR[kl.octave];
      const int cxL = (int)scaleduL, cy = (int)scaledvL, cxR = (int)scaleduR0;
      if (cy - w < 0 || cy + w >= H || cxL - w < 0 || cxL + w >= W) continue;   // cv::Mat range assert in the reference
      const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
      if (iniu < 0 || endu >= W) continue;
      if (cxR - L - w < 0) continue;   // would read the pyramid border / assert in the reference
      int bestSad = INT_MAX, bestinc = 0;
      float vDists[2 * 5 + 1];
      const int cL = IL[(size_t)cy * W + cxL];
      for (int inc = -L; inc <= L; ++inc) {
        const int cR = IR[(size_t)cy * W + cxR + inc];
        int sad = 0;
        for (int dy = -w; dy <= w; ++dy)
          for (int dx = -w; dx <= w; ++dx) {
            const int a = IL[(size_t)(cy + dy) * W + cxL + dx] - cL;
            const int c = IR[(size_t)(cy + dy) * W + cxR + inc + dx] - cR;
            sad += std::abs(a - c);
          }
        const float dist = (float)sad;
        if (dist < (float)bestSad) { bestSad = (int)dist; bestinc = inc; }
        vDists[L + inc] = dist;
      }
      if (bestinc == -L || bestinc == L) continue;
      const float d1 = vDists[L + bestinc - 1], d2 = vDists[L + bestinc], d3 = vDists[L + bestinc + 1];
      const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
      if (deltaR < -1 || deltaR > 1) continue;
      float bestuR = scaleFactors[kl.octave] * ((float)scaleduR0 + (float)bestinc + deltaR);
      float disparity = uL - bestuR;
      if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) {
          disparity = 0.01;
          bestuR = uL - 0.01;
        }
        depth[iL] = bf / disparity;
        uright[iL] = bestuR;
        distIdx.push_back({bestSad, iL});
      }
    }
  }
  if (distIdx.empty()) return ORBX_OK;
  std::sort(distIdx.begin(), distIdx.end());
  const float median = (float)distIdx[distIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  for (int i = (int)distIdx.size() - 1; i >= 0; --i) {
    if ((float)distIdx[i].first < thDist) break;
    uright[distIdx[i].second] = -1;
    depth[distIdx[i].second] = -1;
  }
  return ORBX_OK;
}

int ork_search_by_projection_map(const orbx_frame_desc* F, const uint8_t* kp_blocked, int nq, const float* projX,
                                 const float* projY, const float* projXR, const int32_t* level, const float* viewCos,
                                 const uint8_t* mpDesc, const uint8_t* flags, float th, float nnratio,
                                 const float* scaleFactors, int nlevels, int32_t* best_idx, int32_t* nmatches) {
  (void)nlevels;
  Grid g(F);
  std::vector<uint8_t> blocked(kp_blocked, kp_blocked + F->n);
  std::vector<int> cand;
  int n = 0;
  const bool bFactor = th != 1.0;
  for (int q = 0; q < nq; ++q) {
    best_idx[q] = -1;
    if (!(flags[q] & 1)) continue;
    const int lvl = level[q];
    float r = viewCos[q] > 0.998 ? 2.5f : 4.0f;   // RadiusByViewingCos (float vs double literal compare)
    if (bFactor) r *= th;
    g.query(projX[q], projY[q], r * scaleFactors[lvl], lvl - 1, lvl, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cand) {
      if (blocked[idx]) continue;
      if (F->uright && F->uright[idx] > 0) {
        const float er = std::fabs(projXR[q] - F->uright[idx]);
        if (er > r * scaleFactors[lvl]) continue;
      }
      const int dist = descriptor_distance(mpDesc + 32 * (size_t)q, F->desc + 32 * (size_t)idx);
      if (dist < bestDist) {
        bestDist2 = bestDist;
        bestDist = dist;
        bestLevel2 = bestLevel;
        bestLevel = F->kps[idx].octave;
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = F->kps[idx].octave;
        bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      best_idx[q] = bestIdx;
      blocked[bestIdx] = (flags[q] & 2) ? 1 : 0;   // F.mvpMapPoints[bestIdx] = pMP (Observations()>0 blocks)
      ++n;
    }
  }
  *nmatches = n;
  return ORBX_OK;
}

int ork_search_by_projection_frame(const orbx_frame_desc* C, const uint8_t* cur_blocked, const orbx_camera* cam,
                                   const float* Tc, const float* Tl, int nq, const uint8_t* flags, const float* xw,
                                   const int32_t* octave, const float* angle, const uint8_t* mpDesc, float th,
                                   int bMono, int checkOri, const float* scaleFactors, int nlevels,
                                   int32_t* match_idx, uint8_t* kept, int32_t* cur_match, int32_t* nmatches) {
  (void)nlevels;
  Grid g(C);
  std::vector<uint8_t> blocked(cur_blocked, cur_blocked + C->n);
  for (int i = 0; i < C->n; ++i) cur_match[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  // twc = -Rcw.t()*tcw: a transposed operand sends cv::gemm down its general path (products and sum in double, rounded to
  // float once; tests/test_ref_stub.py); tlc = Rlw*twc + tlw: the small-matrix fp32 path, fixed order
  float twc[3], tlc[3];
  for (int i = 0; i < 3; ++i)
    twc[i] = (float)(-((double)Tc[0 * 4 + i] * (double)Tc[3] + (double)Tc[1 * 4 + i] * (double)Tc[7] + (double)Tc[2 * 4 + i] * (double)Tc[11]));
  for (int i = 0; i < 3; ++i) tlc[i] = Tl[i * 4 + 0] * twc[0] + Tl[i * 4 + 1] * twc[1] + Tl[i * 4 + 2] * twc[2] + Tl[i * 4 + 3];
  const bool bForward = tlc[2] > cam->b && !bMono;
  const bool bBackward = -tlc[2] > cam->b && !bMono;
  std::vector<int> cand;
  int n = 0;
  for (int q = 0; q < nq; ++q) {
    match_idx[q] = -1;
    kept[q] = 0;
    if (!(flags[q] & 1)) continue;
    const float* X = xw + 3 * (size_t)q;
    const float xc = Tc[0] * X[0] + Tc[1] * X[1] + Tc[2] * X[2] + Tc[3];
    const float yc = Tc[4] * X[0] + Tc[5] * X[1] + Tc[6] * X[2] + Tc[7];
    const float zc = Tc[8] * X[0] + Tc[9] * X[1] + Tc[10] * X[2] + Tc[11];
    const float invzc = (float)(1.0 / zc);
    if (invzc < 0) continue;
    const float u = cam->fx * xc / zc + cam->cx, v = cam->fy * yc / zc + cam->cy;
    if (u < C->min_x || u > C->max_x) continue;
    if (v < C->min_y || v > C->max_y) continue;
    const int oct = octave[q];
    const float radius = th * scaleFactors[oct];
    if (bForward) g.query(u, v, radius, oct, -1, cand);
    else if (bBackward) g.query(u, v, radius, 0, oct, cand);
    else g.query(u, v, radius, oct - 1, oct + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (blocked[i2]) continue;
      if (C->uright && C->uright[i2] > 0) {
        const float ur = u - cam->bf * invzc;
        const float er = std::fabs(ur - C->uright[i2]);
        if (er > radius) continue;
      }
      const int dist = descriptor_distance(mpDesc + 32 * (size_t)q, C->desc + 32 * (size_t)i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= TH_HIGH) {
      cur_match[bestIdx2] = q;
      blocked[bestIdx2] = (flags[q] & 2) ? 1 : 0;
      match_idx[q] = bestIdx2;
      kept[q] = 1;
      ++n;
      if (checkOri) rotHist[rot_bin(angle[q], C->kps[bestIdx2].angle)].push_back(q);
    }
  }
  if (checkOri) {
    int h[HISTO_LENGTH], i1, i2, i3;
    for (int i = 0; i < HISTO_LENGTH; ++i) h[i] = (int)rotHist[i].size();
    three_maxima(h, HISTO_LENGTH, i1, i2, i3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == i1 || i == i2 || i == i3) continue;
      for (int q : rotHist[i]) {
        cur_match[match_idx[q]] = -1;   // CurrentFrame.mvpMapPoints[idx] = NULL
        kept[q] = 0;
        --n;
      }
    }
  }
  *nmatches = n;
  return ORBX_OK;
}

int ork_search_for_triangulation(const orbx_frame_desc* K1, const orbx_frame_desc* K2, const uint8_t* has1,
                                 const uint8_t* has2, int nn1, const int32_t* n1id, const int32_t* n1off,
                                 const int32_t* n1idx, int nn2, const int32_t* n2id, const int32_t* n2off,
                                 const int32_t* n2idx, const orbx_camera* cam1, const orbx_camera* cam2,
                                 const float* R1w, const float* t1w, const float* R2w, const float* t2w,
                                 const float* sigma2, const float* scaleFactors, int nlevels, int onlyStereo,
                                 int coarse, int checkOri, int32_t* match12, int32_t* nmatches) {
  (void)nlevels;
  // Cw = -R1w^T t1w ; C2 = R2w Cw + t2w ; ep = project(C2)
  float Cw[3], C2[3];
  for (int i = 0; i < 3; ++i) Cw[i] = -(R1w[0 * 3 + i] * t1w[0] + R1w[1 * 3 + i] * t1w[1] + R1w[2 * 3 + i] * t1w[2]);
  for (int i = 0; i < 3; ++i) C2[i] = R2w[i * 3 + 0] * Cw[0] + R2w[i * 3 + 1] * Cw[1] + R2w[i * 3 + 2] * Cw[2] + t2w[i];
  const float epx = cam2->fx * C2[0] / C2[2] + cam2->cx, epy = cam2->fy * C2[1] / C2[2] + cam2->cy;
  // R12 = R1w R2w^T ; t12 = -R12 t2w + t1w
  float R12[9], t12[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)   // R1w*R2w.t(): gemm's general path (double accumulation) because of the transposed operand
      R12[i * 3 + j] = (float)((double)R1w[i * 3 + 0] * (double)R2w[j * 3 + 0] + (double)R1w[i * 3 + 1] * (double)R2w[j * 3 + 1] +
                               (double)R1w[i * 3 + 2] * (double)R2w[j * 3 + 2]);
  for (int i = 0; i < 3; ++i) t12[i] = -(R12[i * 3 + 0] * t2w[0] + R12[i * 3 + 1] * t2w[1] + R12[i * 3 + 2] * t2w[2]) + t1w[i];
  // F12 = K1^-T [t12]x R12 K2^-1   (Pinhole::epipolarConstrain, computed once instead of per pair)
  float A[9];   // [t12]x R12
  const float tx[9] = {0, -t12[2], t12[1], t12[2], 0, -t12[0], -t12[1], t12[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i * 3 + j] = tx[i * 3 + 0] * R12[0 * 3 + j] + tx[i * 3 + 1] * R12[1 * 3 + j] + tx[i * 3 + 2] * R12[2 * 3 + j];
  // K^-1 = [1/fx 0 -cx/fx; 0 1/fy -cy/fy; 0 0 1]
  const float i1x = 1.0f / cam1->fx, i1y = 1.0f / cam1->fy, c1x = -cam1->cx * i1x, c1y = -cam1->cy * i1y;
  const float i2x = 1.0f / cam2->fx, i2y = 1.0f / cam2->fy, c2x = -cam2->cx * i2x, c2y = -cam2->cy * i2y;
  float Bm[9];  // K1^-T A : row0 = i1x*A0 ; row1 = i1y*A1 ; row2 = c1x*A0 + c1y*A1 + A2
  for (int j = 0; j < 3; ++j) {
    Bm[0 * 3 + j] = i1x * A[0 * 3 + j];
    Bm[1 * 3 + j] = i1y * A[1 * 3 + j];
    Bm[2 * 3 + j] = c1x * A[0 * 3 + j] + c1y * A[1 * 3 + j] + A[2 * 3 + j];
  }
  float F12[9];  // Bm K2^-1 : col0 = Bm[:,0]*i2x ; col1 = Bm[:,1]*i2y ; col2 = Bm[:,0]*c2x + Bm[:,1]*c2y + Bm[:,2]
  for (int i = 0; i < 3; ++i) {
    F12[i * 3 + 0] = Bm[i * 3 + 0] * i2x;
    F12[i * 3 + 1] = Bm[i * 3 + 1] * i2y;
    F12[i * 3 + 2] = Bm[i * 3 + 0] * c2x + Bm[i * 3 + 1] * c2y + Bm[i * 3 + 2];
  }
  for (int i = 0; i < K1->n; ++i) match12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  int n = 0;
  int a = 0, bq = 0;
  while (a < nn1 && bq < nn2) {
    if (n1id[a] == n2id[bq]) {
      for (int i1 = n1off[a]; i1 < n1off[a + 1]; ++i1) {
        const int idx1 = n1idx[i1];
        if (has1[idx1]) continue;
        const bool bStereo1 = K1->uright && K1->uright[idx1] >= 0;
        if (onlyStereo && !bStereo1) continue;
        const orbx_keypoint& kp1 = K1->kps[idx1];
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int i2 = n2off[bq]; i2 < n2off[bq + 1]; ++i2) {
          const int idx2 = n2idx[i2];
          if (has2[idx2]) continue;
          const bool bStereo2 = K2->uright && K2->uright[idx2] >= 0;
          if (onlyStereo && !bStereo2) continue;
          const int dist = descriptor_distance(K1->desc + 32 * (size_t)idx1, K2->desc + 32 * (size_t)idx2);
          if (dist > TH_LOW || dist > bestDist) continue;
          const orbx_keypoint& kp2 = K2->kps[idx2];
          if (!bStereo1 && !bStereo2) {
            const float dex = epx - kp2.x, dey = epy - kp2.y;
            if (dex * dex + dey * dey < 100 * scaleFactors[kp2.octave]) continue;
          }
          bool ok = coarse != 0;
          if (!ok) {
            const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
            const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
            const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
            const float num = la * kp2.x + lb * kp2.y + lc;
            const float den = la * la + lb * lb;
            if (den != 0) {
              const float dsqr = num * num / den;
              ok = dsqr < 3.84 * sigma2[kp2.octave];
            }
          }
          if (ok) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
          match12[idx1] = bestIdx2;
          ++n;
          if (checkOri) rotHist[rot_bin(kp1.angle, K2->kps[bestIdx2].angle)].push_back(idx1);
        }
      }
      ++a;
      ++bq;
    } else if (n1id[a] < n2id[bq]) {
      a = (int)(std::lower_bound(n1id, n1id + nn1, n2id[bq]) - n1id);
    } else {
      bq = (int)(std::lower_bound(n2id, n2id + nn2, n1id[a]) - n2id);
    }
  }
  if (checkOri) {
    int h[HISTO_LENGTH], i1, i2, i3;
    for (int i = 0; i < HISTO_LENGTH; ++i) h[i] = (int)rotHist[i].size();
    three_maxima(h, HISTO_LENGTH, i1, i2, i3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == i1 || i == i2 || i == i3) continue;
      for (int idx1 : rotHist[i]) { match12[idx1] = -1; --n; }
    }
  }
  *nmatches = n;
  return ORBX_OK;
}

}  // extern "C"

// ================================================================================================
// the callers either side of the path
// ================================================================================================
extern "C" {

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:323-591),
// pinhole / Nleft == -1 configuration.
//   kf_has_mp[i]  : vpMapPointsKF[i] != NULL && !isBad()
//   fv*           : FeatureVectors as CSR (ascending node ids)
// Out: match_f[F.n] = keyframe feature whose MapPoint was assigned to frame keypoint i, or -1; returns nmatches.
int ork_search_by_bow(const orbx_frame_desc* KF, const orbx_frame_desc* F, const uint8_t* kf_has_mp, int nnK,
                      const int32_t* fvK_node, const int32_t* fvK_off, const int32_t* fvK_idx, int nnF,
                      const int32_t* fvF_node, const int32_t* fvF_off, const int32_t* fvF_idx, float nnratio,
                      int check_orientation, int32_t* match_f, int32_t* nmatches) {
  for (int i = 0; i < F->n; ++i) match_f[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  int n = 0;
  int a = 0, b = 0;
  while (a < nnK && b < nnF) {
    if (fvK_node[a] == fvF_node[b]) {
      for (int iK = fvK_off[a]; iK < fvK_off[a + 1]; ++iK) {
        const int realIdxKF = fvK_idx[iK];
        if (!kf_has_mp[realIdxKF]) continue;
        const uint8_t* dKF = KF->desc + 32 * (size_t)realIdxKF;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int iF = fvF_off[b]; iF < fvF_off[b + 1]; ++iF) {
          const int realIdxF = fvF_idx[iF];
          if (match_f[realIdxF] >= 0) continue;
          const int dist = descriptor_distance(dKF, F->desc + 32 * (size_t)realIdxF);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            match_f[bestIdxF] = realIdxKF;
            if (check_orientation) rotHist[rot_bin(KF->kps[realIdxKF].angle, F->kps[bestIdxF].angle)].push_back(bestIdxF);
            ++n;
          }
        }
      }
      ++a; ++b;
    } else if (fvK_node[a] < fvF_node[b]) {
      a = (int)(std::lower_bound(fvK_node, fvK_node + nnK, fvF_node[b]) - fvK_node);
    } else {
      b = (int)(std::lower_bound(fvF_node, fvF_node + nnF, fvK_node[a]) - fvF_node);
    }
  }
  if (check_orientation) {
    int cnt[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) cnt[i] = (int)rotHist[i].size();
    int i1, i2, i3;
    three_maxima(cnt, HISTO_LENGTH, i1, i2, i3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == i1 || i == i2 || i == i3) continue;
      for (int idx : rotHist[i]) { match_f[idx] = -1; --n; }
    }
  }
  *nmatches = n;
  return ORBX_OK;
}

// ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th, bRight = false) (src/ORBmatcher.cc:1630-1883):
// the search half (projection, gates, best descriptor).  The map surgery that follows a hit (Replace / AddObservation /
// AddMapPoint, :1844-1867) stays with the caller, which replays it over best_idx in increasing i.
//   flags[i] bit0 : pMP && !isBad() && !IsInKeyFrame(pKF)
//   mp_max_dist / mp_min_dist : the raw members mfMaxDistance / mfMinDistance (the 1.2 / 0.8 invariance factors and
//                               PredictScale's ratio are applied here, src/MapPoint.cc:566-593)
//   Rcw[9], tcw[3], Ow[3]     : GetRotation / GetTranslation / GetCameraCenter, float32
//   log_scale_factor          : pKF->mfLogScaleFactor
// Out: best_idx[i] = keyframe keypoint (bestDist <= TH_LOW) or -1; returns the number of hits (nFused).
// cv::Mat float algebra (Rcw*p+tcw) is evaluated in fp32 left to right; cv::norm / Mat::dot accumulate in double like
// OpenCV's normL2_32f / dotProd_32f; PredictScale's log is evaluated in double (the reference's `log(float)` resolves
// to either overload depending on headers).
int ork_fuse(const orbx_frame_desc* KF, const orbx_camera* cam, const float* Rcw, const float* tcw, const float* Ow,
             int nmp, const uint8_t* flags, const float* xw, const float* mp_max_dist, const float* mp_min_dist,
             const float* mp_normal, const uint8_t* mp_desc, float th, const float* scale_factors,
             const float* inv_level_sigma2, int nlevels, float log_scale_factor, int32_t* best_idx, int32_t* nfused) {
  Grid g(KF);
  std::vector<int> v;
  int n = 0;
  for (int i = 0; i < nmp; ++i) {
    best_idx[i] = -1;
    if (!(flags[i] & 1)) continue;
    const float X = xw[3 * i], Y = xw[3 * i + 1], Z = xw[3 * i + 2];
    const float xc = Rcw[0] * X + Rcw[1] * Y + Rcw[2] * Z + tcw[0];
    const float yc = Rcw[3] * X + Rcw[4] * Y + Rcw[5] * Z + tcw[1];
    const float zc = Rcw[6] * X + Rcw[7] * Y + Rcw[8] * Z + tcw[2];
    if (zc < 0.0f) continue;
    const float invz = 1 / zc;
    const float u = cam->fx * xc / zc + cam->cx, vv = cam->fy * yc / zc + cam->cy;   // Pinhole::project
    if (!(u >= KF->min_x && u < KF->max_x && vv >= KF->min_y && vv < KF->max_y)) continue;   // IsInImage
    const float ur = u - cam->bf * invz;
    const float maxDistance = 1.2f * mp_max_dist[i], minDistance = 0.8f * mp_min_dist[i];
    const float PO[3] = {X - Ow[0], Y - Ow[1], Z - Ow[2]};
    const float dist3D = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    const double dot = (double)PO[0] * mp_normal[3 * i] + (double)PO[1] * mp_normal[3 * i + 1] + (double)PO[2] * mp_normal[3 * i + 2];
    if (dot < 0.5 * dist3D) continue;
    const float ratio = mp_max_dist[i] / dist3D;
    int level = (int)std::ceil(std::log((double)ratio) / (double)log_scale_factor);
    if (level < 0) level = 0; else if (level >= nlevels) level = nlevels - 1;
    const float radius = th * scale_factors[level];
    g.query(u, vv, radius, -1, -1, v);
    if (v.empty()) continue;
    const uint8_t* dMP = mp_desc + 32 * (size_t)i;
    int bestDist = 256, bestIdx = -1;
    for (int idx : v) {
      const orbx_keypoint& kp = KF->kps[idx];
      const int kpLevel = kp.octave;
      if (kpLevel < level - 1 || kpLevel > level) continue;
      if (KF->uright && KF->uright[idx] >= 0) {
        const float ex = u - kp.x, ey = vv - kp.y, er = ur - KF->uright[idx];
        const float e2 = ex * ex + ey * ey + er * er;
        if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
      } else {
        const float ex = u - kp.x, ey = vv - kp.y;
        const float e2 = ex * ex + ey * ey;
        if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
      }
      const int dist = descriptor_distance(dMP, KF->desc + 32 * (size_t)idx);
      if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    if (bestDist <= TH_LOW) { best_idx[i] = bestIdx; ++n; }
  }
  *nfused = n;
  return ORBX_OK;
}

}  // extern "C"
