// ork_extractor.cpp — ORACLE (test infrastructure): CPU restatement of the reference's
// ORBextractor (src/ORBextractor.cc).  Sequential, single thread, written for fidelity.
//
// Deliberate, documented choices where the reference's behaviour is not well defined:
//  * DistributeOctTree sorts (size, ExtractorNode*) pairs (src/ORBextractor.cc:682): ties in size
//    are broken by heap address.  The oracle breaks them by node creation sequence (later-created
//    node = higher "address", i.e. a bump allocator).
//  * Descriptor rotation (src/ORBextractor.cc:110-111) resolves to cosf/sinf via <cmath>; the oracle
//    evaluates cos/sin in double and rounds to float (= the correctly rounded float value except
//    with probability ~1e-8), so the device can reproduce it with its own double libm.
//  * x*b + y*a style expressions are evaluated WITHOUT fused multiply-add (-ffp-contract=off).
#include "ork.h"
#include <algorithm>
#include <cstring>
#include <list>

namespace ork {

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

static const int kPatch = 31, kHalfPatch = 15, kEdge = 19;

// ORBextractor::ORBextractor, src/ORBextractor.cc:408-468
Extractor::Extractor(int nf, float sf, int nl, int ini, int mn)
    : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
  scale.assign(nl, 1.f);
  sigma2.assign(nl, 1.f);
  for (int i = 1; i < nl; ++i) {
    scale[i] = (float)(scale[i - 1] * scaleFactor);  // float * double member -> double -> float
    sigma2[i] = scale[i] * scale[i];
  }
  invScale.resize(nl);
  invSigma2.resize(nl);
  for (int i = 0; i < nl; ++i) {
    invScale[i] = 1.0f / scale[i];
    invSigma2[i] = 1.0f / sigma2[i];
  }
  pyramid.resize(nl);
  cand.resize(nl);
  featuresPerLevel.resize(nl);
  float factor = (float)(1.0f / scaleFactor);
  float nDesired = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; ++l) {
    featuresPerLevel[l] = cv_round(nDesired);
    sum += featuresPerLevel[l];
    nDesired *= factor;
  }
  featuresPerLevel[nl - 1] = std::max(nf - sum, 0);

  umax.assign(kHalfPatch + 1, 0);
  int vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
  int vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
  const double hp2 = kHalfPatch * kHalfPatch;
  for (int v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
  for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
    while (umax[v0] == umax[v0 + 1]) ++v0;
    umax[v] = v0;
    ++v0;
  }
}

void Extractor::level_size(int w, int h, int level, int* lw, int* lh) const {
  float s = invScale[level];  // ComputePyramid, src/ORBextractor.cc:1162-1163
  *lw = cv_round((float)w * s);
  *lh = cv_round((float)h * s);
}

// IC_Angle, src/ORBextractor.cc:75-102
float ic_angle(const Gray& img, int x, int y, const std::vector<int>& umax) {
  int m01 = 0, m10 = 0;
  const uint8_t* c = img.row(y) + x;
  const int step = img.w;
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int vsum = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int p = c[u + v * step], m = c[u - v * step];
      vsum += p - m;
      m10 += u * (p + m);
    }
    m01 += v * vsum;
  }
  return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor, src/ORBextractor.cc:106-145
void orb_descriptor(const Gray& img, int x, int y, float angleDeg, uint8_t* desc) {
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
  float angle = angleDeg * factorPI;
  float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
  const uint8_t* c = img.row(y) + x;
  const int step = img.w;
  const int8_t* p = kPattern;
  for (int i = 0; i < 32; ++i) {
    int val = 0;
    for (int k = 0; k < 8; ++k, p += 4) {
      float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
      int t0 = c[cv_round(x0 * b + y0 * a) * step + cv_round(x0 * a - y0 * b)];
      int t1 = c[cv_round(x1 * b + y1 * a) * step + cv_round(x1 * a - y1 * b)];
      val |= (t0 < t1) << k;
    }
    desc[i] = (uint8_t)val;
  }
}

// ---------------------------------------------------------------------------------------------
// DistributeOctTree + ExtractorNode::DivideNode, src/ORBextractor.cc:479-761.
// ---------------------------------------------------------------------------------------------
namespace {
struct Node {
  int x0, x1, y0, y1;  // UL.x, UR.x, UL.y, BR.y
  std::vector<int> keys;  // indices into the candidate array, in candidate order
  bool noMore = false;
  long seq = 0;  // creation sequence = stand-in for the heap address used as sort tie-break
  std::list<Node>::iterator self;
};

void divide(const Node& n, const std::vector<orbx_keypoint>& K, Node out[4]) {
  const int halfX = (int)std::ceil((float)(n.x1 - n.x0) / 2);
  const int halfY = (int)std::ceil((float)(n.y1 - n.y0) / 2);
  const int mx = n.x0 + halfX, my = n.y0 + halfY;
  out[0] = Node{n.x0, mx, n.y0, my};
  out[1] = Node{mx, n.x1, n.y0, my};
  out[2] = Node{n.x0, mx, my, n.y1};
  out[3] = Node{mx, n.x1, my, n.y1};
  for (int k : n.keys) {
    const orbx_keypoint& kp = K[k];
    int q = (kp.x < (float)mx ? 0 : 1) + (kp.y < (float)my ? 0 : 2);
    out[q].keys.push_back(k);
  }
  for (int q = 0; q < 4; ++q)
    if (out[q].keys.size() == 1) out[q].noMore = true;
}
}  // namespace

std::vector<orbx_keypoint> distribute_octree(const std::vector<orbx_keypoint>& K, int minX, int maxX,
                                             int minY, int maxY, int N) {
  std::vector<orbx_keypoint> result;
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  if (nIni < 1) return result;  // the reference would index an empty vector here
  const float hX = (float)(maxX - minX) / nIni;
  std::list<Node> L;
  long seq = 0;
  std::vector<Node*> ini(nIni);
  for (int i = 0; i < nIni; ++i) {
    Node n{(int)(hX * (float)i), (int)(hX * (float)(i + 1)), 0, maxY - minY};
    n.seq = seq++;
    L.push_back(n);
    ini[i] = &L.back();
  }
  for (size_t i = 0; i < K.size(); ++i) ini[(int)(K[i].x / hX)]->keys.push_back((int)i);
  for (auto it = L.begin(); it != L.end();) {
    if (it->keys.size() == 1) { it->noMore = true; ++it; }
    else if (it->keys.empty()) it = L.erase(it);
    else ++it;
  }

  typedef std::pair<int, long> SizeSeq;  // (size, creation seq): the (size, pointer) pair
  std::vector<std::pair<SizeSeq, Node*>> pending;
  auto push_children = [&](Node ch[4], bool count, int& nToExpand) {
    for (int q = 0; q < 4; ++q) {
      if (ch[q].keys.empty()) continue;
      ch[q].seq = seq++;
      L.push_front(ch[q]);
      if (ch[q].keys.size() > 1) {
        if (count) ++nToExpand;
        pending.push_back({{(int)ch[q].keys.size(), L.front().seq}, &L.front()});
        L.front().self = L.begin();
      }
    }
  };

  bool finish = false;
  while (!finish) {
    int prevSize = (int)L.size();
    int nToExpand = 0;
    pending.clear();
    for (auto it = L.begin(); it != L.end();) {
      if (it->noMore) { ++it; continue; }
      Node ch[4];
      divide(*it, K, ch);
      push_children(ch, true, nToExpand);
      it = L.erase(it);
    }
    if ((int)L.size() >= N || (int)L.size() == prevSize) {
      finish = true;
    } else if ((int)L.size() + nToExpand * 3 > N) {
      while (!finish) {
        prevSize = (int)L.size();
        auto prev = pending;
        pending.clear();
        std::sort(prev.begin(), prev.end(),
                  [](const std::pair<SizeSeq, Node*>& a, const std::pair<SizeSeq, Node*>& b) {
                    return a.first < b.first;
                  });
        for (int j = (int)prev.size() - 1; j >= 0; --j) {
          Node ch[4];
          divide(*prev[j].second, K, ch);
          int dummy = 0;
          push_children(ch, false, dummy);
          L.erase(prev[j].second->self);
          if ((int)L.size() >= N) break;
        }
        if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;
      }
    }
  }
  result.reserve(L.size());
  for (const Node& n : L) {
    int best = n.keys[0];
    float maxR = K[best].response;
    for (size_t k = 1; k < n.keys.size(); ++k)
      if (K[n.keys[k]].response > maxR) { best = n.keys[k]; maxR = K[best].response; }
    result.push_back(K[best]);
  }
  return result;
}

// ORBextractor::operator() + ComputePyramid + ComputeKeyPointsOctTree,
// src/ORBextractor.cc:1074-1183, :763-878
int Extractor::extract(const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                       std::vector<orbx_keypoint>& kps, std::vector<uint8_t>& desc, int* monoIndex) {
  kps.clear();
  desc.clear();
  if (monoIndex) *monoIndex = 0;
  if (!img || w <= 0 || h <= 0) return ORBX_EMPTY;
  {
    int lw, lh;
    level_size(w, h, nlevels - 1, &lw, &lh);
    // the reference divides by nCols = (int)((cols-32)/30) and by nIni = round(W/H): undefined below
    if (lw - 2 * (kEdge - 3) < 30 || lh - 2 * (kEdge - 3) < 30) return ORBX_EINVAL;
    if ((int)std::round((float)(lw - 32) / (lh - 32)) < 1) return ORBX_EINVAL;
    if ((int)std::round((float)(w - 32) / (h - 32)) < 1) return ORBX_EINVAL;
  }
  // --- ComputePyramid ---
  for (int l = 0; l < nlevels; ++l) {
    int lw, lh;
    level_size(w, h, l, &lw, &lh);
    pyramid[l] = Gray(lw, lh);
    if (l == 0) {
      for (int y = 0; y < h; ++y) std::memcpy(pyramid[0].row(y), img + (size_t)y * stride, w);
    } else {
      resize_linear_u8(pyramid[l - 1].px.data(), pyramid[l - 1].w, pyramid[l - 1].h, pyramid[l - 1].w,
                       pyramid[l].px.data(), lw, lh, lw);
    }
  }
  // --- ComputeKeyPointsOctTree ---
  std::vector<std::vector<orbx_keypoint>> all(nlevels);
  const float W = 30;
  std::vector<FastPoint> cell;
  for (int l = 0; l < nlevels; ++l) {
    const Gray& im = pyramid[l];
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = im.w - kEdge + 3, maxBY = im.h - kEdge + 3;
    std::vector<orbx_keypoint>& toDist = cand[l];
    toDist.clear();
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    for (int i = 0; i < nRows; ++i) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; ++j) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
        fast9_16(im.row(y0) + x0, cw, ch, im.w, iniTh, true, cell);
        if (cell.empty()) fast9_16(im.row(y0) + x0, cw, ch, im.w, minTh, true, cell);
        for (const FastPoint& p : cell) {
          orbx_keypoint kp;
          kp.x = (float)p.x + (float)(j * wCell);
          kp.y = (float)p.y + (float)(i * hCell);
          kp.size = 7.f;
          kp.angle = -1.f;
          kp.response = (float)p.score;
          kp.octave = 0;
          toDist.push_back(kp);
        }
      }
    }
    all[l] = distribute_octree(toDist, minBX, maxBX, minBY, maxBY, featuresPerLevel[l]);
    const int scaledPatch = (int)(kPatch * scale[l]);
    for (orbx_keypoint& kp : all[l]) {
      kp.x += minBX;
      kp.y += minBY;
      kp.octave = l;
      kp.size = (float)scaledPatch;
    }
  }
  for (int l = 0; l < nlevels; ++l)
    for (orbx_keypoint& kp : all[l]) kp.angle = ic_angle(pyramid[l], cv_round(kp.x), cv_round(kp.y), umax);

  // --- descriptors + output ordering, src/ORBextractor.cc:1088-1155 ---
  int n = 0;
  for (int l = 0; l < nlevels; ++l) n += (int)all[l].size();
  kps.resize(n);
  desc.assign((size_t)n * 32, 0);
  int mono = 0, stereo = n - 1;
  Gray blurred;
  for (int l = 0; l < nlevels; ++l) {
    if (all[l].empty()) continue;
    blurred = Gray(pyramid[l].w, pyramid[l].h);
    gaussian_blur7_s2(pyramid[l].px.data(), pyramid[l].w, pyramid[l].h, pyramid[l].w, blurred.px.data(),
                      blurred.w);
    const float s = scale[l];
    for (orbx_keypoint kp : all[l]) {
      uint8_t d[32];
      orb_descriptor(blurred, cv_round(kp.x), cv_round(kp.y), kp.angle, d);
      if (l != 0) { kp.x *= s; kp.y *= s; }
      int slot = (kp.x >= (float)lap0 && kp.x <= (float)lap1) ? stereo-- : mono++;
      kps[slot] = kp;
      std::memcpy(&desc[(size_t)slot * 32], d, 32);
    }
  }
  if (monoIndex) *monoIndex = mono;
  return ORBX_OK;
}

}  // namespace ork
