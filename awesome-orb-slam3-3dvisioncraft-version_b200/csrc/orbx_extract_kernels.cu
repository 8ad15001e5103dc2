// orbx_extract_kernels.cu — sm_100a kernels of the batched ORB extractor.
//
// Replaces (reference paths): src/ORBextractor.cc ComputePyramid :1158-1183, ComputeKeyPointsOctTree
// :763-878, DistributeOctTree :537-761, IC_Angle :75-102, GaussianBlur call :1120-1121,
// computeOrbDescriptor :106-145 and the output ordering of operator() :1108-1155.
//
// All pixel arithmetic is integer / fixed-point and reproduces OpenCV 4.x bit for bit (see the
// oracle's comments for the formulas); the few fp32 expressions are written with explicit
// __fmul_rn/__fadd_rn so nvcc cannot contract them into FMAs (parity hazard).
#include "orbx_extract.cuh"

__constant__ __align__(16) int8_t c_pattern[1024] = {
#include "orb_pattern.inc"
};
// umax[v]: half-width of the r=15 disc at row v (src/ORBextractor.cc:452-467); same for every extractor.
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// =====================================================================================
// K1  pyramid level l from level l-1: cv::resize(INTER_LINEAR) 8U fixed point.
// grid (ceil(w/4/128), h, B), block 128; each thread produces 4 horizontally adjacent pixels.
// tab layout per level: X: [xofs(w) | a0(w) | a1(w)]  Y: [y0(h) | y1(h) | b0(h) | b1(h)]  (int16)
// =====================================================================================
__global__ void __launch_bounds__(128) pyr_resize_kernel(const __grid_constant__ ExtractParams p, int level) {
  const LevelParams& D = p.lv[level];
  const LevelParams& S = p.lv[level - 1];
  const int dx0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int dy = blockIdx.y;
  if (dx0 >= D.w) return;
  const int16_t* tx = p.tab + D.tabX;
  const int16_t* ty = p.tab + D.tabY;
  const int y0 = ty[dy], y1 = ty[D.h + dy], b0 = ty[2 * D.h + dy], b1 = ty[3 * D.h + dy];
  const uint8_t* src = S.pyr + (size_t)blockIdx.z * S.imgStride;
  const uint8_t* S0 = src + (size_t)y0 * S.pitch;
  const uint8_t* S1 = src + (size_t)y1 * S.pitch;
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int dx = dx0 + i;
    if (dx < D.w) {
      int x0 = tx[dx], a0 = tx[D.w + dx], a1 = tx[2 * D.w + dx];
      int x1 = min(x0 + 1, S.w - 1);
      int T0 = __ldg(S0 + x0) * a0 + __ldg(S0 + x1) * a1;
      int T1 = __ldg(S1 + x0) * a0 + __ldg(S1 + x1) * a1;
      int v = (((b0 * (T0 >> 4)) >> 16) + ((b1 * (T1 >> 4)) >> 16) + 2) >> 2;
      out |= (uint32_t)(v & 0xff) << (8 * i);
    }
  }
  uint8_t* dst = D.pyr + (size_t)blockIdx.z * D.imgStride + (size_t)dy * D.pitch + dx0;
  *reinterpret_cast<uint32_t*>(dst) = out;  // pitch is a multiple of 16 and padded: safe past w
}

// =====================================================================================
// K2  per-cell FAST-9-16 + cell-local 3x3 NMS + iniTh->minTh fallback + candidate emission.
// One CTA = one cell row x up to ORBX_FAST_CELLS cells of one level of one image.
// =====================================================================================
__device__ __forceinline__ int arc9_maxmin(const int (&d)[16]) {
  int t[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) t[k] = min(min(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
  int best = -256;
#pragma unroll
  for (int k = 0; k < 16; ++k) best = max(best, min(min(t[k], t[(k + 3) & 15]), t[(k + 6) & 15]));
  return best;
}

// 16-bit cyclic mask has >= 9 consecutive ones?
__device__ __forceinline__ bool has_arc9(uint32_t m) {
  m |= m << 16;
  uint32_t r = m & (m >> 1);
  r &= r >> 2;
  r &= r >> 4;   // 8 consecutive
  r &= m >> 8;   // 9 consecutive
  return (r & 0xffffu) != 0;
}


__global__ void __launch_bounds__(256) fast_cells_kernel(const __grid_constant__ ExtractParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  // locate (level, cell row, first cell)
  const int tile = blockIdx.x;
  int level = 0;
#pragma unroll 1
  for (int l = 1; l < p.nlevels; ++l)
    if (tile >= p.lv[l].tileStart) level = l;
  const LevelParams& L = p.lv[level];
  const int lt = tile - L.tileStart;
  const int ci = lt / L.tilesPerRow;
  const int j0 = (lt - ci * L.tilesPerRow) * ORBX_FAST_CELLS;
  const int b = blockIdx.y;

  const int iniY = ORBX_MINB + ci * L.hCell;
  if (iniY >= L.maxBY - 3) return;
  const int maxY = min(iniY + L.hCell + 6, L.maxBY);
  const int iniX = ORBX_MINB + j0 * L.wCell;
  if (iniX >= L.maxBX - 6) return;
  // cells j0 .. j0+nc-1 ; the last one of the level may be truncated or skipped
  int nc = min(ORBX_FAST_CELLS, L.nCols - j0);
  while (nc > 0 && ORBX_MINB + (j0 + nc - 1) * L.wCell >= L.maxBX - 6) --nc;
  if (nc <= 0) return;
  const int maxX = min(iniX + nc * L.wCell + 6, L.maxBX);
  const int tw = maxX - iniX, th = maxY - iniY;       // tile incl. 3-px ring halo
  const int wI = tw - 6, hI = th - 6;                 // interior (evaluated) pixels
  if (wI <= 0 || hI <= 0) return;

  // shared: image tile [th][tp] | score [(hI+2)][sp] | flag [hI][sp] | cellHasIni[ORBX_FAST_CELLS]
  const int tp = (tw + 3) & ~3;
  const int sp = wI + 2;
  int* shas = reinterpret_cast<int*>(smem);
  uint8_t* simg = smem + 64;
  uint8_t* ssc = simg + (size_t)p.fastTileBytes;
  uint8_t* sfl = ssc + (size_t)p.fastTileBytes;

  const uint8_t* img = L.pyr + (size_t)b * L.imgStride;
  for (int i = threadIdx.x; i < th * tw; i += blockDim.x) {
    int y = i / tw, x = i - y * tw;
    simg[y * tp + x] = __ldg(img + (size_t)(iniY + y) * L.pitch + iniX + x);
  }
  for (int i = threadIdx.x; i < (hI + 2) * sp; i += blockDim.x) ssc[i] = 0;
  if (threadIdx.x < ORBX_FAST_CELLS) shas[threadIdx.x] = 0;
  __syncthreads();

  const int t = p.minTh;
  for (int i = threadIdx.x; i < hI * wI; i += blockDim.x) {
    const int y = i / wI, x = i - y * wI;
    const uint8_t* c = simg + (y + 3) * tp + (x + 3);
    const int v = c[0];
    int d[16];
    d[0] = v - c[3 * tp];      d[1] = v - c[3 * tp + 1];   d[2] = v - c[2 * tp + 2];   d[3] = v - c[tp + 3];
    d[4] = v - c[3];           d[5] = v - c[-tp + 3];      d[6] = v - c[-2 * tp + 2];  d[7] = v - c[-3 * tp + 1];
    d[8] = v - c[-3 * tp];     d[9] = v - c[-3 * tp - 1];  d[10] = v - c[-2 * tp - 2]; d[11] = v - c[-tp - 3];
    d[12] = v - c[-3];         d[13] = v - c[tp - 3];      d[14] = v - c[2 * tp - 2];  d[15] = v - c[3 * tp - 1];
    uint32_t mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      mb |= (uint32_t)(d[k] > t) << k;
      md |= (uint32_t)(d[k] < -t) << k;
    }
    int score = 0;
    if (has_arc9(mb)) {
      score = arc9_maxmin(d) - 1;
    } else if (has_arc9(md)) {
#pragma unroll
      for (int k = 0; k < 16; ++k) d[k] = -d[k];
      score = arc9_maxmin(d) - 1;
    }
    if (score > 0) ssc[(y + 1) * sp + x + 1] = (uint8_t)score;
  }
  __syncthreads();

  // NMS inside the pixel's own cell: horizontal neighbours across a cell seam count as 0
  for (int i = threadIdx.x; i < hI * wI; i += blockDim.x) {
    const int y = i / wI, x = i - y * wI;
    const uint8_t* s = ssc + (y + 1) * sp + x + 1;
    const int sc = s[0];
    uint8_t keep = 0;
    if (sc > 0) {
      const int cell = x / L.wCell, xin = x - cell * L.wCell;
      const bool hasL = xin != 0, hasR = (xin != L.wCell - 1) && (x + 1 < wI);
      bool ok = sc > s[-sp] && sc > s[sp];
      if (hasL) ok = ok && sc > s[-1] && sc > s[-sp - 1] && sc > s[sp - 1];
      if (hasR) ok = ok && sc > s[1] && sc > s[-sp + 1] && sc > s[sp + 1];
      if (ok) {
        keep = 1;
        if (sc >= p.iniTh) atomicOr(&shas[cell], 1);
      }
    }
    sfl[y * sp + x] = keep;
  }
  __syncthreads();

  uint32_t* cand = p.cand + (size_t)b * p.candPerImage + L.candOfs;
  int* candN = p.candN + b * p.nlevels + level;
  for (int i = threadIdx.x; i < hI * wI; i += blockDim.x) {
    const int y = i / wI, x = i - y * wI;
    if (!sfl[y * sp + x]) continue;
    const int sc = ssc[(y + 1) * sp + x + 1];
    const int cell = x / L.wCell;
    if (shas[cell] && sc < p.iniTh) continue;
    const int slot = atomicAdd(candN, 1);
    if (slot < L.candCap) {
      // coordinates relative to (minBorderX, minBorderY), as the reference stores them (:847-848)
      const int rx = iniX + 3 + x - ORBX_MINB, ry = iniY + 3 + y - ORBX_MINB;
      cand[slot] = (uint32_t)rx | ((uint32_t)ry << 12) | ((uint32_t)sc << 24);
    } else {
      atomicExch(p.err, 1);
    }
  }
}

// =====================================================================================
// K5  7x7 sigma-2 Gaussian, OpenCV 4.x fixed point: [18,34,48,56,48,34,18]/256 per pass,
//     (sum + 2^15) >> 16, BORDER_REFLECT_101.
// =====================================================================================
__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) gauss7_kernel(const __grid_constant__ ExtractParams p) {
  __shared__ uint8_t sin_[(ORBX_BLUR_TH + 6) * (ORBX_BLUR_TW + 8)];
  __shared__ uint16_t sh[(ORBX_BLUR_TH + 6) * ORBX_BLUR_TW];
  const int tile = blockIdx.x;
  int level = 0;
#pragma unroll 1
  for (int l = 1; l < p.nlevels; ++l)
    if (tile >= p.lv[l].blurTileStart) level = l;
  const LevelParams& L = p.lv[level];
  const int lt = tile - L.blurTileStart;
  const int tyi = lt / L.blurTilesX, txi = lt - tyi * L.blurTilesX;
  const int x0 = txi * ORBX_BLUR_TW, y0 = tyi * ORBX_BLUR_TH;
  const uint8_t* img = L.pyr + (size_t)blockIdx.y * L.imgStride;
  const int PW = ORBX_BLUR_TW + 8;
  for (int i = threadIdx.x; i < (ORBX_BLUR_TH + 6) * (ORBX_BLUR_TW + 6); i += blockDim.x) {
    int y = i / (ORBX_BLUR_TW + 6), x = i - y * (ORBX_BLUR_TW + 6);
    int gy = reflect101(y0 + y - 3, L.h), gx = reflect101(x0 + x - 3, L.w);
    // tiles hanging over the right/bottom edge read reflected garbage that is never written out
    gy = min(max(gy, 0), L.h - 1);
    gx = min(max(gx, 0), L.w - 1);
    sin_[y * PW + x] = __ldg(img + (size_t)gy * L.pitch + gx);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (ORBX_BLUR_TH + 6) * ORBX_BLUR_TW; i += blockDim.x) {
    int y = i / ORBX_BLUR_TW, x = i - y * ORBX_BLUR_TW;
    const uint8_t* s = sin_ + y * PW + x;
    int acc = 18 * (s[0] + s[6]) + 34 * (s[1] + s[5]) + 48 * (s[2] + s[4]) + 56 * s[3];
    sh[i] = (uint16_t)acc;
  }
  __syncthreads();
  uint8_t* out = L.blur + (size_t)blockIdx.y * L.imgStride;
  for (int i = threadIdx.x; i < ORBX_BLUR_TH * (ORBX_BLUR_TW / 4); i += blockDim.x) {
    int y = i / (ORBX_BLUR_TW / 4), x = (i - y * (ORBX_BLUR_TW / 4)) * 4;
    if (y0 + y >= L.h || x0 + x >= L.w) continue;
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint16_t* s = sh + y * ORBX_BLUR_TW + x + k;
      uint32_t acc = 18u * (s[0] + s[6 * ORBX_BLUR_TW]) + 34u * (s[ORBX_BLUR_TW] + s[5 * ORBX_BLUR_TW]) +
                     48u * (s[2 * ORBX_BLUR_TW] + s[4 * ORBX_BLUR_TW]) + 56u * s[3 * ORBX_BLUR_TW];
      o |= ((acc + 32768u) >> 16) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(out + (size_t)(y0 + y) * L.pitch + x0 + x) = o;
  }
}

// =====================================================================================
// K3  DistributeOctTree: one CTA per (level, image).  The reference's std::list is kept as an
// array in list order that is rebuilt every pass; a node's "heap address" tie-break is its
// creation order, which in this representation is the reverse of its position among the nodes
// created in the same pass.
// =====================================================================================
#define OCT_NT 256

struct OctSmem {
  short4* bnd[2];     // x0,x1,y0,y1 (double buffered)
  int* cnt[2];
  int* child;         // [nodeCap*4] tentative child counts
  int* a;             // scratch arrays [nodeCap]
  int* bb;
  int* c;
  int* rankPos;       // [nodeCap] positions in processing order
  unsigned long long* best;
};

// exclusive scan over data[0..n) in shared memory, in place; returns the total to every thread
__device__ int block_excl_scan(int* data, int n, int* s_warp /*[33]*/) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (n + OCT_NT - 1) / OCT_NT;
  const int beg = min(tid * per, n), end = min(beg + per, n);
  int sum = 0;
  for (int i = beg; i < end; ++i) sum += data[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = lane < OCT_NT / 32 ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    if (lane < OCT_NT / 32) s_warp[lane] = wi - w;
    if (lane == OCT_NT / 32 - 1) s_warp[32] = wi;
  }
  __syncthreads();
  int run = s_warp[wid] + incl - sum;
  for (int i = beg; i < end; ++i) {
    int v = data[i];
    data[i] = run;
    run += v;
  }
  const int total = s_warp[32];
  __syncthreads();
  return total;
}

__device__ __forceinline__ int oct_quadrant(uint32_t c, short4 b) {
  const int x = c & 0xfff, y = (c >> 12) & 0xfff;
  const int mx = b.x + ((b.y - b.x + 1) >> 1), my = b.z + ((b.w - b.z + 1) >> 1);
  return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}

__device__ __forceinline__ short4 oct_child_bounds(short4 b, int q) {
  const short mx = b.x + ((b.y - b.x + 1) >> 1), my = b.z + ((b.w - b.z + 1) >> 1);
  short4 r;
  r.x = (q & 1) ? mx : b.x;
  r.y = (q & 1) ? b.y : mx;
  r.z = (q & 2) ? my : b.z;
  r.w = (q & 2) ? b.w : my;
  return r;
}

__global__ void __launch_bounds__(OCT_NT) octree_kernel(const __grid_constant__ ExtractParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int level = blockIdx.x, b = blockIdx.y;
  const LevelParams& L = p.lv[level];
  const int NC = p.nodeCap;
  const int tid = threadIdx.x;
  __shared__ int s_warp[33];
  __shared__ int s_i[8];

  OctSmem S;
  {
    uint8_t* q = smem;
    S.bnd[0] = (short4*)q; q += sizeof(short4) * NC;
    S.bnd[1] = (short4*)q; q += sizeof(short4) * NC;
    S.best = (unsigned long long*)q; q += sizeof(unsigned long long) * NC;
    S.cnt[0] = (int*)q; q += 4 * NC;
    S.cnt[1] = (int*)q; q += 4 * NC;
    S.child = (int*)q; q += 16 * NC;
    S.a = (int*)q; q += 4 * NC;
    S.bb = (int*)q; q += 4 * NC;
    S.c = (int*)q; q += 4 * NC;
    S.rankPos = (int*)q; q += 4 * NC;
  }
  const uint32_t* cand = p.cand + (size_t)b * p.candPerImage + L.candOfs;
  uint16_t* keyNode = p.keyNode + (size_t)b * p.candPerImage + L.candOfs;
  const int n = min(p.candN[b * p.nlevels + level], L.candCap);
  uint2* sel = p.sel + (size_t)b * p.selPerImage + L.selOfs;
  const int N = L.nFeat;

  int cur = 0;      // which node buffer holds the current list
  int size = 0;     // number of nodes in the list
  int nNew = 0;     // nodes [0, nNew) of the list were created by the last pass

  // ---- root nodes (src/ORBextractor.cc:541-577) ----
  for (int i = tid; i < L.nIni; i += OCT_NT) {
    short4 bd;
    bd.x = (short)(int)__fmul_rn(L.hX, (float)i);
    bd.y = (short)(int)__fmul_rn(L.hX, (float)(i + 1));
    bd.z = 0;
    bd.w = (short)(L.maxBY - ORBX_MINB);
    S.bnd[0][i] = bd;
    S.a[i] = 0;
  }
  __syncthreads();
  for (int k = tid; k < n; k += OCT_NT) {
    const int x = cand[k] & 0xfff;
    int r = (int)__fdiv_rn((float)x, L.hX);
    r = min(r, L.nIni - 1);
    keyNode[k] = (uint16_t)r;
    atomicAdd(&S.a[r], 1);
  }
  __syncthreads();
  // drop empty roots (list order = root order)
  for (int i = tid; i < L.nIni; i += OCT_NT) S.bb[i] = S.a[i] > 0;
  __syncthreads();
  size = block_excl_scan(S.bb, L.nIni, s_warp);
  for (int i = tid; i < L.nIni; i += OCT_NT)
    if (S.a[i] > 0) {
      S.bnd[1][S.bb[i]] = S.bnd[0][i];
      S.cnt[1][S.bb[i]] = S.a[i];
    }
  __syncthreads();
  for (int k = tid; k < n; k += OCT_NT) keyNode[k] = (uint16_t)S.bb[keyNode[k]];
  cur = 1;
  nNew = 0;
  __syncthreads();

  bool outer = true;   // true: next pass splits every multi-key node; false: inner (largest first)
  bool finish = (n == 0);
  while (!finish) {
    short4* bnd = S.bnd[cur];
    int* cnt = S.cnt[cur];
    short4* nbnd = S.bnd[cur ^ 1];
    int* ncnt = S.cnt[cur ^ 1];
    const int prevSize = size;

    // --- tentative child counts of every node that may be split in this pass ---
    // outer: all nodes with cnt>1 ; inner: nodes created by the last pass with cnt>1
    const int lim = outer ? size : nNew;
    for (int i = tid; i < lim * 4; i += OCT_NT) S.child[i] = 0;
    __syncthreads();
    for (int k = tid; k < n; k += OCT_NT) {
      const int pos = keyNode[k];
      if (pos < lim && cnt[pos] > 1) atomicAdd(&S.child[pos * 4 + oct_quadrant(cand[k], bnd[pos])], 1);
    }
    __syncthreads();

    // --- processing order: rankPos[r] = list position of the r-th node to split; P = how many ---
    int P;
    {
      for (int i = tid; i < lim; i += OCT_NT) S.a[i] = cnt[i] > 1;
      __syncthreads();
      P = block_excl_scan(S.a, lim, s_warp);      // S.a = index among candidates, list order
      if (outer) {
        for (int i = tid; i < lim; i += OCT_NT)
          if (cnt[i] > 1) S.rankPos[S.a[i]] = i;
        __syncthreads();
      } else {
        // sort by (cnt desc, creation seq desc) = (cnt desc, position asc): rank by counting
        for (int i = tid; i < lim; i += OCT_NT)
          if (cnt[i] > 1) S.bb[S.a[i]] = i;      // compact list of pending positions
        __syncthreads();
        for (int i = tid; i < P; i += OCT_NT) {
          const int pi = S.bb[i], ci = cnt[pi];
          int r = 0;
          for (int j = 0; j < P; ++j) {
            const int pj = S.bb[j], cj = cnt[pj];
            r += (cj > ci) || (cj == ci && pj < pi);
          }
          S.rankPos[r] = pi;
        }
        __syncthreads();
        // split in that order until the list reaches N nodes (src/ORBextractor.cc:684-735)
        for (int r = tid; r < P; r += OCT_NT) {
          const int* ch = &S.child[S.rankPos[r] * 4];
          S.c[r] = (ch[0] > 0) + (ch[1] > 0) + (ch[2] > 0) + (ch[3] > 0) - 1;
        }
        if (tid == 0) s_i[0] = P;
        __syncthreads();
        block_excl_scan(S.c, P, s_warp);          // S.c[r] = gain of the nodes before r
        for (int r = tid; r < P; r += OCT_NT) {
          const int* ch = &S.child[S.rankPos[r] * 4];
          const int incl = S.c[r] + (ch[0] > 0) + (ch[1] > 0) + (ch[2] > 0) + (ch[3] > 0) - 1;
          if (size + incl >= N) atomicMin(&s_i[0], r + 1);
        }
        __syncthreads();
        P = s_i[0];
        __syncthreads();
      }
    }
    if (P == 0) break;   // nothing to split: list unchanged (size == prevSize)

    // --- creation order of the children: nodes in processing order, quadrants 0..3, non-empty ---
    for (int r = tid; r < P; r += OCT_NT) {
      const int* ch = &S.child[S.rankPos[r] * 4];
      S.c[r] = (ch[0] > 0) + (ch[1] > 0) + (ch[2] > 0) + (ch[3] > 0);
    }
    __syncthreads();
    const int nChildren = block_excl_scan(S.c, P, s_warp);   // S.c[r] = creation index base
    // mark processed nodes: S.a[pos] = r+1 (0 = kept)
    for (int i = tid; i < size; i += OCT_NT) S.a[i] = 0;
    __syncthreads();
    for (int r = tid; r < P; r += OCT_NT) S.a[S.rankPos[r]] = r + 1;
    __syncthreads();
    // kept nodes keep their relative order behind the new children
    for (int i = tid; i < size; i += OCT_NT) S.bb[i] = S.a[i] == 0;
    __syncthreads();
    const int nKept = block_excl_scan(S.bb, size, s_warp);
    const int newSize = nChildren + nKept;
    if (newSize > NC) {   // cannot happen (size <= N+2), but never write out of bounds
      if (tid == 0) atomicExch(p.err, 2);
      break;
    }
    for (int i = tid; i < size; i += OCT_NT) {
      if (S.a[i] == 0) {
        nbnd[nChildren + S.bb[i]] = bnd[i];
        ncnt[nChildren + S.bb[i]] = cnt[i];
      } else {
        const int r = S.a[i] - 1;
        int ci = S.c[r];
        for (int q = 0; q < 4; ++q) {
          const int cc = S.child[i * 4 + q];
          if (cc > 0) {
            const int np = nChildren - 1 - ci;   // push_front: later-created nodes come first
            nbnd[np] = oct_child_bounds(bnd[i], q);
            ncnt[np] = cc;
            S.child[i * 4 + q] = -(np + 1);     // remember where this child went
            ++ci;
          }
        }
      }
    }
    __syncthreads();
    for (int k = tid; k < n; k += OCT_NT) {
      const int pos = keyNode[k];
      int np;
      if (S.a[pos] == 0) np = nChildren + S.bb[pos];
      else np = -S.child[pos * 4 + oct_quadrant(cand[k], bnd[pos])] - 1;
      keyNode[k] = (uint16_t)np;
    }
    __syncthreads();
    // nToExpand = new children holding more than one key
    if (tid == 0) s_i[1] = 0;
    __syncthreads();
    {
      int local = 0;
      for (int i = tid; i < nChildren; i += OCT_NT) local += ncnt[i] > 1;
      if (local) atomicAdd(&s_i[1], local);
    }
    __syncthreads();
    const int nToExpand = s_i[1];
    __syncthreads();
    size = newSize;
    nNew = nChildren;
    cur ^= 1;
    if (size >= N || size == prevSize) finish = true;
    else if (outer && size + nToExpand * 3 > N) outer = false;
    // (inner mode persists until finish, src/ORBextractor.cc:674-735)
  }

  // ---- best keypoint per node: max response, first in candidate order on ties (:739-758) ----
  // candidate order = (cell row, cell col, y, x) lexicographic
  for (int i = tid; i < size; i += OCT_NT) S.best[i] = 0ull;
  __syncthreads();
  for (int k = tid; k < n; k += OCT_NT) {
    const uint32_t c = cand[k];
    const int x = (c & 0xfff) - 3, y = ((c >> 12) & 0xfff) - 3;   // relative to the first interior pixel
    const int cj = x / L.wCell, ci = y / L.hCell;
    const uint32_t ord = (uint32_t)(((ci * L.nCols + cj) * 64 + (y - ci * L.hCell)) * 64 + (x - cj * L.wCell));
    const unsigned long long key = ((unsigned long long)(c >> 24) << 32) | (unsigned long long)(0xffffffffu - ord);
    atomicMax(&S.best[keyNode[k]], key);
  }
  __syncthreads();
  // lapping flags + their exclusive prefix (operator() output ordering, :1141-1151)
  for (int i = tid; i < size; i += OCT_NT) {
    const unsigned long long key = S.best[i];
    const uint32_t ord = 0xffffffffu - (uint32_t)(key & 0xffffffffu);
    const int xin = ord & 63, yin = (ord >> 6) & 63, cell = ord >> 12;
    const int ci = cell / L.nCols, cj = cell - ci * L.nCols;
    const int px = ORBX_MINB + 3 + cj * L.wCell + xin, py = ORBX_MINB + 3 + ci * L.hCell + yin;
    float fx = (float)px;
    if (level != 0) fx = __fmul_rn(fx, L.scale);
    const int lap = (fx >= (float)p.lap0 && fx <= (float)p.lap1) ? 1 : 0;
    S.a[i] = lap;
    S.bb[i] = px | (py << 16);
    S.c[i] = (int)(key >> 32);
  }
  __syncthreads();
  for (int i = tid; i < size; i += OCT_NT) S.rankPos[i] = S.a[i];
  __syncthreads();
  const int nLap = block_excl_scan(S.rankPos, size, s_warp);
  for (int i = tid; i < size; i += OCT_NT) {
    if (i < L.selCap) {
      uint2 r;
      r.x = (uint32_t)S.bb[i];
      r.y = (uint32_t)S.c[i] | ((uint32_t)S.a[i] << 8) | ((uint32_t)S.rankPos[i] << 16);
      sel[i] = r;
    }
  }
  if (tid == 0) {
    if (size > L.selCap) atomicExch(p.err, 3);
    p.selN[b * p.nlevels + level] = min(size, L.selCap);
    p.selLap[b * p.nlevels + level] = nLap;
  }
}

// =====================================================================================
// K4+K6  orientation (IC_Angle + fastAtan2), steered rBRIEF and the final scatter into the
// caller's arrays.  One warp per selected keypoint.
// =====================================================================================
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale),
              p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
  const float eps = 2.2204460492503131e-16f;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

__global__ void __launch_bounds__(128) describe_kernel(const __grid_constant__ ExtractParams p, orbx_keypoint* kps,
                                                       uint8_t* desc, int cap, int* nOut, int* monoOut) {
  // lanes of a warp read 32 different pattern rows: stage the table in shared memory (constant
  // memory would serialise the divergent addresses)
  __shared__ int s_pat[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x)   // transposed: word (byte row r, pair k) at [k*32 + r]
    s_pat[(i & 7) * 32 + (i >> 3)] = reinterpret_cast<const int*>(c_pattern)[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  // locate (level, i) of this warp's keypoint and the counts it needs for its output slot
  int level = -1, idx = 0, base = 0, lapBefore = 0, total = 0, totalLap = 0;
  {
    int acc = 0, lapAcc = 0;
    for (int l = 0; l < p.nlevels; ++l) {
      const int nl = p.selN[b * p.nlevels + l];
      if (level < 0 && warp < acc + nl) {
        level = l;
        idx = warp - acc;
        base = acc;
        lapBefore = lapAcc;
      }
      acc += nl;
      lapAcc += p.selLap[b * p.nlevels + l];
    }
    total = acc;
    totalLap = lapAcc;
  }
  if (warp == 0 && lane == 0) {
    nOut[b] = min(total, cap);
    monoOut[b] = total - totalLap;
    if (total > cap) atomicExch(p.err, 4);
  }
  if (level < 0) return;
  const LevelParams& L = p.lv[level];
  const uint2 rec = p.sel[(size_t)b * p.selPerImage + L.selOfs + idx];
  const int px = rec.x & 0xffff, py = rec.x >> 16;
  const int score = rec.y & 0xff, lap = (rec.y >> 8) & 1, lapPrefix = rec.y >> 16;

  // --- IC_Angle on the un-blurred level: lane = patch row ---
  const uint8_t* img = L.pyr + (size_t)b * L.imgStride;
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int dy = lane - 15;
    const int d = c_umax[dy < 0 ? -dy : dy];
    const uint8_t* row = img + (size_t)(py + dy) * L.pitch + px;
    int rs = 0;
    for (int u = -d; u <= d; ++u) {
      const int v = __ldg(row + u);
      rs += v;
      m10 += u * v;
    }
    m01 = dy * rs;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // --- steered BRIEF on the blurred level: lane = descriptor byte ---
  const float factorPI = (float)(3.14159265358979323846 / 180.0);
  const float ang = __fmul_rn(angle, factorPI);
  const float a = (float)cos((double)ang), bsn = (float)sin((double)ang);
  const uint8_t* bl = L.blur + (size_t)b * L.imgStride + (size_t)py * L.pitch + px;
  int val = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int pw = s_pat[k * 32 + lane];   // (x0,y0,x1,y1) packed int8
    const float x0 = (float)(int8_t)(pw & 0xff), y0 = (float)(int8_t)((pw >> 8) & 0xff);
    const float x1 = (float)(int8_t)((pw >> 16) & 0xff), y1 = (float)(int8_t)((pw >> 24) & 0xff);
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, bsn), __fmul_rn(y0, a)));
    const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, bsn)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, bsn), __fmul_rn(y1, a)));
    const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, bsn)));
    const int t0 = __ldg(bl + r0 * L.pitch + c0), t1 = __ldg(bl + r1 * L.pitch + c1);
    val |= (t0 < t1) << k;
  }
  // --- output slot: non-lapping keypoints fill from the front, lapping ones from the back ---
  const int slot = lap ? (total - 1 - (lapBefore + lapPrefix)) : (base - lapBefore + idx - lapPrefix);
  if (slot < 0 || slot >= cap) return;
  desc[((size_t)b * cap + slot) * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    orbx_keypoint kp;
    kp.x = (float)px;
    kp.y = (float)py;
    if (level != 0) {
      kp.x = __fmul_rn(kp.x, L.scale);
      kp.y = __fmul_rn(kp.y, L.scale);
    }
    kp.size = L.kpSize;
    kp.angle = angle;
    kp.response = (float)score;
    kp.octave = level;
    kps[(size_t)b * cap + slot] = kp;
  }
}

// ------------------------------------------------------------------------------------
// host-side launchers (called from orbx_extract.cu)
// ------------------------------------------------------------------------------------
size_t orbx_fast_smem_bytes(int fastTileBytes) { return (size_t)3 * fastTileBytes + 64; }
size_t orbx_octree_smem_bytes(int nodeCap) {
  return (size_t)nodeCap * (2 * sizeof(short4) + sizeof(unsigned long long) + 4 * 2 + 16 + 4 * 4);
}

int orbx_extract_configure(int nodeCap, int fastTileBytes) {
  ORBX_CUDA(cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)orbx_fast_smem_bytes(fastTileBytes)));
  ORBX_CUDA(cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)orbx_octree_smem_bytes(nodeCap)));
  return ORBX_OK;
}

int orbx_extract_launch(orbx_ctx* ctx, cudaStream_t st, const ExtractParams& p, orbx_keypoint* d_kps, uint8_t* d_desc,
                        int cap, int* d_n, int* d_mono, cudaEvent_t* ev /* [ORBX_EXT_STAGES+1] or null */) {
  const int B = p.B;
  ORBX_CUDA(cudaMemsetAsync(p.candN, 0, sizeof(int) * B * p.nlevels, st));
#define ORBX_EV(i) do { if (ev) ORBX_CUDA(cudaEventRecord(ev[i], st)); } while (0)
  ORBX_EV(0);
  for (int l = 1; l < p.nlevels; ++l) {
    dim3 grid(div_up(div_up(p.lv[l].w, 4), 128), p.lv[l].h, B);
    pyr_resize_kernel<<<grid, 128, 0, st>>>(p, l);
    ORBX_LAUNCH(ctx);
  }
  ORBX_EV(1);
  fast_cells_kernel<<<dim3(p.totalFastTiles, B), 256, orbx_fast_smem_bytes(p.fastTileBytes), st>>>(p);
  ORBX_LAUNCH(ctx);
  ORBX_EV(2);
  gauss7_kernel<<<dim3(p.totalBlurTiles, B), 256, 0, st>>>(p);
  ORBX_LAUNCH(ctx);
  ORBX_EV(3);
  octree_kernel<<<dim3(p.nlevels, B), OCT_NT, orbx_octree_smem_bytes(p.nodeCap), st>>>(p);
  ORBX_LAUNCH(ctx);
  ORBX_EV(4);
  const int warpsPerImage = p.selPerImage;
  describe_kernel<<<dim3(div_up(warpsPerImage * 32, 128), B), 128, 0, st>>>(p, d_kps, d_desc, cap, d_n, d_mono);
  ORBX_LAUNCH(ctx);
  ORBX_EV(5);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}
