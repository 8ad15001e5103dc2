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
// grid (nbx, ceil(h/PYR_STRIP), B); the block covers one strip of PYR_STRIP output rows with one thread per 4-pixel word
// (block width = the row's words split evenly over nbx blocks, rounded to whole warps).  The horizontal pass of a SOURCE
// row (T = S[x0]*a0 + S[x0+1]*a1, kept as T >> 4) is computed once and reused by every output row that blends it (with
// the 1.2 scale factor a source row feeds 1.67 output rows on average): the thread streams down the source rows of its
// strip, keeps the previous and the current horizontal pass in registers and has the raw bytes of the next row in flight.
// The vertical pass is
//   ((b0 * (T0 >> 4)) >> 16) + ((b1 * (T1 >> 4)) >> 16) + 2) >> 2
// with each ">> 16" product taken as a high multiply by b << 16 (coefficients are in [0, 2048], so this is exact).
// Packed coefficient tables (int16 x 4 per destination column / row, one 8-byte load each):
//   X[dx] = { xofs, a0, a1, 0 }     Y[dy] = { y0, y1, b0, b1 }
// =====================================================================================
#define PYR_STRIP 8

// Per-thread constants of its 4 output columns.
// WORDS path (every level whose source rows are 4-byte aligned and whose 4 columns span <= 8 source bytes, i.e. scale
// factors up to 2): the 4 x 2 source bytes of a row are cut out of three aligned 32-bit loads -- two funnel shifts bring
// the 8-byte window that starts at x0[0] into (u0, u1), two PRMTs with per-thread selectors line the pixel pairs up as
// (S[x0], S[x0+1]) byte pairs, and one DP2A per pixel forms S[x0]*a0 + S[x0+1]*a1 against the packed coefficient pair
// (a0 | a1 << 16): 3 loads + 12 ALU instructions per 4 pixels instead of 8 byte loads with 64-bit address arithmetic.
// Where a1 == 0 (left / right clamp) the selector repeats byte x0, so nothing right of the last source pixel is touched.
template <bool WORDS>
struct PyrCols;
template <>
struct PyrCols<false> {
  int x0[4];
  int a0[4], a1[4];
};
template <>
struct PyrCols<true> {
  int base;                  // byte offset of the first aligned word (x0[0] & ~3)
  unsigned shift;            // 8 * (x0[0] & 3)
  unsigned sel01, sel23;     // PRMT selectors of the pixel pairs (0,1) and (2,3) inside (u0, u1)
  unsigned coef[4];          // a0 | a1 << 16
  bool need1, need2;         // the window reaches into the 2nd / 3rd word
};

__device__ __forceinline__ void pyr_cols_init(PyrCols<false>& C, const int (&ent)[8], int, int) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    C.x0[i] = ent[2 * i] & 0xffff;
    C.a0[i] = ent[2 * i] >> 16;
    C.a1[i] = (short)(ent[2 * i + 1] & 0xffff);
  }
}
__device__ __forceinline__ void pyr_cols_init(PyrCols<true>& C, const int (&ent)[8], int dx0, int dw) {
  const int x00 = ent[0] & 0xffff;
  C.base = x00 & ~3;
  const int mis = x00 & 3;
  C.shift = 8u * (unsigned)mis;
  unsigned sel[4];
  int last = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool live = dx0 + i < dw;                       // table padding past the last column: weight 0, byte x0[0]
    const int a0 = live ? (ent[2 * i] >> 16) : 0, a1 = live ? (int)(short)(ent[2 * i + 1] & 0xffff) : 0;
    const int o = live ? (ent[2 * i] & 0xffff) - x00 : 0;
    const int o1 = o + (a1 != 0 ? 1 : 0);
    sel[i] = (unsigned)o | ((unsigned)o1 << 4);
    C.coef[i] = (unsigned)a0 | ((unsigned)a1 << 16);
    last = max(last, o1);
  }
  C.sel01 = sel[0] | (sel[1] << 8);
  C.sel23 = sel[2] | (sel[3] << 8);
  C.need1 = mis + last >= 4;
  C.need2 = mis + last >= 8;
}

// raw source bytes of one row for the thread's 4 columns (loaded one source row ahead of their use)
template <bool WORDS>
struct PyrRaw;
template <>
struct PyrRaw<false> { uint32_t s0, s1; };       // bytes S[x0[i]] and S[x0[i] + 1], packed
template <>
struct PyrRaw<true> { uint32_t w0, w1, w2; };    // the aligned words around x0[0]

__device__ __forceinline__ void pyr_load(const uint8_t* __restrict__ row, const PyrCols<false>& C, PyrRaw<false>& R) {
  // x0 + 1 may be one past the last source pixel at the right edge: there a1 == 0 (OpenCV clamps fx to 0) and the
  // byte read lies inside the row's pitch padding
  R.s0 = R.s1 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    R.s0 |= (uint32_t)__ldg(row + C.x0[i]) << (8 * i);
    R.s1 |= (uint32_t)__ldg(row + C.x0[i] + 1) << (8 * i);
  }
}
__device__ __forceinline__ void pyr_load(const uint8_t* __restrict__ row, const PyrCols<true>& C, PyrRaw<true>& R) {
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(row + C.base);
  R.w0 = __ldg(wp);
  R.w1 = C.need1 ? __ldg(wp + 1) : 0u;
  R.w2 = C.need2 ? __ldg(wp + 2) : 0u;
}
__device__ __forceinline__ void pyr_finish(const PyrRaw<false>& R, const PyrCols<false>& C, int (&T)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    T[i] = ((int)((R.s0 >> (8 * i)) & 0xff) * C.a0[i] + (int)((R.s1 >> (8 * i)) & 0xff) * C.a1[i]) >> 4;
}
__device__ __forceinline__ void pyr_finish(const PyrRaw<true>& R, const PyrCols<true>& C, int (&T)[4]) {
  const uint32_t u0 = __funnelshift_r(R.w0, R.w1, C.shift), u1 = __funnelshift_r(R.w1, R.w2, C.shift);
  const uint32_t p01 = __byte_perm(u0, u1, C.sel01), p23 = __byte_perm(u0, u1, C.sel23);
  T[0] = (int)(__dp2a_lo(C.coef[0], p01, 0u) >> 4);
  T[1] = (int)(__dp2a_hi(C.coef[1], p01, 0u) >> 4);
  T[2] = (int)(__dp2a_lo(C.coef[2], p23, 0u) >> 4);
  T[3] = (int)(__dp2a_hi(C.coef[3], p23, 0u) >> 4);
}

// The thread walks down the SOURCE rows its strip of output rows blends (y0 of the first .. y1 of the last; with scale
// factors below 2 every one of them is used): the raw bytes of row sy + 1 are requested before row sy is reduced, and an
// output row is emitted as soon as its lower source row y1 is there (its upper one, y0 = y1 - 1 or y1 at the clamped
// border, is the previous / the same horizontal pass).  All control flow is uniform across the block.
template <bool WORDS>
__global__ void __launch_bounds__(256) pyr_resize_kernel(const __grid_constant__ ExtractParams p, int level) {
  const LevelParams& D = p.lv[level];
  const LevelParams& S = p.lv[level - 1];
  const int dx0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (dx0 >= D.w) return;
  const int dyBeg = blockIdx.y * PYR_STRIP, dyEnd = min(dyBeg + PYR_STRIP, D.h);
  const uint8_t* src = S.pyr + (size_t)blockIdx.z * S.imgStride;
  uint8_t* drow = D.pyr + (size_t)blockIdx.z * D.imgStride + dx0 + (size_t)dyBeg * D.pitch;
  const int dpitch = D.pitch, spitch = S.pitch;
  PyrCols<WORDS> C;
  {
    // the table is padded to a multiple of 4 entries, so the two 16-byte loads never leave it
    const int4* tx4 = reinterpret_cast<const int4*>(p.tab + D.tabX) + (dx0 >> 1);
    const int4 e01 = __ldg(tx4), e23 = __ldg(tx4 + 1);
    const int ent[8] = {e01.x, e01.y, e01.z, e01.w, e23.x, e23.y, e23.z, e23.w};
    pyr_cols_init(C, ent, dx0, D.w);
  }
  const short4* tabY = reinterpret_cast<const short4*>(p.tab + D.tabY);
  const int syBeg = tabY[dyBeg].x, syEnd = tabY[dyEnd - 1].y;
  int dy = dyBeg;
  short4 ty = tabY[dy];                             // uniform across the block
  const uint8_t* srow = src + (size_t)syBeg * spitch;
  PyrRaw<WORDS> R;
  pyr_load(srow, C, R);
  int TP[4] = {0, 0, 0, 0}, TC[4];
#pragma unroll 1
  for (int sy = syBeg; sy <= syEnd; ++sy) {
    srow += spitch;
    PyrRaw<WORDS> N = R;
    if (sy < syEnd) pyr_load(srow, C, N);          // next source row, in flight while this one is reduced
    pyr_finish(R, C, TC);
    R = N;
    while (dy < dyEnd && ty.y == sy) {
      const int B0 = (int)ty.z << 16, B1 = (int)ty.w << 16;
      const bool same = ty.x == sy;                 // y0 == y1 (clamped border)
      int v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        v[i] = (__mulhi(B0, same ? TC[i] : TP[i]) + __mulhi(B1, TC[i]) + 2) >> 2;   // <= 255: the weights sum to 2048 twice
      const uint32_t out = __byte_perm(__byte_perm((uint32_t)v[0], (uint32_t)v[1], 0x0040),
                                       __byte_perm((uint32_t)v[2], (uint32_t)v[3], 0x0040), 0x5410);
      *reinterpret_cast<uint32_t*>(drow) = out;     // pitch is a multiple of 16 and padded: safe past w
      drow += dpitch;
      if (++dy < dyEnd) ty = tabY[dy];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) TP[i] = TC[i];
  }
}

// =====================================================================================
// K2  per-cell FAST-9-16 + cell-local 3x3 NMS + iniTh->minTh fallback + candidate emission.
// One CTA = one cell row x up to ORBX_FAST_CELLS cells of one level of one image.
//
// The reference runs cv::FAST(cell, iniTh) and, only for a cell that yields nothing, cv::FAST(cell, minTh)
// (src/ORBextractor.cc:808-828).  The kernel does the same in two passes over the shared-memory tile: pass 0 at iniTh
// over every pixel, pass 1 at minTh restricted to the pixels of the cells pass 0 left empty.  The kernel is bound by
// instruction issue, not by memory, so every stage is built to spend as few issue slots per pixel as
// exactness allows:
//   A. tile -> shared memory: ONE TMA bulk tensor copy (box ORBX_FAST_TP x fastTH; 32-bit loads when the level has no
//      tensor map).  The shared pitch is the compile-time constant ORBX_FAST_TP for every level, so all ring / neighbour
//      addresses below are immediate offsets.
//   B. dense early reject, 4 pixels per instruction: a thread owns one 4-pixel word column and an 8-row strip, keeps the
//      14 rows it needs in registers, and tests VABSDIFF4 of the centre word against the four compass ring positions
//      (0,4,8,12) with SWAR ">t" masks: "two adjacent compass points differ by more than t" is necessary for a corner
//      (every 9-arc of the 16-ring contains two adjacent compass points).  The 32 flags of the strip (4 px x 8 rows)
//      accumulate in ONE register; the survivors are appended to list 1 with one warp prefix sum + one shared atomic
//      per strip (the previous version paid 4 ballots + 4 prefix popcounts per word: 70 of its 115 instructions/word).
//   C. exact score of every survivor: the 16 ring differences are packed as (256+d, 256-d) in one s16x2 register by a
//      single IMAD each, and the max-over-arcs-of-min network runs for both polarities at once on VIMNMX(3).S16x2
//      (36 instructions: two neighbouring arcs share their 8 common elements).  corner <=> arcmax > t.
//   D. cell-local 3x3 NMS over the corners (neighbours across a cell seam count as 0, like the borders of the
//      per-cell cv::FAST call); a kept pass-0 keypoint marks its cell non-empty.
//   E. emission: the kept keypoints of a warp (compacted in place once more) go straight to the global candidate
//      list with their scores, one global atomic per warp and pass.
// =====================================================================================
#define FAST_TP ORBX_FAST_TP
#define FAST_TPW (ORBX_FAST_TP / 4)
#define FAST_STRIP 8
// row of a pixel code (code / FAST_TP by multiplication; exact for code < 23831, codes stay below 48 * FAST_TP)
#define FAST_CODE_Y(code) ((int)(((unsigned)(code) * (unsigned)((4194304 + FAST_TP - 1) / FAST_TP)) >> 22))
static_assert(FAST_TP == 240, "FAST_CODE_Y's exactness bound was derived for a pitch of 240");

__device__ __forceinline__ unsigned swar_gt_u8(unsigned x, unsigned k) {   // k = (0x7f - t) * 0x01010101, t < 128
  return (((x & 0x7f7f7f7fu) + k) | x) & 0x80808080u;
}

__device__ __forceinline__ unsigned vmin2(unsigned a, unsigned b) { return __vmins2(a, b); }
__device__ __forceinline__ unsigned vmax2(unsigned a, unsigned b) { return __vmaxs2(a, b); }
__device__ __forceinline__ unsigned vmin3(unsigned a, unsigned b, unsigned c) { return __vimin3_s16x2(a, b, c); }
__device__ __forceinline__ unsigned vmax3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }

// X[k] = (256 + d_k) | (256 - d_k) << 16 ; returns max over the 16 cyclic 9-arcs of the lane-wise minimum.
// Arcs j and j+1 (j even) share X[j+1..j+8]: max(arc_j, arc_j+1) = min(min(X[j+1..j+8]), max(X[j], X[j+9])).
__device__ __forceinline__ unsigned arc9_maxmin_x2(const unsigned (&X)[16]) {
  unsigned pr[8], s4[8], r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) pr[i] = vmin2(X[2 * i + 1], X[(2 * i + 2) & 15]);            // min X[2i+1 .. 2i+2]
#pragma unroll
  for (int i = 0; i < 8; ++i) s4[i] = vmin2(pr[i], pr[(i + 1) & 7]);                        // min X[2i+1 .. 2i+4]
#pragma unroll
  for (int i = 0; i < 8; ++i)                                                               // j = 2i
    r[i] = vmin3(s4[i], s4[(i + 2) & 7], vmax2(X[2 * i], X[(2 * i + 9) & 15]));
  const unsigned a = vmax3(r[0], r[1], r[2]), b = vmax3(r[3], r[4], r[5]);
  return vmax2(vmax3(r[6], r[7], a), b);
}

// atom.shared.add issued exactly as written (nvcc wraps atomicAdd in a warp-aggregation sequence of its own; the callers
// below already aggregate per warp)
__device__ __forceinline__ int smem_atomic_add(int* addr, int v) {
  int old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(addr)), "r"(v) : "memory");
  return old;
}

// shared-memory flag words of fast_cells_kernel
enum { FF_CELL = 0 /* [0..7] cell has a pass-0 keypoint */, FF_NCAND = 8,
       FF_WCORN = 16 /* [16..23] corners found by warp w (they sit at the front of its share of list 1) */ };

__global__ void __launch_bounds__(256) fast_cells_kernel(const __grid_constant__ FastTmaMaps maps,
                                                         const __grid_constant__ ExtractParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // tile record (host-built: only tiles that own at least one evaluated pixel are listed):
  //   x = level | nc << 8 | ceil(65536 / ncw) << 16, y = iniX | iniY << 16, z = tw | th << 16, w = ceil(65536 / wCell)
  const int4 trec = __ldg(p.fastTiles + blockIdx.x);
  const int level = trec.x & 0xff, nc = (trec.x >> 8) & 0xff;
  const LevelParams& L = p.lv[level];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int iniX = trec.y & 0xffff, iniY = trec.y >> 16;
  const int tw = trec.z & 0xffff, th = trec.z >> 16;  // tile incl. 3-px ring halo
  const int wI = tw - 6, hI = th - 6;                 // interior (evaluated) pixels; wI <= FAST_TP - 25

  // shared layout: [flags 128 B][mbarrier][image plane FAST_TP x (rows)][score plane FAST_TP x (rows)][column table 256 B]
  //                [word masks 256 B][list 1: survivors, compacted in place to the corners, then to the kept keypoints]
  // A pixel is named by code = y * FAST_TP + x (interior coordinates); image byte = code + 3*FAST_TP + 3 + off, score
  // byte = code + FAST_TP + 1.
  const int xa = iniX & ~15;                          // tile origin: TMA needs a 16-byte aligned innermost coordinate
  const int off = iniX - xa;                          // 0..15: byte column of tile column 0
  const int CC = p.fastCandCap;                       // >= interior pixels of any tile, multiple of 64
  int* sflag = reinterpret_cast<int*>(smem);
  unsigned long long* smbar = reinterpret_cast<unsigned long long*>(smem + 128);  // TMA completion barrier
  uint8_t* simg = smem + 256;                         // 128-byte aligned: TMA destination
  uint8_t* ssc = simg + (size_t)p.fastTileBytes;      // 16-byte aligned (both plane sizes are multiples of FAST_TP)
  uint8_t* scell = ssc + (size_t)p.fastScoreBytes;    // [wI] cell index of interior column x | 16: first | 32: last column of its cell
  uint32_t* smask = reinterpret_cast<uint32_t*>(scell + 256);   // [ncw] bytes of word c that are tested in this pass
  uint16_t* scand = reinterpret_cast<uint16_t*>(scell + 512);   // [CC]   list 1

  // ---- A: load ----
  const uint8_t* img = L.pyr + (size_t)b * L.imgStride;
  if (L.useTma) {
    // one elected thread arms the mbarrier with the byte count and issues ONE bulk tensor copy for the whole tile
    // (box FAST_TP x fastTH at (xa, iniY, b); bytes outside the tensor are zero-filled by the TMA unit)
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(smbar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const unsigned bytes = (unsigned)(FAST_TP * L.fastTH);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(simg);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
      // the descriptor is addressed in place in the kernel-parameter bank (levels compared against constants so the
      // address stays a parameter-space constant + select)
      const CUtensorMap* tm = &maps.m[0];
#pragma unroll
      for (int l = 1; l < ORBX_TMA_LEVELS; ++l)
        if (level == l) tm = &maps.m[l];
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"(dst), "l"(tm), "r"(xa), "r"(iniY), "r"(b), "r"(mbar)
          : "memory");
    }
  } else {
    uint32_t* simg32 = reinterpret_cast<uint32_t*>(simg);
    const int rowWords = (L.pitch - xa) / 4;          // words readable in a row without leaving the pitch
    const int cw = min(FAST_TPW, rowWords);
    for (int r = wid; r < th; r += 8) {               // one warp per tile row: coalesced 128-byte segments, no division
      const uint32_t* g = reinterpret_cast<const uint32_t*>(img + (size_t)(iniY + r) * L.pitch + xa);
      uint32_t* d = simg32 + r * FAST_TPW;
      for (int c = lane; c < FAST_TPW; c += 32) d[c] = c < cw ? __ldg(g + c) : 0u;
    }
  }
  const int c0 = (off + 3) / 4, c1 = (off + 3 + wI - 1) / 4;   // words that hold interior pixels
  const int ncw = c1 - c0 + 1;                                 // <= 58
  {
    uint4* ssc128 = reinterpret_cast<uint4*>(ssc);
    for (int i = tid; i < (hI + 2) * (FAST_TP / 16); i += 256) ssc128[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int x = tid; x < wI; x += 256) {
      const int cell = (x * trec.w) >> 16;            // x / wCell, exact for x < 512
      const int cl = x > 0 ? ((x - 1) * trec.w) >> 16 : -1, cr = x + 1 < wI ? ((x + 1) * trec.w) >> 16 : -1;
      scell[x] = (uint8_t)(cell | (cl != cell ? 16 : 0) | (cr != cell ? 32 : 0));
    }
    if (tid < 32) sflag[tid] = 0;
    // pass 0 tests every interior pixel: the mask only clips the first / last word to the interior columns
    for (int k = tid; k < ncw; k += 256) {
      const int xb = (c0 + k) * 4 - (off + 3);        // interior x of byte 0 (may be negative)
      unsigned m = 0x80808080u;
      if (xb < 0) m &= 0xffffffffu << (8 * (-xb));
      if (xb + 3 >= wI) m &= 0xffffffffu >> (8 * (xb + 4 - wI));
      smask[k] = m;
    }
  }
  __syncthreads();                                    // (also publishes the mbarrier initialisation)
  if (L.useTma) {
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(smbar);
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(mbar)
          : "memory");
    }
  }

  const uint32_t* simg32 = reinterpret_cast<const uint32_t*>(simg);
  const int nStrips = (hI + FAST_STRIP - 1) / FAST_STRIP;
  const int nItems = nStrips * ncw;                   // (strip, word column) items of stage B, <= 5 * 58
  const int rcpNcw = (int)((unsigned)trec.x >> 16);   // ceil(65536 / ncw): item / ncw == (item * rcpNcw) >> 16 for item < 512, ncw <= 64
  const int nPass = p.minTh < p.iniTh ? 2 : 1;        // a second pass at a threshold >= iniTh could not add anything
#pragma unroll 1
  for (int pass = 0; pass < nPass; ++pass) {
    const int t = pass == 0 ? p.iniTh : p.minTh;
    const unsigned kGt = (unsigned)(0x7f - t) * 0x01010101u;
    if (pass == 1) {
      // which cells did pass 0 leave empty?  (all threads read the flags written before the last barrier of pass 0)
      int anyEmpty = 0;
      for (int cidx = 0; cidx < nc; ++cidx) anyEmpty |= sflag[FF_CELL + cidx] == 0;
      if (!anyEmpty) break;
      __syncthreads();
      if (tid == 0) sflag[FF_NCAND] = 0;
      // restrict the word masks to the pixels of empty cells
      for (int k = tid; k < ncw; k += 256) {
        const int xb = (c0 + k) * 4 - (off + 3);
        unsigned m = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int x = xb + q;
          if (x >= 0 && x < wI && sflag[FF_CELL + (scell[x] & 15)] == 0) m |= 0x80u << (8 * q);
        }
        smask[k] = m;
      }
      __syncthreads();
    }

    // ---- B: compass early reject; one (8-row strip, 4-pixel word column) item per thread-iteration ----
#pragma unroll 1
    for (int it0 = 0; it0 < nItems; it0 += 256) {       // warp-uniform trip count (the warp scan below needs all lanes)
      const int item = it0 + tid;
      unsigned M = 0;                                   // bit 8*q + u: pixel q of the word, row u of the strip survives
      int code0 = 0;
      if (item < nItems) {
        const int s = (item * rcpNcw) >> 16, k = item - s * ncw;
        const unsigned m = smask[k];
        const int y0 = s * FAST_STRIP;
        code0 = y0 * FAST_TP + (c0 + k) * 4 - (off + 3);   // code of (row y0, byte 0); byte 0 may lie left of the interior
        if (m) {
          const uint32_t* colp = simg32 + y0 * FAST_TPW + (c0 + k);   // tile row y0 = ring row -3 of interior row y0
          const int nrow = hI - y0;                     // rows of this strip that exist (>= 1)
          uint32_t A[FAST_STRIP + 6];
#pragma unroll
          for (int j = 0; j < FAST_STRIP + 6; ++j) A[j] = colp[j * FAST_TPW];   // rows past the tile: ignored below
#pragma unroll
          for (int u = 0; u < FAST_STRIP; ++u) {
            const uint32_t V = A[u + 3];
            const uint32_t Lw = colp[(u + 3) * FAST_TPW - 1], Rw = colp[(u + 3) * FAST_TPW + 1];
            const uint32_t r4 = __funnelshift_r(V, Rw, 24);      // bytes +3..+6
            const uint32_t r12 = __funnelshift_r(Lw, V, 8);      // bytes -3..0
            const unsigned b0 = swar_gt_u8(__vabsdiffu4(A[u + 6], V), kGt), b4 = swar_gt_u8(__vabsdiffu4(r4, V), kGt);
            const unsigned b8 = swar_gt_u8(__vabsdiffu4(A[u], V), kGt), b12 = swar_gt_u8(__vabsdiffu4(r12, V), kGt);
            const unsigned mu = u < nrow ? m : 0u;
            const unsigned cand = ((b0 | b8) & (b4 | b12)) & mu;   // (b0&b4)|(b4&b8)|(b8&b12)|(b12&b0)
            M = (M >> 1) | cand;
          }
        }
      }
      // append: warp prefix sum of the per-thread survivor counts, one shared atomic per warp
      const int cnt = __popc(M);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total == 0) continue;
      int wbase = 0;
      if (lane == 31) wbase = smem_atomic_add(&sflag[FF_NCAND], total);
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      uint16_t* q = scand + wbase + incl - cnt;
#pragma unroll
      for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int u = 0; u < FAST_STRIP; ++u)
          if (M & (1u << (8 * px + u))) *q++ = (uint16_t)(code0 + u * FAST_TP + px);
    }
    __syncthreads();

    // ---- C: exact score of the survivors; corners go to the score plane and are compacted IN PLACE: warp w owns the
    //      survivors [w*chunk, (w+1)*chunk) of list 1 and moves its corners to the front of that range (the write
    //      position never passes the read position, and a warp is convergent at the ballot), so no second list, no
    //      atomics and no capacity check are needed ----
    const int nCand = sflag[FF_NCAND];
    const int chunk = ((nCand + 255) >> 8) << 5;        // per-warp share, a multiple of 32
    const int cbeg = wid * chunk, cend = min(cbeg + chunk, nCand);
    {
      const int K0 = 256 * 65537;
      const uint8_t* cbase = simg + 3 * FAST_TP + 3 + off;
      const unsigned ltm = (1u << lane) - 1;
      int wcnt = 0;
      for (int i0 = cbeg; i0 < cend; i0 += 32) {
        const int i = i0 + lane;
        bool corner = false;
        int code = 0;
        if (i < cend) {
          code = scand[i];
          const uint8_t* c = cbase + code;
          const int Kv = K0 - 65535 * (int)c[0];          // X = 65535*r + Kv = (256 + v - r) | (256 - v + r) << 16
          unsigned X[16];
          X[0] = 65535u * c[3 * FAST_TP] + Kv;       X[1] = 65535u * c[3 * FAST_TP + 1] + Kv;   X[2] = 65535u * c[2 * FAST_TP + 2] + Kv;
          X[3] = 65535u * c[FAST_TP + 3] + Kv;       X[4] = 65535u * c[3] + Kv;                 X[5] = 65535u * c[-FAST_TP + 3] + Kv;
          X[6] = 65535u * c[-2 * FAST_TP + 2] + Kv;  X[7] = 65535u * c[-3 * FAST_TP + 1] + Kv;  X[8] = 65535u * c[-3 * FAST_TP] + Kv;
          X[9] = 65535u * c[-3 * FAST_TP - 1] + Kv;  X[10] = 65535u * c[-2 * FAST_TP - 2] + Kv; X[11] = 65535u * c[-FAST_TP - 3] + Kv;
          X[12] = 65535u * c[-3] + Kv;               X[13] = 65535u * c[FAST_TP - 3] + Kv;      X[14] = 65535u * c[2 * FAST_TP - 2] + Kv;
          X[15] = 65535u * c[3 * FAST_TP - 1] + Kv;
          const unsigned am = arc9_maxmin_x2(X);
          const int arcmax = max((int)(am & 0xffffu), (int)(am >> 16)) - 256;
          corner = arcmax > t && arcmax > 1;
          if (corner) ssc[code + FAST_TP + 1] = (uint8_t)(arcmax - 1);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, corner);
        if (corner) scand[cbeg + wcnt + __popc(bal & ltm)] = (uint16_t)code;
        wcnt += __popc(bal);
      }
      if (lane == 0) sflag[FF_WCORN + wid] = wcnt;
    }
    __syncthreads();

    // ---- D: NMS inside the pixel's own cell over the corners (warp w: the corners at the front of warp w's share),
    //      kept keypoints compacted in place again, then written straight to the global candidate list: one global
    //      atomic per warp and pass (every kept keypoint is final; the order of the list is irrelevant downstream) ----
    {
      const int n = cbeg + sflag[FF_WCORN + wid];
      const unsigned ltm = (1u << lane) - 1;
      int wkept = 0;
      for (int i0 = cbeg; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false;
        int code = 0;
        if (i < n) {
          code = scand[i];
          const uint8_t* sp0 = ssc + code + FAST_TP + 1;
          const int sc = sp0[0];
          // branch-free 3x3 maximum of the neighbours inside the pixel's own cell (others count as 0)
          const int cf = scell[code - FAST_CODE_Y(code) * FAST_TP];
          const int mL = (cf & 16) ? 0 : 0xff, mR = (cf & 32) ? 0 : 0xff;
          const int nL = max(max((int)sp0[-1], (int)sp0[-FAST_TP - 1]), (int)sp0[FAST_TP - 1]) & mL;
          const int nR = max(max((int)sp0[1], (int)sp0[-FAST_TP + 1]), (int)sp0[FAST_TP + 1]) & mR;
          const int nmax = max(max((int)sp0[-FAST_TP], (int)sp0[FAST_TP]), max(nL, nR));
          keep = sc > nmax;
          if (keep && pass == 0) sflag[FF_CELL + (cf & 15)] = 1;   // (benign race)
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) scand[cbeg + wkept + __popc(m & ltm)] = (uint16_t)code;
        wkept += __popc(m);
      }
      if (wkept > 0) {                                    // warp-uniform
        int gbase = 0;
        if (lane == 0) gbase = atomicAdd(p.candN + b * p.nlevels + level, wkept);
        gbase = __shfl_sync(0xffffffffu, gbase, 0);
        uint32_t* cand = p.cand + (size_t)b * p.candPerImage + L.candOfs;
        for (int i = lane; i < wkept; i += 32) {
          const int code = scand[cbeg + i];
          const int y = FAST_CODE_Y(code), x = code - y * FAST_TP;
          const int sc = ssc[code + FAST_TP + 1];
          const int slot = gbase + i;
          if (slot < L.candCap) {
            // coordinates relative to (minBorderX, minBorderY), as the reference stores them (:847-848)
            const int rx = iniX + 3 + x - ORBX_MINB, ry = iniY + 3 + y - ORBX_MINB;
            cand[slot] = (uint32_t)rx | ((uint32_t)ry << 12) | ((uint32_t)sc << 24);
          } else {
            atomicExch(p.err, 1);
          }
        }
      }
    }
    __syncthreads();
  }
}

// =====================================================================================
// K5  7x7 sigma-2 Gaussian, OpenCV 4.x fixed point: [18,34,48,56,48,34,18]/256 per pass,
//     (sum + 2^15) >> 16, BORDER_REFLECT_101.
// Register-sliding separable filter, no shared memory, VERTICAL PASS FIRST.  Both passes are exact integer sums, so their
// order is free: out = (sum_i wh_i * V(x+i-3, y) + 2^15) >> 16 with V(x, y) = sum_j wv_j * p(x, y+j-3) <= 255 * 256.
// A thread owns one 32-bit word (4 adjacent columns) and walks down a strip of rows keeping the last 7 rows of ITS word
// unpacked to 16-bit lanes (14 registers; the horizontal-first form kept 28 horizontal sums and three words per row):
//   * vertical: the symmetric taps are added as packed 16-bit lanes and scaled by 4 packed multiply-adds per register
//     (no lane overflows: V < 2^16): 16 instructions give V of the 4 columns as two u16x2 registers;
//   * horizontal: the neighbours' V registers come by 4 warp shuffles, and every output column is 4 DP2A (u16 x u8)
//     against weight words whose unused lane is 0 -- the register pairs (-4,-3) (-2,-1) (0,1) (2,3) (4,5) (6,7) line up
//     with every window without any re-alignment;
//   * borders: a lane's word is "virtual": columns outside [0, w) are reflected when the ROW IS LOADED (two words + one
//     PRMT with a thread-invariant selector), so V of a reflected column is right by construction.  Lanes 0 and 31 of a
//     warp only carry the halo words; a warp tile is 30 output words wide.
// =====================================================================================
#define BLUR_STRIP ORBX_BLUR_STRIP   // output rows per warp
#define BLUR_OW (ORBX_BLUR_TW / 4)   // output words per warp tile

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(256, 4) gauss7_kernel(const __grid_constant__ ExtractParams p) {
  static_assert(BLUR_OW == 30, "a warp tile is 30 output words + 2 halo lanes");
  const int tile = blockIdx.x;
  int level = 0;
#pragma unroll 1
  for (int l = 1; l < p.nlevels; ++l)
    if (tile >= p.lv[l].blurTileStart) level = l;
  // everything the row loop needs from the (dynamically indexed) level record, once, in registers
  const int w = p.lv[level].w, h = p.lv[level].h, pitch = p.lv[level].pitch, opitch = p.lv[level].blurPitch;
  // warp tiles (120 columns x BLUR_STRIP rows) are numbered row-major and a CTA takes 8 consecutive ones: its warps sit
  // side by side on the same rows, so the CTA streams whole contiguous image rows (DRAM-page friendly)
  const int lt = tile - p.lv[level].blurTileStart, tilesX = p.lv[level].blurTilesX;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wt = lt * 8 + wid;
  const int tyi = wt / tilesX, txi = wt - tyi * tilesX;
  const int y0 = tyi * BLUR_STRIP;                     // first output row of this warp's strip
  const int E = (w - 1) >> 2;                          // last word of a row
  if (txi * BLUR_OW > E || y0 >= h) return;            // warp-uniform
  const int cv = txi * BLUR_OW + lane - 1;             // this lane's (virtual) word: -1 and E+1 are the reflected halos
  const bool writes = lane >= 1 && lane <= BLUR_OW && cv <= E;
  const uint8_t* img = p.lv[level].pyr + (size_t)blockIdx.y * p.lv[level].imgStride;
  uint8_t* out = p.lv[level].blur + (size_t)blockIdx.y * p.lv[level].blurStride + 4 * max(cv, 0);
  const int y1 = min(y0 + BLUR_STRIP, h);
  // source of the virtual word: its 4 reflected columns lie within two adjacent words
  int base;
  unsigned sel = 0;
  {
    int sc[4], lo = w;
#pragma unroll
    for (int k = 0; k < 4; ++k) { sc[k] = reflect101(4 * cv + k, w); lo = min(lo, sc[k]); }
    base = lo >> 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) sel |= (unsigned)(sc[k] - 4 * base) << (4 * k);
  }
  const bool fix = sel != 0x3210u;
  const int second = min(base + 1, E) - base;          // 0 or 1 words to the right
  const uint32_t* col = reinterpret_cast<const uint32_t*>(img) + base;
  // raw loads and border fix-up are separate so that the loads of the NEXT row are issued before the arithmetic of the
  // current one
  auto load_raw = [&](int yy, uint32_t& a0, uint32_t& b0) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(col) + (size_t)yy * pitch);
    a0 = __ldg(r);
    b0 = fix ? __ldg(r + second) : 0u;
  };
  auto unpack = [&](uint32_t a0, uint32_t b0, uint32_t (&u)[2]) {
    const uint32_t v = fix ? __byte_perm(a0, b0, sel) : a0;
    u[0] = __byte_perm(v, 0u, 0x4140);                 // columns 0,1 as 16-bit lanes
    u[1] = __byte_perm(v, 0u, 0x4342);                 // columns 2,3
  };
  auto bottom = [&](int yy) { return yy >= h ? max(2 * (h - 1) - yy, 0) : yy; };   // rows below the image reflect
  uint32_t U[7][2];   // rows y-3..y+3 of this lane's word (statically indexed: the row loop is unrolled by 7)
  // prologue: rows y0-3 .. y0+2 (six independent rows: all loads are issued before the first use)
  {
    uint32_t ra[6], rb[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) load_raw(reflect101(y0 - 3 + r, h), ra[r], rb[r]);
#pragma unroll
    for (int r = 0; r < 6; ++r) unpack(ra[r], rb[r], U[r]);
  }
  // row y+3 of the next output row is already on its way (prefetching three rows ahead measured 6 % SLOWER: 60 registers,
  // 1.23 vs 1.16 ms per 888 images)
  uint32_t na, nb;
  load_raw(bottom(y0 + 3), na, nb);
  const unsigned WA = 0u | (18u << 8) | (34u << 16) | (48u << 24);     // lane weights (0,18 | 34,48)
  const unsigned WB = 56u | (48u << 8) | (34u << 16) | (18u << 24);    // (56,48 | 34,18)
  const unsigned WC = 18u | (34u << 8) | (48u << 16) | (56u << 24);    // (18,34 | 48,56)
  const unsigned WD = 48u | (34u << 8) | (18u << 16) | (0u << 24);     // (48,34 | 18,0)
  for (int yb = y0; yb < y1; yb += 7) {
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      const int y = yb + u;
      if (y < y1) {                                      // warp-uniform
        // newest row y+3 goes to slot (6+u)%7; rows y-3..y+3 sit in slots (u+j)%7, j = 0..6
        unpack(na, nb, U[(6 + u) % 7]);
        if (y + 1 < y1) load_raw(bottom(y + 4), na, nb);   // prefetch for the next output row
        uint32_t V[2];
#pragma unroll
        for (int q = 0; q < 2; ++q)                      // packed 16-bit lanes, every partial result < 2^16
          V[q] = 18u * (U[u % 7][q] + U[(u + 6) % 7][q]) + 34u * (U[(u + 1) % 7][q] + U[(u + 5) % 7][q]) +
                 48u * (U[(u + 2) % 7][q] + U[(u + 4) % 7][q]) + 56u * U[(u + 3) % 7][q];
        const uint32_t A = __shfl_up_sync(0xffffffffu, V[0], 1), B = __shfl_up_sync(0xffffffffu, V[1], 1);
        const uint32_t Eo = __shfl_down_sync(0xffffffffu, V[0], 1), F = __shfl_down_sync(0xffffffffu, V[1], 1);
        const uint32_t C = V[0], D = V[1];
        // column k: window k-3..k+3 over the pairs (-4,-3)=A (-2,-1)=B (0,1)=C (2,3)=D (4,5)=Eo (6,7)=F
        const unsigned acc0 = __dp2a_hi(D, WB, __dp2a_lo(C, WB, __dp2a_hi(B, WA, __dp2a_lo(A, WA, 32768u))));
        const unsigned acc1 = __dp2a_hi(Eo, WD, __dp2a_lo(D, WD, __dp2a_hi(C, WC, __dp2a_lo(B, WC, 32768u))));
        const unsigned acc2 = __dp2a_hi(Eo, WB, __dp2a_lo(D, WB, __dp2a_hi(C, WA, __dp2a_lo(B, WA, 32768u))));
        const unsigned acc3 = __dp2a_hi(F, WD, __dp2a_lo(Eo, WD, __dp2a_hi(D, WC, __dp2a_lo(C, WC, 32768u))));
        // < 2^24; the rounded result is byte 2 of the accumulator
        const uint32_t o = __byte_perm(__byte_perm(acc0, acc1, 0x0062), __byte_perm(acc2, acc3, 0x0062), 0x5410);
        if (writes) *reinterpret_cast<uint32_t*>(out + (size_t)y * opitch) = o;   // pitch padding absorbs the tail
      }
    }
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
  extern __shared__ __align__(16) uint8_t oct_smem[];
  const int level = blockIdx.x, b = blockIdx.y;
  const LevelParams& L = p.lv[level];
  const int NC = p.nodeCap;
  const int tid = threadIdx.x;
  __shared__ int s_warp[33];
  __shared__ int s_i[8];

  OctSmem S;
  {
    uint8_t* q = oct_smem;
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

#define DESC_NT 256

__device__ __forceinline__ int dp4a_us(unsigned a, unsigned b_s8, int c) {   // unsigned bytes x signed bytes
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_s8), "r"(c));
  return d;
}

#define DESC_KPW 4                        // keypoints a warp processes one after the other
#define DESC_OW 9                         // words per staged row of the un-blurred 31x31 patch (31 + <=3 bytes of misalignment)
#define DESC_BW 20                        // words per staged row of the blurred 37x37 patch (the rotated pattern reaches +-18):
                                          // four 16-byte quads from the 16-byte aligned column at or left of px - 18 (<= 15 + 37
                                          // bytes), padded to 20 words so that the rows spread over the banks
#define DESC_OWORDS (31 * DESC_OW + 1)
#define DESC_BWORDS (37 * DESC_BW)

// The kernel used to be bound by the L1 data pipe (87 % of its wavefront peak, profiles/r02m): a warp's 9 patch-row loads
// touched 31 cache lines each and its 16 byte gathers ~25.  Both patches are now brought in by row-major word loads (a
// request covers 3-4 rows) into per-warp shared memory, and the orientation sums and the 512 pattern samples are read from
// there.
__global__ void __launch_bounds__(DESC_NT, 5) describe_kernel(const __grid_constant__ ExtractParams p, orbx_keypoint* kps,
                                                           uint8_t* desc, int cap, int* nOut, int* monoOut) {
  // lanes of a warp read 32 different pattern rows: stage the table in shared memory (constant
  // memory would serialise the divergent addresses)
  __shared__ float4 s_pat[256];
  // IC_Angle weights: row |v| of the r=15 disc covers u in [-umax[v], umax[v]].  For the 32 bytes u = -15..16 of a
  // patch row, s_icw[v*17 + j] packs the signed weights u (0 outside the disc) of bytes 4j..4j+3 and
  // s_icw[v*17 + 8 + j] the 0/1 membership mask (stride 17: rows of different lanes fall into different banks)
  __shared__ unsigned s_icw[16 * 17];
  __shared__ int s_cnt[2 * ORBX_MAX_LEVELS];
  __shared__ int s_pre[2 * (ORBX_MAX_LEVELS + 1)];     // exclusive prefix sums of s_cnt: keypoints / lapping keypoints before level l
  __shared__ uint32_t s_org[DESC_NT / 32][DESC_OWORDS];
  __shared__ __align__(16) uint32_t s_blr[DESC_NT / 32][DESC_BWORDS];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {  // transposed: pair (byte row r, pair k) at [k*32 + r], as floats
    const int pw = reinterpret_cast<const int*>(c_pattern)[i];
    s_pat[(i & 7) * 32 + (i >> 3)] = make_float4((float)(int8_t)(pw & 0xff), (float)(int8_t)((pw >> 8) & 0xff),
                                                  (float)(int8_t)((pw >> 16) & 0xff), (float)(int8_t)((pw >> 24) & 0xff));
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const int v = i >> 4, j = i & 7, isMask = (i >> 3) & 1;
    const int d = c_umax[v];
    unsigned w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int u = 4 * j + k - 15;
      if (u >= -d && u <= d) w |= (unsigned)((isMask ? 1 : u) & 0xff) << (8 * k);
    }
    s_icw[v * 17 + (isMask ? 8 : 0) + j] = w;
  }
  if (threadIdx.x < 2 * p.nlevels) {
    const int l = threadIdx.x >> 1;
    s_cnt[threadIdx.x] = (threadIdx.x & 1) ? p.selLap[b * p.nlevels + l] : p.selN[b * p.nlevels + l];
  }
  __syncthreads();
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 2) {
    int acc = 0;
    for (int l = 0; l <= ORBX_MAX_LEVELS; ++l) {
      s_pre[2 * l + threadIdx.x] = acc;
      if (l < p.nlevels) acc += s_cnt[2 * l + threadIdx.x];
    }
  }
  __syncthreads();
  const int total = s_pre[2 * ORBX_MAX_LEVELS], totalLap = s_pre[2 * ORBX_MAX_LEVELS + 1];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    nOut[b] = min(total, cap);
    monoOut[b] = total - totalLap;
    if (total > cap) atomicExch(p.err, 4);
  }
  // lane -> (row, word) of the first staging round of either patch; every further round advances by 32 words
  const int oRow0 = lane / DESC_OW, oCol0 = lane - oRow0 * DESC_OW;
  uint32_t* so = s_org[wid];
  uint32_t* sb = s_blr[wid];
  const uint8_t* sbBytes = reinterpret_cast<const uint8_t*>(sb);

#pragma unroll 1
  for (int it = 0; it < DESC_KPW; ++it) {
    const int kp = blockIdx.x * ((DESC_NT / 32) * DESC_KPW) + it * (DESC_NT / 32) + wid;   // warp-uniform
    if (kp >= total) break;
    // locate (level, i) of this keypoint and the counts it needs for its output slot
    int level = 0;
#pragma unroll
    for (int l = 1; l < ORBX_MAX_LEVELS; ++l) level += (kp >= s_pre[2 * l]) ? 1 : 0;   // prefix entries past nlevels equal total > kp
    const int base = s_pre[2 * level], lapBefore = s_pre[2 * level + 1];
    const int idx = kp - base;
    const LevelParams& L = p.lv[level];
    const uint2 rec = p.sel[(size_t)b * p.selPerImage + L.selOfs + idx];
    const int px = rec.x & 0xffff, py = rec.x >> 16;
    const int score = rec.y & 0xff, lap = (rec.y >> 8) & 1, lapPrefix = rec.y >> 16;

    // --- stage both patches: word q = row * W + col of the patch goes to shared word q.  The un-blurred level may be
    //     caller memory with any stride, so each row is aligned on its own; the blurred plane has a 64-byte pitch ---
    const uint8_t* img = L.pyr + (size_t)b * L.imgStride + (size_t)(py - 15) * L.pitch + (px - 15);
    const uint8_t* blb = L.blur + (size_t)b * L.blurStride + (size_t)(py - 18) * L.blurPitch + (px - 18);
    const unsigned bmis = (unsigned)(reinterpret_cast<uintptr_t>(blb) & 15);    // the blurred plane is 64-byte aligned
    blb -= bmis;
    const unsigned imgLo = (unsigned)reinterpret_cast<uintptr_t>(img);
    __syncwarp();                                 // the previous keypoint's readers are done with so / sb
    {
      int r = oRow0, c = oCol0;
#pragma unroll
      for (int t = 0; t < (31 * DESC_OW + 31) / 32; ++t) {
        if (r < 31) {
          const unsigned ro = (unsigned)r * (unsigned)L.pitch;
          const int off = (int)ro - (int)((imgLo + ro) & 3u) + 4 * c;      // may be negative (-3..) in row 0
          so[r * DESC_OW + c] = __ldg(reinterpret_cast<const uint32_t*>(img + off));
        }
        c += 32 % DESC_OW;
        r += 32 / DESC_OW;
        if (c >= DESC_OW) { c -= DESC_OW; ++r; }
      }
#pragma unroll
      for (int t = 0; t < (37 * 4 + 31) / 32; ++t) {
        const int q = t * 32 + lane, br = q >> 2, bc = q & 3;           // quad bc of patch row br
        if (br < 37)
          *reinterpret_cast<uint4*>(sb + br * DESC_BW + 4 * bc) =
              __ldg(reinterpret_cast<const uint4*>(blb + ((unsigned)br * (unsigned)L.blurPitch + 16u * (unsigned)bc)));
      }
    }
    __syncwarp();

    // --- IC_Angle on the un-blurred level: lane = patch row.  The 31 pixels of the row sit in nine aligned words
    //     (the keypoint lies >= 19 px inside the image, so they never leave it), re-aligned with funnel shifts and
    //     reduced with DP4A against the disc weights: m10 = sum u*I, row sum = sum I (exact integers) ---
    int m10 = 0, m01 = 0;
    if (lane < 31) {
      const int dy = lane - 15;
      const int v = dy < 0 ? -dy : dy;
      const unsigned mis = (imgLo + (unsigned)lane * (unsigned)L.pitch) & 3u;
      const uint32_t* wrow = so + lane * DESC_OW;
      uint32_t w[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) w[j] = wrow[j];
      const unsigned* wt = s_icw + v * 17;
      int rs = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t pix = __funnelshift_r(w[j], w[j + 1], 8 * mis);   // bytes u = 4j-15 .. 4j-12
        m10 = dp4a_us(pix, wt[j], m10);
        rs = (int)__dp4a(pix, wt[8 + j], (unsigned)rs);
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
    double sd, cd;
    sincos((double)ang, &sd, &cd);            // (float)cos((double)angle), (float)sin((double)angle): DESIGN.md §3
    const float a = (float)cd, bsn = (float)sd;
    const uint8_t* bl = sbBytes + 18 * (DESC_BW * 4) + 18 + bmis;   // the keypoint inside the staged patch
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 pw = s_pat[k * 32 + lane];   // (x0,y0,x1,y1)
      const float x0 = pw.x, y0 = pw.y, x1 = pw.z, y1 = pw.w;
      const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, bsn), __fmul_rn(y0, a)));
      const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, bsn)));
      const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, bsn), __fmul_rn(y1, a)));
      const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, bsn)));
      const int t0 = bl[r0 * (DESC_BW * 4) + c0], t1 = bl[r1 * (DESC_BW * 4) + c1];
      val |= (t0 < t1) << k;
    }
    // --- output slot: non-lapping keypoints fill from the front, lapping ones from the back ---
    const int slot = lap ? (total - 1 - (lapBefore + lapPrefix)) : (base - lapBefore + idx - lapPrefix);
    if (slot < 0 || slot >= cap) continue;
    desc[((size_t)b * cap + slot) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
      orbx_keypoint kpo;
      kpo.x = (float)px;
      kpo.y = (float)py;
      if (level != 0) {
        kpo.x = __fmul_rn(kpo.x, L.scale);
        kpo.y = __fmul_rn(kpo.y, L.scale);
      }
      kpo.size = L.kpSize;
      kpo.angle = angle;
      kpo.response = (float)score;
      kpo.octave = level;
      kps[(size_t)b * cap + slot] = kpo;
    }
  }
}

// ------------------------------------------------------------------------------------
// host-side launchers (called from orbx_extract.cu)
// ------------------------------------------------------------------------------------
size_t orbx_fast_smem_bytes(int fastTileBytes, int fastScoreBytes, int fastCandCap) {
  return (size_t)fastTileBytes + fastScoreBytes + 256 + 512 + (size_t)2 * fastCandCap;   // planes + tables + list 1 (uint16)
}
size_t orbx_octree_smem_bytes(int nodeCap) {
  return (size_t)nodeCap * (2 * sizeof(short4) + sizeof(unsigned long long) + 4 * 2 + 16 + 4 * 4);
}

// The dynamic shared-memory opt-in is a per-kernel, process-wide attribute while every extractor configures its own
// need: it is only ever RAISED (a later, smaller extractor must not lower it under an earlier one that is still live).
int orbx_extract_configure(int nodeCap, int fastTileBytes, int fastScoreBytes, int fastCandCap) {
  static std::mutex mu;
  static int fastMax[64] = {}, octMax[64] = {};          // per device
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  ORBX_CUDA(cudaGetDevice(&dev));
  dev &= 63;
  const int fastNeed = (int)orbx_fast_smem_bytes(fastTileBytes, fastScoreBytes, fastCandCap), octNeed = (int)orbx_octree_smem_bytes(nodeCap);
  if (fastNeed > fastMax[dev]) {
    ORBX_CUDA(cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fastNeed));
    fastMax[dev] = fastNeed;
  }
  if (octNeed > octMax[dev]) {
    ORBX_CUDA(cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, octNeed));
    octMax[dev] = octNeed;
  }
  return ORBX_OK;
}

int orbx_extract_launch(orbx_ctx* ctx, cudaStream_t st, const ExtractParams& p, const FastTmaMaps& maps, orbx_keypoint* d_kps,
                        uint8_t* d_desc, int cap, int* d_n, int* d_mono, cudaEvent_t* ev /* [ORBX_EXT_STAGES+1] or null */) {
  const int B = p.B;
  ORBX_CUDA(cudaMemsetAsync(p.candN, 0, sizeof(int) * B * p.nlevels, st));
#define ORBX_EV(i) do { if (ev) ORBX_CUDA(cudaEventRecord(ev[i], st)); } while (0)
  ORBX_EV(0);
  for (int l = 1; l < p.nlevels; ++l) {
    // block width: the row's 4-pixel words split evenly over as few blocks of <= 256 threads as possible, rounded up to
    // whole warps (752 -> 627 px: one block of 160 threads instead of two of 128 with 99 idle lanes)
    const int words = div_up(p.lv[l].w, 4), nbx = div_up(words, 256), nt = div_up(div_up(words, nbx), 32) * 32;
    dim3 grid(nbx, div_up(p.lv[l].h, PYR_STRIP), B);
    // word path: source rows 4-byte aligned (level 0 may be the caller's buffer) and <= 8 source bytes per 4 columns
    const bool wordPath = p.lv[l].pyrSpan <= 8 && (p.lv[l - 1].pitch & 3) == 0 && (p.lv[l - 1].imgStride & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(p.lv[l - 1].pyr) & 3) == 0;
    if (wordPath) pyr_resize_kernel<true><<<grid, nt, 0, st>>>(p, l);
    else pyr_resize_kernel<false><<<grid, nt, 0, st>>>(p, l);
    ORBX_LAUNCH(ctx);
  }
  ORBX_EV(1);
  fast_cells_kernel<<<dim3(p.totalFastTiles, B), 256, orbx_fast_smem_bytes(p.fastTileBytes, p.fastScoreBytes, p.fastCandCap), st>>>(maps, p);
  ORBX_LAUNCH(ctx);
  ORBX_EV(2);
  gauss7_kernel<<<dim3(p.totalBlurTiles, B), 256, 0, st>>>(p);
  ORBX_LAUNCH(ctx);
  ORBX_EV(3);
  octree_kernel<<<dim3(p.nlevels, B), OCT_NT, orbx_octree_smem_bytes(p.nodeCap), st>>>(p);
  ORBX_LAUNCH(ctx);
  ORBX_EV(4);
  const int warpsPerImage = p.selPerImage;
  describe_kernel<<<dim3(div_up(warpsPerImage, (DESC_NT / 32) * DESC_KPW), B), DESC_NT, 0, st>>>(p, d_kps, d_desc, cap, d_n, d_mono);
  ORBX_LAUNCH(ctx);
  ORBX_EV(5);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}
