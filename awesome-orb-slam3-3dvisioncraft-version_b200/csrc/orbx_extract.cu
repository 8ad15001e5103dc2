// orbx_extract.cu — host side of the ORB extractor behind the C ABI (include/orbx.h).
//
// Mirrors the reference's ORBextractor object (include/ORBextractor.h:43-109): the constructor
// tables (src/ORBextractor.cc:408-468) are computed here on the host with the same float/double
// expressions; everything per-image runs on the device (orbx_extract_kernels.cu).
#include "orbx_extract.cuh"
#include <vector>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <cstdlib>

int orbx_extract_configure(int nodeCap, int fastTileBytes, int fastScoreBytes, int fastCandCap);
size_t orbx_octree_smem_bytes(int nodeCap);
size_t orbx_fast_smem_bytes(int fastTileBytes, int fastScoreBytes, int fastCandCap);
int orbx_extract_launch(orbx_ctx* ctx, cudaStream_t st, const ExtractParams& p, const FastTmaMaps& maps, orbx_keypoint* d_kps,
                        uint8_t* d_desc, int cap, int* d_n, int* d_mono, cudaEvent_t* ev);

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_floor_d(double v) { int i = (int)v; return i - (i > v); }

struct orbx_ext {
  orbx_ctx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
  double scaleFactor = 0;
  int maxW = 0, maxH = 0, maxB = 0;
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> nFeat;
  int maxKeypoints = 0;

  // geometry currently configured
  int curW = -1, curH = -1, curStride = -1;
  ExtractParams P{};
  FastTmaMaps maps{};
  bool tmaOk = false;              // the driver exposes cuTensorMapEncodeTiled and ORBX_NO_TMA is unset
  const uint8_t* tmaL0 = nullptr;  // level-0 binding the level-0 tensor map was encoded for
  int tmaL0Pitch = -1;
  // device allocations (sized for maxW x maxH x maxB at creation)
  uint8_t* d_pyr = nullptr;      // levels 0..L-1 + blurred levels
  size_t pyrBytes = 0;
  int16_t* d_tab = nullptr;
  size_t tabElems = 0;
  uint32_t* d_cand = nullptr;
  uint16_t* d_keyNode = nullptr;
  size_t candElems = 0;
  uint2* d_sel = nullptr;
  size_t selElems = 0;
  int* d_counts = nullptr;       // candN | selN | selLap | err
  int4* d_tiles = nullptr;       // FAST tile records of the current geometry
  size_t tilesCap = 0;
  orbx_keypoint* d_kps = nullptr;
  uint8_t* d_desc = nullptr;
  int* d_nOut = nullptr;         // nOut[maxB] | mono[maxB]
  // pinned staging for the host-pointer API
  uint8_t* h_img = nullptr;
  size_t hImgBytes = 0;
  orbx_keypoint* h_kps = nullptr;
  uint8_t* h_desc = nullptr;
  int* h_nOut = nullptr;
  int lastB = 0;
  bool level0External = false;
  bool profiling = false;
  bool profiled = false;
  cudaEvent_t ev[ORBX_EXT_STAGES + 1] = {};
};

static void level_dims(const orbx_ext* e, int w, int h, int l, int* lw, int* lh) {
  const float s = e->invScale[l];  // src/ORBextractor.cc:1162-1163
  *lw = cv_round_f((float)w * s);
  *lh = cv_round_f((float)h * s);
}

// Validity of an image size: the reference divides by nCols/nRows/nIni, which must be >= 1 on
// every level (src/ORBextractor.cc:779-782, :541).
static bool size_supported(const orbx_ext* e, int w, int h) {
  for (int l = 0; l < e->nlevels; ++l) {
    int lw, lh;
    level_dims(e, w, h, l, &lw, &lh);
    if (lw - 32 < 30 || lh - 32 < 30) return false;
    if ((int)std::round((float)(lw - 32) / (float)(lh - 32)) < 1) return false;
    if (lw > 4095 + 16 || lh > 4095 + 16) return false;  // 12-bit packed candidate coordinates
  }
  return true;
}

static size_t pitch_for(int w) { return align_up((size_t)w + 4, 64); }

// ---- TMA tensor maps for the FAST tiles (cuTensorMapEncodeTiled through the runtime's driver entry point: no
// link-time dependency on libcuda) ----
typedef CUresult (*orbx_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static orbx_tmap_encode_fn tmap_encoder() {
  static orbx_tmap_encode_fn fn = []() -> orbx_tmap_encode_fn {
    if (std::getenv("ORBX_NO_TMA")) return nullptr;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (orbx_tmap_encode_fn)p;
  }();
  return fn;
}

// Level l of a batch as a 3-D tensor (bytes of a row, rows, images); box = one FAST tile.  Returns false when the
// layout does not satisfy TMA's 16-byte rules (the kernel then falls back to 32-bit loads for that level).
static bool encode_fast_map(CUtensorMap* m, const LevelParams& L, const uint8_t* base, int pitch, size_t imgStride, int B) {
  orbx_tmap_encode_fn enc = tmap_encoder();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (imgStride & 15) || L.fastTP > 256 || L.fastTH > 256)
    return false;
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)L.h, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)imgStride};
  const cuuint32_t box[3] = {(cuuint32_t)L.fastTP, (cuuint32_t)L.fastTH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Fill P.lv[] geometry for (w,h); level-0 pointer/pitch are set by the caller.
static int configure_geometry(orbx_ext* e, int w, int h, int B) {
  ExtractParams& P = e->P;
  // a failed configuration must not leave a half-written geometry behind a still-valid (curW, curH)
  e->curW = e->curH = -1;
  e->tmaL0 = nullptr;
  P.nlevels = e->nlevels;
  P.iniTh = e->iniTh;
  P.minTh = e->minTh;
  P.nodeCap = 0;
  size_t off = 0;
  int tile = 0, btile = 0, cand = 0, sel = 0, fastBytes = 0, fastScore = 0, fastCand = 0;
  std::vector<int16_t> htab;
  std::vector<int4> htiles;
  for (int l = 0; l < e->nlevels; ++l) {
    LevelParams& L = P.lv[l];
    level_dims(e, w, h, l, &L.w, &L.h);
    L.pitch = (int)pitch_for(L.w);
    L.imgStride = (size_t)L.pitch * L.h;
    L.pyr = e->d_pyr + off;
    off += align_up(L.imgStride * e->maxB, 256);
    L.maxBX = L.w - ORBX_EDGE + 3;
    L.maxBY = L.h - ORBX_EDGE + 3;
    const float width = (float)(L.maxBX - ORBX_MINB), height = (float)(L.maxBY - ORBX_MINB);
    L.nCols = (int)(width / 30.f);
    L.nRows = (int)(height / 30.f);
    L.wCell = (int)std::ceil(width / (float)L.nCols);
    L.hCell = (int)std::ceil(height / (float)L.nRows);
    // cells per tile: as many as fit a TMA box (<= 256 bytes wide).  The box starts at the tile's x origin rounded
    // down to 16 bytes (the TMA unit faults on an unaligned innermost coordinate -- tools/tma_probe), so up to 15
    // leading bytes are dead: pitch = align16(15 + n*wCell + 6 + 4)
    L.fastCells = std::min(ORBX_FAST_CELLS, L.nCols);
    while (L.fastCells > 1 && 15 + L.fastCells * L.wCell + 6 + 4 > ORBX_FAST_TP) --L.fastCells;
    if (15 + L.fastCells * L.wCell + 6 + 4 > ORBX_FAST_TP) {
      orbx_set_error("orbx: FAST cell of %d px does not fit the %d-byte tile", L.wCell, ORBX_FAST_TP);
      e->curW = e->curH = -1;
      return ORBX_ECAP;
    }
    // balance the tiles of a cell row (24 columns: 6+6+6+6 instead of 7+7+7+3)
    L.fastCells = div_up(L.nCols, div_up(L.nCols, L.fastCells));
    L.fastTP = ORBX_FAST_TP;
    L.fastTH = L.hCell + 6;
    L.useTma = 0;
    L.tilesPerRow = div_up(L.nCols, L.fastCells);
    L.tileStart = tile;
    // tile records: one cell row x up to fastCells cells each; cells/rows the reference skips (src/ORBextractor.cc:
    // 792-804: iniY >= maxBorderY-3, iniX >= maxBorderX-6) are dropped here, so every CTA has work
    for (int ci = 0; ci < L.nRows; ++ci) {
      const int iniY = ORBX_MINB + ci * L.hCell;
      if (iniY >= L.maxBY - 3) continue;
      const int maxY = std::min(iniY + L.hCell + 6, L.maxBY);
      for (int j0 = 0; j0 < L.nCols; j0 += L.fastCells) {
        const int iniX = ORBX_MINB + j0 * L.wCell;
        if (iniX >= L.maxBX - 6) continue;
        int nc = std::min(L.fastCells, L.nCols - j0);
        while (nc > 0 && ORBX_MINB + (j0 + nc - 1) * L.wCell >= L.maxBX - 6) --nc;
        if (nc <= 0) continue;
        const int maxX = std::min(iniX + nc * L.wCell + 6, L.maxBX);
        const int tw = maxX - iniX, th = maxY - iniY;
        if (tw - 6 <= 0 || th - 6 <= 0) continue;
        // words of a tile row that hold interior pixels (the kernel's ncw) and the reciprocal it divides items by
        const int off = iniX & 15, ncw = (off + 3 + tw - 6 - 1) / 4 - (off + 3) / 4 + 1;
        htiles.push_back(make_int4(l | (nc << 8) | (((65536 + ncw - 1) / ncw) << 16), iniX | (iniY << 16), tw | (th << 16),
                                   (65536 + L.wCell - 1) / L.wCell));
      }
    }
    tile = (int)htiles.size();
    // one shared-memory plane holds the image tile (fastTP x fastTH) or the score plane ((wI+2) x (hI+2))
    // (the 8-row strips of stage B read up to 7 + 6 rows past the last interior row)
    fastBytes = std::max(fastBytes, ORBX_FAST_TP * (L.hCell + 14));
    fastScore = std::max(fastScore, ORBX_FAST_TP * (L.hCell + 2));
    fastCand = std::max(fastCand, (int)align_up((size_t)(L.fastCells * L.wCell) * L.hCell, 64));
    L.blurTilesX = div_up(L.w, ORBX_BLUR_TW);        // warp tiles per row: 120 columns per warp (30 output words + 2 halo lanes)
    L.blurTilesY = div_up(L.h, ORBX_BLUR_STRIP);     // warp tiles per column: 32-row strips
    L.blurTileStart = btile;
    btile += div_up(L.blurTilesX * L.blurTilesY, 8); // a CTA takes 8 consecutive warp tiles (row-major)
    L.nFeat = e->nFeat[l];
    L.nIni = (int)std::round(width / height);
    L.hX = width / (float)L.nIni;
    L.candOfs = cand;
    L.candCap = (int)((size_t)L.w * L.h * 3 / 10) + 64;
    cand += L.candCap;
    L.selOfs = sel;
    L.selCap = L.nFeat + 4 * L.nIni + 8;
    sel += L.selCap;
    P.nodeCap = std::max(P.nodeCap, L.selCap);
    L.scale = e->scale[l];
    L.kpSize = (float)(int)(31 * e->scale[l]);  // src/ORBextractor.cc:862
  }
  for (int l = 0; l < e->nlevels; ++l) {
    LevelParams& L = P.lv[l];
    L.blur = e->d_pyr + off;
    L.blurPitch = L.pitch;             // internal pitch, also for level 0 (whose `pyr` may be rebound per call)
    L.blurStride = L.imgStride;
    off += align_up(L.imgStride * e->maxB, 256);
  }
  if (off > e->pyrBytes || (size_t)cand > e->candElems / e->maxB || (size_t)sel > e->selElems / e->maxB) {
    orbx_set_error("orbx: internal sizing error (pyr %zu/%zu cand %d sel %d)", off, e->pyrBytes, cand, sel);
    return ORBX_ECAP;
  }
  // resize tables: cv::resize INTER_LINEAR 8U, 11-bit fixed-point coefficients (kernel K1)
  for (int l = 1; l < e->nlevels; ++l) {
    LevelParams& L = P.lv[l];
    const LevelParams& S = P.lv[l - 1];
    const double sx = 1.0 / ((double)L.w / S.w), sy = 1.0 / ((double)L.h / S.h);
    // packed tables: X[dx] = {xofs, a0, a1, 0}, padded to a multiple of 4 columns; Y[dy] = {y0, y1, b0, b1}
    htab.resize(align_up(htab.size(), 8));            // 16-byte alignment for the vector loads
    L.tabX = (int)htab.size();
    const int wpad = (int)align_up((size_t)L.w, 4);
    htab.resize(htab.size() + 4 * (size_t)wpad, 0);
    int16_t* tx = htab.data() + L.tabX;
    for (int dx = 0; dx < L.w; ++dx) {
      float fx = (float)((dx + 0.5) * sx - 0.5);
      int s = cv_floor_d(fx);
      fx -= s;
      if (s < 0) { fx = 0; s = 0; }
      if (s >= S.w - 1) { fx = 0; s = S.w - 1; }
      tx[4 * dx] = (int16_t)s;
      tx[4 * dx + 1] = (int16_t)cv_round_f((1.f - fx) * 2048.f);
      tx[4 * dx + 2] = (int16_t)cv_round_f(fx * 2048.f);
    }
    L.pyrSpan = 0;
    for (int dx = 0; dx < L.w; dx += 4) {
      const int lastCol = std::min(dx + 3, L.w - 1);
      L.pyrSpan = std::max(L.pyrSpan, (int)tx[4 * lastCol] + 1 - (int)tx[4 * dx] + 1);
    }
    L.tabY = (int)htab.size();
    htab.resize(htab.size() + 4 * (size_t)L.h);
    int16_t* ty = htab.data() + L.tabY;
    for (int dy = 0; dy < L.h; ++dy) {
      float fy = (float)((dy + 0.5) * sy - 0.5);
      int s = cv_floor_d(fy);
      fy -= s;
      ty[4 * dy] = (int16_t)std::min(std::max(s, 0), S.h - 1);
      ty[4 * dy + 1] = (int16_t)std::min(std::max(s + 1, 0), S.h - 1);
      ty[4 * dy + 2] = (int16_t)cv_round_f((1.f - fy) * 2048.f);
      ty[4 * dy + 3] = (int16_t)cv_round_f(fy * 2048.f);
    }
  }
  if (htab.size() > e->tabElems) {
    orbx_set_error("orbx: resize table overflow");
    return ORBX_ECAP;
  }
  if (!htab.empty())
    ORBX_CUDA(cudaMemcpyAsync(e->d_tab, htab.data(), htab.size() * sizeof(int16_t), cudaMemcpyHostToDevice, e->stream));
  if (htiles.size() > e->tilesCap) {
    cudaFree(e->d_tiles);
    e->d_tiles = nullptr;
    e->tilesCap = 0;
    ORBX_CUDA(cudaMalloc(&e->d_tiles, htiles.size() * sizeof(int4)));
    e->tilesCap = htiles.size();
  }
  if (!htiles.empty())
    ORBX_CUDA(cudaMemcpyAsync(e->d_tiles, htiles.data(), htiles.size() * sizeof(int4), cudaMemcpyHostToDevice, e->stream));
  ORBX_CUDA(cudaStreamSynchronize(e->stream));  // htab / htiles are locals
  P.fastTiles = e->d_tiles;
  P.tab = e->d_tab;
  P.candPerImage = cand;
  P.selPerImage = sel;
  P.totalFastTiles = tile;
  P.totalBlurTiles = btile;
  P.fastTileBytes = fastBytes;
  P.fastScoreBytes = fastScore;
  P.fastCandCap = fastCand;
  P.cand = e->d_cand;
  P.keyNode = e->d_keyNode;
  P.sel = e->d_sel;
  P.candN = e->d_counts;
  P.selN = e->d_counts + e->maxB * e->nlevels;
  P.selLap = e->d_counts + 2 * e->maxB * e->nlevels;
  P.err = e->d_counts + 3 * e->maxB * e->nlevels;
  if (orbx_fast_smem_bytes(fastBytes, fastScore, fastCand) > 200 * 1024 || orbx_octree_smem_bytes(P.nodeCap) > 200 * 1024) {
    orbx_set_error("orbx: shared-memory budget exceeded (fast %zu, octree %zu)", orbx_fast_smem_bytes(fastBytes, fastScore, fastCand),
                   orbx_octree_smem_bytes(P.nodeCap));
    return ORBX_ECAP;
  }
  int rc = orbx_extract_configure(P.nodeCap, fastBytes, fastScore, fastCand);
  if (rc != ORBX_OK) return rc;
  for (int l = 1; l < e->nlevels; ++l) {
    LevelParams& L = P.lv[l];
    L.useTma = (l < ORBX_TMA_LEVELS && encode_fast_map(&e->maps.m[l], L, L.pyr, L.pitch, L.imgStride, e->maxB)) ? 1 : 0;
  }
  e->tmaL0 = nullptr;   // level 0 is (re)bound per call
  e->curW = w;
  e->curH = h;
  (void)B;
  return ORBX_OK;
}

extern "C" {

orbx_ext* orbx_extractor_create(orbx_ctx* ctx, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                                int max_w, int max_h, int max_batch) {
  if (!ctx || nfeatures < 1 || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !(scaleFactor > 1.f) || iniTh < 1 ||
      minTh < 1 || iniTh > 254 || minTh > iniTh || max_w < 1 || max_h < 1 || max_batch < 1) {
    orbx_set_error("orbx_extractor_create: invalid argument");
    return nullptr;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
  orbx_ext* e = new orbx_ext();
  e->ctx = ctx;
  e->nfeatures = nfeatures;
  e->nlevels = nlevels;
  e->iniTh = iniTh;
  e->minTh = minTh;
  e->scaleFactor = scaleFactor;  // float ctor argument stored in a double member (ORBextractor.h:97)
  e->maxW = max_w;
  e->maxH = max_h;
  e->maxB = max_batch;
  // --- src/ORBextractor.cc:413-444 ---
  e->scale.assign(nlevels, 1.f);
  e->sigma2.assign(nlevels, 1.f);
  for (int i = 1; i < nlevels; ++i) {
    e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);
    e->sigma2[i] = e->scale[i] * e->scale[i];
  }
  e->invScale.resize(nlevels);
  e->invSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; ++i) {
    e->invScale[i] = 1.0f / e->scale[i];
    e->invSigma2[i] = 1.0f / e->sigma2[i];
  }
  e->nFeat.resize(nlevels);
  {
    float factor = (float)(1.0f / e->scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
      e->nFeat[l] = cv_round_f(nDesired);
      sum += e->nFeat[l];
      nDesired *= factor;
    }
    e->nFeat[nlevels - 1] = std::max(nfeatures - sum, 0);
  }
  if (!size_supported(e, max_w, max_h)) {
    orbx_set_error("orbx_extractor_create: %dx%d too small/elongated for %d levels at scale %.3f", max_w, max_h,
                   nlevels, scaleFactor);
    delete e;
    return nullptr;
  }
  // --- device memory, sized for the maximum geometry ---
  size_t pyr = 0, cand = 0, sel = 0, tab = 0;
  for (int l = 0; l < nlevels; ++l) {
    int lw, lh;
    level_dims(e, max_w, max_h, l, &lw, &lh);
    pyr += 2 * align_up(pitch_for(lw) * (size_t)lh * max_batch, 256);
    cand += (size_t)lw * lh * 3 / 10 + 64;
    const int nIni = std::max(1, (int)std::round((float)(lw - 32) / (float)(lh - 32)));
    sel += e->nFeat[l] + 4 * nIni + 8;
    tab += 4 * ((size_t)lw + 4) + 4 * (size_t)lh + 16;
  }
  // smaller images than max may have a larger aspect-driven nIni; keep some slack
  sel += 16 * nlevels;
  e->maxKeypoints = (int)sel;
  e->pyrBytes = pyr + 4096;
  e->candElems = cand * max_batch;
  e->selElems = sel * max_batch;
  e->tabElems = tab + 64;
  bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_pyr, e->pyrBytes) == cudaSuccess;
  // the kernels read whole aligned words / quads that may include a row's pitch padding (never used, never written): give
  // those bytes a defined value once, so that initcheck stays quiet and a leaked padding byte could not go unnoticed
  ok = ok && cudaMemset(e->d_pyr, 0, e->pyrBytes) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_tab, e->tabElems * sizeof(int16_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_cand, e->candElems * sizeof(uint32_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_keyNode, e->candElems * sizeof(uint16_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_sel, e->selElems * sizeof(uint2)) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_counts, (3 * (size_t)max_batch * nlevels + 16) * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_kps, (size_t)max_batch * sel * sizeof(orbx_keypoint)) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_desc, (size_t)max_batch * sel * 32) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_nOut, 2 * (size_t)max_batch * sizeof(int)) == cudaSuccess;
  e->hImgBytes = pitch_for(max_w) * (size_t)max_h * max_batch;
  ok = ok && cudaMallocHost(&e->h_img, e->hImgBytes) == cudaSuccess;
  ok = ok && cudaMallocHost(&e->h_kps, (size_t)max_batch * sel * sizeof(orbx_keypoint)) == cudaSuccess;
  ok = ok && cudaMallocHost(&e->h_desc, (size_t)max_batch * sel * 32) == cudaSuccess;
  ok = ok && cudaMallocHost(&e->h_nOut, (2 * (size_t)max_batch + 4) * sizeof(int)) == cudaSuccess;
  if (ok) ok = cudaMemset(e->d_counts, 0, (3 * (size_t)max_batch * nlevels + 16) * sizeof(int)) == cudaSuccess;
  // the host-buffer entry points copy whole [B][cap] output blocks back and use the first n entries of each: defined bytes
  if (ok) ok = cudaMemset(e->d_kps, 0, (size_t)max_batch * sel * sizeof(orbx_keypoint)) == cudaSuccess;
  if (ok) ok = cudaMemset(e->d_desc, 0, (size_t)max_batch * sel * 32) == cudaSuccess;
  if (ok) ok = cudaMemset(e->d_cand, 0, e->candElems * sizeof(uint32_t)) == cudaSuccess;
  if (ok) ok = cudaMemset(e->d_sel, 0, e->selElems * sizeof(uint2)) == cudaSuccess;
  if (ok) ok = cudaStreamSynchronize(0) == cudaSuccess;   // the memsets ran on the legacy stream; e->stream is non-blocking
  if (!ok) {
    orbx_set_error("orbx_extractor_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    orbx_extractor_destroy(e);
    return nullptr;
  }
  return e;
}

void orbx_extractor_destroy(orbx_ext* e) {
  if (!e) return;
  cudaSetDevice(e->ctx->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  cudaFree(e->d_pyr);
  cudaFree(e->d_tab);
  cudaFree(e->d_cand);
  cudaFree(e->d_keyNode);
  cudaFree(e->d_sel);
  cudaFree(e->d_counts);
  cudaFree(e->d_tiles);
  cudaFree(e->d_kps);
  cudaFree(e->d_desc);
  cudaFree(e->d_nOut);
  cudaFreeHost(e->h_img);
  cudaFreeHost(e->h_kps);
  cudaFreeHost(e->h_desc);
  cudaFreeHost(e->h_nOut);
  for (int i = 0; i <= ORBX_EXT_STAGES; ++i)
    if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int orbx_extractor_levels(const orbx_ext* e) { return e ? e->nlevels : ORBX_EINVAL; }

int orbx_extractor_scale_tables(const orbx_ext* e, float* scale, float* inv, float* s2, float* is2) {
  if (!e) return ORBX_EINVAL;
  for (int l = 0; l < e->nlevels; ++l) {
    if (scale) scale[l] = e->scale[l];
    if (inv) inv[l] = e->invScale[l];
    if (s2) s2[l] = e->sigma2[l];
    if (is2) is2[l] = e->invSigma2[l];
  }
  return e->nlevels;
}

int orbx_extractor_features_per_level(const orbx_ext* e, int* n) {
  if (!e || !n) return ORBX_EINVAL;
  for (int l = 0; l < e->nlevels; ++l) n[l] = e->nFeat[l];
  return e->nlevels;
}

int orbx_extractor_set_profiling(orbx_ext* e, int enable) {
  if (!e) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(e->ctx->device));
  if (enable && !e->ev[0])
    for (int i = 0; i <= ORBX_EXT_STAGES; ++i) ORBX_CUDA(cudaEventCreate(&e->ev[i]));
  e->profiling = enable != 0;
  e->profiled = false;
  return ORBX_OK;
}

int orbx_extractor_stage_ms(orbx_ext* e, float* ms, int* launches) {
  if (!e || !ms || !e->profiled) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(e->ctx->device));
  ORBX_CUDA(cudaEventSynchronize(e->ev[ORBX_EXT_STAGES]));
  for (int i = 0; i < ORBX_EXT_STAGES; ++i) ORBX_CUDA(cudaEventElapsedTime(&ms[i], e->ev[i], e->ev[i + 1]));
  if (launches) {
    launches[0] = e->nlevels - 1;
    for (int i = 1; i < ORBX_EXT_STAGES; ++i) launches[i] = 1;
  }
  return ORBX_OK;
}

int orbx_extractor_max_keypoints(const orbx_ext* e) { return e ? e->maxKeypoints : ORBX_EINVAL; }
void* orbx_extractor_stream(orbx_ext* e) { return e ? (void*)e->stream : nullptr; }

// shared tail of the three extract entry points: geometry, level-0 binding, launch
static int run_device(orbx_ext* e, int B, const uint8_t* d_level0, int w, int h, int stride, int lap0, int lap1,
                      orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_n, int* d_mono) {
  if (w > e->maxW || h > e->maxH || B > e->maxB || B < 1) {
    orbx_set_error("orbx_extract: %dx%d x%d exceeds the extractor's capacity %dx%d x%d", w, h, B, e->maxW, e->maxH,
                   e->maxB);
    return ORBX_EINVAL;
  }
  if (!size_supported(e, w, h)) {
    orbx_set_error("orbx_extract: image %dx%d unsupported (a pyramid level is narrower than one FAST cell)", w, h);
    return ORBX_EINVAL;
  }
  if (w != e->curW || h != e->curH) {
    int rc = configure_geometry(e, w, h, B);
    if (rc != ORBX_OK) return rc;
  }
  ExtractParams& P = e->P;
  P.B = B;
  P.lap0 = lap0;
  P.lap1 = lap1;
  P.lv[0].pyr = const_cast<uint8_t*>(d_level0);
  P.lv[0].pitch = stride;
  P.lv[0].imgStride = (size_t)stride * h;
  if (e->tmaL0 != d_level0 || e->tmaL0Pitch != stride) {   // level-0 tensor map follows the caller's buffer
    P.lv[0].useTma = encode_fast_map(&e->maps.m[0], P.lv[0], d_level0, stride, (size_t)stride * h, e->maxB) ? 1 : 0;
    e->tmaL0 = d_level0;
    e->tmaL0Pitch = stride;
  }
  e->lastB = B;
  e->profiled = e->profiling;
  return orbx_extract_launch(e->ctx, e->stream, P, e->maps, d_kps, d_desc, cap, d_n, d_mono, e->profiling ? e->ev : nullptr);
}

int orbx_extract_batch_device(orbx_ext* e, int B, const uint8_t* d_imgs, int w, int h, int stride, int lap0, int lap1,
                              orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_n_out, int* d_mono_out) {
  if (!e || !d_imgs || !d_kps || !d_desc || !d_n_out || !d_mono_out || cap < 1) return ORBX_EINVAL;
  if (w <= 0 || h <= 0) return ORBX_EMPTY;
  if (stride < w) return ORBX_EINVAL;
  if ((stride & 3) || (reinterpret_cast<uintptr_t>(d_imgs) & 3)) {
    orbx_set_error("orbx_extract_batch_device: image base and stride must be multiples of 4 bytes (32-bit tile loads)");
    return ORBX_EINVAL;
  }
  ORBX_CUDA(cudaSetDevice(e->ctx->device));
  e->level0External = true;
  return run_device(e, B, d_imgs, w, h, stride, lap0, lap1, d_kps, d_desc, cap, d_n_out, d_mono_out);
}

int orbx_extract_batch(orbx_ext* e, int B, const uint8_t* const* imgs, int w, int h, int stride, int lap0, int lap1,
                       orbx_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out) {
  if (!e || !imgs || B < 1) return ORBX_EINVAL;
  if (w <= 0 || h <= 0) return ORBX_EMPTY;
  for (int b = 0; b < B; ++b)
    if (!imgs[b]) return ORBX_EMPTY;
  if (stride < w || !kps || !desc || !n_out || cap < 1) return ORBX_EINVAL;
  if (B > e->maxB || w > e->maxW || h > e->maxH) {
    orbx_set_error("orbx_extract_batch: request exceeds extractor capacity");
    return ORBX_EINVAL;
  }
  ORBX_CUDA(cudaSetDevice(e->ctx->device));
  // stage into pinned memory with the device pitch, one H2D copy for the whole batch
  const int pitch = (int)pitch_for(w);
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < h; ++y)
      std::memcpy(e->h_img + ((size_t)b * h + y) * pitch, imgs[b] + (size_t)y * stride, w);
  // level-0 lives at the start of d_pyr (configure_geometry lays levels out in order)
  uint8_t* d_l0 = e->d_pyr;
  ORBX_CUDA(cudaMemcpyAsync(d_l0, e->h_img, (size_t)B * h * pitch, cudaMemcpyHostToDevice, e->stream));
  e->level0External = false;
  const int dcap = e->maxKeypoints;
  int rc = run_device(e, B, d_l0, w, h, pitch, lap0, lap1, e->d_kps, e->d_desc, dcap, e->d_nOut, e->d_nOut + e->maxB);
  if (rc != ORBX_OK) return rc;
  ORBX_CUDA(cudaMemcpyAsync(e->h_nOut, e->d_nOut, sizeof(int) * 2 * e->maxB, cudaMemcpyDeviceToHost, e->stream));
  ORBX_CUDA(cudaMemcpyAsync(e->h_nOut + 2 * e->maxB, e->P.err, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  ORBX_CUDA(cudaMemcpyAsync(e->h_kps, e->d_kps, sizeof(orbx_keypoint) * (size_t)B * dcap, cudaMemcpyDeviceToHost,
                            e->stream));
  ORBX_CUDA(cudaMemcpyAsync(e->h_desc, e->d_desc, (size_t)32 * B * dcap, cudaMemcpyDeviceToHost, e->stream));
  ORBX_CUDA(cudaStreamSynchronize(e->stream));
  if (e->h_nOut[2 * e->maxB] != 0) {
    orbx_set_error("orbx_extract: device capacity overflow (code %d)", e->h_nOut[2 * e->maxB]);
    ORBX_CUDA(cudaMemsetAsync(e->P.err, 0, sizeof(int), e->stream));
    return ORBX_ECAP;
  }
  for (int b = 0; b < B; ++b) {
    const int n = e->h_nOut[b];
    if (n > cap) {
      orbx_set_error("orbx_extract: %d keypoints exceed caller capacity %d", n, cap);
      return ORBX_ECAP;
    }
    std::memcpy(kps + (size_t)b * cap, e->h_kps + (size_t)b * dcap, sizeof(orbx_keypoint) * n);
    std::memcpy(desc + (size_t)b * cap * 32, e->h_desc + (size_t)b * dcap * 32, (size_t)32 * n);
    n_out[b] = n;
    if (mono_out) mono_out[b] = e->h_nOut[e->maxB + b];
  }
  return ORBX_OK;
}

int orbx_extract(orbx_ext* e, const uint8_t* img, int w, int h, int stride, int lap0, int lap1, orbx_keypoint* kps,
                 uint8_t* desc, int cap, int* n_out, int* mono_out) {
  if (!e) return ORBX_EINVAL;
  if (!img || w <= 0 || h <= 0) return ORBX_EMPTY;
  int n = 0, mono = 0;
  int rc = orbx_extract_batch(e, 1, &img, w, h, stride, lap0, lap1, kps, desc, cap, &n, &mono);
  if (n_out) *n_out = n;
  if (mono_out) *mono_out = mono;
  return rc;
}

int orbx_pyramid_level(orbx_ext* e, int b, int level, uint8_t* dst, int dst_stride, int* w_out, int* h_out) {
  if (!e || e->curW < 0 || level < 0 || level >= e->nlevels || b < 0 || b >= e->lastB) return ORBX_EINVAL;
  const LevelParams& L = e->P.lv[level];
  if (w_out) *w_out = L.w;
  if (h_out) *h_out = L.h;
  if (!dst) return ORBX_OK;
  if (dst_stride < L.w) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(e->ctx->device));
  ORBX_CUDA(cudaMemcpy2DAsync(dst, dst_stride, L.pyr + (size_t)b * L.imgStride, L.pitch, L.w, L.h,
                              cudaMemcpyDeviceToHost, e->stream));
  ORBX_CUDA(cudaStreamSynchronize(e->stream));
  return ORBX_OK;
}

int orbx_debug_blur_level(orbx_ext* e, int b, int level, uint8_t* dst, int dst_stride, int* w_out, int* h_out) {
  if (!e || e->curW < 0 || level < 0 || level >= e->nlevels || b < 0 || b >= e->lastB) return ORBX_EINVAL;
  const LevelParams& L = e->P.lv[level];
  if (w_out) *w_out = L.w;
  if (h_out) *h_out = L.h;
  if (!dst) return ORBX_OK;
  if (dst_stride < L.w) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(e->ctx->device));
  ORBX_CUDA(cudaMemcpy2DAsync(dst, dst_stride, L.blur + (size_t)b * L.blurStride, L.blurPitch, L.w, L.h,
                              cudaMemcpyDeviceToHost, e->stream));
  ORBX_CUDA(cudaStreamSynchronize(e->stream));
  return ORBX_OK;
}

int orbx_debug_candidates(orbx_ext* e, int b, int level, int16_t* xy, uint8_t* score, int cap, int* n_out) {
  if (!e || e->curW < 0 || level < 0 || level >= e->nlevels || b < 0 || b >= e->lastB) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(e->ctx->device));
  const LevelParams& L = e->P.lv[level];
  int n = 0;
  ORBX_CUDA(cudaMemcpy(&n, e->P.candN + b * e->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  n = std::min(n, L.candCap);
  if (n_out) *n_out = n;
  if (n > cap) return ORBX_ECAP;
  std::vector<uint32_t> tmp(n);
  if (n)
    ORBX_CUDA(cudaMemcpy(tmp.data(), e->P.cand + (size_t)b * e->P.candPerImage + L.candOfs, sizeof(uint32_t) * n,
                         cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) {
    xy[2 * i] = (int16_t)(tmp[i] & 0xfff);
    xy[2 * i + 1] = (int16_t)((tmp[i] >> 12) & 0xfff);
    score[i] = (uint8_t)(tmp[i] >> 24);
  }
  return ORBX_OK;
}

}  // extern "C"

// Internal: the extractor's own level-0 storage (where host-pointer calls upload their images); the asynchronous tracker
// pipeline copies its staged inputs there so that level 0 is stable while the next step's H2D is already running.
uint8_t* orbx_ext_level0_storage(orbx_ext* e, size_t* bytes) {
  if (!e) return nullptr;
  if (bytes) *bytes = align_up((size_t)pitch_for(e->maxW) * e->maxH * e->maxB, 256);
  return e->d_pyr;
}

// Internal: device view of image b's pyramid of the last extract call (used by the stereo matcher).
int orbx_ext_pyramid_view(orbx_ext* e, int b, int* nlevels, const uint8_t** ptr, int* w, int* h, int* pitch, float* scale,
                          float* invScale, cudaStream_t* st) {
  if (!e || e->curW < 0 || b < 0 || b >= e->lastB) {
    orbx_set_error("orbx: extractor has no pyramid for image %d (run an extraction first)", b);
    return ORBX_EINVAL;
  }
  *nlevels = e->nlevels;
  for (int l = 0; l < e->nlevels; ++l) {
    const LevelParams& L = e->P.lv[l];
    ptr[l] = L.pyr + (size_t)b * L.imgStride;
    w[l] = L.w;
    h[l] = L.h;
    pitch[l] = L.pitch;
    scale[l] = e->scale[l];
    invScale[l] = e->invScale[l];
  }
  *st = e->stream;
  return ORBX_OK;
}
