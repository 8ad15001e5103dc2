// orbx_voc.cu — DBoW2 vocabulary tree on the device: `transform` of a frame's descriptors into BowVector and
// FeatureVector.
//
// Replaces (reference paths, vendored DBoW2): TemplatedVocabulary::loadFromBinaryFile
// Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1442-1478, transform(features, BowVector, FeatureVector, levelsup)
// :1140-1219 over transform(feature, word, weight, nid, levelsup) :1231-1271, BowVector::addWeight / addIfNotExist /
// normalize Thirdparty/DBoW2/DBoW2/BowVector.cpp:32-84, FeatureVector::addFeature FeatureVector.cpp:31-45,
// FORB::distance FORB.cpp:81-101.  Callers: Frame::ComputeBoW src/Frame.cc:865-872, KeyFrame::ComputeBoW
// src/KeyFrame.cc:125-134 (levelsup = 4).
//
// B200 layout: the whole tree (ORBvoc: 1 082 073 nodes x 32-byte descriptors = 35 MB, children CSR, weights) lives in
// HBM and fits the 126 MB L2, so the k-ary descent of every feature is an L2-resident gather + __popc.
//   K13 voc_descend_kernel : one thread per feature walks root -> leaf (first minimum wins ties, like `d < best_d`)
//   K14 voc_collect_kernel : one CTA per frame; bitonic sort of (word, feature) and (node, feature) keys in shared
//       memory, run-length -> BowVector values (repeated addition of the word's weight, exactly the sums addWeight
//       produces), sequential L1/L2 norm in ascending word order (the order std::map iterates), FeatureVector CSR.
#include <algorithm>
#include <cstring>
#include <vector>
#include "orbx_match.cuh"

struct orbx_voc {
  orbx_ctx* ctx = nullptr;
  int k = 0, L = 0, scoring = 0, weighting = 0, nNodes = 0, nWords = 0;
  uint8_t* d_desc = nullptr;     // [nNodes][32]
  int* d_childOfs = nullptr;     // [nNodes + 1] CSR of children (node-id order)
  int* d_childIdx = nullptr;     // [nNodes - 1]
  int* d_wordId = nullptr;       // [nNodes] (-1 for inner nodes)
  double* d_weight = nullptr;    // [nNodes]
};

struct VocDev {
  const uint8_t* desc;
  const int* childOfs;
  const int* childIdx;
  const int* wordId;
  const double* weight;
  int L, scoring, weighting;
};

__device__ __forceinline__ int voc_hamming(const uint4 a0, const uint4 a1, const uint8_t* b) {
  const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(b)), b1 = __ldg(reinterpret_cast<const uint4*>(b) + 1);
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
         __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// K13: feature i of frame blockIdx.y -> leaf node reached (or -1 when the word's weight is 0: "stopped") and the node
// passed at level L - levelsup
__global__ void __launch_bounds__(128) voc_descend_kernel(VocDev V, const uint8_t* __restrict__ desc, const int* __restrict__ nDev,
                                                          int n, int cap, int levelsup, int* __restrict__ leaf, int* __restrict__ node) {
  const int f = blockIdx.y;
  const int nf = nDev ? min(nDev[f], cap) : n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nf) return;
  const uint8_t* d = desc + ((size_t)f * cap + i) * 32;
  const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(d)), a1 = __ldg(reinterpret_cast<const uint4*>(d) + 1);
  const int nidLevel = V.L - levelsup;
  int nid = 0, cur = 0, level = 0;
  int beg = V.childOfs[0], end = V.childOfs[1];
  while (end > beg) {                      // isLeaf() == children.empty()
    ++level;
    int best = V.childIdx[beg];
    int bestD = voc_hamming(a0, a1, V.desc + (size_t)best * 32);
    for (int c = beg + 1; c < end; ++c) {  // `d < best_d`: the first minimum wins
      const int id = V.childIdx[c];
      const int dd = voc_hamming(a0, a1, V.desc + (size_t)id * 32);
      if (dd < bestD) { bestD = dd; best = id; }
    }
    cur = best;
    if (level == nidLevel) nid = cur;
    beg = V.childOfs[cur];
    end = V.childOfs[cur + 1];
  }
  const bool stopped = cur == 0 || !(V.weight[cur] > 0);
  leaf[(size_t)f * cap + i] = stopped ? -1 : cur;
  node[(size_t)f * cap + i] = nid;
}

// in-place ascending bitonic sort of P (power of two) 64-bit keys in shared memory
__device__ void bitonic_sort_u64(unsigned long long* s, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s[i] = b; s[ixj] = a; }
        }
      }
    }
  }
  __syncthreads();
}

// s_aux[i] = ordinal of the run that starts at sorted position i (runs = equal high words), -1 elsewhere; returns #runs
__device__ int run_ordinals(const unsigned long long* s_key, int* s_aux, int P, int* s_cnt) {
  const int tid = threadIdx.x;
  const unsigned long long SENT = ~0ull;
  for (int i = tid; i < P; i += blockDim.x)
    s_aux[i] = s_key[i] != SENT && (i == 0 || (unsigned)(s_key[i] >> 32) != (unsigned)(s_key[i - 1] >> 32));
  __syncthreads();
  if (tid < 32) {                       // exclusive scan of the head flags by one warp, 32 positions per step
    int carry = 0;
    for (int b0 = 0; b0 < P; b0 += 32) {
      const int v = s_aux[b0 + tid];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += t;
      }
      s_aux[b0 + tid] = v ? (carry + incl - 1) : -1;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (tid == 0) *s_cnt = carry;
  }
  __syncthreads();
  return *s_cnt;
}

struct VocOut {
  int* bowWord;     // [F][cap]
  double* bowVal;   // [F][cap]
  int* nBow;        // [F]
  int* fvNode;      // [F][cap]
  int* fvOff;       // [F][cap + 1]
  int* fvIdx;       // [F][cap]
  int* nFv;         // [F]
};

// K14: one CTA per frame.  P = power of two >= cap; dynamic shared memory = 12 * P bytes.
__global__ void __launch_bounds__(256) voc_collect_kernel(VocDev V, const int* __restrict__ nDev, int n, int cap, int P,
                                                          const int* __restrict__ leaf, const int* __restrict__ node, VocOut O) {
  extern __shared__ __align__(16) unsigned long long s_key[];   // [P] keys (reused as [P] doubles), then [P] ints
  int* s_aux = reinterpret_cast<int*>(s_key + P);
  double* s_val = reinterpret_cast<double*>(s_key);
  __shared__ int s_cnt;
  __shared__ double s_norm;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int nf = nDev ? min(nDev[f], cap) : n;
  const int* lf = leaf + (size_t)f * cap;
  const int* nd = node + (size_t)f * cap;
  int* bowWord = O.bowWord + (size_t)f * cap;
  double* bowVal = O.bowVal + (size_t)f * cap;
  int* fvNode = O.fvNode + (size_t)f * cap;
  int* fvOff = O.fvOff + (size_t)f * (cap + 1);
  int* fvIdx = O.fvIdx + (size_t)f * cap;
  const unsigned long long SENT = ~0ull;
  // ScoringObject::mustNormalize: L1_NORM 0 -> L1, L2_NORM 1 -> L2, CHI_SQUARE 2 / KL 3 / BHATTACHARYYA 4 -> L1,
  // DOT_PRODUCT 5 -> no normalisation
  const bool must = V.scoring != 5, l2 = V.scoring == 1;
  const bool tf = V.weighting == 0 || V.weighting == 1;     // TF_IDF, TF: addWeight ; IDF, BINARY: addIfNotExist

  // ---------------- BowVector: word ids grow with the leaf's node id (both follow file order), so sorting by leaf
  // is the std::map<WordId, WordValue> iteration order ----------------
  for (int i = tid; i < P; i += 256)
    s_key[i] = (i < nf && lf[i] >= 0) ? (((unsigned long long)(unsigned)lf[i] << 32) | (unsigned)i) : SENT;
  bitonic_sort_u64(s_key, P);
  const int nBow = run_ordinals(s_key, s_aux, P, &s_cnt);
  for (int i = tid; i < P; i += 256) {
    const int r = s_aux[i];
    if (r < 0) continue;
    const unsigned lid = (unsigned)(s_key[i] >> 32);
    int cnt = 1;
    while (i + cnt < P && s_key[i + cnt] != SENT && (unsigned)(s_key[i + cnt] >> 32) == lid) ++cnt;
    const double wt = V.weight[lid];
    double v = wt;                                   // first addWeight inserts w, every further one does += w
    if (tf)
      for (int c = 1; c < cnt; ++c) v += wt;
    if (tf && !must) v /= (double)nBow;              // (:1177-1183, only without normalisation)
    bowWord[r] = V.wordId[lid];
    bowVal[r] = v;
  }
  __syncthreads();                                   // keys are dead from here: the buffer becomes s_val
  for (int r = tid; r < nBow; r += 256) s_val[r] = bowVal[r];
  __syncthreads();
  if (must) {
    if (tid == 0) {                                  // BowVector::normalize: sequential sum in ascending word order
      double norm = 0.0;
      if (!l2) { for (int r = 0; r < nBow; ++r) norm += fabs(s_val[r]); }
      else { for (int r = 0; r < nBow; ++r) norm += s_val[r] * s_val[r]; norm = sqrt(norm); }
      s_norm = norm;
    }
    __syncthreads();
    const double norm = s_norm;
    if (norm > 0.0)
      for (int r = tid; r < nBow; r += 256) bowVal[r] = s_val[r] / norm;
  }
  if (tid == 0) O.nBow[f] = nBow;
  __syncthreads();

  // ---------------- FeatureVector: node -> feature indices (ascending), only for features that were not stopped ------
  for (int i = tid; i < P; i += 256)
    s_key[i] = (i < nf && lf[i] >= 0) ? (((unsigned long long)(unsigned)nd[i] << 32) | (unsigned)i) : SENT;
  bitonic_sort_u64(s_key, P);
  const int nFv = run_ordinals(s_key, s_aux, P, &s_cnt);
  int nValid = 0;
  for (int i = tid; i < P; i += 256) {
    if (s_key[i] == SENT) continue;
    fvIdx[i] = (int)(s_key[i] & 0xffffffffu);        // valid keys sort in front of the sentinels: position = CSR slot
    const int r = s_aux[i];
    if (r >= 0) { fvNode[r] = (int)(s_key[i] >> 32); fvOff[r] = i; }
    ++nValid;
  }
  // total number of valid features -> closing offset
  __shared__ int s_total;
  if (tid == 0) s_total = 0;
  __syncthreads();
  if (nValid) atomicAdd(&s_total, nValid);
  __syncthreads();
  if (tid == 0) { fvOff[nFv] = s_total; O.nFv[f] = nFv; }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static void voc_free(orbx_voc* v) {
  if (!v) return;
  cudaFree(v->d_desc);
  cudaFree(v->d_childOfs);
  cudaFree(v->d_childIdx);
  cudaFree(v->d_wordId);
  cudaFree(v->d_weight);
  delete v;
}

static VocDev voc_dev(const orbx_voc* v) {
  VocDev D;
  D.desc = v->d_desc; D.childOfs = v->d_childOfs; D.childIdx = v->d_childIdx; D.wordId = v->d_wordId; D.weight = v->d_weight;
  D.L = v->L; D.scoring = v->scoring; D.weighting = v->weighting;
  return D;
}

static int voc_launch(orbx_voc* v, cudaStream_t st, int F, const uint8_t* d_desc, const int* d_n, int n, int cap, int levelsup,
                      int* d_leaf, int* d_node, const VocOut& O) {
  int P = 32;
  while (P < cap) P <<= 1;
  if (P > 8192) { orbx_set_error("orbx_vocabulary_transform: more than 8192 features per frame"); return ORBX_ECAP; }
  const size_t smem = (size_t)12 * P;
  static int configured = 0;
  if (smem > 48 * 1024 && configured < (int)smem) {
    ORBX_CUDA(cudaFuncSetAttribute(voc_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 8192));
    configured = 12 * 8192;
  }
  const VocDev D = voc_dev(v);
  voc_descend_kernel<<<dim3(div_up(cap, 128), F), 128, 0, st>>>(D, d_desc, d_n, n, cap, levelsup, d_leaf, d_node);
  ORBX_LAUNCH(v->ctx);
  voc_collect_kernel<<<F, 256, smem, st>>>(D, d_n, n, cap, P, d_leaf, d_node, O);
  ORBX_LAUNCH(v->ctx);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

extern "C" {

orbx_voc* orbx_vocabulary_from_memory(orbx_ctx* ctx, const void* data, size_t bytes) {
  if (!ctx || !data || bytes < 24) { orbx_set_error("orbx_vocabulary_from_memory: bad arguments"); return nullptr; }
  const uint8_t* p = (const uint8_t*)data;
  uint32_t nb, sz;
  int32_t k, L, sc, we;
  memcpy(&nb, p, 4); memcpy(&sz, p + 4, 4); memcpy(&k, p + 8, 4); memcpy(&L, p + 12, 4); memcpy(&sc, p + 16, 4); memcpy(&we, p + 20, 4);
  if (sz != 41 || nb < 1 || sc < 0 || sc > 5 || we < 0 || we > 3) {
    orbx_set_error("orbx_vocabulary: not a DBoW2 binary ORB vocabulary (node size %u, scoring %d, weighting %d)", sz, sc, we);
    return nullptr;
  }
  // the header's node count includes the root; the file holds one 41-byte record per non-root node:
  // parent id (int32) | descriptor (32 bytes) | weight (float32) | is-leaf (1 byte)   (:1456-1473)
  const size_t nrec = (bytes - 24) / sz;
  const int N = (int)nrec + 1;
  std::vector<int> parent(N, 0), wordId(N, -1), count(N + 1, 0);
  std::vector<uint8_t> desc((size_t)N * 32, 0);
  std::vector<double> weight(N, 0.0);
  int nWords = 0;
  for (size_t r = 0; r < nrec; ++r) {
    const uint8_t* rec = p + 24 + r * sz;
    const int nid = (int)r + 1;
    int32_t par;
    float w;
    memcpy(&par, rec, 4);
    memcpy(&w, rec + 36, 4);
    if (par < 0 || par >= N) { orbx_set_error("orbx_vocabulary: node %d has parent %d", nid, par); return nullptr; }
    parent[nid] = par;
    ++count[par];
    memcpy(&desc[(size_t)nid * 32], rec + 4, 32);
    weight[nid] = (double)w;                        // WordValue is double; the file stores float
    if (rec[40]) wordId[nid] = nWords++;            // words are numbered in file order
  }
  std::vector<int> ofs(N + 1, 0), idx(std::max(N - 1, 1), 0), fill(N, 0);
  for (int i = 0; i < N; ++i) ofs[i + 1] = ofs[i] + count[i];
  for (int nid = 1; nid < N; ++nid) idx[ofs[parent[nid]] + fill[parent[nid]]++] = nid;   // children in node-id order
  orbx_voc* v = new orbx_voc;
  v->ctx = ctx; v->k = k; v->L = L; v->scoring = sc; v->weighting = we; v->nNodes = N; v->nWords = nWords;
  cudaSetDevice(ctx->device);
  bool ok = cudaMalloc(&v->d_desc, desc.size()) == cudaSuccess;
  ok = ok && cudaMalloc(&v->d_childOfs, ofs.size() * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc(&v->d_childIdx, idx.size() * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc(&v->d_wordId, wordId.size() * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc(&v->d_weight, weight.size() * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMemcpy(v->d_desc, desc.data(), desc.size(), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaMemcpy(v->d_childOfs, ofs.data(), ofs.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaMemcpy(v->d_childIdx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaMemcpy(v->d_wordId, wordId.data(), wordId.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaMemcpy(v->d_weight, weight.data(), weight.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(0) == cudaSuccess;   // pageable H2D on the legacy stream: wait for the DMA (the context's stream is non-blocking)
  if (!ok) {
    orbx_set_error("orbx_vocabulary: device allocation/copy failed (%s)", cudaGetErrorString(cudaGetLastError()));
    voc_free(v);
    return nullptr;
  }
  return v;
}

orbx_voc* orbx_vocabulary_load(orbx_ctx* ctx, const char* path) {
  if (!ctx || !path) { orbx_set_error("orbx_vocabulary_load: bad arguments"); return nullptr; }
  FILE* f = fopen(path, "rb");
  if (!f) { orbx_set_error("orbx_vocabulary_load: cannot open %s", path); return nullptr; }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)std::max(n, 0L));
  const size_t got = fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (got != buf.size()) { orbx_set_error("orbx_vocabulary_load: short read on %s", path); return nullptr; }
  return orbx_vocabulary_from_memory(ctx, buf.data(), buf.size());
}

void orbx_vocabulary_destroy(orbx_voc* v) {
  if (v) cudaSetDevice(v->ctx->device);
  voc_free(v);
}

int orbx_vocabulary_info(const orbx_voc* v, int* k, int* L, int* n_nodes, int* n_words, int* scoring, int* weighting) {
  if (!v) return ORBX_EINVAL;
  if (k) *k = v->k;
  if (L) *L = v->L;
  if (n_nodes) *n_nodes = v->nNodes;
  if (n_words) *n_words = v->nWords;
  if (scoring) *scoring = v->scoring;
  if (weighting) *weighting = v->weighting;
  return ORBX_OK;
}

int orbx_vocabulary_transform(orbx_voc* v, const uint8_t* desc, int n, int levelsup, int32_t* bow_word, double* bow_value,
                              int32_t* n_bow, int32_t* fv_node, int32_t* fv_off, int32_t* fv_idx, int32_t* n_fv) {
  if (!v || n < 0 || (n > 0 && !desc) || !bow_word || !bow_value || !n_bow || !fv_node || !fv_off || !fv_idx || !n_fv)
    return ORBX_EINVAL;
  *n_bow = 0;
  *n_fv = 0;
  fv_off[0] = 0;
  if (n == 0 || v->nNodes <= 1) return ORBX_OK;     // empty(): both outputs stay empty (:1147-1150)
  ORBX_CUDA(cudaSetDevice(v->ctx->device));
  cudaStream_t st = v->ctx->stream;
  DevScope S(v->ctx, st);
  const uint8_t* d_desc = S.upload(desc, (size_t)n * 32);
  int* d_leaf = S.alloc<int>(n);
  int* d_node = S.alloc<int>(n);
  VocOut O;
  O.bowWord = S.alloc<int>(n); O.bowVal = S.alloc<double>(n); O.nBow = S.alloc<int>(1);
  O.fvNode = S.alloc<int>(n); O.fvOff = S.alloc<int>(n + 1); O.fvIdx = S.alloc<int>(n); O.nFv = S.alloc<int>(1);
  if (S.failed) return ORBX_ECUDA;
  int rc = voc_launch(v, st, 1, d_desc, nullptr, n, n, levelsup, d_leaf, d_node, O);
  if (rc != ORBX_OK) return rc;
  // one synchronisation: every output array has room for n entries, the counts say how many are meaningful
  S.download(bow_word, (const int32_t*)O.bowWord, (size_t)n);
  S.download(bow_value, (const double*)O.bowVal, (size_t)n);
  S.download(fv_node, (const int32_t*)O.fvNode, (size_t)n);
  S.download(fv_off, (const int32_t*)O.fvOff, (size_t)n + 1);
  S.download(fv_idx, (const int32_t*)O.fvIdx, (size_t)n);
  S.download(n_bow, (const int32_t*)O.nBow, (size_t)1);
  S.download(n_fv, (const int32_t*)O.nFv, (size_t)1);
  return S.finish();
}

int orbx_vocabulary_transform_batch_device(orbx_voc* v, int F, const uint8_t* d_desc, const int32_t* d_n, int cap, int levelsup,
                                           int32_t* d_leaf, int32_t* d_node, int32_t* d_bow_word, double* d_bow_value,
                                           int32_t* d_n_bow, int32_t* d_fv_node, int32_t* d_fv_off, int32_t* d_fv_idx,
                                           int32_t* d_n_fv) {
  if (!v || F <= 0 || cap <= 0 || !d_desc || !d_n || !d_leaf || !d_node || !d_bow_word || !d_bow_value || !d_n_bow || !d_fv_node ||
      !d_fv_off || !d_fv_idx || !d_n_fv)
    return ORBX_EINVAL;
  if (v->nNodes <= 1) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(v->ctx->device));
  VocOut O;
  O.bowWord = d_bow_word; O.bowVal = d_bow_value; O.nBow = d_n_bow;
  O.fvNode = d_fv_node; O.fvOff = d_fv_off; O.fvIdx = d_fv_idx; O.nFv = d_n_fv;
  return voc_launch(v, v->ctx->stream, F, d_desc, d_n, 0, cap, levelsup, d_leaf, d_node, O);
}

}  // extern "C"
