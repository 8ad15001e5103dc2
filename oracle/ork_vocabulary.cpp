// ork_vocabulary.cpp — ORACLE (test infrastructure): DBoW2 vocabulary tree and `transform`, restated without OpenCV.
//
// Follows (reference tree, vendored DBoW2):
//   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h  loadFromBinaryFile :1442-1478, transform(features, BowVector,
//       FeatureVector, levelsup) :1140-1219, transform(feature, word_id, weight, nid, levelsup) :1231-1271
//   Thirdparty/DBoW2/DBoW2/BowVector.cpp  addWeight :32-45, addIfNotExist :49-57, normalize :61-84
//   Thirdparty/DBoW2/DBoW2/FeatureVector.cpp  addFeature :31-45
//   Thirdparty/DBoW2/DBoW2/FORB.cpp  distance :81-101 (256-bit Hamming)
//   Thirdparty/DBoW2/DBoW2/ScoringObject.h  mustNormalize per scoring type
// Call sites: src/Frame.cc:865-872 (ComputeBoW), src/KeyFrame.cc:125-134, both with levelsup = 4.
//
// Loader quirk restated, not reproduced: `while(!f.eof())` reads one record past the end, which re-parses the stale
// buffer, i.e. appends a duplicate of the LAST node under the same parent.  The duplicate has the same descriptor as
// an earlier sibling and `d < best_d` is strict, so it can never be selected; it is left out here.
//
// PARITY UNPINNED by the reference (no tests; DBoW2 needs OpenCV to build).  Pinned to a numpy brute-force
// restatement (tests/test_oracle_vocabulary.py), also run on the reference's real Vocabulary/ORBvoc.bin when present.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#include "ork.h"

namespace {

struct Voc {
  int k = 0, L = 0, scoring = 0, weighting = 0;
  int nNodes = 0;                         // including the root (id 0)
  std::vector<int> parent;
  std::vector<uint8_t> desc;              // [nNodes][32]
  std::vector<double> weight;             // WordValue (double), assigned from the file's float
  std::vector<uint8_t> leaf;
  std::vector<int> wordId;                // leaves, numbered in file order
  std::vector<std::vector<int>> children; // in file (= node id) order
  int nWords = 0;
};

int hamming256(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return d;
}

Voc* parse(const uint8_t* p, size_t bytes) {
  if (bytes < 24) return nullptr;
  uint32_t nb, sz;
  int32_t k, L, sc, we;
  std::memcpy(&nb, p, 4); std::memcpy(&sz, p + 4, 4); std::memcpy(&k, p + 8, 4);
  std::memcpy(&L, p + 12, 4); std::memcpy(&sc, p + 16, 4); std::memcpy(&we, p + 20, 4);
  if (sz != 4 + 32 + 4 + 1 || nb < 1) return nullptr;
  const size_t nrec = (bytes - 24) / sz;   // the header counts the root too: nrec == nb - 1 for DBoW2-written files
  Voc* v = new Voc;
  v->k = k; v->L = L; v->scoring = sc; v->weighting = we;
  v->nNodes = (int)nrec + 1;
  v->parent.assign(v->nNodes, 0);
  v->desc.assign((size_t)v->nNodes * 32, 0);
  v->weight.assign(v->nNodes, 0.0);
  v->leaf.assign(v->nNodes, 0);
  v->wordId.assign(v->nNodes, -1);
  v->children.resize(v->nNodes);
  for (size_t r = 0; r < nrec; ++r) {
    const uint8_t* rec = p + 24 + r * sz;
    const int nid = (int)r + 1;
    int32_t par;
    float w;
    std::memcpy(&par, rec, 4);
    std::memcpy(&w, rec + 36, 4);
    if (par < 0 || par >= v->nNodes) { delete v; return nullptr; }
    v->parent[nid] = par;
    v->children[par].push_back(nid);
    std::memcpy(&v->desc[(size_t)nid * 32], rec + 4, 32);
    v->weight[nid] = (double)w;
    if (rec[40]) {
      v->leaf[nid] = 1;
      v->wordId[nid] = v->nWords++;
    }
  }
  return v;
}

// transform(feature, word_id, weight, nid, levelsup), :1231-1271
void transform_one(const Voc& V, const uint8_t* f, int levelsup, int* word, double* weight, int* nid) {
  const int nidLevel = V.L - levelsup;
  if (nidLevel <= 0) *nid = 0;
  int finalId = 0, level = 0;
  do {
    ++level;
    const std::vector<int>& nodes = V.children[finalId];
    finalId = nodes[0];
    double best = (double)hamming256(f, &V.desc[(size_t)finalId * 32]);
    for (size_t c = 1; c < nodes.size(); ++c) {
      const double d = (double)hamming256(f, &V.desc[(size_t)nodes[c] * 32]);
      if (d < best) { best = d; finalId = nodes[c]; }
    }
    if (level == nidLevel) *nid = finalId;
  } while (!V.children[finalId].empty());   // isLeaf(): children.empty()
  *word = V.wordId[finalId];
  *weight = V.weight[finalId];
}

}  // namespace

extern "C" {

void* ork_voc_from_memory(const uint8_t* data, size_t bytes) { return parse(data, bytes); }
void* ork_voc_load(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return nullptr;
  std::vector<uint8_t> buf;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf.resize((size_t)n);
  const size_t got = std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  return got == buf.size() ? parse(buf.data(), buf.size()) : nullptr;
}
void ork_voc_destroy(void* v) { delete (Voc*)v; }
int ork_voc_info(void* vp, int* k, int* L, int* nNodes, int* nWords, int* scoring, int* weighting) {
  Voc* v = (Voc*)vp;
  if (!v) return ORBX_EINVAL;
  if (k) *k = v->k;
  if (L) *L = v->L;
  if (nNodes) *nNodes = v->nNodes;
  if (nWords) *nWords = v->nWords;
  if (scoring) *scoring = v->scoring;
  if (weighting) *weighting = v->weighting;
  return ORBX_OK;
}

// transform(features, BowVector, FeatureVector, levelsup), :1140-1219.  Outputs:
//   word_id/node_id [n]: per-feature word and node-at-level (diagnostics; -1 word = stopped, weight 0)
//   bow_word/bow_value [<= n]: the BowVector in ascending word order; fv_* : the FeatureVector as CSR in ascending
//   node order (feature indices ascending inside a node).
int ork_voc_transform(void* vp, const uint8_t* desc, int n, int levelsup, int32_t* word_id, int32_t* node_id,
                      int32_t* bow_word, double* bow_value, int32_t* n_bow, int32_t* fv_node, int32_t* fv_off,
                      int32_t* fv_idx, int32_t* n_fv) {
  Voc* v = (Voc*)vp;
  if (!v || n < 0) return ORBX_EINVAL;
  std::map<int, double> bow;
  std::map<int, std::vector<int>> fv;
  // ScoringObject::mustNormalize: L1_NORM 0 -> L1, L2_NORM 1 -> L2, CHI_SQUARE 2 / KL 3 / BHATTACHARYYA 4 -> L1,
  // DOT_PRODUCT 5 -> none
  const bool must = v->scoring != 5;
  const bool l2 = v->scoring == 1;
  const bool tf = v->weighting == 0 /*TF_IDF*/ || v->weighting == 1 /*TF*/;
  if (v->nNodes > 1) {
    for (int i = 0; i < n; ++i) {
      int w = -1, nid = 0;
      double wt = 0;
      transform_one(*v, desc + 32 * (size_t)i, levelsup, &w, &wt, &nid);
      if (word_id) word_id[i] = wt > 0 ? w : -1;
      if (node_id) node_id[i] = nid;
      if (wt > 0) {   // not stopped
        if (tf) {     // addWeight
          auto it = bow.lower_bound(w);
          if (it != bow.end() && it->first == w) it->second += wt; else bow.insert(it, {w, wt});
        } else {      // addIfNotExist
          auto it = bow.lower_bound(w);
          if (it == bow.end() || it->first != w) bow.insert(it, {w, wt});
        }
        fv[nid].push_back(i);
      }
    }
    if (tf && !bow.empty() && !must) {
      const double nd = (double)bow.size();
      for (auto& kv : bow) kv.second /= nd;
    }
    if (must) {   // BowVector::normalize
      double norm = 0.0;
      if (!l2) { for (auto& kv : bow) norm += std::fabs(kv.second); }
      else { for (auto& kv : bow) norm += kv.second * kv.second; norm = std::sqrt(norm); }
      if (norm > 0.0) for (auto& kv : bow) kv.second /= norm;
    }
  } else if (word_id || node_id) {
    for (int i = 0; i < n; ++i) { if (word_id) word_id[i] = -1; if (node_id) node_id[i] = 0; }
  }
  int nb = 0;
  for (auto& kv : bow) { bow_word[nb] = kv.first; bow_value[nb] = kv.second; ++nb; }
  *n_bow = nb;
  int nn = 0, pos = 0;
  for (auto& kv : fv) {
    fv_node[nn] = kv.first;
    fv_off[nn] = pos;
    for (int i : kv.second) fv_idx[pos++] = i;
    ++nn;
  }
  fv_off[nn] = pos;
  *n_fv = nn;
  return ORBX_OK;
}

}  // extern "C"
