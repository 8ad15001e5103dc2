// ref_dbow2_glue.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's UNMODIFIED DBoW2
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h + FORB.cpp + BowVector.cpp + FeatureVector.cpp + ScoringObject.cpp,
// compiled in place against the OpenCV stand-in) — the ORBVocabulary typedef of include/ORBVocabulary.h:
// loadFromBinaryFile / loadFromTextFile and transform(features, BowVector, FeatureVector, levelsup) as
// Frame::ComputeBoW calls it (src/Frame.cc:504-512).
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
#include <cstdint>
#include <cstring>

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocabulary;   // include/ORBVocabulary.h:31

extern "C" {

void* ref_voc_load(const char* path, int binary) {
  RefVocabulary* v = new RefVocabulary();
  const bool ok = binary ? v->loadFromBinaryFile(path) : v->loadFromTextFile(path);
  if (!ok) { delete v; return nullptr; }
  return v;
}
void ref_voc_destroy(void* v) { delete (RefVocabulary*)v; }

int ref_voc_info(void* v_, int* k, int* L, int* words, int* scoring, int* weighting) {
  RefVocabulary* v = (RefVocabulary*)v_;
  *k = v->getBranchingFactor(); *L = v->getDepthLevels(); *words = (int)v->size();
  *scoring = (int)v->getScoringType(); *weighting = (int)v->getWeightingType();
  return 0;
}

// desc: n x 32 bytes.  Outputs: BowVector as (word id, value) pairs in std::map order; FeatureVector as CSR
// (node id, offsets into feat[]).  Returns 0, or -1 when a capacity is too small (counts are still written).
int ref_voc_transform(void* v_, const uint8_t* desc, int n, int levelsup, uint32_t* bowId, double* bowVal, int bowCap,
                      int* nBow, uint32_t* fvNode, int* fvOfs, int fvCap, int* nFv, uint32_t* feat, int featCap) {
  RefVocabulary* v = (RefVocabulary*)v_;
  std::vector<cv::Mat> features(n);
  for (int i = 0; i < n; ++i) {
    features[i] = cv::Mat(1, 32, CV_8U);
    std::memcpy(features[i].ptr(), desc + (size_t)32 * i, 32);
  }
  DBoW2::BowVector bow;
  DBoW2::FeatureVector fv;
  v->transform(features, bow, fv, levelsup);
  *nBow = (int)bow.size();
  *nFv = (int)fv.size();
  size_t nfeat = 0;
  for (auto& kv : fv) nfeat += kv.second.size();
  if ((int)bow.size() > bowCap || (int)fv.size() > fvCap || (int)nfeat > featCap) return -1;
  int i = 0;
  for (auto& kv : bow) { bowId[i] = kv.first; bowVal[i] = kv.second; ++i; }
  i = 0;
  int o = 0;
  for (auto& kv : fv) {
    fvNode[i] = kv.first;
    fvOfs[i] = o;
    for (unsigned f : kv.second) feat[o++] = f;
    ++i;
  }
  fvOfs[i] = o;
  return 0;
}

double ref_voc_score(void* v_, const uint32_t* idA, const double* valA, int nA, const uint32_t* idB, const double* valB, int nB) {
  RefVocabulary* v = (RefVocabulary*)v_;
  DBoW2::BowVector a, b;
  for (int i = 0; i < nA; ++i) a.insert(std::make_pair(idA[i], valA[i]));
  for (int i = 0; i < nB; ++i) b.insert(std::make_pair(idB[i], valB[i]));
  return v->score(a, b);
}

// DBoW2::FORB::distance (Thirdparty/DBoW2/DBoW2/FORB.cpp:80-110), the same bit trick as ORBmatcher::DescriptorDistance
int ref_forb_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat A(1, 32, CV_8U, (void*)a), B(1, 32, CV_8U, (void*)b);
  return DBoW2::FORB::distance(A, B);
}

}  // extern "C"
