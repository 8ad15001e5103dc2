// Vocabulary_orbx.cc — drop-in replacements for Frame::ComputeBoW (src/Frame.cc:865-872) and KeyFrame::ComputeBoW
// (src/KeyFrame.cc:125-134), SURVEY.md §8 f1: DBoW2's transform(vCurrentDesc, mBowVec, mFeatVec, 4) runs on the device
// against a vocabulary that is uploaded once per process (orbx_shim::vocabulary() loads the same ORBvoc.bin the reference's
// System constructor loads, path taken from ORBX_VOCABULARY or the System's strVocFile).
#include "orbx_shim_config.h"

namespace orbx_shim { orbx_voc* vocabulary(); }

namespace {
void compute_bow(const cv::Mat& descriptors, int n, DBoW2::BowVector& bow, DBoW2::FeatureVector& fv) {
  std::vector<int32_t> w(n > 0 ? n : 1), fn(n > 0 ? n : 1), fo(n + 1), fi(n > 0 ? n : 1);
  std::vector<double> v(n > 0 ? n : 1);
  int32_t nb = 0, nn = 0;
  orbx_shim::check("orbx_vocabulary_transform",
                   orbx_vocabulary_transform(orbx_shim::vocabulary(), descriptors.data, n, 4, w.data(), v.data(), &nb, fn.data(),
                                             fo.data(), fi.data(), &nn));
  bow.clear();
  fv.clear();
  for (int i = 0; i < nb; ++i) bow.insert(bow.end(), std::make_pair((unsigned int)w[i], v[i]));           // ascending: O(1) each
  for (int k = 0; k < nn; ++k)
    fv.insert(fv.end(), std::make_pair((unsigned int)fn[k], std::vector<unsigned int>(fi.begin() + fo[k], fi.begin() + fo[k + 1])));
}
}  // namespace

namespace ORB_SLAM3 {

void Frame::ComputeBoW() {
  if (mBowVec.empty()) compute_bow(mDescriptors, mDescriptors.rows, mBowVec, mFeatVec);
}

void KeyFrame::ComputeBoW() {
  if (mBowVec.empty() || mFeatVec.empty()) compute_bow(mDescriptors, mDescriptors.rows, mBowVec, mFeatVec);
}

}  // namespace ORB_SLAM3
