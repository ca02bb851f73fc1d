// TEST INFRASTRUCTURE — C entry point around the reference's vendored DBoW2 (Thirdparty/DBoW2, compiled in place against
// oracle/refshim): loads a vocabulary with the reference's own loadFromTextFile and runs the reference's own
// TemplatedVocabulary<FORB>::transform, exactly as Frame::ComputeBoW / KeyFrame::ComputeBoW do (src/Frame.cc:880-896).
// Built into oracle/_ref/libref_dbow2.so (git-ignored); used only by tests/test_oracle_reference_dbow2.py.
#include <cstdint>
#include <cstring>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> VocBase;
struct Voc : public VocBase {
  // the per-feature overload is protected (TemplatedVocabulary.h:371-373)
  void one(const cv::Mat& f, DBoW2::WordId& id, DBoW2::WordValue& w, DBoW2::NodeId* nid, int levelsup) const {
    VocBase::transform(f, id, w, nid, levelsup);
  }
};
}  // namespace

extern "C" {
// returns the number of BowVector entries, -1: load failure, -2: bow_cap too small.
// word_id / weight / node_id: per feature (the protected overload); bow_*: the BowVector (std::map order);
// fv_node[i]: node the FeatureVector files feature i under, -1: not filed; fv_pos[i]: its position in that node's list.
int ref_bow_transform(const char* vocab_txt, const uint8_t* feats, int n, int levelsup, int32_t* word_id, double* weight,
                      int32_t* node_id, int32_t* bow_word, double* bow_val, int bow_cap, int32_t* fv_node, int32_t* fv_pos,
                      int32_t* n_words) {
  Voc voc;
  if (!voc.loadFromTextFile(vocab_txt)) return -1;
  *n_words = (int)voc.size();
  std::vector<cv::Mat> f(n);
  for (int i = 0; i < n; i++) {
    f[i] = cv::Mat(1, 32, CV_8U);
    std::memcpy(f[i].data, feats + 32 * (size_t)i, 32);
  }
  for (int i = 0; i < n; i++) {
    DBoW2::WordId id = 0;
    DBoW2::WordValue w = 0;
    DBoW2::NodeId nid = 0;
    voc.one(f[i], id, w, &nid, levelsup);
    word_id[i] = (int)id;
    weight[i] = w;
    node_id[i] = (int)nid;
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  voc.transform(f, bv, fv, levelsup);
  if ((int)bv.size() > bow_cap) return -2;
  int k = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) {
    bow_word[k] = (int)it->first;
    bow_val[k] = it->second;
  }
  for (int i = 0; i < n; i++) fv_node[i] = fv_pos[i] = -1;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
    for (size_t j = 0; j < it->second.size(); j++) {
      fv_node[it->second[j]] = (int)it->first;
      fv_pos[it->second[j]] = (int)j;
    }
  return k;
}
}
