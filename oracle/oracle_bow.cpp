// ORACLE - TEST INFRASTRUCTURE ONLY (see oracle_common.h).  CPU restatement of
//   DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
//     Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1194 (+ per-feature descent :1217-1259)
//   FORB::distance  Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
//   BowVector::addWeight / addIfNotExist / normalize  Thirdparty/DBoW2/DBoW2/BowVector.cpp:25-84
//   FeatureVector::addFeature  Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:31-45
// as used by Frame::ComputeBoW (src/Frame.cc:1115-1122).  The reference ships no vocabulary file
// (ORBvoc.txt is a missing large blob) and no test vectors: parity for this row is pinned by an
// independent dictionary-based Python model on small trees (tests/test_bow_cpu.py).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "oracle_common.h"

namespace plvio {

static int forb_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t x, y;
    memcpy(&x, a + 4 * i, 4);
    memcpy(&y, b + 4 * i, 4);
    unsigned int v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

}  // namespace plvio

using namespace plvio;

extern "C" {

// Tree given as in loadFromTextFile: parent[i] for nodes 1..n-1 in id order.  Outputs sized n_feat:
// per-feature word / weight / node, BowVector (count, words, values), FeatureVector as CSR.
void plvio_bow_transform(int L, int scoring, int weighting, int n_nodes, const int* parent, const uint8_t* node_desc,
                         const double* node_weight, const uint8_t* feat, int n_feat, int levelsup, int* word_id,
                         double* word_weight, int* node_id, int* bow_count, int* bow_words, double* bow_values,
                         int* fv_count, int* fv_nodes, int* fv_start, int* fv_features) {
  std::vector<std::vector<int>> children(n_nodes);
  for (int i = 1; i < n_nodes; i++) children[parent[i]].push_back(i);
  std::vector<int> wid(n_nodes, -1);
  int nw = 0;
  for (int i = 1; i < n_nodes; i++)
    if (children[i].empty()) wid[i] = nw++;
  std::map<int, double> v;
  std::map<int, std::vector<unsigned>> fv;
  const int nid_level = L - levelsup;
  for (int f = 0; f < n_feat; f++) {
    const uint8_t* q = feat + 32 * (size_t)f;
    int nid = 0, final_id = 0, current_level = 0;
    do {
      ++current_level;
      const std::vector<int>& nodes = children[final_id];
      final_id = nodes[0];
      double best_d = forb_distance(q, node_desc + 32 * (size_t)final_id);
      for (size_t c = 1; c < nodes.size(); c++) {
        const int id = nodes[c];
        const double d = forb_distance(q, node_desc + 32 * (size_t)id);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (current_level == nid_level) nid = final_id;
    } while (!children[final_id].empty());
    const double w = node_weight[final_id];
    word_id[f] = wid[final_id];
    word_weight[f] = w;
    node_id[f] = nid;
    if (w > 0) {
      if (weighting <= 1) {          // TF_IDF, TF: addWeight
        auto it = v.lower_bound(wid[final_id]);
        if (it != v.end() && !(v.key_comp()(wid[final_id], it->first))) it->second += w;
        else v.insert(it, std::make_pair(wid[final_id], w));
      } else {                       // IDF, BINARY: addIfNotExist
        auto it = v.lower_bound(wid[final_id]);
        if (it == v.end() || v.key_comp()(wid[final_id], it->first)) v.insert(it, std::make_pair(wid[final_id], w));
      }
      fv[nid].push_back((unsigned)f);
    }
  }
  const int norm_mode = scoring == 1 ? 2 : (scoring == 5 ? 0 : 1);
  if (weighting <= 1 && !v.empty() && norm_mode == 0) {
    const double nd = (double)v.size();
    for (auto& kv : v) kv.second /= nd;
  }
  if (norm_mode != 0) {
    double norm = 0.0;
    if (norm_mode == 1) {
      for (auto& kv : v) norm += fabs(kv.second);
    } else {
      for (auto& kv : v) norm += kv.second * kv.second;
      norm = sqrt(norm);
    }
    if (norm > 0.0)
      for (auto& kv : v) kv.second /= norm;
  }
  int r = 0;
  for (auto& kv : v) { bow_words[r] = kv.first; bow_values[r] = kv.second; r++; }
  *bow_count = r;
  int nn = 0, pos = 0;
  for (auto& kv : fv) {
    fv_nodes[nn] = kv.first;
    fv_start[nn] = pos;
    for (unsigned i : kv.second) fv_features[pos++] = (int)i;
    nn++;
  }
  fv_start[nn] = pos;
  *fv_count = nn;
}

}  // extern "C"
