// Bag-of-words transform of ORB descriptors:
//   Frame::ComputeBoW (src/Frame.cc:1115-1122) ->
//   DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&,
//   levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1194, per-feature descent :1217-1259),
//   FORB::distance (FORB.cpp:81-101), BowVector::addWeight/addIfNotExist/normalize (BowVector.cpp),
//   FeatureVector::addFeature (FeatureVector.cpp:31-45).
// One CTA per frame: thread per feature walks the k-ary tree (first child with the strictly smallest
// Hamming distance wins), then the CTA sorts (word, feature) and (node, feature) keys in shared memory
// and emits the two std::maps as sorted arrays.  Sums follow the reference's order: a word's weight is
// accumulated feature by feature, the norm word by word in ascending word id.
#include <vector>

#include "plvi_internal.cuh"

struct plvi_vocab {
  int device = 0;
  int k = 0, L = 0, nNodes = 0, nWords = 0;
  int weighting = 0;   // DBoW2::WeightingType: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY
  int normMode = 1;    // 0 none, 1 L1, 2 L2 (ScoringObject::mustNormalize)
  uint4* dDesc = nullptr;        // [nNodes][2]
  int* dChildStart = nullptr;    // [nNodes + 1]
  int* dChildIds = nullptr;      // [nNodes - 1]
  int* dWordId = nullptr;        // [nNodes] (-1: inner node)
  double* dWeight = nullptr;     // [nNodes]
};

namespace plvi {

struct BowArgs {
  const uint4* desc; const int* childStart; const int* childIds; const int* wordId; const double* weight;
  int L, weighting, normMode, levelsup;
  const uint8_t* feat; const int* counts; int stride, p2;
  int* outWord; double* outWeight; int* outNode;
  int* bowCount; int* bowWords; double* bowValues;
  int* fvCount; int* fvNodes; int* fvStart; int* fvFeatures;
};

__device__ __forceinline__ int hamming256_regs(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ void bitonic_sort_u64(unsigned long long* key, int p2, int tid, int nt) {
  for (int k = 2; k <= p2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < p2; i += nt) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = key[i], b = key[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { key[i] = b; key[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// exclusive block scan of one int per thread (256 threads); returns the exclusive prefix, *total = sum
__device__ int block_scan_256(int v, int tid, int* wtmp, int* total) {
  const int lane = tid & 31, wid = tid >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wtmp[wid] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < wid; w++) base += wtmp[w];
  int tot = 0;
  for (int w = 0; w < 8; w++) tot += wtmp[w];
  *total = tot;
  __syncthreads();
  return base + incl - v;
}

__global__ void __launch_bounds__(256) k_bow_transform(const BowArgs a) {
  extern __shared__ __align__(16) unsigned long long skey[];   // [p2]
  __shared__ int wtmp[8];
  __shared__ double s_norm;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n = min(a.counts[f], a.stride);
  const size_t fo = (size_t)f * a.stride;
  const int nidLevel = a.L - a.levelsup;

  // ---- per-feature descent (TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup))
  for (int i = tid; i < n; i += 256) {
    const uint4* fp = reinterpret_cast<const uint4*>(a.feat + (fo + i) * 32);
    const uint4 q0 = fp[0], q1 = fp[1];
    int node = 0, level = 0, nid = 0;
    do {
      ++level;
      const int cs = a.childStart[node], ce = a.childStart[node + 1];
      int best = __ldg(a.childIds + cs);
      int bestD = hamming256_regs(q0, q1, __ldg(a.desc + 2 * (size_t)best), __ldg(a.desc + 2 * (size_t)best + 1));
      for (int c = cs + 1; c < ce; c++) {
        const int id = __ldg(a.childIds + c);
        const int d = hamming256_regs(q0, q1, __ldg(a.desc + 2 * (size_t)id), __ldg(a.desc + 2 * (size_t)id + 1));
        if (d < bestD) { bestD = d; best = id; }
      }
      node = best;
      if (level == nidLevel) nid = node;
    } while (a.childStart[node + 1] > a.childStart[node]);
    a.outWord[fo + i] = a.wordId[node];
    a.outWeight[fo + i] = a.weight[node];
    a.outNode[fo + i] = nid;
  }
  __syncthreads();

  // ---- BowVector: sort (word, feature); stopped words (weight <= 0) and padding go to the end
  for (int i = tid; i < a.p2; i += 256) {
    unsigned long long key = ~0ull;
    if (i < n && a.outWeight[fo + i] > 0) key = ((unsigned long long)(unsigned)a.outWord[fo + i] << 32) | (unsigned)i;
    skey[i] = key;
  }
  __syncthreads();
  bitonic_sort_u64(skey, a.p2, tid, 256);
  {
    // run starts -> rank among the unique words
    const int per = (a.p2 + 255) / 256;
    const int beg = tid * per, end = min(beg + per, a.p2);
    int cnt = 0;
    for (int p = beg; p < end; p++) {
      const unsigned long long kk = skey[p];
      if (kk != ~0ull && (p == 0 || (unsigned)(skey[p - 1] >> 32) != (unsigned)(kk >> 32))) cnt++;
    }
    int total;
    int rank = block_scan_256(cnt, tid, wtmp, &total);
    for (int p = beg; p < end; p++) {
      const unsigned long long kk = skey[p];
      if (kk == ~0ull) continue;
      const unsigned word = (unsigned)(kk >> 32);
      if (p != 0 && (unsigned)(skey[p - 1] >> 32) == word) continue;
      const double w = a.outWeight[fo + (unsigned)(kk & 0xffffffffu)];
      double s = w;
      if (a.weighting <= 1)   // TF_IDF / TF: addWeight accumulates once per feature; IDF / BINARY: addIfNotExist
        for (int q = p + 1; q < a.p2 && skey[q] != ~0ull && (unsigned)(skey[q] >> 32) == word; q++) s = __dadd_rn(s, w);
      a.bowWords[fo + rank] = (int)word;
      a.bowValues[fo + rank] = s;
      rank++;
    }
    if (tid == 0) a.bowCount[f] = total;
    __syncthreads();
    // normalisation (BowVector::normalize): the norm is summed in ascending word id
    if (tid == 0) {
      double norm = 0.0;
      if (a.normMode == 1) {
        for (int r = 0; r < total; r++) norm = __dadd_rn(norm, fabs(a.bowValues[fo + r]));
      } else if (a.normMode == 2) {
        for (int r = 0; r < total; r++) norm = __dadd_rn(norm, __dmul_rn(a.bowValues[fo + r], a.bowValues[fo + r]));
        norm = __dsqrt_rn(norm);
      } else if (a.weighting <= 1) {
        norm = (double)total;      // "unnecessary when normalizing": vit->second /= nd
      }
      s_norm = norm;
    }
    __syncthreads();
    const double norm = s_norm;
    if (norm > 0.0)
      for (int r = tid; r < total; r += 256) a.bowValues[fo + r] = __ddiv_rn(a.bowValues[fo + r], norm);
  }
  __syncthreads();

  // ---- FeatureVector: sort (node, feature) of the non-stopped features
  for (int i = tid; i < a.p2; i += 256) {
    unsigned long long key = ~0ull;
    if (i < n && a.outWeight[fo + i] > 0) key = ((unsigned long long)(unsigned)a.outNode[fo + i] << 32) | (unsigned)i;
    skey[i] = key;
  }
  __syncthreads();
  bitonic_sort_u64(skey, a.p2, tid, 256);
  {
    const int per = (a.p2 + 255) / 256;
    const int beg = tid * per, end = min(beg + per, a.p2);
    int cnt = 0, valid = 0;
    for (int p = beg; p < end; p++) {
      const unsigned long long kk = skey[p];
      if (kk == ~0ull) continue;
      valid++;
      a.fvFeatures[fo + p] = (int)(unsigned)(kk & 0xffffffffu);
      if (p == 0 || (unsigned)(skey[p - 1] >> 32) != (unsigned)(kk >> 32)) cnt++;
    }
    int total, totalValid;
    int rank = block_scan_256(cnt, tid, wtmp, &total);
    block_scan_256(valid, tid, wtmp, &totalValid);
    for (int p = beg; p < end; p++) {
      const unsigned long long kk = skey[p];
      if (kk == ~0ull) continue;
      if (p != 0 && (unsigned)(skey[p - 1] >> 32) == (unsigned)(kk >> 32)) continue;
      a.fvNodes[fo + rank] = (int)(unsigned)(kk >> 32);
      a.fvStart[(size_t)f * (a.stride + 1) + rank] = p;
      rank++;
    }
    if (tid == 0) {
      a.fvCount[f] = total;
      a.fvStart[(size_t)f * (a.stride + 1) + total] = totalValid;
    }
  }
}

}  // namespace plvi

using namespace plvi;

extern "C" {

int plvi_vocab_create(plvi_vocab** out, int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                      const uint8_t* is_leaf, const uint8_t* desc, const double* weight, int device) {
  if (!out || n_nodes < 2 || !parent || !is_leaf || !desc || !weight || k < 1 || L < 1 || scoring < 0 || scoring > 5 ||
      weighting < 0 || weighting > 3) {
    set_error("plvi_vocab_create: invalid argument");
    return PLVI_ERR_INVALID;
  }
  // children in id order (loadFromTextFile: m_nodes[pid].children.push_back(nid)), words = leaves in id order
  std::vector<int> start(n_nodes + 1, 0), ids(n_nodes - 1), word(n_nodes, -1), cur(n_nodes, 0);
  for (int i = 1; i < n_nodes; i++) {
    if (parent[i] < 0 || parent[i] >= n_nodes || parent[i] == i) { set_error("plvi_vocab_create: bad parent id"); return PLVI_ERR_INVALID; }
    start[parent[i] + 1]++;
  }
  for (int i = 0; i < n_nodes; i++) start[i + 1] += start[i];
  for (int i = 1; i < n_nodes; i++) ids[start[parent[i]] + cur[parent[i]]++] = i;
  int nw = 0;
  for (int i = 1; i < n_nodes; i++) {
    const bool leaf = start[i + 1] == start[i];
    if (leaf != (is_leaf[i] != 0)) { set_error("plvi_vocab_create: is_leaf does not match the tree"); return PLVI_ERR_INVALID; }
    if (leaf) word[i] = nw++;
  }
  if (start[1] == start[0] || nw == 0) { set_error("plvi_vocab_create: empty vocabulary"); return PLVI_ERR_EMPTY; }
  PLVI_CUDA_TRY(cudaSetDevice(device));
  plvi_vocab* v = new plvi_vocab();
  v->device = device; v->k = k; v->L = L; v->nNodes = n_nodes; v->nWords = nw; v->weighting = weighting;
  v->normMode = scoring == 1 ? 2 : (scoring == 5 ? 0 : 1);   // L2_NORM -> L2, DOT_PRODUCT -> none, else L1 (ScoringObject.h)
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes + 64); };
  A((void**)&v->dDesc, (size_t)n_nodes * 32);
  A((void**)&v->dChildStart, (size_t)(n_nodes + 1) * sizeof(int));
  A((void**)&v->dChildIds, (size_t)n_nodes * sizeof(int));
  A((void**)&v->dWordId, (size_t)n_nodes * sizeof(int));
  A((void**)&v->dWeight, (size_t)n_nodes * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(v->dDesc, desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->dChildStart, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->dChildIds, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->dWordId, word.data(), word.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->dWeight, weight, (size_t)n_nodes * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error(std::string("plvi_vocab_create: ") + cudaGetErrorString(e));
    plvi_vocab_destroy(v);
    return PLVI_ERR_CUDA;
  }
  *out = v;
  return PLVI_OK;
}

void plvi_vocab_destroy(plvi_vocab* v) {
  if (!v) return;
  cudaSetDevice(v->device);
  cudaFree(v->dDesc); cudaFree(v->dChildStart); cudaFree(v->dChildIds); cudaFree(v->dWordId); cudaFree(v->dWeight);
  delete v;
}

int plvi_vocab_words(const plvi_vocab* v) { return v ? v->nWords : PLVI_ERR_INVALID; }

int plvi_bow_transform(plvi_vocab* v, void* stream, const uint8_t* d_desc, const int* d_counts, int n_frames, int stride,
                       int levelsup, int* d_word_id, double* d_word_weight, int* d_node_id, int* d_bow_count,
                       int* d_bow_words, double* d_bow_values, int* d_fv_count, int* d_fv_nodes, int* d_fv_start,
                       int* d_fv_features) {
  if (!v || !d_desc || !d_counts || n_frames < 1 || stride < 1 || !d_word_id || !d_word_weight || !d_node_id || !d_bow_count ||
      !d_bow_words || !d_bow_values || !d_fv_count || !d_fv_nodes || !d_fv_start || !d_fv_features) {
    set_error("plvi_bow_transform: invalid argument");
    return PLVI_ERR_INVALID;
  }
  if (stride > 16384) { set_error("plvi_bow_transform: more than 16384 features per frame"); return PLVI_ERR_CAPACITY; }
  PLVI_CUDA_TRY(cudaSetDevice(v->device));
  BowArgs a;
  a.desc = v->dDesc; a.childStart = v->dChildStart; a.childIds = v->dChildIds; a.wordId = v->dWordId; a.weight = v->dWeight;
  a.L = v->L; a.weighting = v->weighting; a.normMode = v->normMode; a.levelsup = levelsup;
  a.feat = d_desc; a.counts = d_counts; a.stride = stride;
  int p2 = 256;
  while (p2 < stride) p2 <<= 1;
  a.p2 = p2;
  a.outWord = d_word_id; a.outWeight = d_word_weight; a.outNode = d_node_id;
  a.bowCount = d_bow_count; a.bowWords = d_bow_words; a.bowValues = d_bow_values;
  a.fvCount = d_fv_count; a.fvNodes = d_fv_nodes; a.fvStart = d_fv_start; a.fvFeatures = d_fv_features;
  const size_t smem = (size_t)p2 * sizeof(unsigned long long);
  if (smem > 48 * 1024) PLVI_CUDA_TRY(cudaFuncSetAttribute(k_bow_transform, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bow_transform<<<n_frames, 256, smem, (cudaStream_t)stream>>>(a);
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

}  // extern "C"
