// ORACLE (test infrastructure, NOT product code) — reference-backed kNN.
//
// Compiles the reference's OWN vendored nanoflann 1.7.1 header, in place from
// /root/reference/include/cilantro/3rd_party/nanoflann/nanoflann.hpp (never copied into this
// repo), behind a small C ABI so tests and the CPU baseline can run the exact kd-tree build and
// search code cilantro runs. Output: oracle/_ref/libcilantro_ref_knn.so (git-ignored, travels
// to the GPU box). Recipe: oracle/Makefile target `ref`.
//
// What is restated here (cilantro glue that needs Eigen and so cannot be compiled):
//   * the data adaptor over a packed 3xN float array   core/kd_tree.hpp:12-37
//   * KDTree ctor parameters: leaf 10, flags None, 1 build thread, eps 0, sorted
//                                                     core/kd_tree.hpp:162-170
//   * KNNSearchResultAdaptor (k slots, last slot pre-seeded with the squared radius, insertion
//     shifts only entries with stored value STRICTLY greater)   core/kd_tree.hpp:63-109
//   * the batched query loop `omp parallel for schedule(dynamic, 256)`
//                             correspondence_search/correspondence_search_kd_tree_utilities.hpp:26
#include <cstdint>
#include <cstddef>
#include <limits>
#include <vector>
#include <nanoflann.hpp>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

struct PackedXYZ {
  std::vector<float> xyz;
  size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(size_t idx, size_t dim) const { return xyz[3 * idx + dim]; }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const { return false; }
};

using Metric = nanoflann::L2_Adaptor<float, PackedXYZ, float, size_t>;
using Tree = nanoflann::KDTreeSingleIndexAdaptor<Metric, PackedXYZ, 3, size_t>;

struct RefTree {
  PackedXYZ data;
  Tree* tree;
};

// k best (value, index) pairs in ascending value order; see header comment.
class BoundedKBest {
public:
  using DistanceType = float;
  using IndexType = size_t;
  BoundedKBest(float* vals, size_t* inds, size_t k, float bound) : v_(vals), i_(inds), k_(k), n_(0) {
    v_[k_ - 1] = bound;
  }
  size_t size() const { return n_; }
  bool full() const { return n_ == k_; }
  float worstDist() const { return v_[k_ - 1]; }
  void sort() const {}
  bool addPoint(float dist, size_t index) {
    size_t pos = n_;
    while (pos > 0 && v_[pos - 1] > dist) {
      if (pos < k_) {
        v_[pos] = v_[pos - 1];
        i_[pos] = i_[pos - 1];
      }
      --pos;
    }
    if (pos < k_) {
      v_[pos] = dist;
      i_[pos] = index;
    }
    if (n_ < k_) ++n_;
    return true;
  }

private:
  float* v_;
  size_t* i_;
  size_t k_, n_;
};

}  // namespace

REF_API void* ref_tree_build(const float* xyz, size_t n, size_t max_leaf) {
  RefTree* t = new RefTree;
  t->data.xyz.assign(xyz, xyz + 3 * n);
  t->data.n = n;
  t->tree = new Tree(3, t->data,
                     nanoflann::KDTreeSingleIndexAdaptorParams(
                         max_leaf, nanoflann::KDTreeSingleIndexAdaptorFlags::None, 1));
  return t;
}

REF_API void ref_tree_free(void* h) {
  RefTree* t = (RefTree*)h;
  if (!t) return;
  delete t->tree;
  delete t;
}

// kNNInRadiusSearch for one query, any k (core/kd_tree.hpp:284-291). Returns the count found.
REF_API size_t ref_knn_in_radius(void* h, const float* q, size_t k, float r2, uint64_t* idx, float* d2) {
  RefTree* t = (RefTree*)h;
  if (t->data.n == 0 || k == 0) return 0;
  std::vector<float> v(k);
  std::vector<size_t> ix(k);
  BoundedKBest rs(v.data(), ix.data(), k, r2);
  t->tree->findNeighbors(rs, q, nanoflann::SearchParameters(0.0f, true));
  for (size_t i = 0; i < rs.size(); i++) {
    idx[i] = ix[i];
    d2[i] = v[i];
  }
  return rs.size();
}

// Batched radius-bounded 1-NN; signature matches orc_knn_fn in cilantro_oracle.cpp.
REF_API void ref_knn1_radius_cb(void* h, const float* qry, size_t nq, float max_d2, int64_t* idx, float* d2) {
  RefTree* t = (RefTree*)h;
  if (t->data.n == 0) {
    for (size_t i = 0; i < nq; i++) {
      idx[i] = -1;
      d2[i] = max_d2;
    }
    return;
  }
  const nanoflann::SearchParameters sp(0.0f, true);
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < nq; i++) {
    float v;
    size_t ix;
    BoundedKBest rs(&v, &ix, 1, max_d2);
    t->tree->findNeighbors(rs, qry + 3 * i, sp);
    if (rs.size() == 1) {
      idx[i] = (int64_t)ix;
      d2[i] = v;
    } else {
      idx[i] = -1;
      d2[i] = max_d2;
    }
  }
}

// nearestNeighborSearch (unbounded, nanoflann's own KNNResultSet) — core/kd_tree.hpp:181-204.
REF_API void ref_nn1(void* h, const float* qry, size_t nq, int64_t* idx, float* d2) {
  RefTree* t = (RefTree*)h;
#pragma omp parallel for
  for (size_t i = 0; i < nq; i++) {
    size_t ix = 0;
    float v = std::numeric_limits<float>::max();
    size_t found = t->tree->knnSearch(qry + 3 * i, 1, &ix, &v);
    idx[i] = found ? (int64_t)ix : -1;
    d2[i] = v;
  }
}

// Batched kNNInRadiusSearch / kNNSearch (core/kd_tree.hpp:233-240, 302-309): row i of idx/d2 holds
// the neighbours of query i in ascending distance, cnt[i] of them (unused slots: -1 / r2).
REF_API void ref_knn_in_radius_batch(void* h, const float* qry, size_t nq, size_t k, float r2, int64_t* idx,
                                     float* d2, uint32_t* cnt) {
  RefTree* t = (RefTree*)h;
  const nanoflann::SearchParameters sp(0.0f, true);
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < nq; i++) {
    std::vector<float> v(k);
    std::vector<size_t> ix(k);
    size_t m = 0;
    if (t->data.n > 0 && k > 0) {
      BoundedKBest rs(v.data(), ix.data(), k, r2);
      t->tree->findNeighbors(rs, qry + 3 * i, sp);
      m = rs.size();
    }
    for (size_t j = 0; j < k; j++) {
      idx[i * k + j] = j < m ? (int64_t)ix[j] : -1;
      d2[i * k + j] = j < m ? v[j] : r2;
    }
    cnt[i] = (uint32_t)m;
  }
}

// Batched radiusSearch (core/kd_tree.hpp:251-272) through nanoflann's own radiusSearch, which
// collects every point with d2 < r2 and sorts by distance like RadiusSearchResultAdaptor::sort
// (kd_tree.hpp:130-133). Rows are truncated to `stride` entries; cnt[i] is the full count.
REF_API void ref_radius_batch(void* h, const float* qry, size_t nq, float r2, size_t stride, int64_t* idx,
                              float* d2, uint32_t* cnt) {
  RefTree* t = (RefTree*)h;
  const nanoflann::SearchParameters sp(0.0f, true);
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < nq; i++) {
    std::vector<nanoflann::ResultItem<size_t, float>> out;
    size_t m = 0;
    if (t->data.n > 0) m = t->tree->radiusSearch(qry + 3 * i, r2, out, sp);
    for (size_t j = 0; j < stride; j++) {
      idx[i * stride + j] = j < m ? (int64_t)out[j].first : -1;
      d2[i * stride + j] = j < m ? out[j].second : r2;
    }
    cnt[i] = (uint32_t)m;
  }
}

// FIRST_TO_SECOND search of the correspondence engine (correspondence_search_kd_tree.hpp:195-204): the reference
// builds a fresh kd-tree over the TRANSFORMED source points on every call and queries it with the destination
// points. Same signature as orc_knn1_brute (cilantro_oracle.cpp) so that the oracle can use either.
REF_API void ref_knn1_build_query(const float* ref_pts, size_t nref, const float* qry, size_t nq, float max_d2,
                                  int64_t* idx, float* d2) {
  void* t = ref_tree_build(ref_pts, nref, 10);
  ref_knn1_radius_cb(t, qry, nq, max_d2, idx, d2);
  ref_tree_free(t);
}

REF_API unsigned ref_nanoflann_version() { return NANOFLANN_VERSION; }
