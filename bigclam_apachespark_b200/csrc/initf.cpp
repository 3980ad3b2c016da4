// initf.cpp — the caller in front of the hot path: conductance seeding and the initial F
// (SURVEY.md §8f-2).  One-off integer graph work, done on the host like the reference does it on
// the Spark driver/executors; the result is handed to bigclam_set_F.
//
//   conductanceLocalMin()   codes/bigclam4-7.scala:58-73   -> bigclam_conductance_seeds
//   initNeighborComF(K)     codes/bigclam4-7.scala:81-108  -> bigclam_init_neighbor_com_F
//
// Reproduced AS CODED, including the quirks:
//   * ego net of x = [x] ++ neighbours(x); z = concatenation of the neighbour lists of all ego members;
//     cut_S = #{i in z : i not in ego}; vol_S = |z| - cut_S; vol_T = sigmaDegrees - vol_S - 2 cut_S with
//     sigmaDegrees = sum of in+out degrees = number of neighbour-list entries (:60,:64-66);
//     conductance = 0 if vol_S == 0, 1 if vol_T == 0, else cut_S / min(vol_S, vol_T) (:67).
//   * the candidate of node x is `y.map(v => (v, cond(v))).min` (:70): tuples order by the neighbour id
//     first, so it is the MIN-ID neighbour of x with that neighbour's conductance (not the neighbour of
//     minimal conductance); a node without neighbours contributes (x, 10.0).
//   * candidates are de-duplicated by id (reduceByKey) and sorted by conductance ascending (:70).  Spark
//     leaves ties and the zipWithIndex order (:85-86) partition dependent; here ties break by node id and
//     community c is the c-th selected seed in ascending id order (documented deviation: deterministic).
//   * F column c = 1.0 on the neighbours of seed c, the seed itself excluded (:85-86; Bigclamv2.scala:70
//     includes it: include_self); fewer seeds than K => the remaining columns are random 0/1 (:90-102,
//     unseeded scala.util.Random in the reference; a seeded xorshift64* here).
#include "../../include/bigclam_b200.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

// Candidates (:70): (min-id neighbour, its conductance) per node, or (x, 10.0) without neighbours; de-duplicated by
// id (reduceByKey keeps the minimum), sorted by conductance ascending, ties by id.  Shared by the host and the GPU
// path (csrc/initf_gpu.cu).  A conductance can be negative as coded (vol_T < 0 on multigraph input): candidates
// are tracked by a flag, not by the sign of the key.
void bigclam_select_seeds_internal(int64_t n, const int64_t *rowptr, const int32_t *col, const double *cond,
                                   int32_t *seeds_out, int64_t *n_seeds_out) {
    std::vector<double> key((size_t)n, 0.0);
    std::vector<char> is_cand((size_t)n, 0);
    auto offer = [&](int64_t v, double c) {
        if (!is_cand[(size_t)v] || c < key[(size_t)v]) key[(size_t)v] = c;
        is_cand[(size_t)v] = 1;
    };
    for (int64_t x = 0; x < n; ++x) {
        if (rowptr[x + 1] > rowptr[x]) {
            int32_t v = col[rowptr[x]];
            for (int64_t e = rowptr[x] + 1; e < rowptr[x + 1]; ++e) v = std::min(v, col[e]);
            offer(v, cond[(size_t)v]);
        } else {
            offer(x, 10.0);
        }
    }
    std::vector<int32_t> cand;
    for (int64_t v = 0; v < n; ++v)
        if (is_cand[(size_t)v]) cand.push_back((int32_t)v);
    std::stable_sort(cand.begin(), cand.end(), [&](int32_t a, int32_t b) { return key[(size_t)a] < key[(size_t)b]; });
    std::memcpy(seeds_out, cand.data(), sizeof(int32_t) * cand.size());
    *n_seeds_out = (int64_t)cand.size();
}

extern "C" int bigclam_conductance_seeds(int64_t n, const int64_t *rowptr, const int32_t *col,
                                         double *conductance_out /* n, optional */,
                                         int32_t *seeds_out /* n */, int64_t *n_seeds_out) {
    if (n <= 0 || rowptr == nullptr || seeds_out == nullptr || n_seeds_out == nullptr) return BIGCLAM_EINVAL;
    if (rowptr[n] > 0 && col == nullptr) return BIGCLAM_EINVAL;
    const double sigma = (double)rowptr[n];          // sum of in+out degrees == neighbour-list entries (:60)
    std::vector<double> cond((size_t)n);
    std::vector<int32_t> ego;
    for (int64_t x = 0; x < n; ++x) {
        ego.assign(col + rowptr[x], col + rowptr[x + 1]);
        ego.push_back((int32_t)x);
        std::sort(ego.begin(), ego.end());
        int64_t zsize = 0, cut = 0;
        // z = y.flatMap(u => Neightborbc.value(u)) over y = [x] ++ neighbours (with multiplicity) (:63)
        auto scan = [&](int32_t u) {
            for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {
                ++zsize;
                if (!std::binary_search(ego.begin(), ego.end(), col[e])) ++cut;
            }
        };
        scan((int32_t)x);
        for (int64_t e = rowptr[x]; e < rowptr[x + 1]; ++e) scan(col[e]);
        const double cut_S = (double)cut, vol_S = (double)(zsize - cut), vol_T = sigma - vol_S - cut_S * 2;
        cond[(size_t)x] = (vol_S == 0) ? 0.0 : (vol_T == 0) ? 1.0 : cut_S / std::min(vol_S, vol_T);   // :67
    }
    if (conductance_out != nullptr) std::memcpy(conductance_out, cond.data(), sizeof(double) * (size_t)n);

    bigclam_select_seeds_internal(n, rowptr, col, cond.data(), seeds_out, n_seeds_out);
    return BIGCLAM_OK;
}

static inline uint64_t xorshift64s(uint64_t &s) {
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return s * 2685821657736338717ULL;
}

extern "C" int bigclam_init_neighbor_com_F(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k,
                                           const int32_t *ranked_seeds, int64_t n_ranked, int32_t include_self,
                                           uint64_t pad_seed, double *F_out /* n x k, row-major */) {
    if (n <= 0 || k <= 0 || rowptr == nullptr || F_out == nullptr || (n_ranked > 0 && ranked_seeds == nullptr))
        return BIGCLAM_EINVAL;
    std::memset(F_out, 0, sizeof(double) * (size_t)n * (size_t)k);
    // S = Sbc.value.take(K) (:83); community index = position among the selected seeds in id order (:85-86)
    std::vector<int32_t> S(ranked_seeds, ranked_seeds + std::min<int64_t>(k, n_ranked));
    std::sort(S.begin(), S.end());
    for (size_t c = 0; c < S.size(); ++c) {
        const int32_t s = S[c];
        if (s < 0 || s >= n) return BIGCLAM_EINVAL;
        for (int64_t e = rowptr[s]; e < rowptr[s + 1]; ++e) F_out[(size_t)col[e] * k + c] = 1.0;
        if (include_self) F_out[(size_t)s * k + c] = 1.0;                 // Bigclamv2.scala:70
    }
    // padding columns: Array.fill(n)(Random.nextInt(2).toDouble) (:77-79, :90-102)
    uint64_t st = pad_seed ? pad_seed : 0x9E3779B97F4A7C15ULL;
    for (int32_t c = (int32_t)S.size(); c < k; ++c)
        for (int64_t u = 0; u < n; ++u) F_out[(size_t)u * k + c] = (double)((xorshift64s(st) >> 33) & 1ULL);
    return BIGCLAM_OK;
}
