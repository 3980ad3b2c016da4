// emu_driver.cpp — TEST INFRASTRUCTURE ONLY (see include/cuda_emu.h): runs the kernel source of
// csrc/bigclam_sparse.cuh / csrc/bigclam_kernels.cuh on the host for tests/test_emu_kernels.py.
#include <algorithm>
#include <cstdio>
#include <numeric>

#ifdef BIGCLAM_EMU_SPARSE            // the transformed copies made by build.sh
#include "bigclam_sparse.cuh"        // (includes bigclam_kernels.cuh)
#else
#include "bigclam_kernels.cuh"
#endif

thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
namespace emu {
thread_local WarpCtx *warp = nullptr;
thread_local BlockCtx *block = nullptr;
thread_local int lane = 0;
thread_local int xpar = 0;
unsigned char *g_dyn_smem = nullptr;
}  // namespace emu

using namespace bigclam;

namespace {
struct Problem {
    int64_t n;
    int32_t k, ld, nsteps;
    std::vector<NodeMeta> meta;
    std::vector<double> F, sumF, partials;
    std::vector<int8_t> accepted;
    StepArgs a;
};

// mirrors fill_args / rebuild_order_list of csrc/bigclam_capi.cu
void setup(Problem &P, int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k, const double *F_in, const double *sumF,
           const uint8_t *mask, int do_linesearch, int max_inter, double alpha, double beta, unsigned *work_counter,
           unsigned init_positions) {
    P.n = n;
    P.k = k;
    P.ld = (k + 3) & ~3;
    P.nsteps = max_inter + 1;
    const int ld = P.ld;
    P.F.assign((size_t)n * ld, 0.0);
    for (int64_t u = 0; u < n; ++u) std::copy(F_in + u * k, F_in + (u + 1) * k, P.F.begin() + u * ld);
    P.sumF.assign(ld, 0.0);
    std::copy(sumF, sumF + k, P.sumF.begin());
    P.partials.assign(2 * ld + 2, 0.0);
    P.accepted.assign(n, -1);
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int32_t x, int32_t y) { return (rowptr[x + 1] - rowptr[x]) > (rowptr[y + 1] - rowptr[y]); });
    P.meta.resize(n);
    for (int64_t i = 0; i < n; ++i) P.meta[i] = NodeMeta{order[i], (int32_t)(rowptr[order[i] + 1] - rowptr[order[i]]), rowptr[order[i]]};
    StepArgs &a = P.a;
    std::memset(&a, 0, sizeof(a));
    a.n = n;
    a.rowptr = rowptr;
    a.col = col;
    a.sumF = P.sumF.data();
    a.k = k;
    a.ld = ld;
    a.nsteps = P.nsteps;
    double s = 1.0;
    a.steps[0] = s;
    for (int i = 1; i <= max_inter; ++i) { s *= beta; a.steps[i] = s; }
    a.alpha = alpha;
    a.min_p = 0.0001; a.max_p = 0.9999; a.min_f = 0.0; a.max_f = 1000.0;
    a.x_lo = -std::log(a.max_p) * (1.0 - 1e-12);
    a.x_hi = -std::log(a.min_p) * (1.0 + 1e-12);
    a.t_lo = std::log(1.0 - a.max_p);
    a.t_hi = std::log(1.0 - a.min_p);
    a.w_lo = 1.0 / (1.0 - a.max_p);
    a.w_hi = 1.0 / (1.0 - a.min_p);
    a.meta = P.meta.data();
    a.order_n = n;
    a.maxm = std::min<int32_t>(ld, kMaxActiveCap);
    *work_counter = init_positions;
    a.work_counter = work_counter;
    a.node_mask = mask;
    a.partials = P.partials.data();
    a.accepted = P.accepted.data();
    a.do_linesearch = do_linesearch;
}
}  // namespace

// One step of the dense kernels (step_kernel<C2, R, false, false>: no hub phase, no peer pushes), same outputs.
template <int C2>
static void run_dense(Problem &P, int grid) {
    constexpr int R = RowsInFlight<C2>::value;
    emu::launch(step_kernel<C2, R, false, false>, (unsigned)grid, (unsigned)kBlockThreads, block_smem_bytes(P.ld, P.a.maxm), P.a);
}

// hub phase (step_kernel<C2, R, true, false>, one block: the emulation runs blocks one after the other and the
// phases of a multi-block hub wait for each other)
template <int C2>
static void run_dense_hubs(Problem &P) {
    constexpr int R = RowsInFlight<C2>::value;
    emu::launch(step_kernel<C2, R, true, false>, 1u, (unsigned)kBlockThreads, block_smem_bytes(P.ld, P.a.maxm), P.a);
}

// hub_deg > 0 (grid 1, K <= 256): nodes of at least that degree go through the block-cooperative hub phase, those
// above kHubSlice edges as multi-phase "mega" hubs — the item list of rebuild_order_list (csrc/bigclam_capi.cu).
static int dense_step_impl(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k, const double *F_in,
                           const double *sumF, const uint8_t *mask, int do_linesearch, int max_inter, double alpha,
                           double beta, int grid, int hub_deg, double *F_out, double *partials_out, int8_t *accepted_out) {
    Problem P;
    unsigned work = 0;
    setup(P, n, rowptr, col, k, F_in, sumF, mask, do_linesearch, max_inter, alpha, beta, &work, 3u * (unsigned)grid * kWarpsPerBlock);
    const int ld = P.ld;
    std::vector<double> Fo(P.F);                 // a PRE-only launch leaves F_out alone
    P.a.F_in = P.F.data();
    P.a.F_out = Fo.data();
    const int c2raw = (ld / 2 + 31) / 32;
    const int c2 = c2raw <= 1 ? 1 : c2raw <= 2 ? 2 : c2raw <= 4 ? 4 : c2raw <= 8 ? 8 : 16;
    std::vector<HubItem> items;
    std::vector<double> scratch(ld + 32, 0.0);
    std::vector<unsigned int> counters(2, 0u);
    int nh = 0;
    if (hub_deg > 0 && grid == 1 && c2 <= 4 && P.nsteps <= 16) {
        while (nh < n && P.meta[nh].deg >= hub_deg) ++nh;
        std::vector<HubItem> i1, i0, i2, i3;
        int n_mega = 0;
        for (int i = 0; i < nh; ++i) {
            HubItem it{};
            it.hub = i;
            const int nsl = (P.meta[i].deg + kHubSlice - 1) / kHubSlice;
            if (nsl > 1) {
                it.mslot = n_mega++;
                it.nslices = nsl;
                for (int sl = 0; sl < nsl; ++sl) {
                    it.slice = sl;
                    it.phase = 1; i1.push_back(it);
                    it.phase = 2; i2.push_back(it);
                }
                it.slice = 0;
                it.phase = 3; i3.push_back(it);
            } else {
                it.phase = 0; it.nslices = 1; it.mslot = 0;
                i0.push_back(it);
            }
        }
        items.insert(items.end(), i1.begin(), i1.end());
        items.insert(items.end(), i0.begin(), i0.end());
        items.insert(items.end(), i2.begin(), i2.end());
        items.insert(items.end(), i3.begin(), i3.end());
        scratch.assign((size_t)std::max(1, n_mega) * (ld + 32), 0.0);
        counters.assign(2 * (size_t)std::max(1, n_mega), 0u);
        P.a.n_hubs = nh;
        P.a.n_hub_items = (int32_t)items.size();
        P.a.hub_items = items.data();
        P.a.hub_scratch = scratch.data();
        P.a.hub_counters = counters.data();
        work = (unsigned)nh + 3u * kWarpsPerBlock;
        switch (c2) {
            case 1: run_dense_hubs<1>(P); break;
            case 2: run_dense_hubs<2>(P); break;
            default: run_dense_hubs<4>(P); break;
        }
    } else
    switch (c2) {
        case 1: run_dense<1>(P, grid); break;
        case 2: run_dense<2>(P, grid); break;
        case 4: run_dense<4>(P, grid); break;
        case 8: run_dense<8>(P, grid); break;
        default: run_dense<16>(P, grid); break;
    }
    for (int64_t u = 0; u < n; ++u) std::copy(Fo.begin() + u * ld, Fo.begin() + u * ld + k, F_out + u * k);
    std::copy(P.partials.begin(), P.partials.end(), partials_out);
    std::copy(P.accepted.begin(), P.accepted.end(), accepted_out);
    return nh > 0 ? 1000 + nh : 0;
}

extern "C" int emu_dense_step(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k, const double *F_in,
                              const double *sumF, const uint8_t *mask, int do_linesearch, int max_inter, double alpha,
                              double beta, int grid, double *F_out, double *partials_out, int8_t *accepted_out) {
    return dense_step_impl(n, rowptr, col, k, F_in, sumF, mask, do_linesearch, max_inter, alpha, beta, grid, 0, F_out,
                           partials_out, accepted_out);
}

// returns 1000 + number of hub nodes on success
extern "C" int emu_dense_step_hubs(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k, const double *F_in,
                                   const double *sumF, const uint8_t *mask, int do_linesearch, int max_inter, double alpha,
                                   double beta, int hub_deg, double *F_out, double *partials_out, int8_t *accepted_out) {
    return dense_step_impl(n, rowptr, col, k, F_in, sumF, mask, do_linesearch, max_inter, alpha, beta, 1, hub_deg, F_out,
                           partials_out, accepted_out);
}

#ifdef BIGCLAM_EMU_SPARSE
// One step over sparse rows: dense F_in -> dense_to_sparse_kernel -> sparse_step_kernel -> sparse_to_dense_kernel.
// partials_out: [D(ld) | unused(ld) | llh | n_updated] as in the library.
// hub_deg > 0: nodes of at least that degree are split into kSpHubSeg-edge segments over the warps (grid must be
// 1 here: the emulation runs blocks one after the other and the hub phases wait for each other).
static int sparse_step_impl(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k, const double *F_in,
                            const double *sumF, const uint8_t *mask, int do_linesearch, int max_inter, double alpha,
                            double beta, int grid, int hub_deg, double *F_out, double *partials_out, int8_t *accepted_out,
                            int64_t *pool_words_out) {
    Problem P;
    unsigned work = 0;
    setup(P, n, rowptr, col, k, F_in, sumF, mask, do_linesearch, max_inter, alpha, beta, &work, 3u * (unsigned)grid * (unsigned)sp_warps_per_block((k + 3) & ~3));
    const int ld = P.ld;
    if (ld > 1024) return -1;
    const uint64_t cap8 = (uint64_t)n * sp_words((uint32_t)ld);
    std::vector<uint64_t> hdr0(n, 0), hdr1(n, 0);
    std::vector<double> pool0(cap8 + 8, 0.0), pool1(cap8 + 8, 0.0);
    unsigned long long top[2] = {0, 0};
    int32_t overflow = 0;
    emu::launch(dense_to_sparse_kernel, (unsigned)((n + 7) / 8), 256u, (size_t)0, (const double *)P.F.data(), n, ld, hdr0.data(),
                pool0.data(), &top[0], cap8, &overflow);
    SparseArgs sp;
    sp.hdr_in = hdr0.data();
    sp.pool_in = pool0.data();
    sp.hdr_out = hdr1.data();
    sp.pool_out = pool1.data();
    sp.pool_top = &top[1];
    sp.pool_cap8 = cap8;
    sp.overflow = &overflow;
    P.a.F_in = P.F.data();
    P.a.F_out = nullptr;
    sp.region_base8 = 0;
    sp.n_peers = 0;
    // split hubs: same item list as rebuild_order_list (csrc/bigclam_capi.cu)
    std::vector<HubItem> items;
    std::vector<double> scratch;
    std::vector<unsigned int> counters;
    int nh = 0;
    if (hub_deg > 0 && grid == 1 && P.nsteps <= 16) {
        while (nh < n && P.meta[nh].deg >= hub_deg) ++nh;
        std::vector<HubItem> i1, i2, i3;
        for (int i = 0; i < nh; ++i) {
            HubItem it{};
            it.hub = i;
            it.mslot = i;
            it.nslices = (P.meta[i].deg + kSpHubSeg - 1) / kSpHubSeg;
            for (int sl = 0; sl < it.nslices; ++sl) {
                it.slice = sl;
                it.phase = 1; i1.push_back(it);
                it.phase = 2; i2.push_back(it);
            }
            it.slice = 0;
            it.phase = 3; i3.push_back(it);
        }
        items.insert(items.end(), i1.begin(), i1.end());
        items.insert(items.end(), i2.begin(), i2.end());
        items.insert(items.end(), i3.begin(), i3.end());
    }
    scratch.assign((size_t)std::max(1, nh) * (ld + 32), 0.0);
    counters.assign(2 * (size_t)std::max(1, nh) + 1, 0u);
    P.a.n_hubs = nh;
    P.a.n_hub_items = (int32_t)items.size();
    P.a.hub_items = items.data();
    P.a.hub_scratch = scratch.data();
    P.a.hub_counters = counters.data();
    sp.hub_work = counters.data() + 2 * (size_t)std::max(1, nh);
    work = (unsigned)nh + 3u * (unsigned)grid * (unsigned)sp_warps_per_block((k + 3) & ~3);
    if (nh > 0) emu::launch(sparse_step_kernel<false, true>, (unsigned)grid, 32u * (unsigned)sp_warps_per_block(ld), sp_block_smem_bytes(ld, sp_warps_per_block(ld)), P.a, sp);
    else emu::launch(sparse_step_kernel<false, false>, (unsigned)grid, 32u * (unsigned)sp_warps_per_block(ld), sp_block_smem_bytes(ld, sp_warps_per_block(ld)), P.a, sp);
    std::vector<double> Fo((size_t)n * ld, 0.0);
    if (do_linesearch)
        emu::launch(sparse_to_dense_kernel, (unsigned)((n + 7) / 8), 256u, (size_t)0, (const uint64_t *)hdr1.data(),
                    (const double *)pool1.data(), n, ld, Fo.data());
    else
        Fo = P.F;
    for (int64_t u = 0; u < n; ++u) std::copy(Fo.begin() + u * ld, Fo.begin() + u * ld + k, F_out + u * k);
    std::copy(P.partials.begin(), P.partials.end(), partials_out);
    std::copy(P.accepted.begin(), P.accepted.end(), accepted_out);
    if (pool_words_out) *pool_words_out = (int64_t)top[1];
    return overflow ? -2 : (nh > 0 ? 1000 + nh : 0);
}

extern "C" int emu_sparse_step(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k, const double *F_in,
                               const double *sumF, const uint8_t *mask, int do_linesearch, int max_inter, double alpha,
                               double beta, int grid, double *F_out, double *partials_out, int8_t *accepted_out,
                               int64_t *pool_words_out) {
    return sparse_step_impl(n, rowptr, col, k, F_in, sumF, mask, do_linesearch, max_inter, alpha, beta, grid, 0, F_out,
                            partials_out, accepted_out, pool_words_out);
}

// returns 1000 + number of split hubs on success
extern "C" int emu_sparse_step_hubs(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k, const double *F_in,
                                    const double *sumF, const uint8_t *mask, int do_linesearch, int max_inter, double alpha,
                                    double beta, int hub_deg, double *F_out, double *partials_out, int8_t *accepted_out) {
    return sparse_step_impl(n, rowptr, col, k, F_in, sumF, mask, do_linesearch, max_inter, alpha, beta, 1, hub_deg, F_out,
                            partials_out, accepted_out, nullptr);
}

// Node-partitioned step over sparse rows, `world` ranks emulated one after the other: every rank owns the nodes
// order[r::world] of the degree-sorted list, allocates in its own region of the output pool and pushes its rows
// into all replicas (sparse_step_kernel<true>).  Returns every replica's dense view of the new F
// (F_out: world x n x k) and the summed partials.
extern "C" int emu_sparse_step_ranks(int64_t n, const int64_t *rowptr, const int32_t *col, int32_t k, const double *F_in,
                                     const double *sumF, int max_inter, double alpha, double beta, int world,
                                     double *F_out, double *partials_out, int8_t *accepted_out) {
    if (world < 1 || world > 8) return -1;
    Problem P0;
    unsigned work = 0;
    setup(P0, n, rowptr, col, k, F_in, sumF, nullptr, 1, max_inter, alpha, beta, &work, 3u * (unsigned)sp_warps_per_block((k + 3) & ~3));
    const int ld = P0.ld;
    if (ld > 1024) return -1;
    const uint64_t cap8 = (uint64_t)n * sp_words((uint32_t)ld);
    // one input replica is enough here (it is only read); every rank has its own output replica
    std::vector<uint64_t> hdr_in(n, 0);
    std::vector<double> pool_in(cap8 + 8, 0.0);
    unsigned long long top_in = 0;
    int32_t overflow = 0;
    emu::launch(dense_to_sparse_kernel, (unsigned)((n + 7) / 8), 256u, (size_t)0, (const double *)P0.F.data(), n, ld, hdr_in.data(),
                pool_in.data(), &top_in, cap8, &overflow);
    std::vector<std::vector<uint64_t>> hdr_out(world, std::vector<uint64_t>(n, ~0ull));
    std::vector<std::vector<double>> pool_out(world, std::vector<double>(cap8 + 8, -7.0));
    std::vector<double> partials(2 * ld + 2, 0.0);
    std::vector<int8_t> accepted(n, -5);
    uint64_t base = 0;
    for (int r = 0; r < world; ++r) {
        std::vector<NodeMeta> mine;
        for (int64_t i = r; i < n; i += world) mine.push_back(P0.meta[i]);        // degree-sorted list dealt round-robin
        Problem P = P0;
        P.a.meta = mine.data();
        P.a.order_n = (int64_t)mine.size();
        P.a.sumF = P.sumF.data();
        P.a.partials = P.partials.data();
        P.a.accepted = accepted.data();
        unsigned w = 3u * (unsigned)sp_warps_per_block((k + 3) & ~3);
        P.a.work_counter = &w;
        unsigned long long top = 0;
        SparseArgs sp;
        sp.hdr_in = hdr_in.data();
        sp.pool_in = pool_in.data();
        sp.hdr_out = hdr_out[r].data();
        sp.pool_out = pool_out[r].data();
        sp.pool_top = &top;
        sp.region_base8 = base;
        sp.pool_cap8 = (uint64_t)mine.size() * sp_words((uint32_t)ld);
        base += sp.pool_cap8;
        sp.overflow = &overflow;
        sp.n_peers = 0;
        for (int q = 0; q < world; ++q)
            if (q != r) {
                sp.peer_hdr[sp.n_peers] = hdr_out[q].data();
                sp.peer_pool[sp.n_peers] = pool_out[q].data();
                ++sp.n_peers;
            }
        sp.hub_work = nullptr;
        if (sp.n_peers > 0) emu::launch(sparse_step_kernel<true, false>, 1u, 32u * (unsigned)sp_warps_per_block(ld), sp_block_smem_bytes(ld, sp_warps_per_block(ld)), P.a, sp);
        else emu::launch(sparse_step_kernel<false, false>, 1u, 32u * (unsigned)sp_warps_per_block(ld), sp_block_smem_bytes(ld, sp_warps_per_block(ld)), P.a, sp);
        for (size_t i = 0; i < partials.size(); ++i) partials[i] += P.partials[i];          // the all-reduce
    }
    for (int r = 0; r < world; ++r) {
        std::vector<double> Fo((size_t)n * ld, 0.0);
        emu::launch(sparse_to_dense_kernel, (unsigned)((n + 7) / 8), 256u, (size_t)0, (const uint64_t *)hdr_out[r].data(),
                    (const double *)pool_out[r].data(), n, ld, Fo.data());
        for (int64_t u = 0; u < n; ++u) std::copy(Fo.begin() + u * ld, Fo.begin() + u * ld + k, F_out + ((size_t)r * n + u) * k);
    }
    std::copy(partials.begin(), partials.end(), partials_out);
    std::copy(accepted.begin(), accepted.end(), accepted_out);
    return overflow ? -2 : 0;
}

// Host packer of the sparse layout (bigclam_set_F_csr) -> sparse_to_dense_kernel -> dense, and unpack again.
extern "C" int64_t emu_pack_roundtrip(int64_t n, int32_t k, const int64_t *indptr, const int32_t *indices, const double *values,
                                      double *F_out, double *colsum_out, int64_t *indptr_out, int32_t *indices_out,
                                      double *values_out) {
    const int ld = (k + 3) & ~3;
    const uint64_t cap8 = (uint64_t)n * sp_words((uint32_t)ld);
    std::vector<uint64_t> hdr(n);
    std::vector<double> pool(cap8 + 8, -3.0), colsum(ld, 0.0);
    const int64_t used = sp_host_pack(n, k, ld, indptr, indices, values, hdr.data(), pool.data(), cap8, colsum.data());
    if (used < 0) return used;
    std::vector<double> Fo((size_t)n * ld, 0.0);
    emu::launch(sparse_to_dense_kernel, (unsigned)((n + 7) / 8), 256u, (size_t)0, (const uint64_t *)hdr.data(), (const double *)pool.data(),
                n, ld, Fo.data());
    for (int64_t u = 0; u < n; ++u) std::copy(Fo.begin() + u * ld, Fo.begin() + u * ld + k, F_out + u * k);
    std::copy(colsum.begin(), colsum.begin() + k, colsum_out);
    sp_host_unpack(n, hdr.data(), pool.data(), indptr_out, indices_out, values_out);
    return used;
}
#endif  // BIGCLAM_EMU_SPARSE
