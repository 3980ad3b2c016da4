// emu_driver.cpp — TEST INFRASTRUCTURE ONLY (see include/cuda_emu.h): runs the kernel source of
// csrc/bigclam_kernels.cuh (the dense kernels) on the host for tests/test_emu_kernels.py.
#include <algorithm>
#include <cstdio>
#include <numeric>

#include "bigclam_kernels.cuh"        // the transformed copy made by build.sh

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


