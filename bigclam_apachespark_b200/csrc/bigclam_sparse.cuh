// bigclam_sparse.cuh — the same step (codes/bigclam4-7.scala:152-223) over SPARSE rows of F.
//
// Why: the reference keeps F as Breeze sparse vectors (`BSV[Double]`, bigclam4-7.scala:97-104) because the rows
// ARE sparse: on the bench workload (com-amazon, K = 200) a row holds ~9 non-zeros of 200 through the whole
// run and ~17 components are "active" in a line search.  The dense kernel (bigclam_kernels.cuh) moves and
// multiplies 95 % zeros.  Here a row is (count, ascending component indices, values) in a per-step pool, every
// neighbour row is read once per step (~120 bytes instead of 1.6 KB), and all per-edge work is proportional
// to the row's non-zeros.
//
// Layout (per F buffer; double-buffered like the dense F):
//   hdr[u]      uint64: (offset in 8-byte words << 24) | count
//   pool        row block at `offset`: pad4(count) doubles, then pad4(count) uint16 indices (ascending)
//   pool_top    bump allocator of the OUTPUT pool (words), zeroed before every step; a warp takes its new
//               row's block with one atomicAdd.  The input pool is only read (Jacobi), so a step whose pool
//               overflowed can simply be repeated with a larger pool.
//
// Per node (one warp), with fu scattered into a dense shared-memory vector fu_d[ld]:
//   PRE   the entries of up to 32 neighbour rows are staged in shared memory; lane e walks row e:
//         x_e = sum_i val_i * fu_d[idx_i] (no reduction: the dot lands in the lane that evaluates exp/log for
//         that edge, 32 edges per call); the gradient sum g_d[idx] += w_e * val goes neighbour by neighbour
//         (indices are unique inside a row: no conflicts, fixed order);
//   scan  one pass over the ld components turns g_d into the gradient (:168), sums |g|^2 and lists the
//         active components (fu > 0 or g > 0);
//   LS    lane (trial j, edge parity h) walks the staged entries of its edges:
//         D = sum_i clamp(fu_d[idx_i] + s_j * g_d[idx_i]) * val_i — an inactive component clamps to 0 and adds
//         exactly nothing, so no pair list / intersection is needed; two edges per lane in flight;
//   SWAP  the accepted candidate's non-zeros are compacted (ascending) into the staging buffer, a block is
//         taken from the output pool, the row and its header are written.
//
// Limits of this version: ld <= 256 (K <= 256), MIN_F_ == 0.
#pragma once
#include "bigclam_kernels.cuh"

namespace bigclam {

// Build-time knobs for A/B runs (tools/build_variant.sh <name> -DBIGCLAM_SP_BLOCKS=2 -DBIGCLAM_SP_PREFETCH=1):
//   BIGCLAM_SP_BLOCKS    blocks per SM the kernel is compiled for: 2 (default) = 16 warps/SM, ~120 registers, no
//                        spills; 3 = 24 warps/SM at 80 registers with 72-112 bytes of spills (the dense kernel's
//                        128-register builds lost 40 % to far fewer spilled bytes: to be measured, not assumed);
//   BIGCLAM_SP_PREFETCH  load the next node's header, neighbour ids, neighbour headers and own entries one node
//                        ahead (two dependent round trips less per node, ~16 more live registers).
#ifndef BIGCLAM_SP_BLOCKS
#define BIGCLAM_SP_BLOCKS 2
#endif
#ifndef BIGCLAM_SP_PREFETCH
#define BIGCLAM_SP_PREFETCH 0
#endif
constexpr int kSpBlocksPerSM = BIGCLAM_SP_BLOCKS;
constexpr bool kSpPrefetch = BIGCLAM_SP_PREFETCH != 0;
constexpr int kSpWarps = 8;            // warps per block at most (ld <= 256); wide rows run fewer (sp_warps_per_block)
constexpr int kSpThreads = kSpWarps * 32;
// staged neighbour entries per chunk: at least one full row always fits
__host__ __device__ inline int sp_entries(int ld) { return ld > 512 ? ld : 512; }

__host__ __device__ inline uint64_t sp_pack(uint64_t off8, uint32_t cnt) { return (off8 << 24) | (uint64_t)cnt; }
__host__ __device__ inline uint32_t sp_cnt(uint64_t h) { return (uint32_t)(h & 0xffffffull); }
__host__ __device__ inline uint64_t sp_off8(uint64_t h) { return h >> 24; }
__host__ __device__ inline uint32_t sp_pad(uint32_t cnt) { return (cnt + 3u) & ~3u; }
__host__ __device__ inline uint64_t sp_words(uint32_t cnt) { return (uint64_t)sp_pad(cnt) * 5u / 4u; }   // 8-byte words of a row block

struct SparseArgs {
    const uint64_t *hdr_in;
    const double *pool_in;
    uint64_t *hdr_out;
    double *pool_out;
    unsigned long long *pool_top;      // words used of this rank's region of pool_out
    uint64_t pool_cap8;                // capacity of that region in words
    uint64_t region_base8;             // where the region starts in pool_out (0 on a single GPU)
    int32_t *overflow;                 // set when a row did not fit (the step must be repeated with a larger pool)
    // node-partitioned multi-GPU: the owners' new rows go to the same offsets of every replica's output pool
    // (plain stores to IPC-mapped peer memory over NVLink); each rank allocates only inside its own region, so
    // all replicas end up with the same layout and no remote atomics are needed.  Every owned row is written
    // (and pushed) every step: the output pool is rebuilt from scratch each step.
    int32_t n_peers;
    uint64_t *peer_hdr[7];
    double *peer_pool[7];
    unsigned int *hub_work;            // next hub item to hand out (zeroed per launch); items: StepArgs::hub_items
};

// The dense per-warp / per-block vectors are padded to a multiple of 32 components (zeros: a padding component has
// fu = sumF = 0, hence gradient 0, never active), so that the loops over components need no bounds checks.
__host__ __device__ inline int sp_ldp(int ld) { return (ld + 31) & ~31; }
// per-warp shared memory: fu_d[ldp] | g_d[ldp] | ent_val[E] | ent_idx[E] u16 | aidx[max(ld, 256)] u16 | poff[40] u16 |
//                         cbal[32] u32 | ccum[32] u16 (+ pad)   (ballots / running counts of the entry compaction)
__host__ __device__ inline size_t sp_warp_bytes(int ld) {
    return sizeof(double) * 2 * (size_t)sp_ldp(ld) + (size_t)sp_entries(ld) * 10 + 2 * (size_t)(ld > 256 ? ld : 256) + 2 * 40 + 4 * 32 + 2 * 32;
}
// block: steps[kMaxSteps] | sumF[ldp] | D[ldp] | wpb x warp area
__host__ __device__ inline size_t sp_block_smem_bytes(int ld, int wpb) {
    return sizeof(double) * (kMaxSteps + 2 * (size_t)sp_ldp(ld)) + (size_t)wpb * sp_warp_bytes(ld);
}
// warps per block: as many resident warps per SM as the shared memory (227 KB, 1 KB reserved per block) allows
inline int sp_warps_per_block(int ld) {
    int best = 1, best_warps = 0;
    for (int wpb = kSpWarps; wpb >= 1; wpb >>= 1) {
        const size_t bytes = sp_block_smem_bytes(ld, wpb) + 1024 + 256;
        const int blocks = (int)((size_t)233472 / bytes);
        const int warps = (blocks > kSpBlocksPerSM ? kSpBlocksPerSM : blocks) * wpb;         // the kernel is built for that many blocks per SM
        if (warps > best_warps) { best_warps = warps; best = wpb; }
    }
    return best;
}

// Stages the rows of up to 32 neighbours (ids colp[0 .. cnt32)) of one node into the warp's entry buffer: the
// longest prefix of them whose entries fit the `cap` entries of the buffer (at least one: a row has at most ld <= cap).
// Returns the number ne of staged neighbours; poff[e] .. poff[e + 1] is row e's range in the buffer.
// (noinline, scalar arguments only: one copy in the code, called from PRE and from the line search.)
__device__ __noinline__ int sp_stage_chunk(const uint64_t *__restrict__ hdr_in, const double *__restrict__ pool_in,
                                           const int32_t *__restrict__ colp, int cnt32, int lane, int cap, double *ent_val,
                                           unsigned short *ent_idx, unsigned short *poff, int use_pre = 0,
                                           unsigned long long pre_hv = 0ull) {
    // use_pre: the caller already holds this chunk's row headers (loaded one node ahead)
    uint64_t hv = pre_hv;
    if (!use_pre) {
        const int v = (lane < cnt32) ? colp[lane] : 0;
        hv = (lane < cnt32) ? __ldg(hdr_in + v) : 0ull;
    }
    const int cv = (int)sp_cnt(hv);
    int incl = cv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned fit = __ballot_sync(0xffffffffu, (lane < cnt32) && (incl <= cap));
    const int ne = __popc(fit);                     // incl is monotone: the fitting lanes are 0 .. ne-1
    if (lane == 0) poff[0] = 0;
    if (lane < ne) poff[lane + 1] = (unsigned short)incl;
    __syncwarp();
    // the T entries of the ne rows, 32 at a time, one per lane: all loads of a round are in flight together
    // (the rows are ~9 entries long: going row by row would serialise a round trip per row)
    const int T = (ne > 0) ? __shfl_sync(0xffffffffu, incl, ne - 1) : 0;
    for (int base = 0; base < T; base += 32) {
        const int j = base + lane;
        int lo = 0, hi = ne;                                   // row of entry j: the largest e with poff[e] <= j
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int mid = (lo + hi) >> 1;
            const bool le = (int)poff[mid] <= j;
            if (hi - lo > 1) { if (le) lo = mid; else hi = mid; }
        }
        const uint64_t he = __shfl_sync(0xffffffffu, hv, lo);
        if (j < T) {
            const int ce = (int)sp_cnt(he);
            const double *vals = pool_in + sp_off8(he);
            const unsigned short *idxp = reinterpret_cast<const unsigned short *>(vals + sp_pad((uint32_t)ce));
            const int i = j - (int)poff[lo];
            ent_val[j] = __ldg(vals + i);
            ent_idx[j] = __ldg(idxp + i);
        }
    }
    __syncwarp();
    return ne;
}

// Hubs.  A node whose neighbour list is long enough to dominate a launch when one warp walks it (host: a sizeable
// fraction of a warp's share of the launch) is split into segments of kSpHubSeg edges that different warps work
// on; the pieces meet in a global scratch row per hub (same layout as the dense kernels' mega hubs:
// G[ld] | S1 | ST[16], stride ld + 32 doubles, two counters per hub):
//   phase 1  PRE of one segment: its share of sum_v w_v fv (atomic adds into G) and of S1;
//   phase 2  line search of one segment, once all phase-1 segments of the hub are in: the segment's share of the
//            16 per-trial sums (atomic adds into ST);
//   phase 3  once per hub, after its phase-2 segments: gradient, active set, Armijo decision, new row.
// Items are handed out in that order by one counter and every warp holds one item at a time on a grid whose
// warps are all resident, so a waiting warp only waits for items that are being processed: no deadlock.
// (G is summed by atomics in arrival order, so a hub's gradient is reproducible only to rounding.)
constexpr int kSpHubSeg = 256;

// kPush: multi-GPU launch, the peers' replicas are written too; kHub: the launch has split hubs.  Both are
// compile-time so that the plain single-GPU kernel carries none of that code.
template <bool kPush, bool kHub>
__global__ void __launch_bounds__(kSpThreads, kSpBlocksPerSM) sparse_step_kernel(const StepArgs a, const SparseArgs sp) {
    if (a.done_flag != nullptr && *a.done_flag != 0) return;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = a.ld;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wpb = (int)(blockDim.x >> 5), nthreads = (int)blockDim.x;
    const int ecap = sp_entries(ld);
    const int ldp = sp_ldp(ld);
    double *s_steps = reinterpret_cast<double *>(smem_raw);
    double *s_sumF = s_steps + kMaxSteps;
    double *s_D = s_sumF + ldp;
    unsigned char *wbase = reinterpret_cast<unsigned char *>(s_D + ldp) + (size_t)wib * sp_warp_bytes(ld);
    double *fu_d = reinterpret_cast<double *>(wbase);
    double *g_d = fu_d + ldp;
    double *ent_val = g_d + ldp;
    unsigned short *ent_idx = reinterpret_cast<unsigned short *>(ent_val + ecap);
    unsigned short *aidx = ent_idx + ecap;
    unsigned short *poff = aidx + (ld > 256 ? ld : 256);
    unsigned int *cbal = reinterpret_cast<unsigned int *>(poff + 40);
    unsigned short *ccum = reinterpret_cast<unsigned short *>(cbal + 32);

#pragma unroll 1
    for (int i = threadIdx.x; i < ldp; i += nthreads) { s_sumF[i] = (i < ld) ? a.sumF[i] : 0.0; s_D[i] = 0.0; }
#pragma unroll 1
    for (int i = threadIdx.x; i < kMaxSteps; i += nthreads) s_steps[i] = a.steps[i];
#pragma unroll 1
    for (int i = lane; i < ldp; i += 32) { fu_d[i] = 0.0; g_d[i] = 0.0; }
    __syncthreads();

    const EdgeConst ec = {a.x_lo, a.x_hi, a.t_lo, a.t_hi, a.w_lo, a.w_hi};
    const double max_f = a.max_f;
    const int nsteps = a.nsteps;
    const int64_t order_n = a.order_n;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int j16 = lane & 15, h = lane >> 4;
    double llh_acc = 0.0, nupd_acc = 0.0;

    // ---- pieces shared by the plain node path and the hub phases ----
    // PRE over the edges [eb, ee) of a node whose fu is in fu_d: returns this lane's share of S1; with `axpy`
    // the weighted neighbour rows are added into g_d.  `single` = the range fitted one staged chunk (its
    // entries are still in the buffer, `ne_last` rows).
    auto pre_range = [&](int64_t e0, int eb, int ee, bool axpy, int &nchunks, int &ne_last, int use_pre = 0,
                         unsigned long long pre_hv = 0ull) -> double {
        double S1 = 0.0;
        nchunks = 0;
        ne_last = 0;
        for (int cb = eb; cb < ee;) {
            const int ne = sp_stage_chunk(sp.hdr_in, sp.pool_in, a.col + e0 + cb, min(32, ee - cb), lane, ecap, ent_val, ent_idx, poff,
                                          (use_pre && cb == eb) ? 1 : 0, pre_hv);
            double x = 0.0;
            if (lane < ne) {
                const int end = poff[lane + 1];
                for (int i = poff[lane]; i < end; ++i) x = fma(ent_val[i], fu_d[ent_idx[i]], x);
            }
            double w;
            const double t = edge_term<true>(x, ec, w);
            S1 += (lane < ne) ? t : 0.0;
            if (axpy) {
                for (int e = 0; e < ne; ++e) {
                    const double we = __shfl_sync(0xffffffffu, w, e);
                    const int pe = poff[e], pn = poff[e + 1];
                    for (int i = pe + lane; i < pn; i += 32) {
                        const int c = ent_idx[i];
                        g_d[c] = fma(we, ent_val[i], g_d[c]);
                    }
                    __syncwarp();
                }
            }
            cb += ne;
            ne_last = ne;
            ++nchunks;
        }
        return S1;
    };
    // g_d (sum of weighted neighbour rows) -> gradient (:168) in place; returns |g|^2, lists the active
    // components in aidx (m of them) and tells whether any candidate can reach MAX_F_.
    auto scan_gradient = [&](int &m, bool &need_hi) -> double {
        double G2 = 0.0;
        bool hi_lane = false;
        m = 0;
        for (int c0 = 0; c0 < ldp; c0 += 32) {          // (padding components: f = g = 0, inactive)
            const int c = c0 + lane;
            const double f = fu_d[c];
            const double g = (g_d[c] - s_sumF[c]) + f;
            g_d[c] = g;
            G2 = fma(g, g, G2);
            const bool act = (f > 0.0 || g > 0.0);
            const unsigned bal = __ballot_sync(0xffffffffu, act);
            if (act) {
                aidx[m + __popc(bal & lt_mask)] = (unsigned short)c;
                hi_lane |= (f + g > max_f);
            }
            m += __popc(bal);
        }
        G2 = warp_sum(G2);
        need_hi = __any_sync(0xffffffffu, hi_lane);
        __syncwarp();
        return G2;
    };
    // The staged entries of `ne` rows shrink, in place, to those on ACTIVE components (fu > 0 or grad > 0): only
    // they can contribute to a candidate's dot (an inactive component clamps to 0), and a neighbour row typically
    // keeps ~3 of its ~9 entries.  Order inside a row is kept, poff is rewritten.
    auto compact_active = [&](int ne) {
        const int T = poff[ne];
        int total = 0;
        for (int base = 0; base < T; base += 32) {
            const int j = base + lane;
            const bool in = j < T;
            const int c = in ? (int)ent_idx[j] : 0;
            const double v = in ? ent_val[j] : 0.0;
            const bool act = in && (fu_d[c] > 0.0 || g_d[c] > 0.0);
            const unsigned bal = __ballot_sync(0xffffffffu, act);
            if (lane == 0) { cbal[base >> 5] = bal; ccum[base >> 5] = (unsigned short)total; }
            __syncwarp();                       // this block's reads are done; writes land at or below them
            if (act) {
                const int p = total + __popc(bal & lt_mask);
                ent_idx[p] = (unsigned short)c;
                ent_val[p] = v;
            }
            total += __popc(bal);
        }
        __syncwarp();
        int newp = 0;
        if (lane < ne) {
            const int p = poff[lane];
            newp = (p >= T) ? total : (int)ccum[p >> 5] + __popc(cbal[p >> 5] & ((1u << (p & 31)) - 1u));
        }
        __syncwarp();
        if (lane < ne) poff[lane] = (unsigned short)newp;
        if (lane == 0) poff[ne] = (unsigned short)total;
        __syncwarp();
    };
    // Line search over the edges [eb, ee): lane (j, h) returns the sum over its edges of the clamped edge term
    // for candidate step s; `staged` rows of a single chunk may still be in the buffer from PRE.
    auto ls_range = [&](int64_t e0, int eb, int ee, double s, bool need_hi, int staged) -> double {
        double sumterms = 0.0;
        for (int cb = eb; cb < ee;) {
            const int ne = (staged > 0) ? staged
                                        : sp_stage_chunk(sp.hdr_in, sp.pool_in, a.col + e0 + cb, min(32, ee - cb), lane, ecap, ent_val, ent_idx, poff);
            compact_active(ne);
#pragma unroll 1
            for (int e2 = 0; e2 < ne; e2 += 4) {
                const int eA = e2 + h, eB = e2 + 2 + h;
                const bool vA = eA < ne, vB = eB < ne;
                const int iA = vA ? (int)poff[eA] : 0, nA = vA ? (int)poff[eA + 1] - iA : 0;
                const int iB = vB ? (int)poff[eB] : 0, nB = vB ? (int)poff[eB + 1] - iB : 0;
                const int nmax = max(nA, nB);
                double DA = 0.0, DB = 0.0;
#pragma unroll 1
                for (int k = 0; k < nmax; ++k) {
                    const bool ka = k < nA, kb = k < nB;
                    const int ca = ka ? (int)ent_idx[iA + k] : 0, cb2 = kb ? (int)ent_idx[iB + k] : 0;
                    const double pa = ka ? ent_val[iA + k] : 0.0, pb = kb ? ent_val[iB + k] : 0.0;
                    const double fa = fu_d[ca], ga = g_d[ca], fb = fu_d[cb2], gb = g_d[cb2];
                    if (need_hi) {
                        DA = fma(clamp_step0(fa, s, ga, max_f), pa, DA);
                        DB = fma(clamp_step0(fb, s, gb, max_f), pb, DB);
                    } else {
                        DA = fma(clamp_step0_lo(fa, s, ga), pa, DA);
                        DB = fma(clamp_step0_lo(fb, s, gb), pb, DB);
                    }
                }
                double tA, tB;
                edge_term2(DA, DB, ec, tA, tB);
                sumterms += vA ? tA : 0.0;
                sumterms += vB ? tB : 0.0;
            }
            __syncwarp();
            cb += ne;
        }
        return sumterms;
    };
    // Armijo decision for the 16 candidates tg .. tg+15 given each lane's edge-term sum (already summed over h).
    auto decide = [&](int tg, double s, bool jok, double sumterms, int m, bool need_hi, double llh_u, double G2) -> int {
        // - newfu.sfT + newfu.newfu with sfT = (sumF - fu) + newfu   (:176,:180)
        double oa = 0.0, ob = 0.0;
        for (int t = h; t < m; t += 2) {
            const int c = aidx[t];
            const double f = fu_d[c], g = g_d[c];
            const double nf = need_hi ? clamp_step0(f, s, g, max_f) : clamp_step0_lo(f, s, g);
            const double sf = (s_sumF[c] - f) + nf;
            oa = fma(nf, sf, oa);
            ob = fma(nf, nf, ob);
        }
        oa += __shfl_xor_sync(0xffffffffu, oa, 16);
        ob += __shfl_xor_sync(0xffffffffu, ob, 16);
        const double result = (sumterms - oa) + ob;
        const double rhs = llh_u + (a.alpha * s) * G2;
        const unsigned pass = __ballot_sync(0xffffffffu, jok && (result >= rhs)) & 0xffffu;
        return pass ? tg + __ffs(pass) - 1 : -1;          // lowest j == largest step (:182 max)
    };
    // SWAP (:183-190): the accepted candidate's non-zeros (or the old row) go to the output pool(s).
    auto swap_row = [&](int64_t u, int jstar, int m, int cu, const double *uval, const unsigned short *uidx) {
        int cnt_new = 0;
        if (jstar >= 0) {
            const double s = s_steps[jstar];
            for (int t0 = 0; t0 < m; t0 += 32) {
                const int t = t0 + lane;
                const bool ok = t < m;
                const int c = ok ? (int)aidx[t] : 0;
                const double f = fu_d[c], g = g_d[c];
                const double nr = clamp_step(f, s, g, a.min_f, max_f);
                const bool nz = ok && (nr != 0.0);
                const unsigned bal = __ballot_sync(0xffffffffu, nz);
                if (nz) {
                    const int p = cnt_new + __popc(bal & lt_mask);
                    ent_val[p] = nr;
                    ent_idx[p] = (unsigned short)c;
                }
                if (ok && f != nr) atomicAdd(s_D + c, f - nr);       // :191-192, sum over accepted nodes of old - new
                cnt_new += __popc(bal);
            }
            nupd_acc += 1.0;
        } else {
            for (int i = lane; i < cu; i += 32) {
                ent_val[i] = __ldg(uval + i);
                ent_idx[i] = __ldg(uidx + i);
            }
            cnt_new = cu;
        }
        __syncwarp();
        const unsigned long long words = sp_words((uint32_t)cnt_new);
        unsigned long long rel = 0;
        if (lane == 0 && cnt_new > 0) rel = atomicAdd(sp.pool_top, words);
        rel = __shfl_sync(0xffffffffu, rel, 0);
        if (rel + words > sp.pool_cap8) {
            if (lane == 0) { *sp.overflow = 1; sp.hdr_out[u] = sp_pack(0, 0); }
        } else {
            const unsigned long long off = sp.region_base8 + rel;
            const uint64_t hnew = sp_pack(off, (uint32_t)cnt_new);
            double *ov = sp.pool_out + off;
            unsigned short *oi = reinterpret_cast<unsigned short *>(ov + sp_pad((uint32_t)cnt_new));
            for (int i = lane; i < cnt_new; i += 32) {
                ov[i] = ent_val[i];
                oi[i] = ent_idx[i];
            }
            if (lane == 0) sp.hdr_out[u] = hnew;
            if (kPush) {
                for (int pr = 0; pr < sp.n_peers; ++pr) {
                    double *pv = sp.peer_pool[pr] + off;
                    unsigned short *pi = reinterpret_cast<unsigned short *>(pv + sp_pad((uint32_t)cnt_new));
                    for (int i = lane; i < cnt_new; i += 32) {
                        pv[i] = ent_val[i];
                        pi[i] = ent_idx[i];
                    }
                    if (lane == 0) sp.peer_hdr[pr][u] = hnew;
                }
            }
        }
    };

    // ---------------- split hubs (see above), then one warp per node ----------------
    if constexpr (kHub) {
        for (;;) {
            unsigned int it = 0;
            if (lane == 0) it = atomicAdd(sp.hub_work, 1u);
            it = __shfl_sync(0xffffffffu, it, 0);
            if (it >= (unsigned int)a.n_hub_items) break;
            const HubItem item = a.hub_items[it];
            const NodeMeta nm = a.meta[item.hub];
            const int64_t u = nm.u, e0 = nm.e0;
            const int deg = nm.deg;
            double *scr = a.hub_scratch + (size_t)item.mslot * (ld + 32);
            unsigned int *cnt = a.hub_counters + 2 * (size_t)item.mslot;
            const int sb = item.slice * kSpHubSeg, se = min(deg, sb + kSpHubSeg);
            const uint64_t hu = __ldg(sp.hdr_in + u);
            const int cu = (int)sp_cnt(hu);
            const double *uval = sp.pool_in + sp_off8(hu);
            const unsigned short *uidx = reinterpret_cast<const unsigned short *>(uval + sp_pad((uint32_t)cu));
            double fusf = 0.0, fufu = 0.0;
            for (int i = lane; i < cu; i += 32) {
                const double v = __ldg(uval + i);
                const int c = __ldg(uidx + i);
                fu_d[c] = v;
                fusf = fma(v, s_sumF[c], fusf);
                fufu = fma(v, v, fufu);
            }
            fusf = warp_sum(fusf);
            fufu = warp_sum(fufu);
            __syncwarp();
            const bool in_uset = (a.node_mask == nullptr) || (a.node_mask[u] != 0);
            const bool want_ls = a.do_linesearch && in_uset;
            if (item.phase == 1) {
                int nch, nel;
                double S1 = warp_sum(pre_range(e0, sb, se, want_ls, nch, nel));
                if (want_ls) {
                    for (int c = lane; c < ld; c += 32) {
                        const double v = g_d[c];
                        if (v != 0.0) atomicAdd(scr + c, v);
                        g_d[c] = 0.0;
                    }
                }
                if (lane == 0) atomicAdd(scr + ld, S1);
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(cnt, 1u);
            } else {
                if (lane == 0) {
                    const unsigned int *c = cnt + (item.phase == 2 ? 0 : 1);
                    while (*reinterpret_cast<const volatile unsigned int *>(c) < (unsigned int)item.nslices) __nanosleep(200);
                    __threadfence();
                }
                __syncwarp();
                const double llh_u = (__ldcg(scr + ld) - fusf) + fufu;
                int m = 0, jstar = -1;
                bool need_hi = false;
                double G2 = 0.0;
                if (want_ls) {
                    for (int c = lane; c < ldp; c += 32) g_d[c] = (c < ld) ? __ldcg(scr + c) : 0.0;
                    __syncwarp();
                    G2 = scan_gradient(m, need_hi);
                }
                const double s = s_steps[j16 < nsteps ? j16 : 0];       // hubs are only split when nsteps <= 16
                if (item.phase == 2) {
                    if (want_ls) {
                        double st = ls_range(e0, sb, se, s, need_hi, 0);
                        st += __shfl_xor_sync(0xffffffffu, st, 16);
                        if (lane < 16) atomicAdd(scr + ld + 1 + lane, st);
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) atomicAdd(cnt + 1, 1u);
                } else {
                    if (want_ls) jstar = decide(0, s, j16 < nsteps, __ldcg(scr + ld + 1 + j16), m, need_hi, llh_u, G2);
                    llh_acc += llh_u;
                    if (a.do_linesearch) swap_row(u, jstar, m, cu, uval, uidx);
                    if (a.accepted != nullptr && lane == 0) a.accepted[u] = (int8_t)jstar;
                }
                __syncwarp();
                if (want_ls)
                    for (int c = lane; c < ldp; c += 32) g_d[c] = 0.0;
            }
            __syncwarp();
            for (int i = lane; i < cu; i += 32) fu_d[__ldg(uidx + i)] = 0.0;
            __syncwarp();
        }
    }

    // positions n_hubs .. n_hubs + 3*#warps - 1 are pre-assigned, the rest is handed out by the work counter two
    // nodes ahead
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    int64_t pos = (kHub ? (int64_t)a.n_hubs : 0) + (int64_t)blockIdx.x * wpb + wib;
    int64_t pos_n = pos + nwarps, pos_nn = pos + 2 * nwarps;
    NodeMeta cur = {0, 0, 0}, nxt = {0, 0, 0};
    if (pos < order_n) cur = a.meta[pos];
    if (pos_n < order_n) nxt = a.meta[pos_n];
    // kSpPrefetch: what the current node needs first is loaded while the previous one is processed — its header
    // (c_hu), its first 32 entries (lane i holds entry i), the headers of its first 32 neighbours (c_hv)
    unsigned long long c_hu = 0ull, c_hv = 0ull;
    double c_val = 0.0;
    int c_idx = 0;
    if (kSpPrefetch && pos < order_n) {
        c_hu = __ldg(sp.hdr_in + cur.u);
        const int v0 = (lane < min(32, cur.deg)) ? a.col[cur.e0 + lane] : 0;
        c_hv = (lane < min(32, cur.deg)) ? __ldg(sp.hdr_in + v0) : 0ull;
        const int cu0 = (int)sp_cnt(c_hu);
        const double *uv0 = sp.pool_in + sp_off8(c_hu);
        if (lane < cu0) {
            c_val = __ldg(uv0 + lane);
            c_idx = __ldg(reinterpret_cast<const unsigned short *>(uv0 + sp_pad((uint32_t)cu0)) + lane);
        }
    }

    while (pos < order_n) {
        const int64_t u = cur.u, e0 = cur.e0;
        const int deg = cur.deg;
        NodeMeta nn = {0, 0, 0};
        if (pos_nn < order_n) nn = a.meta[pos_nn];
        unsigned int fetched = 0;
        if (lane == 0) fetched = atomicAdd(a.work_counter, 1u);
        // stage 1 of the prefetch for the next node: header and neighbour ids
        const bool has_next = pos_n < order_n;
        unsigned long long n_hu = 0ull, n_hv = 0ull;
        int n_v = 0, n_idx = 0;
        double n_val = 0.0;
        if (kSpPrefetch && has_next) {
            n_hu = __ldg(sp.hdr_in + nxt.u);
            n_v = (lane < min(32, nxt.deg)) ? a.col[nxt.e0 + lane] : 0;
        }

        // ---- own row: scatter into fu_d ----
        const uint64_t hu = kSpPrefetch ? (uint64_t)c_hu : __ldg(sp.hdr_in + u);
        const int cu = (int)sp_cnt(hu);
        const double *uval = sp.pool_in + sp_off8(hu);
        const unsigned short *uidx = reinterpret_cast<const unsigned short *>(uval + sp_pad((uint32_t)cu));
        double fusf = 0.0, fufu = 0.0;
        for (int i = lane; i < cu; i += 32) {
            const bool first = kSpPrefetch && i < 32;
            const double v = first ? c_val : __ldg(uval + i);
            const int c = first ? c_idx : (int)__ldg(uidx + i);
            fu_d[c] = v;
            fusf = fma(v, s_sumF[c], fusf);
            fufu = fma(v, v, fufu);
        }
        fusf = warp_sum(fusf);
        fufu = warp_sum(fufu);
        __syncwarp();

        const bool in_uset = (a.node_mask == nullptr) || (a.node_mask[u] != 0);
        const bool want_ls = a.do_linesearch && in_uset && deg > 0;

        // ---------------- PRE (:157-169) ----------------
        int nchunks, ne_last;
        const double S1 = warp_sum(pre_range(e0, 0, deg, want_ls, nchunks, ne_last, kSpPrefetch ? 1 : 0, c_hv));
        const double llh_u = (S1 - fusf) + fufu;
        llh_acc += llh_u;
        // stage 2 of the prefetch: the next node's first entries and its neighbours' headers (stage 1 has landed)
        if (kSpPrefetch && has_next) {
            n_hv = (lane < min(32, nxt.deg)) ? __ldg(sp.hdr_in + n_v) : 0ull;
            const int cun = (int)sp_cnt(n_hu);
            const double *uvn = sp.pool_in + sp_off8(n_hu);
            if (lane < cun) {
                n_val = __ldg(uvn + lane);
                n_idx = __ldg(reinterpret_cast<const unsigned short *>(uvn + sp_pad((uint32_t)cun)) + lane);
            }
        }

        int jstar = -1, m = 0;
        if (want_ls) {
            bool need_hi;
            const double G2 = scan_gradient(m, need_hi);
            // ---------------- LS (:172-182) ----------------
            for (int tg = 0; tg < nsteps && jstar < 0; tg += 16) {
                const int j = tg + j16;
                const bool jok = j < nsteps;
                const double s = s_steps[jok ? j : 0];
                // a node whose neighbours fitted one chunk still has them staged from PRE
                double sumterms = ls_range(e0, 0, deg, s, need_hi, (nchunks == 1 && tg == 0) ? ne_last : 0);
                sumterms += __shfl_xor_sync(0xffffffffu, sumterms, 16);
                jstar = decide(tg, s, jok, sumterms, m, need_hi, llh_u, G2);
            }
        }
        if (a.do_linesearch) swap_row(u, jstar, m, cu, uval, uidx);
        if (a.accepted != nullptr && lane == 0) a.accepted[u] = (int8_t)jstar;

        // ---- leave the warp's dense vectors at zero for the next node ----
        __syncwarp();
        for (int i = lane; i < cu; i += 32) fu_d[__ldg(uidx + i)] = 0.0;
        if (want_ls)
            for (int c = lane; c < ldp; c += 32) g_d[c] = 0.0;
        __syncwarp();

        cur = nxt;
        nxt = nn;
        pos = pos_n;
        pos_n = pos_nn;
        pos_nn = (int64_t)__shfl_sync(0xffffffffu, fetched, 0);
        if (kSpPrefetch) { c_hu = n_hu; c_hv = n_hv; c_val = n_val; c_idx = n_idx; }
    }

    // ---------------- block reduction of the partials ----------------
    __syncthreads();
    if (a.do_linesearch) {
        for (int i = threadIdx.x; i < ld; i += nthreads) {
            const double v = s_D[i];
            if (v != 0.0) atomicAdd(a.partials + i, v);
        }
    }
    __shared__ double s_red[2 * kSpWarps];
    if (lane == 0) { s_red[wib] = llh_acc; s_red[kSpWarps + wib] = nupd_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double l = 0.0, c = 0.0;
#pragma unroll
        for (int w = 0; w < kSpWarps; ++w)
            if (w < wpb) { l += s_red[w]; c += s_red[kSpWarps + w]; }
        atomicAdd(a.partials + 2 * ld, l);
        if (c != 0.0) atomicAdd(a.partials + 2 * ld + 1, c);
    }
}

// Dense n x ld rows -> sparse rows (one warp per row; non-zeros in ascending component order).
__global__ void dense_to_sparse_kernel(const double *F, int64_t n, int ld, uint64_t *hdr, double *pool,
                                       unsigned long long *pool_top, uint64_t pool_cap8, int32_t *overflow) {
    const int lane = threadIdx.x & 31;
    const int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= n) return;
    const unsigned lt_mask = (1u << lane) - 1u;
    const double *row = F + (size_t)u * ld;
    int cnt = 0;
    for (int c0 = 0; c0 < ld; c0 += 32) {
        const int c = c0 + lane;
        cnt += __popc(__ballot_sync(0xffffffffu, c < ld && row[c] != 0.0));
    }
    const unsigned long long words = sp_words((uint32_t)cnt);
    unsigned long long off = 0;
    if (lane == 0 && cnt > 0) off = atomicAdd(pool_top, words);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + words > pool_cap8) {
        if (lane == 0) { *overflow = 1; hdr[u] = sp_pack(0, 0); }
        return;
    }
    double *ov = pool + off;
    unsigned short *oi = reinterpret_cast<unsigned short *>(ov + sp_pad((uint32_t)cnt));
    int p = 0;
    for (int c0 = 0; c0 < ld; c0 += 32) {
        const int c = c0 + lane;
        const double v = (c < ld) ? row[c] : 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) {
            const int q = p + __popc(bal & lt_mask);
            ov[q] = v;
            oi[q] = (unsigned short)c;
        }
        p += __popc(bal);
    }
    if (lane == 0) hdr[u] = sp_pack(off, (uint32_t)cnt);
}

// Host side of the layout: rows given as CSR (indptr, ascending-or-not indices, values; explicit zeros are dropped)
// -> header + pool image, and back.  Used by bigclam_set_F_csr / bigclam_get_F_csr (and by the emulation tests).
// Returns the number of 8-byte words used, or -1 for an index outside [0, k) / a row longer than ld.
inline int64_t sp_host_pack(int64_t n, int32_t k, int32_t ld, const int64_t *indptr, const int32_t *indices, const double *values,
                            uint64_t *hdr, double *pool, uint64_t pool_cap8, double *colsum /* k, optional */) {
    uint64_t top = 0;
    if (colsum != nullptr)
        for (int32_t c = 0; c < k; ++c) colsum[c] = 0.0;
    for (int64_t u = 0; u < n; ++u) {
        uint32_t cnt = 0;
        for (int64_t i = indptr[u]; i < indptr[u + 1]; ++i) {
            if (indices[i] < 0 || indices[i] >= k) return -1;
            if (values[i] != 0.0) ++cnt;
        }
        if (cnt > (uint32_t)ld) return -1;
        const uint64_t words = sp_words(cnt);
        if (top + words > pool_cap8) return -2;
        double *ov = pool + top;
        unsigned short *oi = reinterpret_cast<unsigned short *>(ov + sp_pad(cnt));
        uint32_t q = 0;
        for (int64_t i = indptr[u]; i < indptr[u + 1]; ++i) {
            if (values[i] == 0.0) continue;
            // insertion keeps the indices ascending (rows arrive sorted in practice: one comparison per entry)
            uint32_t p = q;
            while (p > 0 && oi[p - 1] > (unsigned short)indices[i]) { oi[p] = oi[p - 1]; ov[p] = ov[p - 1]; --p; }
            oi[p] = (unsigned short)indices[i];
            ov[p] = values[i];
            ++q;
            if (colsum != nullptr) colsum[indices[i]] += values[i];
        }
        for (uint32_t z = cnt; z < sp_pad(cnt); ++z) { ov[z] = 0.0; oi[z] = 0; }
        hdr[u] = sp_pack(cnt ? top : 0, cnt);
        top += words;
    }
    return (int64_t)top;
}

inline int64_t sp_host_nnz(int64_t n, const uint64_t *hdr) {
    int64_t t = 0;
    for (int64_t u = 0; u < n; ++u) t += sp_cnt(hdr[u]);
    return t;
}

inline void sp_host_unpack(int64_t n, const uint64_t *hdr, const double *pool, int64_t *indptr, int32_t *indices, double *values) {
    int64_t t = 0;
    for (int64_t u = 0; u < n; ++u) {
        indptr[u] = t;
        const uint32_t cnt = sp_cnt(hdr[u]);
        const double *ov = pool + sp_off8(hdr[u]);
        const unsigned short *oi = reinterpret_cast<const unsigned short *>(ov + sp_pad(cnt));
        for (uint32_t i = 0; i < cnt; ++i) { indices[t] = oi[i]; values[t] = ov[i]; ++t; }
    }
    indptr[n] = t;
}

// Sparse rows -> dense n x ld (rows are zeroed here, no separate memset).
__global__ void sparse_to_dense_kernel(const uint64_t *hdr, const double *pool, int64_t n, int ld, double *F) {
    const int lane = threadIdx.x & 31;
    const int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= n) return;
    double *row = F + (size_t)u * ld;
    for (int c = lane; c < ld; c += 32) row[c] = 0.0;
    __syncwarp();
    const uint64_t h = hdr[u];
    const int cnt = (int)sp_cnt(h);
    const double *vals = pool + sp_off8(h);
    const unsigned short *idx = reinterpret_cast<const unsigned short *>(vals + sp_pad((uint32_t)cnt));
    for (int i = lane; i < cnt; i += 32) row[idx[i]] = vals[i];
}

}  // namespace bigclam
