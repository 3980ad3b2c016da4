// bigclam_kernels.cuh — sm_100a kernels of the BigCLAM hot path (one warp per node).
//
// What one launch of step_kernel computes, per node u (reference: codes/bigclam4-7.scala):
//   PRE   :157-169  x_uv = fu.fv, p = clamp(exp(-x)), grad_u = sum_v fv/(1-p) - sumF + fu,
//                   llh_u = sum_v (log(1-p)+x) - fu.sumF + fu.fu
//   LS    :172-182  16 candidates s_j, nf_j = clamp(fu + s_j*grad), Armijo test, max passing s
//   SWAP  :183-190  F_out[u] = nf_{j*} (or fu when nothing passes)  — Jacobi: only F_in is read
//   partial reductions for :191-192 (sum of old rows, sum of new rows) and for the LLH
//   (sum_u llh_u == the LLH the previous call returns, :196-219).
//
// Layout: F is n x ld fp64 row-major, ld = K rounded up to a multiple of 4 (32-byte sectors,
// 16-byte double2 loads); padding columns are zero and stay zero.  Lane l of the warp owns the
// double2 chunks q = l + 32c (components 2q, 2q+1), so a row load is C2 coalesced LDG.128.
//
// Line search, sparse path: a component can only matter in nf_j . fv if nf_j is non-zero for
// some j, i.e. fu_i > 0 or grad_i > 0 (MIN_F_ = 0).  Those "active" components (typically the
// node's few communities) are compacted into shared memory; lanes then re-map to (trial j, edge
// parity) and each lane evaluates its own trial for its own edges — every exp/log of the
// 16 x deg grid is computed by exactly one lane.  Rows with more than MAXM active components,
// or MIN_F_ != 0, take the dense path (lane-owned components, one trial at a time, early exit).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bigclam {

constexpr int kWarpsPerBlock = 8;
constexpr int kBlockThreads = kWarpsPerBlock * 32;
constexpr int kMaxActive = 64;      // active-set capacity of the sparse line-search path
constexpr int kMaxSteps = 64;       // MaxInter + 1 <= kMaxSteps

struct StepArgs {
    int64_t n;
    const int64_t *rowptr;
    const int32_t *col;
    const double *F_in;
    double *F_out;
    const double *sumF;
    int32_t k, ld;
    int32_t nsteps;
    double steps[kMaxSteps];
    double alpha, min_p, max_p, min_f, max_f;
    // thresholds/constants of the clamped edge term (exact shortcuts, see edge_eval)
    double x_lo, x_hi, t_lo, t_hi, w_lo, w_hi;
    const int32_t *order;     // processing order (degree descending) over the owned nodes
    int64_t order_n;
    const uint8_t *node_mask; // optional uset
    double *partials;         // [A(ld) | B(ld) | llh | n_updated]
    int8_t *accepted;         // optional, n
    const int32_t *done_flag; // optional: non-zero -> the launch is a no-op
    int32_t do_linesearch;    // 0: PRE/LLH only (loglikelihood())
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double2 ldg2(const double *p) {
    return __ldg(reinterpret_cast<const double2 *>(p));
}

// step(), bigclam4-7.scala:110-113: product and sum rounded separately (the JVM never fuses).
__device__ __forceinline__ double clamp_step(double f, double s, double g, double lo, double hi) {
    return fmin(fmax(__dadd_rn(f, __dmul_rn(s, g)), lo), hi);
}

// log(1 - clamp(exp(-x), MIN_P, MAX_P)) + x and 1/(1 - p)   (bigclam4-7.scala:166-167).
// Shortcuts are exact: x == 0 or x <= x_lo gives p == MAX_P after the clamp, x >= x_hi gives
// p == MIN_P; x_lo/x_hi carry a 1e-12 safety margin so the clamp outcome is never in doubt.
template <bool kNeedW>
__device__ __forceinline__ void edge_eval(double x, const StepArgs &a, double &t, double &w) {
    if (x <= a.x_lo) { t = a.t_lo + x; if (kNeedW) w = a.w_lo; return; }
    if (x >= a.x_hi) { t = a.t_hi + x; if (kNeedW) w = a.w_hi; return; }
    double p = fmin(fmax(exp(-x), a.min_p), a.max_p);
    double omp = 1.0 - p;
    t = log(omp) + x;
    if (kNeedW) w = 1.0 / omp;
}

template <int C2> struct RowsInFlight { static constexpr int value = (C2 <= 4) ? 4 : (C2 <= 8 ? 2 : 1); };

// Dots of `vec` (lane-owned components) with the rows of up to 32 neighbours; lane e returns
// the dot for neighbour e of the chunk.
template <int C2>
__device__ __forceinline__ double chunk_dots(const double2 (&vec)[C2], const double *__restrict__ F,
                                             int ld, int ld2, int lane, int myv, int cnt) {
    constexpr int R = RowsInFlight<C2>::value;
    double myx = 0.0;
    for (int eb = 0; eb < cnt; eb += R) {
        double part[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int e = eb + r;
            const int v = __shfl_sync(0xffffffffu, myv, e & 31);
            const double *fv = F + (size_t)v * ld;
            double p = 0.0;
            if (e < cnt) {
#pragma unroll
                for (int c = 0; c < C2; ++c) {
                    const int q = lane + 32 * c;
                    if (q < ld2) {
                        const double2 x = ldg2(fv + 2 * q);
                        p = fma(vec[c].x, x.x, p);
                        p = fma(vec[c].y, x.y, p);
                    }
                }
            }
            part[r] = p;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (lane == eb + r) myx = part[r];
    }
    return myx;
}

template <int C2>
__global__ void __launch_bounds__(kBlockThreads) step_kernel(const StepArgs a) {
    if (a.done_flag != nullptr && *a.done_flag != 0) return;

    extern __shared__ double smem[];
    const int ld = a.ld, ld2 = a.ld >> 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *s_sumF = smem;                                  // ld
    double *s_A = smem + ld + (size_t)wib * 2 * ld;         // per warp: A(ld) | B(ld)
    double *s_B = s_A + ld;
    double *s_lists = smem + ld + (size_t)kWarpsPerBlock * 2 * ld;
    double *s_afu = s_lists + (size_t)wib * 3 * kMaxActive; // per warp: fu | g | idx (as int)
    double *s_ag = s_afu + kMaxActive;
    int *s_aidx = reinterpret_cast<int *>(s_ag + kMaxActive);

    for (int i = threadIdx.x; i < ld; i += kBlockThreads) s_sumF[i] = a.sumF[i];
    for (int i = lane; i < 2 * ld; i += 32) s_A[i] = 0.0;
    __syncthreads();

    double llh_acc = 0.0;
    double nupd_acc = 0.0;
    const int64_t warp_global = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const double *__restrict__ F = a.F_in;

    for (int64_t pos = warp_global; pos < a.order_n; pos += nwarps) {
        const int64_t u = a.order[pos];
        const int64_t e0 = a.rowptr[u];
        const int deg = (int)(a.rowptr[u + 1] - e0);
        const double *frow = F + (size_t)u * ld;
        double *orow = a.F_out + (size_t)u * ld;

        double2 fu[C2];
        double fusf = 0.0, fufu = 0.0;
#pragma unroll
        for (int c = 0; c < C2; ++c) {
            const int q = lane + 32 * c;
            fu[c] = (q < ld2) ? ldg2(frow + 2 * q) : make_double2(0.0, 0.0);
            if (q < ld2) {
                const double2 sf = *reinterpret_cast<const double2 *>(s_sumF + 2 * q);
                fusf = fma(fu[c].x, sf.x, fusf); fusf = fma(fu[c].y, sf.y, fusf);
                fufu = fma(fu[c].x, fu[c].x, fufu); fufu = fma(fu[c].y, fu[c].y, fufu);
            }
        }
        fusf = warp_sum(fusf);
        fufu = warp_sum(fufu);

        // ---------------- PRE (:157-169) ----------------
        double2 g[C2];
#pragma unroll
        for (int c = 0; c < C2; ++c) g[c] = make_double2(0.0, 0.0);
        double S1 = 0.0;
        for (int cb = 0; cb < deg; cb += 32) {
            const int cnt = min(32, deg - cb);
            const int myv = (lane < cnt) ? a.col[e0 + cb + lane] : 0;
            const double myx = chunk_dots<C2>(fu, F, ld, ld2, lane, myv, cnt);
            double t = 0.0, w = 0.0;
            if (lane < cnt) edge_eval<true>(myx, a, t, w);
            S1 += warp_sum(t);
            if (a.do_linesearch) {
                for (int e = 0; e < cnt; ++e) {
                    const int v = __shfl_sync(0xffffffffu, myv, e);
                    const double we = __shfl_sync(0xffffffffu, w, e);
                    const double *fv = F + (size_t)v * ld;
#pragma unroll
                    for (int c = 0; c < C2; ++c) {
                        const int q = lane + 32 * c;
                        if (q < ld2) {
                            const double2 x = ldg2(fv + 2 * q);
                            g[c].x = fma(we, x.x, g[c].x);
                            g[c].y = fma(we, x.y, g[c].y);
                        }
                    }
                }
            }
        }
        const double llh_u = (S1 - fusf) + fufu;
        llh_acc += llh_u;

        const bool in_uset = (a.node_mask == nullptr) || (a.node_mask[u] != 0);
        int jstar = -1;

        if (a.do_linesearch && in_uset && deg > 0) {
            // grad = (sum - sumF) + fu  (:168);  G2 = grad.grad for the Armijo slope (:181)
            double G2 = 0.0;
#pragma unroll
            for (int c = 0; c < C2; ++c) {
                const int q = lane + 32 * c;
                if (q < ld2) {
                    const double2 sf = *reinterpret_cast<const double2 *>(s_sumF + 2 * q);
                    g[c].x = (g[c].x - sf.x) + fu[c].x;
                    g[c].y = (g[c].y - sf.y) + fu[c].y;
                    G2 = fma(g[c].x, g[c].x, G2);
                    G2 = fma(g[c].y, g[c].y, G2);
                }
            }
            G2 = warp_sum(G2);

            // ---- active set: components whose candidate value can be non-zero ----
            int m = 0;
            const bool sparse_ok = (a.min_f == 0.0);
            if (sparse_ok) {
#pragma unroll
                for (int c = 0; c < C2; ++c) {
                    const int q = lane + 32 * c;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const double fval = hh ? fu[c].y : fu[c].x;
                        const double gval = hh ? g[c].y : g[c].x;
                        const bool act = (q < ld2) && (fval > 0.0 || gval > 0.0);
                        const unsigned bal = __ballot_sync(0xffffffffu, act);
                        if (act) {
                            const int posn = m + __popc(bal & ((1u << lane) - 1u));
                            if (posn < kMaxActive) {
                                s_afu[posn] = fval;
                                s_ag[posn] = gval;
                                s_aidx[posn] = 2 * q + hh;
                            }
                        }
                        m += __popc(bal);
                    }
                }
                __syncwarp();
            }

            if (sparse_ok && m <= kMaxActive) {
                // ---------------- LS, sparse path: lane = (trial j, edge parity h) ----------------
                const int j16 = lane & 15, h = lane >> 4;
                for (int tg = 0; tg < a.nsteps && jstar < 0; tg += 16) {
                    const int j = tg + j16;
                    const bool jok = j < a.nsteps;
                    const double s = a.steps[jok ? j : 0];
                    double sumterms = 0.0;
                    for (int cb = 0; cb < deg; cb += 8) {
                        int v[4];
                        bool ok[4];
                        double acc[4];
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const int e = cb + 2 * qq + h;
                            ok[qq] = e < deg;
                            v[qq] = ok[qq] ? a.col[e0 + e] : 0;
                            acc[qq] = 0.0;
                        }
                        for (int t = 0; t < m; ++t) {
                            const double nf = clamp_step(s_afu[t], s, s_ag[t], 0.0, a.max_f);
                            const int idx = s_aidx[t];
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq)
                                if (ok[qq]) acc[qq] = fma(nf, __ldg(F + (size_t)v[qq] * ld + idx), acc[qq]);
                        }
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            if (ok[qq] && jok) {
                                double t, w;
                                edge_eval<false>(acc[qq], a, t, w);
                                sumterms += t;
                            }
                        }
                    }
                    sumterms += __shfl_xor_sync(0xffffffffu, sumterms, 16);
                    // - newfu.sfT + newfu.newfu with sfT = (sumF - fu) + newfu   (:176,:180)
                    double oa = 0.0, ob = 0.0;
                    for (int t = h; t < m; t += 2) {
                        const double fut = s_afu[t];
                        const double nf = clamp_step(fut, s, s_ag[t], 0.0, a.max_f);
                        const double sf = (s_sumF[s_aidx[t]] - fut) + nf;
                        oa = fma(nf, sf, oa);
                        ob = fma(nf, nf, ob);
                    }
                    oa += __shfl_xor_sync(0xffffffffu, oa, 16);
                    ob += __shfl_xor_sync(0xffffffffu, ob, 16);
                    const double result = (sumterms - oa) + ob;
                    const double rhs = llh_u + (a.alpha * s) * G2;
                    const unsigned pass = __ballot_sync(0xffffffffu, jok && (result >= rhs)) & 0xffffu;
                    if (pass) jstar = tg + __ffs(pass) - 1;   // lowest j == largest step (:182 max)
                }
                __syncwarp();
            } else {
                // ---------------- LS, dense path: one candidate at a time, descending ----------------
                for (int j = 0; j < a.nsteps && jstar < 0; ++j) {
                    const double s = a.steps[j];
                    double2 nf[C2];
                    double oa = 0.0, ob = 0.0;
#pragma unroll
                    for (int c = 0; c < C2; ++c) {
                        const int q = lane + 32 * c;
                        nf[c] = make_double2(0.0, 0.0);
                        if (q < ld2) {
                            const double2 sf = *reinterpret_cast<const double2 *>(s_sumF + 2 * q);
                            nf[c].x = clamp_step(fu[c].x, s, g[c].x, a.min_f, a.max_f);
                            nf[c].y = clamp_step(fu[c].y, s, g[c].y, a.min_f, a.max_f);
                            oa = fma(nf[c].x, (sf.x - fu[c].x) + nf[c].x, oa);
                            oa = fma(nf[c].y, (sf.y - fu[c].y) + nf[c].y, oa);
                            ob = fma(nf[c].x, nf[c].x, ob);
                            ob = fma(nf[c].y, nf[c].y, ob);
                        }
                    }
                    oa = warp_sum(oa);
                    ob = warp_sum(ob);
                    double sumterms = 0.0;
                    for (int cb = 0; cb < deg; cb += 32) {
                        const int cnt = min(32, deg - cb);
                        const int myv = (lane < cnt) ? a.col[e0 + cb + lane] : 0;
                        const double myx = chunk_dots<C2>(nf, F, ld, ld2, lane, myv, cnt);
                        double t = 0.0, w;
                        if (lane < cnt) edge_eval<false>(myx, a, t, w);
                        sumterms += warp_sum(t);
                    }
                    const double result = (sumterms - oa) + ob;
                    const double rhs = llh_u + (a.alpha * s) * G2;
                    if (result >= rhs) jstar = j;
                }
            }
        }

        // ---------------- SWAP (:183-190) + partial sums for :191-192 ----------------
        if (jstar >= 0) {
            const double s = a.steps[jstar];
#pragma unroll
            for (int c = 0; c < C2; ++c) {
                const int q = lane + 32 * c;
                if (q < ld2) {
                    double2 nr;
                    nr.x = clamp_step(fu[c].x, s, g[c].x, a.min_f, a.max_f);
                    nr.y = clamp_step(fu[c].y, s, g[c].y, a.min_f, a.max_f);
                    *reinterpret_cast<double2 *>(orow + 2 * q) = nr;
                    double2 *pa = reinterpret_cast<double2 *>(s_A + 2 * q);
                    double2 *pb = reinterpret_cast<double2 *>(s_B + 2 * q);
                    double2 va = *pa, vb = *pb;
                    va.x += fu[c].x; va.y += fu[c].y;
                    vb.x += nr.x; vb.y += nr.y;
                    *pa = va; *pb = vb;
                }
            }
            nupd_acc += 1.0;
        } else if (a.do_linesearch) {
#pragma unroll
            for (int c = 0; c < C2; ++c) {
                const int q = lane + 32 * c;
                if (q < ld2) *reinterpret_cast<double2 *>(orow + 2 * q) = fu[c];
            }
        }
        if (a.accepted != nullptr && lane == 0) a.accepted[u] = (int8_t)jstar;
    }

    // ---------------- block reduction of the partials, one RED per address per block ----------------
    __syncthreads();
    if (a.do_linesearch) {
        for (int i = threadIdx.x; i < 2 * ld; i += kBlockThreads) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kWarpsPerBlock; ++w) v += smem[ld + (size_t)w * 2 * ld + i];
            if (v != 0.0) atomicAdd(a.partials + i, v);
        }
    }
    __shared__ double s_red[2 * kWarpsPerBlock];
    if (lane == 0) { s_red[wib] = llh_acc; s_red[kWarpsPerBlock + wib] = nupd_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double l = 0.0, c = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) { l += s_red[w]; c += s_red[kWarpsPerBlock + w]; }
        atomicAdd(a.partials + 2 * ld, l);
        if (c != 0.0) atomicAdd(a.partials + 2 * ld + 1, c);
    }
}

// Device-side bookkeeping between two step kernels (single block).
struct RunState {
    double llhold;          // LLHold of the reference's loop
    double last_llh;        // most recent call's LLH
    double ret_llh;         // value the loop returns
    long long calls_seen;   // number of calls whose LLH is known
    long long conv_call;    // call index at which |1 - new/old| < tol, or 0
    long long n_updated;    // of the most recent applied step
    int done;
    int pad;
};

struct FinishArgs {
    double *partials;        // [A | B | llh | nupd] of the kernel that just ran
    const double *sumF_cur;
    double *sumF_next;
    int32_t ld;
    RunState *st;
    int32_t *done_flag;
    double *trace;           // optional device trace
    long long trace_cap;
    long long kernel_index;  // c: this is the finish of step kernel c (1-based); 0 = plain apply
    int32_t variant;
    double rel_tol;
    int32_t apply;           // apply the sumF update of this kernel
    int32_t llh_is_final;    // partials.llh is the LLH of the final state (tail LLH kernel)
};

__global__ void finish_kernel(const FinishArgs f) {
    __shared__ int s_done;
    const int ld = f.ld;
    if (threadIdx.x == 0) {
        RunState *st = f.st;
        int done = st->done;
        if (!done && f.kernel_index > 0) {
            // partials.llh == LLH(state before kernel c) == LLH returned by call c-1
            const double L = f.partials[2 * ld];
            const long long call = f.llh_is_final ? f.kernel_index : f.kernel_index - 1;
            if (call == 0) {
                if (f.variant == 2) st->llhold = L;          // LLHold = loglikelihood()
            } else {
                if (f.trace != nullptr && call - 1 < f.trace_cap) f.trace[call - 1] = L;
                st->last_llh = L;
                st->calls_seen = call;
                bool test = true;
                if (f.variant == 4 && call == 1) { st->llhold = L; test = false; }   // :228
                if (f.variant == 3 && call == 1) { st->llhold = 0.0; }               // v3 :207
                if (test) {
                    if (fabs(1.0 - L / st->llhold) < f.rel_tol) {                    // :237
                        done = 1;
                        st->done = 1;
                        st->conv_call = call;
                        st->ret_llh = (f.variant == 4) ? st->llhold : L;             // :242
                        *f.done_flag = 1;
                    } else {
                        st->llhold = L;
                    }
                }
                if (!done) st->ret_llh = L;
            }
        }
        s_done = done;
        if (!done && f.apply) st->n_updated = (long long)(f.partials[2 * ld + 1] + 0.5);
    }
    __syncthreads();
    if (s_done) return;
    if (f.apply) {
        const bool any = f.partials[2 * ld + 1] > 0.0;
        for (int i = threadIdx.x; i < ld; i += blockDim.x) {
            const double sf = f.sumF_cur[i];
            // sumF = sumF - (changeFu._1 - changeFu._2)   (:192)
            f.sumF_next[i] = any ? sf - (f.partials[i] - f.partials[ld + i]) : sf;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * ld + 2; i += blockDim.x) f.partials[i] = 0.0;
}

// F <-> pitched layout helpers and column sums.
__global__ void colsum_kernel(const double *F, int64_t n, int ld, double *out) {
    // one block per 32 columns slice; deterministic order within a thread, tree across threads
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rlane = threadIdx.x >> 5;             // 0..7
    double acc = 0.0;
    if (col < ld)
        for (int64_t r = rlane; r < n; r += 8) acc += F[(size_t)r * ld + col];
    __shared__ double s[8][33];
    s[rlane][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rlane == 0 && col < ld) {
        double v = 0.0;
        for (int i = 0; i < 8; ++i) v += s[i][threadIdx.x & 31];
        out[col] = v;
    }
}

}  // namespace bigclam
