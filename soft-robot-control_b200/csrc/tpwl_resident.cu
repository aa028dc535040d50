// tpwl_resident.cu -- TPWL nearest-neighbour rollout (sofacontrol/tpwl/tpwl.py:115-126, 160-168, 193-234) on a bank
// that needs no per-step discretisation, exploiting the TEMPORAL COHERENCE of the method: a trajectory stays in the
// region of one stored point for many steps (config-2 workload: the index changes on 6 % of the steps).
//
//   * The selected entry [A_i | B_i | d_i] (44 KB at the Diamond size) is RESIDENT in shared memory, fetched by TMA
//     bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) only when the nearest index changes; the affine step reads
//     it with conflict-free LDS.128.  Four trajectories per CTA (one CTA per SM), 128 threads each, independent
//     (named barriers): warp 0 of a group selects the point of x_t while warps 1-3 already evaluate
//     A_cur x_t + B_cur u_t + d_cur -- the result stands if the index did not change.
//   * EXACT search from a candidate list.  At a refresh step with sub-state c0 every distance d0_p (bit-for-bit the
//     reference's float64 value, tpwl.cuh) is computed; the list holds all p with
//         d0_p <= (d0_min + 2 Delta) (1 + 1e-11)
//     for the largest Delta of a geometric ladder that leaves at most 32 candidates.  While the weighted movement
//     w ||c - c0|| (rounded up) stays <= Delta, the triangle inequality puts every point outside the list strictly
//     behind the list's best:  d_p(c) >= d0_p - Delta > d0_min + Delta >= d_best(c); the factor 1e-11 covers the
//     rounding of the computed float64 distances (relative error < 1e-13 for sums of <= 1024 squares) on both sides,
//     so the reference's computed argmin -- first occurrence on ties included -- is among the candidates.  One lane
//     per candidate re-scores them with the bit-exact numpy-order distance from coordinates cached in shared memory;
//     first-occurrence argmin.  Movement beyond Delta, a non-finite distance or an overfull ladder: the group runs
//     the full float64 search again (np.argmin semantics as in tpwl.cu).
//
// The selected indices and the step arithmetic (summation order of the affine map) are identical to
// tpwl_rollout_nn_screen_kernel's (tests/test_tpwl_gpu.py compares index traces and states bit for bit).
#include <cstdlib>
#include "tpwl.cuh"

namespace srcb {

constexpr int kRG = 4;                  // trajectories (thread groups) per CTA
constexpr int kRT = 128;                // threads per group: warp 0 selects, warps 1-3 step
constexpr int kRThreads = kRG * kRT;
constexpr int kRC = 32;                 // candidate slots (one lane each)
constexpr int kRL = 16;                 // ladder levels: Delta_j = spread 2^-(j+1)
constexpr int kRPts = 8;                // stored points per thread in the full search (P <= kRPts * kRT)

struct ResidentPlan {                   // byte offsets inside one group's block
    size_t A, B, d, cc, c0, x, u, wcnt, redd, redm, redi, cidx, scal, delta, mbar, stride;
};
__host__ __device__ constexpr size_t rp_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ constexpr ResidentPlan make_resident_plan(int n, int m, int r) {
    ResidentPlan S{};
    size_t o = 0;
    S.A = o;     o += rp_up(sizeof(double) * n * n, 16);
    S.B = o;     o += rp_up(sizeof(double) * n * m, 16);
    S.d = o;     o += rp_up(sizeof(double) * n, 16);
    S.cc = o;    o += rp_up(sizeof(double) * r * kRC, 16);
    S.c0 = o;    o += rp_up(sizeof(double) * r, 16);
    S.x = o;     o += 2 * rp_up(sizeof(double) * n, 16);
    S.u = o;     o += 2 * rp_up(sizeof(double) * m, 16);
    S.wcnt = o;  o += sizeof(int) * 4 * kRL;
    S.redd = o;  o += sizeof(double) * 4;
    S.redm = o;  o += sizeof(double) * 4;
    S.redi = o;  o += sizeof(int) * 4;
    S.cidx = o;  o += sizeof(int) * kRC;
    S.scal = o;  o += sizeof(int) * 4;          // cnt, flag, sel, list valid
    S.delta = o; o += sizeof(double);
    S.mbar = o;  o += sizeof(uint64_t);
    S.stride = rp_up(o, 128);
    return S;
}

#ifdef SRCB_RES_STATS
__device__ unsigned long long g_res_stats[8];   // refreshes, refresh cycles, entry loads, load cycles, steps, step cycles
#define RS_ADD(i, v) do { if (ht == 0) atomicAdd(&g_res_stats[i], (unsigned long long)(v)); } while (0)
#else
#define RS_ADD(i, v) do { } while (0)
#endif

__device__ __forceinline__ void group_sync(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(kRT) : "memory");
}

// 2^-k, exactly (the ladder thresholds must be the same bits wherever they are recomputed)
__device__ __forceinline__ double pow2neg(int k) { return __longlong_as_double((long long)(1023 - k) << 52); }

// first-occurrence argmin across a warp
__device__ __forceinline__ void warp_argmin(double& d, int& p) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, d, off);
        const int op = __shfl_xor_sync(0xffffffffu, p, off);
        if (od < d || (od == d && op < p)) { d = od; p = op; }
    }
}

// Refresh step of one group (all 128 threads): full float64 search (the reference's distances, np.argmin order) and
// a new candidate list.  Kept out of line so that its registers do not weigh on the time loop.
template <int RT, int CM>
__device__ __noinline__ int resident_refresh(const TpwlDev& M, const double* __restrict__ xc, unsigned char* gb, int grp,
                             int ht, int useq) {
    constexpr int r = RT, n = 2 * RT, m = CM;
    constexpr ResidentPlan S = make_resident_plan(n, m, r);
    const int P = M.P, lane = ht & 31, hw = ht >> 5;
    double* cc = reinterpret_cast<double*>(gb + S.cc);
    double* c0 = reinterpret_cast<double*>(gb + S.c0);
    int* wcnt = reinterpret_cast<int*>(gb + S.wcnt);
    double* redd = reinterpret_cast<double*>(gb + S.redd);
    double* redm = reinterpret_cast<double*>(gb + S.redm);
    int* redi = reinterpret_cast<int*>(gb + S.redi);
    int* cidx = reinterpret_cast<int*>(gb + S.cidx);
    int* scal = reinterpret_cast<int*>(gb + S.scal);
    double* deltap = reinterpret_cast<double*>(gb + S.delta);
    const double* bankT = useq ? M.qT : M.vT;
    const int xoff = useq ? r : 0;
    // ---- refresh: full float64 search (the reference's distances, np.argmin order) + new candidate list
    double dl[kRPts];
    double best = INFINITY, dmx = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kRPts; ++k) {
        const int p = ht + k * kRT;
        dl[k] = INFINITY;
        if (p < P) {
            const double dd = tpwl_distance(M, xc, p);
            dl[k] = dd;
            if (dd < best) { best = dd; bi = p; }
            if (dd > dmx && dd < INFINITY) dmx = dd;
        }
    }
    warp_argmin(best, bi);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dmx = fmax(dmx, __shfl_xor_sync(0xffffffffu, dmx, off));
    if (lane == 0) { redd[hw] = best; redi[hw] = bi; redm[hw] = dmx; }
    group_sync(grp);
    best = redd[0]; bi = redi[0]; dmx = redm[0];
#pragma unroll
    for (int k = 1; k < kRT / 32; ++k) {
        const double od = redd[k];
        const int op = redi[k];
        if (od < best || (od == best && op < bi)) { best = od; bi = op; }
        dmx = fmax(dmx, redm[k]);
    }
    const int sel = (bi == 0x7fffffff) ? 0 : bi;      // all distances NaN / inf: np.argmin of an all-inf row
    const double spread = 0.5 * (dmx - best);
    const bool ladder = (bi != 0x7fffffff) && (spread > 0.0) && (spread < INFINITY);   // false on NaN
    if (ladder) {
#pragma unroll
        for (int j = 0; j < kRL; ++j) {
            const double thr = (best + 2.0 * (spread * pow2neg(j + 1))) * (1.0 + 1e-11);
            int cj = 0;
#pragma unroll
            for (int k = 0; k < kRPts; ++k) cj += (dl[k] <= thr) ? 1 : 0;
            cj = __reduce_add_sync(0xffffffffu, cj);
            if (lane == 0) wcnt[hw * kRL + j] = cj;
        }
    }
    group_sync(grp);
    int js = -1, total = 0, base = 0;
    if (ladder) {
        for (int j = 0; j < kRL && js < 0; ++j) {
            int tot = 0, bs = 0;
#pragma unroll
            for (int k = 0; k < kRT / 32; ++k) {
                const int v = wcnt[k * kRL + j];
                if (k < hw) bs += v;
                tot += v;
            }
            if (tot <= kRC) { js = j; total = tot; base = bs; }
        }
    }
    if (js >= 0) {
        const double delta = spread * pow2neg(js + 1);
        const double thr = (best + 2.0 * delta) * (1.0 + 1e-11);
        int running = 0;
#pragma unroll
        for (int k = 0; k < kRPts; ++k) {
            const bool pass = dl[k] <= thr;
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            const int slot = base + running + __popc(bal & ((1u << lane) - 1u));
            if (pass && slot < kRC) {
                const int p = ht + k * kRT;
                cidx[slot] = p;
                for (int j = 0; j < r; ++j) cc[j * kRC + slot] = bankT[(size_t)j * P + p];
            }
            running += __popc(bal);
        }
        if (ht < r) c0[ht] = xc[xoff + ht];
        if (ht == 0) { scal[0] = total; *deltap = delta; scal[3] = 1; }
    } else if (ht == 0) {
        scal[3] = 0;
    }
    return sel;
}

template <int RT, int CM>
__global__ void __launch_bounds__(kRThreads, 1)
tpwl_rollout_nn_resident_kernel(TpwlDev M, long long batch, int N, const double* __restrict__ x0,
                                const double* __restrict__ u, double* __restrict__ xo, int* __restrict__ idxo, int useq) {
    constexpr int r = RT, n = 2 * RT, m = CM;
    constexpr int kStepThreads = kRT - 32, kRowGroups = kStepThreads / 4, kRows = n / kRowGroups;
    static_assert(n % 8 == 0 && n % kRowGroups == 0, "affine step layout: 16-byte loads, whole rows per lane group");
    static_assert(kStepThreads >= n && m <= 32, "output / input rows are moved by warps 1-3");
    constexpr ResidentPlan S = make_resident_plan(n, m, r);
    constexpr int mpad = (int)(rp_up(sizeof(double) * m, 16) / sizeof(double));
    extern __shared__ __align__(128) unsigned char smraw[];
    const int tid = threadIdx.x, grp = tid / kRT, ht = tid - grp * kRT, lane = ht & 31, hw = ht >> 5;
    unsigned char* gb = smraw + (size_t)grp * S.stride;
    double* sA = reinterpret_cast<double*>(gb + S.A);
    double* sB = reinterpret_cast<double*>(gb + S.B);
    double* sd = reinterpret_cast<double*>(gb + S.d);
    double* cc = reinterpret_cast<double*>(gb + S.cc);          // r x kRC: coordinates of the candidates
    double* c0 = reinterpret_cast<double*>(gb + S.c0);          // sub-state of the last refresh
    double* sxb = reinterpret_cast<double*>(gb + S.x);          // two state buffers
    double* sub = reinterpret_cast<double*>(gb + S.u);          // two input buffers
    int* cidx = reinterpret_cast<int*>(gb + S.cidx);
    int* scal = reinterpret_cast<int*>(gb + S.scal);            // [0] candidates, [1] refresh flag, [2] selection, [3] list valid
    double* deltap = reinterpret_cast<double*>(gb + S.delta);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(gb + S.mbar);
    const double w = useq ? M.wq : M.wv;
    const int xoff = useq ? r : 0;                              // x = [v; q]
    constexpr unsigned kEntryBytes = (unsigned)(sizeof(double) * (n * n + n * m + n));

    if (ht == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    group_sync(grp);
    unsigned ph = 0;

    // affine step of warps 1-3: four lanes per row, rows rg, rg + 24, rg + 48 (summation order of tpwl_screen.cu)
    const int mt = ht - 32, part = mt & 3, rg = mt >> 2;
    auto affine = [&](const double* __restrict__ xc, const double* __restrict__ uc, double (&y)[kRows]) {
        double ax[kRows], bu[kRows];
#pragma unroll
        for (int k = 0; k < kRows; ++k) { ax[k] = 0.0; bu[k] = 0.0; }
#pragma unroll
        for (int s2 = 0; s2 < n / 8; ++s2) {
            const double2 xv = *reinterpret_cast<const double2*>(xc + 8 * s2 + 2 * part);
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                const double2 v = *reinterpret_cast<const double2*>(sA + (rg + kRowGroups * k) * n + 8 * s2 + 2 * part);
                ax[k] = fma(v.x, xv.x, ax[k]);
                ax[k] = fma(v.y, xv.y, ax[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
            const int row = rg + kRowGroups * k;
            for (int kk = part; kk < m; kk += 4) bu[k] = fma(sB[row * m + kk], uc[kk], bu[k]);
            ax[k] += __shfl_xor_sync(0xffffffffu, ax[k], 1);  bu[k] += __shfl_xor_sync(0xffffffffu, bu[k], 1);
            ax[k] += __shfl_xor_sync(0xffffffffu, ax[k], 2);  bu[k] += __shfl_xor_sync(0xffffffffu, bu[k], 2);
            y[k] = __dadd_rn(__dadd_rn(ax[k], bu[k]), sd[row]);
        }
    };

    for (long long b = blockIdx.x + (long long)gridDim.x * grp; b < batch; b += (long long)gridDim.x * kRG) {
        for (int i = ht; i < n; i += kRT) sxb[i] = x0[b * n + i];
        if (ht < m) sub[ht] = (N > 0) ? u[b * (long long)N * m + ht] : 0.0;
        if (ht == 0) scal[3] = 0;
        int cur = -1;
        group_sync(grp);
        const long long ts0 = clock64();
        for (int t = 0; t < N; ++t) {
            const double* xc = sxb + (t & 1) * n;
            double* xnx = sxb + ((t + 1) & 1) * n;
            const double* uc = sub + (t & 1) * mpad;
            double y[kRows];
            if (hw == 0) {
                // ---- selection from the candidate list (warp 0)
                int need = 1, sl = 0;
                if (scal[3]) {
                    double s2 = 0.0;
                    for (int j = lane; j < r; j += 32) {
                        const double dd = xc[xoff + j] - c0[j];
                        s2 = fma(dd, dd, s2);
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                    const double mv = w * sqrt(s2) * (1.0 + 1e-9);
                    if (mv <= *deltap) {
                        const int c = scal[0];
                        if (c == 1) {
                            need = 0;
                            sl = cidx[0];
                        } else {
                            double best = INFINITY;
                            int bi = 0x7fffffff;
                            if (lane < c) {
                                double ss;
                                if constexpr (RT == 36) ss = np_pairwise_sumsq_36(cc, kRC, lane, xc + xoff);
                                else                    ss = np_pairwise_sumsq(cc, kRC, lane, xc + xoff, 0, r);
                                const double dd = __dmul_rn(w, sqrt(ss));      // + 0.0 of the unused weight: same bits
                                if (dd < best) { best = dd; bi = cidx[lane]; }
                            }
                            warp_argmin(best, bi);
                            if (bi != 0x7fffffff) { need = 0; sl = bi; }
                        }
                    }
                }
                if (lane == 0) { scal[1] = need; scal[2] = sl; }
            } else {
                // ---- x_t leaves, u_{t+1} arrives, and the step with the resident entry is evaluated on speculation
                if (mt < n) xo[(b * (long long)(N + 1) + t) * n + mt] = xc[mt];
                if (mt < m && t + 1 < N) sub[((t + 1) & 1) * mpad + mt] = u[(b * (long long)N + t + 1) * m + mt];
                if (cur >= 0) affine(xc, uc, y);
            }
            group_sync(grp);
            int sel;
            if (scal[1]) {
                const long long tr0 = clock64();
                sel = resident_refresh<RT, CM>(M, xc, gb, grp, ht, useq);
                RS_ADD(0, 1); RS_ADD(1, clock64() - tr0);
            } else {
                sel = scal[2];
            }
            if (ht == 0 && idxo) idxo[b * (long long)N + t] = sel;
            if (sel != cur) {
                // ---- the entry of the new point replaces the resident one (every reader of the old one is past the barrier)
                if (ht == 0) {
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    mbar_expect_tx(mbar, kEntryBytes);
                    tma_bulk_g2s(sA, M.A + (size_t)sel * n * n, (unsigned)(sizeof(double) * n * n), mbar);
                    tma_bulk_g2s(sB, M.B + (size_t)sel * n * m, (unsigned)(sizeof(double) * n * m), mbar);
                    tma_bulk_g2s(sd, M.d + (size_t)sel * n, (unsigned)(sizeof(double) * n), mbar);
                }
                const long long tl0 = clock64();
                mbar_wait(mbar, ph);
                ph ^= 1u;
                cur = sel;
                RS_ADD(2, 1); RS_ADD(3, clock64() - tl0);
                if (hw != 0) affine(xc, uc, y);
            }
            if (hw != 0 && part == 0) {
#pragma unroll
                for (int k = 0; k < kRows; ++k) xnx[rg + kRowGroups * k] = y[k];
            }
            group_sync(grp);
        }
        RS_ADD(4, N); RS_ADD(5, clock64() - ts0);
        {
            const double* xc = sxb + (N & 1) * n;
            for (int i = ht; i < n; i += kRT) xo[(b * (long long)(N + 1) + N) * n + i] = xc[i];
        }
        group_sync(grp);
    }
}

// Dispatch: nn rollout on a bank that needs no discretisation at the Diamond shape (r = 36, m = 4), exactly one
// positive distance weight, P <= 1024 -- and only on request (SRCB200_TPWL_RESIDENT=1).  The kernel pays when
// trajectories move slowly against the spacing of the stored points (candidate lists that live for many steps); on
// the config-2 benchmark bank (isotropic 36-dimensional Gaussian cloud: the 32nd neighbour is 20 % farther than the
// nearest, the state moves 3 % of that distance per step) a list lives 2.6 steps and every refresh streams the 288 KB
// float64 point bank from L2 -- more than the 44 KB gather it saves: 12.5 ms against tpwl_screen.cu's 3.6 ms
// (profiles/experiments/README.md).  The default therefore stays the screened kernel.
int tpwl_rollout_nn_resident_launch(const TpwlDev& M, long long batch, int N, const double* x0, const double* u,
                                    double* x, int* idx, cudaStream_t st, bool* handled) {
    *handled = false;
    const char* env = getenv("SRCB200_TPWL_RESIDENT");
    if (!(env && env[0] == '1')) return 0;
    env = getenv("SRCB200_TPWL_NOSCREEN");
    if (env && env[0] == '1') return 0;
    const bool useq = M.wq != 0.0, usev = M.wv != 0.0;
    if (useq == usev) return 0;                                    // both or none
    if (!((useq ? M.wq : M.wv) > 0.0)) return 0;
    if (M.r != 36 || M.m != 4 || M.n != 72 || M.P < 1 || M.P > kRPts * kRT) return 0;
    if ((reinterpret_cast<uintptr_t>(M.A) | reinterpret_cast<uintptr_t>(M.B) | reinterpret_cast<uintptr_t>(M.d)) & 15)
        return 0;                                                  // TMA bulk copies need 16-byte aligned sources
    if (batch < 1) { *handled = true; return 0; }
    constexpr ResidentPlan S = make_resident_plan(72, 4, 36);
    constexpr size_t smem = S.stride * kRG;
    static_assert(smem <= 227 * 1024, "four resident entries per SM");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(batch < sms ? batch : sms);
    SRCB_CUDA(cudaFuncSetAttribute(tpwl_rollout_nn_resident_kernel<36, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tpwl_rollout_nn_resident_kernel<36, 4><<<grid, kRThreads, smem, st>>>(M, batch, N, x0, u, x, idx, useq ? 1 : 0);
    SRCB_LAUNCH_CHECK("tpwl_rollout_nn_resident_kernel");
#ifdef SRCB_RES_STATS
    {
        unsigned long long h[8];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_res_stats, sizeof(h));
        fprintf(stderr, "resident stats: steps %llu (%.0f cycles each), refreshes %llu (%.0f cycles each), entry loads %llu (%.0f cycles wait each)\n",
                h[4], h[4] ? (double)h[5] / h[4] : 0.0, h[0], h[0] ? (double)h[1] / h[0] : 0.0, h[2], h[2] ? (double)h[3] / h[2] : 0.0);
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_res_stats, h, sizeof(h));
    }
#endif
    *handled = true;
    return 0;
}

}  // namespace srcb
