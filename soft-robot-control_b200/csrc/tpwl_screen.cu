// tpwl_screen.cu -- TPWL nearest-neighbour rollout (sofacontrol/tpwl/tpwl.py:115-126, 160-168, 193-234) on a bank
// that needs no per-step discretisation, with the EXACT two-stage point search:
//
//   stage 1  FP32 distances d^ to all P stored points from an FP32 copy of the point bank that stays in shared
//            memory for the whole kernel (144 KB at the Diamond size), for 8 trajectories at once (every bank value
//            is read once per step and used 8 times), with the rigorous error bound
//                eps_p = 2^-24 (34 d^_p + 2 (max_p ||Q_p|| + ||q||)) + 1e-14 d^_p
//            (derivation: ilqr_fwd_tpwl.cuh); squared distances are compared, no square root per point;
//   stage 2  candidates {p : d^_p - eps_p <= min_p' (d^_p' + eps_p')} -- the FP64 argmin is provably among them --
//            are re-scored with the bit-exact numpy-order FP64 distance (tpwl.cuh) by one warp per trajectory;
//            first-occurrence argmin.  No candidate / more than 32 candidates / NaN: full FP64 search for that
//            trajectory.
//
// The selected indices are identical to the full search's (tests/test_tpwl_gpu.py compares the index trace with
// numpy).  The un-screened kernel spent its time in 3 P r un-fused FP64 operations per step; this one does 2 P r
// FP32 operations from shared memory and ~2 exact evaluations.
#include <cstdlib>
#include "tpwl.cuh"

namespace srcb {

constexpr int kST = 8;              // trajectories per CTA
constexpr int kSThreads = 512;
constexpr int kSWarps = kSThreads / 32;
constexpr int kSCand = 32;          // candidate slots per trajectory (one warp re-scores them)

struct ScreenPlan {
    size_t bank, sx, sxn, su, xfT, red, redf, thr2, xnorm, cnt, cand, sel, fb, red_i, total;
};
__host__ __device__ inline ScreenPlan make_screen_plan(int n, int m, int r, int P) {
    ScreenPlan S;
    size_t o = 0;
    auto take = [&o](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
    S.bank = take(sizeof(float) * (size_t)r * P);
    S.sx = take(sizeof(double) * kST * n);
    S.sxn = take(sizeof(double) * kST * n);
    S.su = take(sizeof(double) * kST * m);
    S.xfT = take(sizeof(float) * (size_t)r * kST);
    S.red = take(sizeof(double) * kST * kSWarps);
    S.redf = take(sizeof(float) * kST * kSWarps);
    S.thr2 = take(sizeof(float) * kST);
    S.xnorm = take(sizeof(double) * (kST + 2));         // + [kST]: bank norm bound
    S.cnt = take(sizeof(int) * kST);
    S.cand = take(sizeof(int) * kST * kSCand);
    S.sel = take(sizeof(int) * kST);
    S.fb = take(sizeof(int) * kST);
    S.red_i = take(sizeof(int) * kSWarps);
    S.total = o;
    return S;
}

template <int RT, int CM>
__global__ void __launch_bounds__(kSThreads, 1)
tpwl_rollout_nn_screen_kernel(TpwlDev M, long long batch, int N, const double* __restrict__ x0,
                              const double* __restrict__ u, double* __restrict__ xo, int* __restrict__ idxo, int useq, int tpg) {
    // tpg <= kST: trajectories per group, chosen by the launcher so that every CTA gets the same number of groups
    extern __shared__ __align__(16) unsigned char smraw[];
    // RT / CM > 0: r (and n = 2r) / m are compile-time constants (Diamond: 36, 4): unrolled loops, constant divisions
    const int r = RT > 0 ? RT : M.r, n = RT > 0 ? 2 * RT : M.n, m = CM > 0 ? CM : M.m, P = M.P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ScreenPlan S = make_screen_plan(n, m, r, P);
    float* bank = reinterpret_cast<float*>(smraw + S.bank);
    double* sx = reinterpret_cast<double*>(smraw + S.sx);
    double* sxn = reinterpret_cast<double*>(smraw + S.sxn);
    double* su = reinterpret_cast<double*>(smraw + S.su);
    float* xfT = reinterpret_cast<float*>(smraw + S.xfT);          // r x kST: the screened half of the 8 states
    double* red = reinterpret_cast<double*>(smraw + S.red);
    float* redf = reinterpret_cast<float*>(smraw + S.redf);
    float* thr2 = reinterpret_cast<float*>(smraw + S.thr2);
    double* xnorm = reinterpret_cast<double*>(smraw + S.xnorm);
    int* cnt = reinterpret_cast<int*>(smraw + S.cnt);
    int* cand = reinterpret_cast<int*>(smraw + S.cand);
    int* sel = reinterpret_cast<int*>(smraw + S.sel);
    int* fb = reinterpret_cast<int*>(smraw + S.fb);
    int* red_i = reinterpret_cast<int*>(smraw + S.red_i);
    const double w = useq ? M.wq : M.wv;
    const double* bankT = useq ? M.qT : M.vT;
    const int xoff = useq ? r : 0;                                  // x = [v; q]

    // ---- FP32 bank + the largest weighted point norm, once per CTA
    {
        double bmax = 0.0;
        for (int p = tid; p < P; p += kSThreads) {
            double sq = 0.0;
            for (int j = 0; j < r; ++j) {
                const double v = bankT[(size_t)j * P + p];
                bank[j * P + p] = (float)v;
                sq = fma(v, v, sq);
            }
            bmax = fmax(bmax, w * sqrt(sq));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, off));
        if (lane == 0) red[warp] = bmax;
        __syncthreads();
        if (tid == 0) {
            double b2 = 0.0;
            for (int k = 0; k < kSWarps; ++k) b2 = fmax(b2, red[k]);
            xnorm[kST] = b2 * (1.0 + 1e-6);
        }
        __syncthreads();
    }
    const double bank_norm = xnorm[kST];
    const bool screen_ok = isfinite(bank_norm) && bank_norm < 1e100;
    const double u24 = 5.9604644775390625e-08;                     // 2^-24
    const int p0 = tid, p1 = tid + kSThreads;
    const bool has0 = p0 < P, has1 = p1 < P;

    const long long groups = (batch + tpg - 1) / tpg;
    for (long long gidx = blockIdx.x; gidx < groups; gidx += gridDim.x) {
        const long long b0 = gidx * tpg;
        const int nt = (int)((batch - b0) < tpg ? (batch - b0) : tpg);
        __syncthreads();
        for (int e = tid; e < kST * n; e += kSThreads) {
            const int tr = e / n, i = e - tr * n;
            const double v = (tr < nt) ? x0[(b0 + tr) * n + i] : 0.0;
            sx[e] = v;
            sxn[e] = 0.0;                   // rows of unused trajectory slots stay zero
            if (tr < nt) xo[(b0 + tr) * (long long)(N + 1) * n + i] = v;
        }
        __syncthreads();
        for (int t = 0; t < N; ++t) {
            // ---- per-step inputs; FP32 copy and norm of the screened half of every state (warp tr)
            for (int e = tid; e < kST * m; e += kSThreads) {
                const int tr = e / m, i = e - tr * m;
                su[e] = (tr < nt) ? u[((b0 + tr) * (long long)N + t) * m + i] : 0.0;
            }
            if (warp < kST) {
                double s2 = 0.0;
                for (int j = lane; j < r; j += 32) {
                    const double v = sx[warp * n + xoff + j];
                    xfT[j * kST + warp] = (float)v;
                    s2 = fma(v, v, s2);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                if (lane == 0) { xnorm[warp] = sqrt(s2) * (1.0 + 1e-9); cnt[warp] = 0; fb[warp] = 0; }
            }
            __syncthreads();
            if (screen_ok) {
                // ---- stage 1: FP32 distances of this thread's two points to the 8 states
                float a0[kST], a1[kST];
#pragma unroll
                for (int tr = 0; tr < kST; ++tr) { a0[tr] = 0.f; a1[tr] = 0.f; }
                const float* bp0 = bank + (has0 ? p0 : 0);
                const float* bp1 = bank + (has1 ? p1 : 0);
#pragma unroll 4
                for (int j = 0; j < r; ++j) {
                    const float q0 = bp0[j * P], q1 = bp1[j * P];
                    const float4 xa = *reinterpret_cast<const float4*>(xfT + j * kST);
                    const float4 xb = *reinterpret_cast<const float4*>(xfT + j * kST + 4);
                    const float xs[kST] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                    for (int tr = 0; tr < kST; ++tr) {
                        const float d0 = q0 - xs[tr], d1 = q1 - xs[tr];
                        a0[tr] = fmaf(d0, d0, a0[tr]);
                        a1[tr] = fmaf(d1, d1, a1[tr]);
                    }
                }
                // With eps(d) = c1 d + c0, c1 = 34 * 2^-24 + 1e-14, c0 = 2^-24 * 2 (bank norm + w |x|): the smallest upper
                // bound is (1 + c1) w sqrt(min a) + c0 =: U, and "lower bound <= U" is  a <= ((U + c0) / (w (1 - c1)))^2.
                // So the per-point work is a float minimum and one float compare; the bound arithmetic runs once per
                // trajectory in double and the threshold is rounded UP (more candidates, never fewer).
                float amin[kST];
#pragma unroll
                for (int tr = 0; tr < kST; ++tr) {
                    float v = INFINITY;
                    if (has0) v = a0[tr];
                    if (has1) v = fminf(v, a1[tr]);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, off));
                    amin[tr] = v;
                }
                if (lane == 0) {
#pragma unroll
                    for (int tr = 0; tr < kST; ++tr) redf[tr * kSWarps + warp] = amin[tr];
                }
                __syncthreads();
                if (warp < kST) {
                    float v = (lane < kSWarps) ? redf[warp * kSWarps + lane] : INFINITY;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, off));
                    if (lane == 0) {
                        const double c1 = 34.0 * u24 + 1e-14;
                        const double c0 = u24 * 2.0 * (bank_norm + w * xnorm[warp]);
                        const double U = (1.0 + c1) * w * sqrt((double)v) + c0;
                        const double T = (U + c0) / (w * (1.0 - c1));
                        const double T2 = T * T * (1.0 + 1e-6);
                        thr2[warp] = (T2 < 3.0e38) ? __double2float_ru(T2) : INFINITY;   // NaN compares false below: fallback
                    }
                }
                __syncthreads();
                // ---- candidates
#pragma unroll
                for (int tr = 0; tr < kST; ++tr) {
                    const float th = thr2[tr];
                    if (has0 && a0[tr] <= th) {
                        const int pos = atomicAdd(&cnt[tr], 1);
                        if (pos < kSCand) cand[tr * kSCand + pos] = p0;
                    }
                    if (has1 && a1[tr] <= th) {
                        const int pos = atomicAdd(&cnt[tr], 1);
                        if (pos < kSCand) cand[tr * kSCand + pos] = p1;
                    }
                }
                __syncthreads();
                // ---- stage 2: warp tr re-scores its candidates with the exact FP64 distance
                if (warp < kST) {
                    const int c = cnt[warp];
                    if (c >= 1 && c <= kSCand) {
                        double best = INFINITY;
                        int bi = 0x7fffffff;
                        if (lane < c) {
                            const int p = cand[warp * kSCand + lane];
                            const double dd = tpwl_distance(M, sx + warp * n, p);
                            if (dd < best) { best = dd; bi = p; }
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            const double od = __shfl_xor_sync(0xffffffffu, best, off);
                            const int op = __shfl_xor_sync(0xffffffffu, bi, off);
                            if (od < best || (od == best && op < bi)) { best = od; bi = op; }
                        }
                        if (lane == 0) {
                            if (bi == 0x7fffffff) fb[warp] = 1; else sel[warp] = bi;
                        }
                    } else if (lane == 0) {
                        fb[warp] = 1;
                    }
                }
                __syncthreads();
            } else {
                if (tid < kST) fb[tid] = 1;
                __syncthreads();
            }
            // ---- anything unusual: full FP64 search for that trajectory (uniform branch on shared flags)
            for (int tr = 0; tr < kST; ++tr) {
                if (fb[tr]) {
                    const int idx = tpwl_nearest<kSThreads>(M, sx + tr * n, nullptr, red, red_i, nullptr);
                    if (tid == 0) sel[tr] = idx;
                    __syncthreads();
                }
            }
            if (idxo && tid < nt) idxo[(b0 + tid) * (long long)N + t] = sel[tid];
            // ---- x+ = (A_i x + B_i u) + d_i (tpwl.py:231-234): four lanes per (trajectory, row)
            {
                const int part = tid & 3;
                for (int rb = 0; rb < tpg * n; rb += kSThreads / 4) {
                    const int row = rb + (tid >> 2);
                    const bool act = row < tpg * n;
                    double ax = 0.0, bu = 0.0;
                    int tr = 0, i = 0;
                    long long p = 0;
                    if (act) {
                        tr = row / n; i = row - tr * n;
                        p = sel[tr];
                        const double* Ap = M.A + (p * n + i) * n;
                        const double* Bp = M.B + (p * n + i) * m;
                        const double* xs = sx + tr * n;
                        for (int k = part; k < n; k += 4) ax = fma(Ap[k], xs[k], ax);
                        for (int k = part; k < m; k += 4) bu = fma(Bp[k], su[tr * m + k], bu);
                    }
                    ax += __shfl_xor_sync(0xffffffffu, ax, 1);
                    bu += __shfl_xor_sync(0xffffffffu, bu, 1);
                    ax += __shfl_xor_sync(0xffffffffu, ax, 2);
                    bu += __shfl_xor_sync(0xffffffffu, bu, 2);
                    if (act && part == 0) sxn[row] = __dadd_rn(__dadd_rn(ax, bu), M.d[p * n + i]);
                }
            }
            __syncthreads();
            for (int e = tid; e < kST * n; e += kSThreads) {
                const int tr = e / n, i = e - tr * n;
                const double v = sxn[e];
                sx[e] = v;
                if (tr < nt) xo[((b0 + tr) * (long long)(N + 1) + t + 1) * n + i] = v;
            }
            __syncthreads();
        }
    }
}

// Dispatch: nn rollout on a bank that needs no discretisation, exactly one non-negative distance weight, P <= 1024
// and an FP32 bank that fits in shared memory.
int tpwl_rollout_nn_screen_launch(const TpwlDev& M, long long batch, int N, const double* x0, const double* u,
                                  double* x, int* idx, cudaStream_t st, bool* handled) {
    *handled = false;
    const char* env = getenv("SRCB200_TPWL_NOSCREEN");
    if (env && env[0] == '1') return 0;
    const bool useq = M.wq != 0.0, usev = M.wv != 0.0;
    if (useq == usev) return 0;                                    // both or none
    if ((useq ? M.wq : M.wv) < 0.0) return 0;
    if (M.P > 2 * kSThreads || M.P < 1 || M.r > 128) return 0;
    const ScreenPlan S = make_screen_plan(M.n, M.m, M.r, M.P);
    if (S.total > 200 * 1024) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (batch < 1) { *handled = true; return 0; }
    // equal work per CTA: rounds = groups of kST per SM, then the smallest group size that still needs that many rounds
    const long long rounds = (batch + (long long)sms * kST - 1) / ((long long)sms * kST);
    int tpg = (int)((batch + sms * rounds - 1) / (sms * rounds));
    if (tpg > kST) tpg = kST;
    if (tpg < 1) tpg = 1;
    const long long groups = (batch + tpg - 1) / tpg;
    const int grid = (int)(groups < sms ? groups : sms);
    if (M.r == 36 && M.n == 72 && M.m == 4) {
        SRCB_CUDA(cudaFuncSetAttribute(tpwl_rollout_nn_screen_kernel<36, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
        tpwl_rollout_nn_screen_kernel<36, 4><<<grid, kSThreads, S.total, st>>>(M, batch, N, x0, u, x, idx, useq ? 1 : 0, tpg);
    } else {
        SRCB_CUDA(cudaFuncSetAttribute(tpwl_rollout_nn_screen_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
        tpwl_rollout_nn_screen_kernel<0, 0><<<grid, kSThreads, S.total, st>>>(M, batch, N, x0, u, x, idx, useq ? 1 : 0, tpg);
    }
    SRCB_LAUNCH_CHECK("tpwl_rollout_nn_screen_kernel");
    *handled = true;
    return 0;
}

}  // namespace srcb
