// tpwl_screen.cu -- TPWL nearest-neighbour rollout (sofacontrol/tpwl/tpwl.py:115-126, 160-168, 193-234) on a bank
// that needs no per-step discretisation, with the EXACT two-stage point search:
//
//   stage 1  FP32 squared distances a^ to all P stored points from an FP32 copy of the point bank that stays in
//            shared memory for the whole kernel (144 KB at the Diamond size), for up to 8 trajectories at once (every
//            bank value is read once per step and used 8 times).  Bank and state are CENTRED on the bank mean mu (in
//            float64, before the conversion) and the squared distance is expanded,
//                a^_p = (|b_p - mu|^2 + |c - mu|^2) - 2 (b_p - mu).(c - mu),
//            so the inner loop is ONE FMA per (point, coordinate, trajectory) instead of a subtraction and an FMA.
//            Rigorous bound (u = 2^-24; b, c the centred vectors): the float32 operands carry relative error u, the
//            FMA chain of length r adds gamma_r sum|b_j c_j| <= r u |b||c|, the two norms are rounded once each, the
//            final add and FMA once each:
//                |a^_p - |b_p - c|^2|  <=  u (|b| + |c|)^2 (r/2 + 4.5)  <=  E2 := (r/2 + 8) u (max_p|b_p| + |c|)^2 (1 + 1e-3)
//            (float64 roundings of the centring and of the reference's own distances are 1e-16-relative: inside the
//            slack).  The reference's float64 argmin p* has |b_p* - c|^2 <= min_p |b_p - c|^2 (1 + 1e-12), hence
//            a^_p* <= min a^ + 2 E2 (+ slack): one float compare per point against a per-trajectory threshold
//            (rounded UP);
//   stage 2  the FP64 argmin is provably among the candidates.  One candidate (the usual case: measured 701
//            candidates per 700 searches): done.  Several: one warp re-scores them with the bit-exact numpy-order
//            FP64 distance (tpwl.cuh), first-occurrence argmin.  None / more than 32 / NaN: that warp runs the full
//            FP64 search.
//
// The selected indices are identical to the full search's (tests/test_tpwl_gpu.py compares the index trace with
// numpy).  A CTA is two independent halves of 256 threads (named barriers), each rolling its own group of
// trajectories against the shared bank copy: the search is FP32-issue bound, the gathered affine step
// x+ = A_i x + B_i u + d_i is bound by the load path (44 KB gathered from L2 per trajectory-step; ncu: L1/LSU 73 %,
// L2 16 % of peak), and the halves drift so that one computes while the other gathers.
#include <cstdlib>
#include "tpwl.cuh"

namespace srcb {

constexpr int kST = 8;              // trajectory slots per half
constexpr int kSHalf = 256;         // threads per half
constexpr int kSThreads = 2 * kSHalf;
constexpr int kSHW = kSHalf / 32;   // warps per half (one per trajectory slot)
constexpr int kSPts = 4;            // stored points per thread (P <= kSPts * kSHalf)
constexpr int kSCand = 32;          // candidate slots per trajectory (one warp re-scores them)
static_assert(kSHW == kST, "one warp per trajectory slot");

struct ScreenPlan {                 // byte offsets; the per-half arrays exist twice, `hstride` bytes apart
    size_t bank, nb, mu, half, hstride, sx, sxn, su, xfT, redf, thr2, xnorm, ncf, cnt, cand, sel, fb, total;
};
__host__ __device__ inline ScreenPlan make_screen_plan(int n, int m, int r, int P) {
    ScreenPlan S;
    size_t o = 0;
    auto take = [&o](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
    S.bank = take(sizeof(float) * (size_t)r * P);
    S.nb = take(sizeof(float) * (size_t)P);
    S.mu = take(sizeof(double) * (size_t)r);
    S.half = o;
    o = 0;
    S.sx = take(sizeof(double) * kST * n);
    S.sxn = take(sizeof(double) * kST * n);
    S.su = take(sizeof(double) * kST * m);
    S.xnorm = take(sizeof(double) * kST);
    S.xfT = take(sizeof(float) * (size_t)r * kST);
    S.redf = take(sizeof(float) * kST * kSHW);
    S.thr2 = take(sizeof(float) * kST);
    S.ncf = take(sizeof(float) * kST);
    S.cnt = take(sizeof(int) * kST);
    S.cand = take(sizeof(int) * kST * kSCand);
    S.sel = take(sizeof(int) * kST);
    S.fb = take(sizeof(int) * kST);
    S.hstride = o;
    S.total = S.half + 2 * S.hstride + 64;
    return S;
}

#ifdef SRCB_NN_PHASES
__device__ unsigned long long g_nn_phase[8];   // cycles per phase of the time loop (thread 0 of half 0 of CTA 0), [7] = steps
#define NN_PH(i) do { if (tid == 0 && blockIdx.x == 0) { const long long tn_ = clock64(); atomicAdd(&g_nn_phase[i], (unsigned long long)(tn_ - tph_)); tph_ = tn_; } } while (0)
#else
#define NN_PH(i) do { } while (0)
#endif


__device__ __forceinline__ void half_sync(int half) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(kSHalf) : "memory");
}

template <int RT, int CM>
__global__ void __launch_bounds__(kSThreads, 1)
tpwl_rollout_nn_screen_kernel(TpwlDev M, long long batch, int N, const double* __restrict__ x0,
                              const double* __restrict__ u, double* __restrict__ xo, int* __restrict__ idxo, int useq,
                              int tpg, int stagger) {
    // RT / CM > 0: r (and n = 2r) / m are compile-time constants (Diamond: 36, 4): unrolled loops, constant divisions.
    // tpg <= kST: trajectories per group, chosen by the launcher so that every half gets the same number of groups.
    extern __shared__ __align__(16) unsigned char smraw[];
    const int r = RT > 0 ? RT : M.r, n = RT > 0 ? 2 * RT : M.n, m = CM > 0 ? CM : M.m, P = M.P;
    const int tid = threadIdx.x, half = tid / kSHalf, ht = tid - half * kSHalf, lane = ht & 31, hw = ht >> 5;
    const ScreenPlan S = make_screen_plan(n, m, r, P);
    float* bank = reinterpret_cast<float*>(smraw + S.bank);
    unsigned char* hb = smraw + S.half + (size_t)half * S.hstride;
    double* sx = reinterpret_cast<double*>(hb + S.sx);
    double* sxn = reinterpret_cast<double*>(hb + S.sxn);
    double* su = reinterpret_cast<double*>(hb + S.su);
    double* xnorm = reinterpret_cast<double*>(hb + S.xnorm);
    float* xfT = reinterpret_cast<float*>(hb + S.xfT);              // r x kST: the screened part of the 8 states
    float* redf = reinterpret_cast<float*>(hb + S.redf);
    float* thr2 = reinterpret_cast<float*>(hb + S.thr2);
    float* ncf = reinterpret_cast<float*>(hb + S.ncf);               // |c - mu|^2 of the 8 states
    float* nbs = reinterpret_cast<float*>(smraw + S.nb);             // |b_p - mu|^2
    double* smu = reinterpret_cast<double*>(smraw + S.mu);           // bank mean
    int* cnt = reinterpret_cast<int*>(hb + S.cnt);
    int* cand = reinterpret_cast<int*>(hb + S.cand);
    int* sel = reinterpret_cast<int*>(hb + S.sel);
    int* fb = reinterpret_cast<int*>(hb + S.fb);
    double* shared_tail = reinterpret_cast<double*>(smraw + S.half + 2 * S.hstride);     // [0]: bank norm bound
    const double* bankT = useq ? M.qT : M.vT;
    const int xoff = useq ? r : 0;                                  // x = [v; q]

    // ---- bank mean, centred FP32 bank, centred squared norms and the largest centred norm: once per CTA
    {
        double* redd = reinterpret_cast<double*>(smraw + S.half);   // scratch before the halves start using their arrays
        for (int j = tid; j < r; j += kSThreads) {
            double acc = 0.0;
            for (int p2 = 0; p2 < P; ++p2) acc += bankT[(size_t)j * P + p2];
            smu[j] = acc / (double)P;       // any centre is valid; every CTA computes the same one
        }
        __syncthreads();
        double bmax = 0.0;
        for (int p = tid; p < P; p += kSThreads) {
            double sq = 0.0;
            for (int j = 0; j < r; ++j) {
                const double v = bankT[(size_t)j * P + p] - smu[j];
                bank[j * P + p] = (float)v;
                sq = fma(v, v, sq);
            }
            nbs[p] = (float)sq;
            bmax = fmax(bmax, sqrt(sq));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, off));
        if ((tid & 31) == 0) redd[tid >> 5] = bmax;
        __syncthreads();
        if (tid == 0) {
            double b2 = 0.0;
            for (int k = 0; k < kSThreads / 32; ++k) b2 = fmax(b2, redd[k]);
            shared_tail[0] = b2 * (1.0 + 1e-6);
        }
        __syncthreads();
    }
    const double bank_norm = shared_tail[0];                       // max_p |b_p - mu|, rounded up
    const bool screen_ok = isfinite(bank_norm) && bank_norm < 1e17;
    const double u24 = 5.9604644775390625e-08;                     // 2^-24
    int pt[kSPts];
    bool has[kSPts];
    float nbf[kSPts];
#pragma unroll
    for (int k = 0; k < kSPts; ++k) { pt[k] = ht + k * kSHalf; has[k] = pt[k] < P; nbf[k] = has[k] ? nbs[pt[k]] : 0.f; }
    __syncthreads();        // the scratch is free again

    // Optional start offset of half 1 (SRCB200_NN_STAGGER cycles): the halves run the same phases with the same period,
    // and an offset would keep one searching while the other gathers -- measured: they do not collide measurably.
    if (half == 1 && stagger > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < stagger) { }
    }
    const long long groups = (batch + tpg - 1) / tpg;
    for (long long gidx = (long long)blockIdx.x * 2 + half; gidx < groups; gidx += (long long)gridDim.x * 2) {
        const long long b0 = gidx * tpg;
        const int nt = (int)((batch - b0) < tpg ? (batch - b0) : tpg);
        half_sync(half);
        for (int e = ht; e < kST * n; e += kSHalf) {
            const int tr = e / n, i = e - tr * n;
            const double v = (tr < nt) ? x0[(b0 + tr) * n + i] : 0.0;
            sx[e] = v;
            sxn[e] = 0.0;                   // rows of unused trajectory slots stay zero
            if (tr < nt) xo[(b0 + tr) * (long long)(N + 1) * n + i] = v;
        }
        for (int e = ht; e < kST * m; e += kSHalf) {
            const int tr = e / m, i = e - tr * m;
            su[e] = (tr < nt && N > 0) ? u[((b0 + tr) * (long long)N) * m + i] : 0.0;
        }
        half_sync(half);
#ifdef SRCB_NN_PHASES
        long long tph_ = clock64();
        if (tid == 0 && blockIdx.x == 0) atomicAdd(&g_nn_phase[7], (unsigned long long)N);
#endif
        // ---- FP32 copy and norm of the screened part of every state (warp hw <-> trajectory slot hw); for step t + 1 it
        //      runs in the state-update phase of step t, straight from the new state (one barrier and one pass less)
        auto prep_state = [&](const double* __restrict__ xs) {
            double s2 = 0.0;
            for (int j = lane; j < r; j += 32) {
                const double v = xs[hw * n + xoff + j] - smu[j];
                xfT[j * kST + hw] = (float)v;
                s2 = fma(v, v, s2);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
            if (lane == 0) {
                xnorm[hw] = sqrt(s2) * (1.0 + 1e-9);
                ncf[hw] = (float)s2;
                cnt[hw] = 0;
                fb[hw] = screen_ok ? 0 : 1;
            }
        };
        prep_state(sx);
        half_sync(half);
        for (int t = 0; t < N; ++t) {
            NN_PH(0);
            if (screen_ok) {
                // ---- stage 1: FP32 dot products of this thread's (centred) points with the 8 (centred) states
                // packed pairs: FFMA2 (fma.rn.f32x2, two IEEE FMAs per instruction, the point coordinate as the broadcast
                // operand) halves the instruction count of the loop the kernel issues most
                float a[kSPts][kST];
                unsigned long long a2[kSPts][kST / 2];
#pragma unroll
                for (int k = 0; k < kSPts; ++k)
#pragma unroll
                    for (int h = 0; h < kST / 2; ++h) a2[k][h] = 0ull;
                const float* bp[kSPts];
#pragma unroll
                for (int k = 0; k < kSPts; ++k) bp[k] = bank + (has[k] ? pt[k] : 0);
#pragma unroll 2
                for (int j = 0; j < r; ++j) {
                    const ulonglong2 xa = *reinterpret_cast<const ulonglong2*>(xfT + j * kST);
                    const ulonglong2 xb = *reinterpret_cast<const ulonglong2*>(xfT + j * kST + 4);
                    const unsigned long long xp[kST / 2] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
                    for (int k = 0; k < kSPts; ++k) {
                        const float q = bp[k][j * P];
                        unsigned long long qq;
                        asm("mov.b64 %0, {%1, %1};" : "=l"(qq) : "f"(q));
#pragma unroll
                        for (int h = 0; h < kST / 2; ++h)
                            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[k][h]) : "l"(qq), "l"(xp[h]));
                    }
                }
#pragma unroll
                for (int k = 0; k < kSPts; ++k)
#pragma unroll
                    for (int h = 0; h < kST / 2; ++h) {
                        float lo, hi;
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a2[k][h]));
                        a[k][2 * h] = lo;
                        a[k][2 * h + 1] = hi;
                    }
                // a^ = (|b|^2 + |c|^2) - 2 b.c
                {
                    const float4 na = *reinterpret_cast<const float4*>(ncf);
                    const float4 nb4 = *reinterpret_cast<const float4*>(ncf + 4);
                    const float ncs[kST] = {na.x, na.y, na.z, na.w, nb4.x, nb4.y, nb4.z, nb4.w};
#pragma unroll
                    for (int k = 0; k < kSPts; ++k)
#pragma unroll
                        for (int tr = 0; tr < kST; ++tr) a[k][tr] = fmaf(-2.f, a[k][tr], nbf[k] + ncs[tr]);
                }
                // ---- per-trajectory minimum of a^ -> threshold (one warp per trajectory does the bound arithmetic)
#pragma unroll
                for (int tr = 0; tr < kST; ++tr) {
                    float v = INFINITY;
#pragma unroll
                    for (int k = 0; k < kSPts; ++k) if (has[k]) v = fminf(v, a[k][tr]);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, off));
                    if (lane == 0) redf[tr * kSHW + hw] = v;
                }
                half_sync(half);
            NN_PH(1);
                {
                    float v = (lane < kSHW) ? redf[hw * kSHW + lane] : INFINITY;
#pragma unroll
                    for (int off = 4; off > 0; off >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, off));
                    if (lane == 0) {
                        const double c2 = (0.5 * r + 8.0) * u24 * 1.001 + 1e-12;
                        const double sn = bank_norm + xnorm[hw];
                        const double E2 = c2 * sn * sn + 1e-36;           // + underflow of the float32 products
                        const double am = (double)v;                      // smallest a^ (may be slightly negative)
                        const double T2 = am + 2.0 * E2 + 1e-6 * (fabs(am) + 2.0 * E2);
                        // not representable / NaN: no candidate -> this trajectory runs the full float64 search
                        thr2[hw] = (sn < 1e17 && T2 < 3.0e38) ? __double2float_ru(T2) : __int_as_float(0x7fc00000);
                    }
                }
                half_sync(half);
            NN_PH(2);
                // ---- candidates
                // (32 compares folded into one hit mask without branches; a thread owns a candidate once in ~250 steps,
                // so the atomics sit behind one rarely taken branch instead of 32 branch regions)
                {
                    const float4 t0 = *reinterpret_cast<const float4*>(thr2);
                    const float4 t1 = *reinterpret_cast<const float4*>(thr2 + 4);
                    const float ths[kST] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                    unsigned hits = 0u;
#pragma unroll
                    for (int tr = 0; tr < kST; ++tr)
#pragma unroll
                        for (int k = 0; k < kSPts; ++k)
                            hits |= (has[k] && a[k][tr] <= ths[tr]) ? (1u << (tr * kSPts + k)) : 0u;
                    while (hits) {
                        const int bit = __ffs(hits) - 1;
                        hits &= hits - 1u;
                        const int tr = bit / kSPts, k = bit - tr * kSPts;
                        const int pos = atomicAdd(&cnt[tr], 1);
                        if (pos < kSCand) cand[tr * kSCand + pos] = ht + k * kSHalf;       // == pt[k]
                    }
                }
                half_sync(half);
            NN_PH(3);
                // ---- stage 2 (warp hw <-> trajectory hw)
                const int c = cnt[hw];
                if (c == 1) {
                    if (lane == 0) sel[hw] = cand[hw * kSCand];      // the argmin is among the candidates: it is this one
                } else if (c > 1 && c <= kSCand) {
                    double best = INFINITY;
                    int bi = 0x7fffffff;
                    if (lane < c) {
                        const int p = cand[hw * kSCand + lane];
                        const double dd = tpwl_distance(M, sx + hw * n, p);
                        if (dd < best) { best = dd; bi = p; }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const double od = __shfl_xor_sync(0xffffffffu, best, off);
                        const int op = __shfl_xor_sync(0xffffffffu, bi, off);
                        if (od < best || (od == best && op < bi)) { best = od; bi = op; }
                    }
                    if (lane == 0) { if (bi == 0x7fffffff) fb[hw] = 1; else sel[hw] = bi; }
                } else if (lane == 0) {
                    fb[hw] = 1;
                }
                __syncwarp();
            }
            // ---- anything unusual: this warp runs the full FP64 search for its trajectory (np.argmin semantics)
            if (fb[hw]) {
                double best = INFINITY;
                int bi = 0x7fffffff;
                for (int p = lane; p < P; p += 32) {
                    const double dd = tpwl_distance(M, sx + hw * n, p);
                    if (dd < best) { best = dd; bi = p; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, best, off);
                    const int op = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (od < best || (od == best && op < bi)) { best = od; bi = op; }
                }
                if (lane == 0) sel[hw] = (bi == 0x7fffffff) ? 0 : bi;
            }
            half_sync(half);
            NN_PH(4);
            if (idxo && ht < nt) idxo[(b0 + ht) * (long long)N + t] = sel[ht];
            // ---- x+ = (A_i x + B_i u) + d_i (tpwl.py:231-234): four lanes per (trajectory, row), two rows per lane
            //      group and pass with all loads of both rows issued first; u of the next step is fetched meanwhile
            double u_next = 0.0;
            if (ht < kST * m && t + 1 < N) {
                const int tr = ht / m, i = ht - tr * m;
                if (tr < nt) u_next = u[((b0 + tr) * (long long)N + t + 1) * m + i];
            }
            if constexpr (RT == 36 && CM == 4) {
                // Warp hw gathers the entry of trajectory hw.  Two consecutive rows of A_i are 1152 contiguous bytes
                // = nine full 128-byte lines: sixteen lanes sweep such a ROW PAIR with five 16-byte loads each (the
                // fifth by eight lanes), so a warp-wide load touches four full lines (the row-per-four-lanes layout
                // touched eight half lines), and because all pairs of a lane belong to ONE trajectory the lane's ten
                // state values stay in registers for the whole step.  The memory-instruction queue of the SM is the
                // limiter of this phase (L1 74 % busy; shared-memory loads and shuffles wait behind the global loads).
                // A lane's elements of a pair fall into row A (offset < 72) or row B; B_i u rides in the same sums; the
                // six totals of a pass (three pairs) are reduced over the sixteen lanes by an eight-shuffle butterfly
                // that leaves total 2 q + r on the lanes whose bits 3..1 spell that index; the even lanes store.
                constexpr int PPT = n / 2, GP = 3, PASSES = PPT / 2 / GP;        // 18 pairs per half-warp, 6 passes
                static_assert(PASSES * GP * 2 == PPT, "pairs split evenly");
                const int l16 = lane & 15, hsel = lane >> 4;
                const bool gact = hw < tpg;                                      // warp-uniform
                const double* xsrc = sx + hw * n;
                const bool toA2 = l16 < 4;
                double2 xr[5];
                xr[0] = *reinterpret_cast<const double2*>(xsrc + 2 * l16);
                xr[1] = *reinterpret_cast<const double2*>(xsrc + 32 + 2 * l16);
                xr[2] = *reinterpret_cast<const double2*>(xsrc + (toA2 ? 64 + 2 * l16 : 2 * l16 - 8));
                xr[3] = *reinterpret_cast<const double2*>(xsrc + 24 + 2 * l16);
                xr[4] = (l16 < 8) ? *reinterpret_cast<const double2*>(xsrc + 56 + 2 * l16) : make_double2(0.0, 0.0);
                const double2 uv = (l16 < 4) ? *reinterpret_cast<const double2*>(su + hw * m + ((2 * l16) & 3)) : make_double2(0.0, 0.0);
                const long long p = sel[hw];
                const double* Ab = M.A + p * n * n + (long long)hsel * (PPT / 2) * 2 * n;
                const double* Bb = M.B + p * n * m + (long long)hsel * (PPT / 2) * 2 * m;
                const double* db = M.d + p * n + hsel * (PPT / 2) * 2;
                const int own = ((l16 >> 3) & 1) * 4 + ((l16 >> 2) & 1) * 2 + ((l16 >> 1) & 1);   // value index 2 q + r
                const bool sown = gact && own < 2 * GP && !(l16 & 1);
#pragma unroll 1
                for (int ps = 0; ps < PASSES; ++ps) {
                    double2 v[GP][5], bv[GP];
#pragma unroll
                    for (int q = 0; q < GP; ++q) {
                        const double* Ap = Ab + (long long)(ps * GP + q) * 2 * n;
#pragma unroll
                        for (int c = 0; c < 5; ++c)
                            v[q][c] = (gact && (c < 4 || l16 < 8)) ? __ldcg(reinterpret_cast<const double2*>(Ap + 32 * c + 2 * l16))
                                                                   : make_double2(0.0, 0.0);
                        bv[q] = (gact && l16 < 4) ? __ldcg(reinterpret_cast<const double2*>(Bb + (ps * GP + q) * 2 * m + 2 * l16))
                                                  : make_double2(0.0, 0.0);
                    }
                    const int rloc = (ps * GP + (own >> 1)) * 2 + (own & 1);     // row (within this half-warp's 36) the lane stores
                    const double down = sown ? __ldcg(db + rloc) : 0.0;
                    double val[8];
#pragma unroll
                    for (int q = 0; q < GP; ++q) {
                        double accA = 0.0, accB = 0.0;
                        accA = fma(v[q][0].x, xr[0].x, accA); accA = fma(v[q][0].y, xr[0].y, accA);
                        accA = fma(v[q][1].x, xr[1].x, accA); accA = fma(v[q][1].y, xr[1].y, accA);
                        {
                            double t = toA2 ? accA : accB;
                            t = fma(v[q][2].x, xr[2].x, t); t = fma(v[q][2].y, xr[2].y, t);
                            if (toA2) accA = t; else accB = t;
                        }
                        accB = fma(v[q][3].x, xr[3].x, accB); accB = fma(v[q][3].y, xr[3].y, accB);
                        if (l16 < 8) { accB = fma(v[q][4].x, xr[4].x, accB); accB = fma(v[q][4].y, xr[4].y, accB); }
                        if (l16 < 4) {   // B_i u: elements 2 l16, 2 l16 + 1 of the pair's 2 x 4 block
                            double t = (l16 < 2) ? accA : accB;
                            t = fma(bv[q].x, uv.x, t); t = fma(bv[q].y, uv.y, t);
                            if (l16 < 2) accA = t; else accB = t;
                        }
                        val[2 * q] = accA;
                        val[2 * q + 1] = accB;
                    }
#pragma unroll
                    for (int i = 2 * GP; i < 8; ++i) val[i] = 0.0;
                    // (issuing the next pass's loads here, ahead of the reduction, was measured: slower -- the shuffles then
                    // queue behind the loads in the SM's memory-instruction pipe)
                    double k4[4], k2[2], k1;
                    {
                        const bool hi = (l16 & 8) != 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double send = hi ? val[i] : val[i + 4];
                            const double recv = __shfl_xor_sync(0xffffffffu, send, 8);
                            k4[i] = (hi ? val[i + 4] : val[i]) + recv;
                        }
                    }
                    {
                        const bool hi = (l16 & 4) != 0;
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const double send = hi ? k4[i] : k4[i + 2];
                            const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
                            k2[i] = (hi ? k4[i + 2] : k4[i]) + recv;
                        }
                    }
                    {
                        const bool hi = (l16 & 2) != 0;
                        const double send = hi ? k2[0] : k2[1];
                        const double recv = __shfl_xor_sync(0xffffffffu, send, 2);
                        k1 = (hi ? k2[1] : k2[0]) + recv;
                    }
                    k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
                    if (sown) sxn[hw * n + hsel * (PPT / 2) * 2 + rloc] = __dadd_rn(k1, down);
                }
            } else
            {
                const int part = ht & 3;
                constexpr int QROWS = kSHalf / 4;
                for (int rb = 0; rb < tpg * n; rb += 2 * QROWS) {
                    // compile-time shape: a lane group takes two ADJACENT rows (same trajectory, n even), so the state
                    // values are read once for both rows -- 9 LDS.128 instead of 36 LDS.64 per pass in a phase whose
                    // limiter is the load / shared-memory instruction queue
                    const int row0 = (RT > 0) ? rb + 2 * (ht >> 2) : rb + (ht >> 2), row1 = (RT > 0) ? row0 + 1 : row0 + QROWS;
                    const bool act0 = row0 < tpg * n, act1 = row1 < tpg * n;
                    const int tr0 = act0 ? row0 / n : 0, tr1 = act1 ? row1 / n : 0;
                    const int i0 = act0 ? row0 - tr0 * n : 0, i1 = act1 ? row1 - tr1 * n : 0;
                    const long long pa = sel[tr0], pb = sel[tr1];
                    const double* A0 = M.A + (pa * n + i0) * n;
                    const double* A1 = M.A + (pb * n + i1) * n;
                    const double* x0s = sx + tr0 * n;
                    const double* x1s = sx + tr1 * n;
                    double ax0 = 0.0, ax1 = 0.0, bu0 = 0.0, bu1 = 0.0;
                    // B / d of both rows are requested together with A: one memory round trip per pass, not two
                    const double* B0 = M.B + (pa * n + i0) * m;
                    const double* B1 = M.B + (pb * n + i1) * m;
                    const double d0 = __ldcg(M.d + pa * n + i0), d1 = __ldcg(M.d + pb * n + i1);
                    double b0v = 0.0, b1v = 0.0;
                    if (CM > 0 && CM <= 4) {
                        if (part < m) { b0v = __ldcg(B0 + part); b1v = __ldcg(B1 + part); }
                    }
                    if constexpr (RT > 0 && (2 * RT) % 8 == 0) {
                        // 16-byte loads: lane `part` takes elements 8 s + 2 part, 8 s + 2 part + 1 (rows are 16-byte
                        // aligned: n even, bank base from cudaMalloc); half the load instructions / L1 tag lookups
                        constexpr int NK = (2 * RT) / 8;
                        double2 v0[NK], v1[NK];
#pragma unroll
                        for (int s2 = 0; s2 < NK; ++s2) {
                            // L2-only loads: with ~173 KB of shared memory the L1 holds far fewer lines than the gather
                            // keeps in flight, and an allocating load waits for a free line
                            v0[s2] = __ldcg(reinterpret_cast<const double2*>(A0 + 8 * s2 + 2 * part));
                            v1[s2] = __ldcg(reinterpret_cast<const double2*>(A1 + 8 * s2 + 2 * part));
                        }
#pragma unroll
                        for (int s2 = 0; s2 < NK; ++s2) {
                            const double2 xv = *reinterpret_cast<const double2*>(x0s + 8 * s2 + 2 * part);   // x1s == x0s
                            ax0 = fma(v0[s2].x, xv.x, ax0); ax0 = fma(v0[s2].y, xv.y, ax0);
                            ax1 = fma(v1[s2].x, xv.x, ax1); ax1 = fma(v1[s2].y, xv.y, ax1);
                        }
                    } else {
                        for (int k = part; k < n; k += 4) { ax0 = fma(A0[k], x0s[k], ax0); ax1 = fma(A1[k], x1s[k], ax1); }
                    }
                    if (CM > 0 && CM <= 4) {
                        if (part < m) {
                            bu0 = fma(b0v, su[tr0 * m + part], bu0);
                            bu1 = fma(b1v, su[tr1 * m + part], bu1);
                        }
                    } else {
                        for (int k = part; k < m; k += 4) {
                            bu0 = fma(B0[k], su[tr0 * m + k], bu0);
                            bu1 = fma(B1[k], su[tr1 * m + k], bu1);
                        }
                    }
                    ax0 += __shfl_xor_sync(0xffffffffu, ax0, 1);  ax1 += __shfl_xor_sync(0xffffffffu, ax1, 1);
                    bu0 += __shfl_xor_sync(0xffffffffu, bu0, 1);  bu1 += __shfl_xor_sync(0xffffffffu, bu1, 1);
                    ax0 += __shfl_xor_sync(0xffffffffu, ax0, 2);  ax1 += __shfl_xor_sync(0xffffffffu, ax1, 2);
                    bu0 += __shfl_xor_sync(0xffffffffu, bu0, 2);  bu1 += __shfl_xor_sync(0xffffffffu, bu1, 2);
                    if (part == 0) {
                        if (act0) sxn[row0] = __dadd_rn(__dadd_rn(ax0, bu0), d0);
                        if (act1) sxn[row1] = __dadd_rn(__dadd_rn(ax1, bu1), d1);
                    }
                }
            }
            half_sync(half);
            NN_PH(5);
            for (int e = ht; e < kST * n; e += kSHalf) {
                const int tr = e / n, i = e - tr * n;
                const double v = sxn[e];
                sx[e] = v;
                if (tr < nt) xo[((b0 + tr) * (long long)(N + 1) + t + 1) * n + i] = v;
            }
            if (ht < kST * m) su[ht] = u_next;
            if (t + 1 < N) prep_state(sxn);
            half_sync(half);
            NN_PH(6);
        }
    }
}

// Dispatch: nn rollout on a bank that needs no discretisation, exactly one non-negative distance weight,
// P <= 1024 and an FP32 bank that fits in shared memory.
int tpwl_rollout_nn_screen_launch(const TpwlDev& M, long long batch, int N, const double* x0, const double* u,
                                  double* x, int* idx, cudaStream_t st, bool* handled) {
    *handled = false;
    const char* env = getenv("SRCB200_TPWL_NOSCREEN");
    if (env && env[0] == '1') return 0;
    const bool useq = M.wq != 0.0, usev = M.wv != 0.0;
    if (useq == usev) return 0;                                    // both or none
    if ((useq ? M.wq : M.wv) < 0.0) return 0;
    if (M.P > kSPts * kSHalf || M.P < 1 || M.r > 128 || M.m > 32 || M.n != 2 * M.r) return 0;
    if ((reinterpret_cast<uintptr_t>(M.A) | reinterpret_cast<uintptr_t>(M.B)) & 15) return 0;   // the gather uses 16-byte loads
    const ScreenPlan S = make_screen_plan(M.n, M.m, M.r, M.P);
    if (S.total > 220 * 1024) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (batch < 1) { *handled = true; return 0; }
    // equal work per half-CTA: rounds = groups of kST per half, then the smallest group size that needs that many rounds
    const long long halves = 2LL * sms;
    const long long rounds = (batch + halves * kST - 1) / (halves * kST);
    int tpg = (int)((batch + halves * rounds - 1) / (halves * rounds));
    if (tpg > kST) tpg = kST;
    if (tpg < 1) tpg = 1;
    const long long groups = (batch + tpg - 1) / tpg;
    const long long ctas = (groups + 1) / 2;
    const int grid = (int)(ctas < sms ? ctas : sms);
    int stagger = 0;                   // measured at 0 / 4000 / 9000 / 14000 cycles: no effect on the 4096 x 100 rollout
    if (const char* e2 = getenv("SRCB200_NN_STAGGER")) stagger = atoi(e2);
    if (M.r == 36 && M.m == 4) {
        SRCB_CUDA(cudaFuncSetAttribute(tpwl_rollout_nn_screen_kernel<36, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
        tpwl_rollout_nn_screen_kernel<36, 4><<<grid, kSThreads, S.total, st>>>(M, batch, N, x0, u, x, idx, useq ? 1 : 0, tpg, stagger);
    } else {
        SRCB_CUDA(cudaFuncSetAttribute(tpwl_rollout_nn_screen_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
        tpwl_rollout_nn_screen_kernel<0, 0><<<grid, kSThreads, S.total, st>>>(M, batch, N, x0, u, x, idx, useq ? 1 : 0, tpg, stagger);
    }
    SRCB_LAUNCH_CHECK("tpwl_rollout_nn_screen_kernel");
#ifdef SRCB_NN_PHASES
    {
        unsigned long long h[8];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_nn_phase, sizeof(h));
        const char* nm[7] = {"state copy + norm", "dot products + min", "threshold", "candidates", "stage 2 / fallback", "gather + affine step", "state update"};
        for (int i = 0; i < 7; ++i) fprintf(stderr, "  nn phase %-22s %8.0f cycles per step\n", nm[i], h[7] ? (double)h[i] / h[7] : 0.0);
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_nn_phase, h, sizeof(h));
    }
#endif
    *handled = true;
    return 0;
}

}  // namespace srcb
