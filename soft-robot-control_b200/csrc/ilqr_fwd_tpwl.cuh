// ilqr_fwd_tpwl.cuh -- forward rollout of the iLQR line search (sofacontrol/lqr/ilqr.py:117-175) specialised for a
// TPWL model in nearest-neighbour mode on a bank that needs no per-step discretisation (tpwl.py:160-168, 236-244:
// the linearisation of a step is an index into the bank).  Included by ilqr_impl.cuh; one CTA of NT threads per problem.
//
// What makes the generic step slow on this model is (1) the exact nearest-point search streaming the whole FP64
// point bank from L2 every step and (2) a chain of small dependent global-memory round trips.  Here
//   * the search is a two-stage EXACT search: an FP32 copy of the point bank lives in shared memory for the whole
//     pass; every step computes FP32 distances d^ with a rigorous error bound eps (below), keeps the candidates
//     {p : d^_p - eps_p <= min_p' (d^_p' + eps_p')} -- the FP64 argmin is provably among them -- and evaluates only
//     those with the bit-exact numpy-order FP64 distance (tpwl.cuh).  The selected index is identical to the full
//     search, ties included (first occurrence).  If the bank does not fit or anything is off (no candidate, NaN,
//     candidate overflow) the step falls back to the full FP64 search.
//   * the inputs of step t+1 (nominal x/u, gains K/k, target) are cp.async-prefetched into shared memory during
//     step t, and [A | B | d] of the current bank index stays in shared memory while the index does not change.
//
// Error bound of the screening distance (u = 2^-24, a_j = Q_pj - q_j exact, a^_j its FP32 evaluation from rounded
// operands): ||a^ - a|| <= u (1+u) (||Q_p|| + ||q||) + u ||a||;  the FP32 sum of 36 squares and the square root add at
// most 20 u ||a^||.  eps_p = 2^-24 (32 d^_p + 2 (max_p ||Q_p|| + ||q||)) + 1e-14 d^_p covers both with margin.
#pragma once

namespace srcb {

// FP32 screening distances of points p0 (and p1): sqrt(sum_j (bank[j*P+p] - x[j])^2) with bank and x rounded to FP32;
// xn2 receives sum_j x[j]^2 (the error bound needs the norm of the state part)
__device__ __forceinline__ void screen2(const float* __restrict__ bank, int P, int r, const float* __restrict__ xf,
                                        int p0, int p1, bool has1, float& d0, float& d1, float& xn2) {
    float s0 = 0.f, s1 = 0.f, sx2 = 0.f;
#pragma unroll 4
    for (int j = 0; j < r; ++j) {
        const float xj = xf[j];
        sx2 = fmaf(xj, xj, sx2);
        const float a0 = bank[j * P + p0] - xj;
        s0 = fmaf(a0, a0, s0);
        if (has1) {
            const float a1 = bank[j * P + p1] - xj;
            s1 = fmaf(a1, a1, s1);
        }
    }
    d0 = sqrtf(s0);
    d1 = sqrtf(s1);
    xn2 = sx2;
}

template <class MP>
__device__ double fwd_pass_tpwl_nn(const TpwlDev& M, const IlqrArgs& a, const Smem& S, double* sm,
                                   const double* __restrict__ nx, const double* __restrict__ nu, double alpha,
                                   const double* __restrict__ K, const double* __restrict__ k, const Rec& tr,
                                   const double* __restrict__ ztar, const double* __restrict__ ulast,
                                   double* __restrict__ dout) {
    constexpr int NT = MP::NT;
    constexpr int NW = NT / 32;
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = M.P, r = n / 2;
    const bool useq = M.wq != 0.0, usev = M.wv != 0.0;
    const FwdNNPlan F = make_fwdnn(n, m, nz, P, r, useq, usev, NT);
    double* base = sm + S.mscr;
    double* sxa = sm + S.x;   double* sxb = sm + S.xn;  double* su = sm + S.u;   double* sup = sm + S.uprev;
    double* sz = sm + S.z;    double* se = sm + S.e;    double* sQe = sm + S.Qe; double* sRdu = sm + S.Rdu;
    const double* sQ = sm + S.Qs; const double* sR = sm + S.Rs; const double* sQf = sm + S.Qfs;
    const double* sH = sm + S.Hcs;                       // constant output matrix H (n_z x n), staged by load_costs
    double* scal = sm + S.scal;
    double* Acur = base + F.Acur;
    const int LDA = F.LDA;
    float* sxf = reinterpret_cast<float*>(base + F.xf);
    float* dh = reinterpret_cast<float*>(base + F.dh);
    int* cand = reinterpret_cast<int*>(base + F.cand);
    double* red_d = base + F.red;
    int* red_i = reinterpret_cast<int*>(red_d + NW);
    double* misc = base + F.misc;                        // [0] = bank norm bound, [1] = min upper bound
    float* qf = reinterpret_cast<float*>(base + F.qf);
    float* vf = reinterpret_cast<float*>(base + F.vf);
    const bool screen = F.screen && M.wq >= 0.0 && M.wv >= 0.0;
    const int PRE_NX = 0, PRE_NU = n, PRE_KK = n + m, PRE_K = n + 2 * m, PRE_ZT = n + 2 * m + m * n;

    auto prefetch = [&](int t) {
        double* dst = base + ((t & 1) ? F.pre1 : F.pre0);
        for (int e = tid; e < n; e += NT) cp_async8(dst + PRE_NX + e, nx + (long long)t * n + e);
        for (int e = tid; e < m; e += NT) cp_async8(dst + PRE_NU + e, nu + t * m + e);
        if (k) for (int e = tid; e < m; e += NT) cp_async8(dst + PRE_KK + e, k + t * m + e);
        if (K) for (int e = tid; e < m * n; e += NT) cp_async8(dst + PRE_K + e, K + (long long)t * m * n + e);
        for (int e = tid; e < nz; e += NT) cp_async8(dst + PRE_ZT + e, ztar + t * nz + e);
    };

    // ---- pass prologue: state, FP32 banks + the largest weighted point norm (the error bound needs it)
    for (int i = tid; i < n; i += NT) { sxa[i] = nx[i]; sxf[i] = (float)nx[i]; tr.x[i] = nx[i]; }
    for (int i = tid; i < m; i += NT) sup[i] = ulast ? ulast[i] : 0.0;
    if (N > 0) prefetch(0);
    if (screen) {
        double bmax = 0.0;
        for (int p = tid; p < P; p += NT) {
            double sq = 0.0, sv = 0.0;
            for (int j = 0; j < r; ++j) {
                if (useq) { const double v = M.qT[(size_t)j * P + p]; qf[j * P + p] = (float)v; sq = fma(v, v, sq); }
                if (usev) { const double v = M.vT[(size_t)j * P + p]; vf[j * P + p] = (float)v; sv = fma(v, v, sv); }
            }
            bmax = fmax(bmax, M.wq * sqrt(sq) + M.wv * sqrt(sv));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, off));
        if (lane == 0) red_d[warp] = bmax;
        __syncthreads();
        if (tid == 0) {
            double b2 = 0.0;
            for (int w = 0; w < NW; ++w) b2 = fmax(b2, red_d[w]);
            misc[0] = b2 * (1.0 + 1e-6);
        }
    }
    __syncthreads();
    const double bank_norm = screen ? misc[0] : 0.0;
    const bool screen_ok = screen && isfinite(bank_norm) && bank_norm < 1e100;
    int cur_idx = -1;
    double cost = 0.0;                                   // meaningful on the cost warp's lane 0
    double* sx = sxa;
    double* sxn = sxb;

    for (int t = 0; t < N; ++t) {
        cp_async_wait_all();
        __syncthreads();                                 // inputs of step t are in shared memory; sx is final
        const double* pre = base + ((t & 1) ? F.pre1 : F.pre0);
        if (t + 1 < N) prefetch(t + 1);
        // ---- phase 1: u_t = u_prev + alpha k + K (x - x_prev) (ilqr.py:140) by warps 0..m-1, z_t = H x + z_ref
        //      (tpwl.py:121-122) by warps m..m+nz-1 (one row each, lanes split the sum), screening by everybody
        for (int row = warp; row < m + nz; row += NW) {
            if (row < m) {
                double acc = 0.0;
                if (K) for (int j = lane; j < n; j += 32) acc = fma(pre[PRE_K + row * n + j], __dsub_rn(sx[j], pre[PRE_NX + j]), acc);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                if (lane == 0) {
                    double v = pre[PRE_NU + row];
                    if (k) v = __dadd_rn(v, __dmul_rn(alpha, pre[PRE_KK + row]));
                    if (K) v = __dadd_rn(v, acc);
                    su[row] = v;
                    tr.u[t * m + row] = v;
                }
            } else {
                const int i = row - m;
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(sH[i * n + j], sx[j], acc);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                if (lane == 0) {
                    const double z = __dadd_rn(acc, M.zref[i]);
                    const double e = __dsub_rn(z, pre[PRE_ZT + i]);
                    sz[i] = z;
                    se[i] = e;
                    tr.e[t * nz + i] = e;
                }
            }
        }
        int idx = -1;
        bool full_search = !screen_ok;
        if (screen_ok && useq != usev && P <= 2 * NT) {
            // ---- one screened bank (the usual case: one non-zero distance weight).  With eps(d) = c1 d + c0 the smallest
            //      upper bound is U = (1 + c1) w sqrt(min a^) + c0 and "lower bound <= U" is a^ <= ((U + c0) / (w (1 - c1)))^2:
            //      per point one float minimum and one float compare on the SQUARED screening distance a^; the bound
            //      arithmetic runs in double on the reduced minimum, the threshold is rounded UP.
            const double u24 = 5.9604644775390625e-08;   // 2^-24
            const float* bk = useq ? qf : vf;
            const float* xs = useq ? sxf + r : sxf;
            const double w = useq ? M.wq : M.wv;
            const int p0 = tid, p1 = tid + NT;
            const bool has0 = p0 < P, has1 = p1 < P;
            const float* b0 = bk + (has0 ? p0 : 0);
            const float* b1 = bk + (has1 ? p1 : 0);
            float a0 = 0.f, a1 = 0.f, xn2 = 0.f;
#pragma unroll 4
            for (int j = 0; j < r; ++j) {
                const float xj = xs[j];
                xn2 = fmaf(xj, xj, xn2);
                const float d0 = b0[j * P] - xj, d1 = b1[j * P] - xj;
                a0 = fmaf(d0, d0, a0);
                a1 = fmaf(d1, d1, a1);
            }
            float amin = INFINITY;
            if (has0) amin = a0;
            if (has1) amin = fminf(amin, a1);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, off));
            float* red_f = reinterpret_cast<float*>(red_d);
            if (lane == 0) red_f[warp] = amin;
            if (tid == 0) cand[0] = 0;
            __syncthreads();
            float vmin = red_f[0];
#pragma unroll
            for (int k2 = 1; k2 < NW; ++k2) vmin = fminf(vmin, red_f[k2]);
            const double c1 = (0.5 * r + 16.0) * u24 + 1e-14;   // r/2 + 2 ulp accumulate over r squares (+ margin); = 34 u at r = 36
            const double c0 = u24 * 2.0 * (bank_norm + w * 1.001 * (double)sqrtf(xn2));
            const double U = (1.0 + c1) * w * sqrt((double)vmin) + c0;
            const double T = (U + c0) / (w * (1.0 - c1));
            const double T2 = T * T * (1.0 + 1e-6);
            const float thr2 = (T2 < 3.0e38) ? __double2float_ru(T2) : INFINITY;       // NaN compares false: full search
            if (has0 && a0 <= thr2) { const int pos = atomicAdd(&cand[0], 1); if (pos < kFwdNNCandCap) cand[2 + pos] = p0; }
            if (has1 && a1 <= thr2) { const int pos = atomicAdd(&cand[0], 1); if (pos < kFwdNNCandCap) cand[2 + pos] = p1; }
            __syncthreads();
            const int cnt = cand[0];
            if (cnt == 1) {
                idx = cand[2];      // the argmin is among the candidates and there is only one: no FP64 evaluation needed
            } else if (cnt >= 1 && cnt <= kFwdNNCandCap && cnt <= NT) {
                double best = INFINITY;
                int bi = 0x7fffffff;
                if (tid < cnt) {
                    const int p = cand[2 + tid];
                    const double dd = tpwl_distance(M, sx, p);       // bit-exact numpy-order FP64 distance
                    if (dd < best) { best = dd; bi = p; }
                }
                cta_argmin<NT>(best, bi, red_d, red_i);
                if (bi == 0x7fffffff) full_search = true;            // NaN distances: let the full search decide
                else idx = bi;
            } else {
                full_search = true;
            }
        } else if (screen_ok) {
            const double u24 = 5.9604644775390625e-08;   // 2^-24
            double ubmin = INFINITY;
            double slack = 0.0;
            for (int p0 = tid; p0 < P; p0 += 2 * NT) {
                const int p1 = p0 + NT;
                const bool has1 = p1 < P;
                float q0 = 0.f, q1 = 0.f, v0 = 0.f, v1 = 0.f, xq2 = 0.f, xv2 = 0.f;
                if (useq) screen2(qf, P, r, sxf + r, p0, has1 ? p1 : p0, has1, q0, q1, xq2);
                if (usev) screen2(vf, P, r, sxf, p0, has1 ? p1 : p0, has1, v0, v1, xv2);
                // 2 (max point norm + state norm), the FP32 norms padded by 1e-3 relative
                slack = 2.0 * (bank_norm + 1.001 * (M.wq * (double)sqrtf(xq2) + M.wv * (double)sqrtf(xv2)));
                const double d0 = M.wq * (double)q0 + M.wv * (double)v0;
                const double d1 = M.wq * (double)q1 + M.wv * (double)v1;
                dh[p0] = (float)d0;
                // (float)d rounds once more: the candidate test below reads the rounded value and pads the bound for it
                const double e0 = u24 * ((0.5 * r + 16.0) * d0 + slack) + 1e-14 * d0;
                ubmin = fmin(ubmin, d0 + e0);
                if (has1) {
                    dh[p1] = (float)d1;
                    const double e1 = u24 * ((0.5 * r + 16.0) * d1 + slack) + 1e-14 * d1;
                    ubmin = fmin(ubmin, d1 + e1);
                }
            }
            if (lane == 0 && warp == 0) misc[2] = slack;        // thread 0 always owns a point (P >= 1)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ubmin = fmin(ubmin, __shfl_xor_sync(0xffffffffu, ubmin, off));
            if (lane == 0) red_d[warp] = ubmin;
            if (tid == 0) cand[0] = 0;
            __syncthreads();
            double U = red_d[0];
#pragma unroll
            for (int w = 1; w < NW; ++w) U = fmin(U, red_d[w]);
            // candidates: lower bound <= smallest upper bound
            const double slack_all = misc[2];
            for (int p = tid; p < P; p += NT) {
                const double d = (double)dh[p];
                const double e = u24 * ((0.5 * r + 18.0) * d + slack_all) + 1e-14 * d;
                if (d - e <= U) {
                    const int pos = atomicAdd(&cand[0], 1);
                    if (pos < kFwdNNCandCap) cand[2 + pos] = p;
                }
            }
            __syncthreads();
            const int cnt = cand[0];
            if (cnt == 1) {
                idx = cand[2];      // the argmin is among the candidates and there is only one: no FP64 evaluation needed
            } else if (cnt >= 1 && cnt <= kFwdNNCandCap && cnt <= NT) {
                double best = INFINITY;
                int bi = 0x7fffffff;
                if (tid < cnt) {
                    const int p = cand[2 + tid];
                    const double dd = tpwl_distance(M, sx, p);       // bit-exact numpy-order FP64 distance
                    if (dd < best) { best = dd; bi = p; }
                }
                cta_argmin<NT>(best, bi, red_d, red_i);
                if (bi == 0x7fffffff) full_search = true;            // NaN distances: let the full search decide
                else idx = bi;
            } else {
                full_search = true;
            }
        } else {
            __syncthreads();
        }
        if (full_search) idx = tpwl_nearest<NT>(M, sx, nullptr, red_d, red_i, nullptr);
        // ---- phase 2: the linearisation of this step: [A | B | d] of bank entry idx (kept while idx is unchanged)
        if (idx != cur_idx) {
            const LinRef b = MP::bank(M, idx);
            for (int e = tid; e < n * n; e += NT) { const int i = e / n; Acur[i * LDA + (e - i * n)] = b.A[e]; }
            for (int e = tid; e < n * m; e += NT) Acur[n * LDA + e] = b.B[e];
            for (int e = tid; e < n; e += NT) Acur[n * LDA + n * m + e] = b.d[e];
            cur_idx = idx;
        }
        __syncthreads();                                  // su, se, Acur ready
        if (tid == 0) tr.idx[t] = idx;
        if (dout) for (int e = tid; e < n; e += NT) dout[(long long)t * n + e] = Acur[n * LDA + n * m + e];
        // ---- step cost (ilqr.py:168-175) on the last warp: .5 e^T Q e + .5 du^T R du, row vector times matrix first
        if (warp == NW - 1) {
            double du = 0.0;
            if (lane < m) {
                du = a.cfg.include_input_var_constraint ? __dsub_rn(su[lane], sup[lane]) : su[lane];
            }
            __syncwarp();
            if (lane < m) sup[lane] = du;
            __syncwarp();
            if (lane < nz) {
                double acc = 0.0;
                for (int i = 0; i < nz; ++i) acc = fma(se[i], sQ[i * nz + lane], acc);
                sQe[lane] = acc;
            }
            if (lane < m) {
                double acc = 0.0;
                for (int i = 0; i < m; ++i) acc = fma(sup[i], sR[i * m + lane], acc);
                sRdu[lane] = acc;
            }
            __syncwarp();
            if (lane == 0) {
                double s1 = 0.0, s2 = 0.0;
                for (int j = 0; j < nz; ++j) s1 = fma(sQe[j], se[j], s1);
                for (int j = 0; j < m; ++j) s2 = fma(sRdu[j], sup[j], s2);
                cost = __dadd_rn(cost, __dadd_rn(__dmul_rn(0.5, s1), __dmul_rn(0.5, s2)));
            }
            __syncwarp();
            if (lane < m) sup[lane] = su[lane];           // u_{t-1} of the next step
        }
        // ---- x_{t+1} = (A x + B u) + d  (tpwl.py:231-234): four lanes per row, each a strided quarter of the sums
        {
            const double* A = Acur;
            const double* B = Acur + n * LDA;
            const double* d = Acur + n * LDA + n * m;
            const int part = tid & 3;
            for (int ib = 0; ib < n; ib += NT / 4) {                 // uniform trip count: the shuffles need whole warps
                const int i = ib + (tid >> 2);
                const bool act = i < n;
                double ax = 0.0, bu = 0.0;
                if (act) {
                    for (int kk = part; kk < n; kk += 4) ax = fma(A[i * LDA + kk], sx[kk], ax);
                    for (int kk = part; kk < m; kk += 4) bu = fma(B[i * m + kk], su[kk], bu);
                }
                ax += __shfl_xor_sync(0xffffffffu, ax, 1);
                bu += __shfl_xor_sync(0xffffffffu, bu, 1);
                ax += __shfl_xor_sync(0xffffffffu, ax, 2);
                bu += __shfl_xor_sync(0xffffffffu, bu, 2);
                if (act && part == 0) {
                    const double v = __dadd_rn(__dadd_rn(ax, bu), d[i]);
                    sxn[i] = v;
                    sxf[i] = (float)v;
                    tr.x[(long long)(t + 1) * n + i] = v;
                }
            }
        }
        double* tmp = sx; sx = sxn; sxn = tmp;
    }
    __syncthreads();
    // ---- terminal cost (ilqr.py:164-166)
    for (int row = warp; row < nz; row += NW) {
        double acc = 0.0;
        for (int j = lane; j < n; j += 32) acc = fma(sH[row * n + j], sx[j], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) {
            const double e = __dsub_rn(__dadd_rn(acc, M.zref[row]), ztar[N * nz + row]);
            se[row] = e;
            tr.e[N * nz + row] = e;
        }
    }
    __syncthreads();
    if (warp == NW - 1) {
        if (lane < nz) {
            double acc = 0.0;
            for (int i = 0; i < nz; ++i) acc = fma(se[i], sQf[i * nz + lane], acc);
            sQe[lane] = acc;
        }
        __syncwarp();
        if (lane == 0) {
            double s1 = 0.0;
            for (int j = 0; j < nz; ++j) s1 = fma(sQe[j], se[j], s1);
            cost = __dadd_rn(cost, __dmul_rn(0.5, s1));
            scal[0] = cost;
        }
    }
    __syncthreads();
    cost = scal[0];
    __syncthreads();
    return cost;
}

}  // namespace srcb
