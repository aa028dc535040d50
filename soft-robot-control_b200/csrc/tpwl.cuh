// tpwl.cuh -- device-side TPWL point selection (nearest neighbour / exponential weights).
// Reference math: sofacontrol/tpwl/tpwl.py:160-191.
#pragma once
#include "common.cuh"

namespace srcb {

struct TpwlDev {
    int n, m, nz, P, r, method, discr;
    double wq, wv, beta;
    const double* qT;
    const double* vT;
    const double* A;
    const double* B;
    const double* d;
    const double* H;
    const double* zref;
};

inline TpwlDev to_dev(const srcb200_tpwl_model& s) {
    TpwlDev t;
    t.n = s.n; t.m = s.m; t.nz = s.nz; t.P = s.P; t.r = s.n / 2; t.method = s.method; t.discr = s.discr_method;
    t.wq = s.wq; t.wv = s.wv; t.beta = s.beta;
    t.qT = s.qT; t.vT = s.vT; t.A = s.A; t.B = s.B; t.d = s.d; t.H = s.H; t.zref = s.z_ref;
    return t;
}

int check_tpwl_model(const srcb200_tpwl_model* mdl);

// sum_j (bankT[j*P + p] - c[j])^2 over j in [lo, hi) in EXACTLY the order numpy's pairwise summation uses for
// np.add.reduce(s, axis=1) on a contiguous row (what np.linalg.norm(axis=1) runs; tpwl.py:166-167): squares and
// differences are rounded separately (no FMA), 8 strided accumulators for 8 <= n <= 128, sequential tail, rows
// longer than 128 split recursively at n/2 rounded down to a multiple of 8 (the model is
// checked against numpy's bits in tests/test_oracle_pairwise.py).
__device__ __forceinline__ double sq_diff(const double* __restrict__ bankT, int P, int p, const double* __restrict__ c, int j) {
    const double t = __dsub_rn(bankT[(size_t)j * P + p], c[j]);
    return __dmul_rn(t, t);
}

__device__ inline double np_pairwise_sumsq(const double* __restrict__ bankT, int P, int p,
                                           const double* __restrict__ c, int lo, int hi) {
    const int n = hi - lo;
    if (n < 8) {
        double res = 0.0;
        for (int i = lo; i < hi; ++i) res = __dadd_rn(res, sq_diff(bankT, P, p, c, i));
        return res;
    }
    if (n <= 128) {
        double r0 = sq_diff(bankT, P, p, c, lo + 0), r1 = sq_diff(bankT, P, p, c, lo + 1);
        double r2 = sq_diff(bankT, P, p, c, lo + 2), r3 = sq_diff(bankT, P, p, c, lo + 3);
        double r4 = sq_diff(bankT, P, p, c, lo + 4), r5 = sq_diff(bankT, P, p, c, lo + 5);
        double r6 = sq_diff(bankT, P, p, c, lo + 6), r7 = sq_diff(bankT, P, p, c, lo + 7);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
            r0 = __dadd_rn(r0, sq_diff(bankT, P, p, c, lo + i + 0));
            r1 = __dadd_rn(r1, sq_diff(bankT, P, p, c, lo + i + 1));
            r2 = __dadd_rn(r2, sq_diff(bankT, P, p, c, lo + i + 2));
            r3 = __dadd_rn(r3, sq_diff(bankT, P, p, c, lo + i + 3));
            r4 = __dadd_rn(r4, sq_diff(bankT, P, p, c, lo + i + 4));
            r5 = __dadd_rn(r5, sq_diff(bankT, P, p, c, lo + i + 5));
            r6 = __dadd_rn(r6, sq_diff(bankT, P, p, c, lo + i + 6));
            r7 = __dadd_rn(r7, sq_diff(bankT, P, p, c, lo + i + 7));
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                               __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
        for (; i < n; ++i) res = __dadd_rn(res, sq_diff(bankT, P, p, c, lo + i));
        return res;
    }
    // rows longer than 128 (numpy splits them at n/2 rounded down to a multiple of 8, two levels at most up to 512): not
    // recursive on the device -- a recursive function has no static stack bound.  check_tpwl_model limits r to 64, so
    // this branch is only reached by direct callers with long rows.
    int n2 = n / 2;
    n2 -= n2 % 8;
    double half[2];
    for (int h = 0; h < 2; ++h) {
        const int l = h ? lo + n2 : lo, u = h ? hi : lo + n2, nn = u - l;
        if (nn <= 128) {
            double r[8];
            for (int k = 0; k < 8; ++k) r[k] = sq_diff(bankT, P, p, c, l + k);
            int i = 8;
            for (; i < nn - (nn % 8); i += 8)
                for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], sq_diff(bankT, P, p, c, l + i + k));
            double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                   __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
            for (; i < nn; ++i) res = __dadd_rn(res, sq_diff(bankT, P, p, c, l + i));
            half[h] = res;
        } else {
            half[h] = __longlong_as_double(0x7ff8000000000000LL);      // rows beyond 256 entries are not supported
        }
    }
    return __dadd_rn(half[0], half[1]);
}

// r = 36 (Diamond) with everything unrolled: 8 accumulators over the first 32 terms, 4 sequential tail terms --
// exactly the order np_pairwise_sumsq takes for n = 36, without the loop / branch overhead.
__device__ __forceinline__ double np_pairwise_sumsq_36(const double* __restrict__ bankT, int P, int p,
                                                       const double* __restrict__ c) {
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = sq_diff(bankT, P, p, c, k);
#pragma unroll
    for (int i = 8; i < 32; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], sq_diff(bankT, P, p, c, i + k));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
#pragma unroll
    for (int i = 32; i < 36; ++i) res = __dadd_rn(res, sq_diff(bankT, P, p, c, i));
    return res;
}

// Four points at once, each in exactly the order of np_pairwise_sumsq_36: the four load streams are independent, so
// four times as many bank loads are in flight (the searches are bound by the latency of these loads).
__device__ __forceinline__ void np_pairwise_sumsq_36x4(const double* __restrict__ bankT, int P, const int (&p)[4],
                                                       const double* __restrict__ c, double (&out)[4]) {
    double r[4][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double cj = c[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double t = __dsub_rn(bankT[(size_t)k * P + p[q]], cj);
            r[q][k] = __dmul_rn(t, t);
        }
    }
#pragma unroll
    for (int i = 8; i < 32; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double cj = c[i + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double t = __dsub_rn(bankT[(size_t)(i + k) * P + p[q]], cj);
                r[q][k] = __dadd_rn(r[q][k], __dmul_rn(t, t));
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        out[q] = __dadd_rn(__dadd_rn(__dadd_rn(r[q][0], r[q][1]), __dadd_rn(r[q][2], r[q][3])),
                           __dadd_rn(__dadd_rn(r[q][4], r[q][5]), __dadd_rn(r[q][6], r[q][7])));
#pragma unroll
    for (int i = 32; i < 36; ++i) {
        const double cj = c[i];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double t = __dsub_rn(bankT[(size_t)i * P + p[q]], cj);
            out[q] = __dadd_rn(out[q], __dmul_rn(t, t));
        }
    }
}

// tpwl_distance for four points (r = 36), bit-identical per point
__device__ __forceinline__ void tpwl_distance_36x4(const TpwlDev& M, const double* __restrict__ x, const int (&p)[4],
                                                   double (&d)[4]) {
    double sq[4], sv[4];
    if (M.wq != 0.0) np_pairwise_sumsq_36x4(M.qT, M.P, p, x + 36, sq);
    if (M.wv != 0.0) np_pairwise_sumsq_36x4(M.vT, M.P, p, x, sv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double dq = (M.wq != 0.0) ? __dmul_rn(M.wq, sqrt(sq[q])) : 0.0;
        const double dv = (M.wv != 0.0) ? __dmul_rn(M.wv, sqrt(sv[q])) : 0.0;
        d[q] = __dadd_rn(dq, dv);
    }
}

// d_p = wq * ||Q_p - q|| + wv * ||V_p - v||  with x = [v; q]  (utils.py:133-142, tpwl.py:165-168)
// A zero weight contributes +0.0 (numpy computes 0 * norm; identical unless the norm is inf/nan).
__device__ __forceinline__ double tpwl_distance(const TpwlDev& M, const double* __restrict__ x, int p) {
    const int r = M.r;
    double dq = 0.0, dv = 0.0;
    if (r == 36) {
        if (M.wq != 0.0) dq = __dmul_rn(M.wq, sqrt(np_pairwise_sumsq_36(M.qT, M.P, p, x + r)));
        if (M.wv != 0.0) dv = __dmul_rn(M.wv, sqrt(np_pairwise_sumsq_36(M.vT, M.P, p, x)));
        return __dadd_rn(dq, dv);
    }
    if (M.wq != 0.0) dq = __dmul_rn(M.wq, sqrt(np_pairwise_sumsq(M.qT, M.P, p, x + r, 0, r)));
    if (M.wv != 0.0) dv = __dmul_rn(M.wv, sqrt(np_pairwise_sumsq(M.vT, M.P, p, x, 0, r)));
    return __dadd_rn(dq, dv);
}

// CTA-wide first-occurrence argmin of (d, p) pairs held one per thread.  red_d/red_i: NT/32 entries of shared
// scratch.  Every thread returns the winning pair.
template <int NT>
__device__ __forceinline__ void cta_argmin(double& d, int& p, double* red_d, int* red_i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, d, off);
        const int op = __shfl_xor_sync(0xffffffffu, p, off);
        if (od < d || (od == d && op < p)) { d = od; p = op; }
    }
    if (NT > 32) {
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        __syncthreads();
        if (l == 0) { red_d[w] = d; red_i[w] = p; }
        __syncthreads();
        d = red_d[0]; p = red_i[0];
#pragma unroll
        for (int k = 1; k < NT / 32; ++k) {
            const double od = red_d[k];
            const int op = red_i[k];
            if (od < d || (od == d && op < p)) { d = od; p = op; }
        }
        __syncthreads();
    }
}

// Nearest stored point to the state x (shared memory, n doubles).  Optionally writes all P distances to `dist`
// (shared or global).  Returns (index, distance) to every thread.
// X4: four points per thread at a time at r = 36 (more bank loads in flight; ~60 more registers -- the stand-alone
// selection / weighting kernels of tpwl.cu use it, the register-bound iLQR kernels do not).
template <int NT, bool X4 = false>
__device__ __forceinline__ int tpwl_nearest(const TpwlDev& M, const double* __restrict__ x, double* __restrict__ dist,
                                            double* red_d, int* red_i, double* dmin_out) {
    double best = INFINITY;
    int bi = 0x7fffffff;
    if (X4 && M.r == 36) {
        for (int p0 = threadIdx.x; p0 < M.P; p0 += 4 * NT) {
            int pp[4];
            double dd[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) pp[q] = (p0 + q * NT < M.P) ? p0 + q * NT : p0;     // out of range: repeat a valid point
            tpwl_distance_36x4(M, x, pp, dd);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (p0 + q * NT < M.P) {
                    if (dist) dist[pp[q]] = dd[q];
                    if (dd[q] < best) { best = dd[q]; bi = pp[q]; }
                }
            }
        }
    } else {
        for (int p = threadIdx.x; p < M.P; p += NT) {
            const double dd = tpwl_distance(M, x, p);
            if (dist) dist[p] = dd;
            if (dd < best) { best = dd; bi = p; }
        }
    }
    cta_argmin<NT>(best, bi, red_d, red_i);
    if (bi == 0x7fffffff) bi = 0;   // all distances NaN/inf: np.argmin returns 0 for an all-inf row
    if (dmin_out) *dmin_out = best;
    return bi;
}

}  // namespace srcb
