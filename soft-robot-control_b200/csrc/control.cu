// control.cu -- batched kernels for the callers either side of the hot path:
//   * DiscreteEKFObserver predict / update        (sofacontrol/tpwl/observer.py:94-126)
//   * infinite-horizon discrete LQR gains          (sofacontrol/lqr/lqr.py:6-31: solve_riccati, dare)
//   * time-varying LQR tracking recursion          (sofacontrol/lqr/traj_tracking_lqr.py:18-48)
//   * TPWL bank construction of one stored point   (sofacontrol/utils.py:251-286 extract_AB,
//                                                   sofacontrol/tpwl/tpwl_utils.py:263-276 add_continuous_TPWL)
//   * GuSTO model-accuracy ratio                   (sofacontrol/scp/gusto.py:203-223 compute_accuracy)
//   * receding-horizon glue: shift plan / targets  (warm-start hooks of sofacontrol/lqr/ilqr.py:24-25,46-47)
// One CTA per problem (grid-stride), matrices in shared memory, the n x n x n products on the FP64 tensor pipe
// (mm_dmma, common.cuh) when n >= 16.  Every reduction is sequential in ascending k: deterministic.
#include "common.cuh"

namespace srcb {
namespace ctl {

constexpr int NT = 256;

__host__ __device__ inline int up2(int v) { return (v + 1) & ~1; }

template <bool TA, bool TB>
__device__ __forceinline__ void mmx(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int N,
                                    int K, const double* D = nullptr, int ldd = 0) {
    if (M >= 16 && N >= 16 && K >= 16) mm_dmma<NT, TA, TB>(C, ldc, A, lda, B, ldb, M, N, K, D, ldd);
    else mm<NT, TA, TB>(C, ldc, A, lda, B, ldb, M, N, K, D, ldd);
}

// ---------------------------------------------------------------------------------------------------------------
// EKF predict (observer.py:94-104): x <- A_d x + B_d u + d_d ; Sigma <- (A_d Sigma) A_d^T + W
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
ekf_predict_kernel(int n, int m, long long batch, const double* __restrict__ Ad, const double* __restrict__ Bd,
                   const double* __restrict__ dd, const double* __restrict__ u, const double* __restrict__ W,
                   double* __restrict__ x, double* __restrict__ Sigma) {
    extern __shared__ __align__(16) double sm[];
    double* sA = sm;                 // n x n
    double* sS = sA + n * n;         // n x n
    double* sT = sS + n * n;         // n x n
    double* sx = sT + n * n;         // n
    double* su = sx + up2(n);        // m
    const int tid = threadIdx.x;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* A = Ad + b * (long long)n * n;
        for (int e = tid; e < n * n; e += NT) { sA[e] = A[e]; sS[e] = Sigma[b * (long long)n * n + e]; }
        for (int i = tid; i < n; i += NT) sx[i] = x[b * n + i];
        for (int i = tid; i < m; i += NT) su[i] = u[b * m + i];
        __syncthreads();
        // x+ = (A x + B u) + d   (tpwl.py:336-339)
        for (int i = tid; i < n; i += NT) {
            double ax = 0.0, bu = 0.0;
            for (int k = 0; k < n; ++k) ax = fma(sA[i * n + k], sx[k], ax);
            for (int k = 0; k < m; ++k) bu = fma(Bd[b * (long long)n * m + i * m + k], su[k], bu);
            x[b * n + i] = __dadd_rn(__dadd_rn(ax, bu), dd[b * n + i]);
        }
        mmx<false, false>(sT, n, sA, n, sS, n, n, n, n);                 // A Sigma
        __syncthreads();
        mmx<false, true>(sS, n, sT, n, sA, n, n, n, n, W, n);            // W + (A Sigma) A^T
        __syncthreads();
        for (int e = tid; e < n * n; e += NT) Sigma[b * (long long)n * n + e] = sS[e];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// EKF update (observer.py:106-126): y_r = y - y_ref ; S = (C Sigma) C^T + V ; K = (Sigma C^T) inv(S) ;
// x <- x + K (y_r - C x) ; Sigma <- (I - K C) Sigma
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
ekf_update_kernel(int n, int p, long long batch, const double* __restrict__ C, const double* __restrict__ V,
                  const double* __restrict__ yref, const double* __restrict__ y, double* __restrict__ x,
                  double* __restrict__ Sigma) {
    extern __shared__ __align__(16) double sm[];
    double* sS = sm;                         // Sigma n x n
    double* sM = sS + n * n;                 // I - K C, n x n
    double* sO = sM + n * n;                 // new Sigma n x n
    double* sC = sO + n * n;                 // p x n
    double* sCS = sC + up2(p * n);           // C Sigma   p x n
    double* sSC = sCS + up2(p * n);          // Sigma C^T n x p
    double* sK = sSC + up2(p * n);           // n x p
    double* sIn = sK + up2(p * n);           // S p x p (LU in place)
    double* sInv = sIn + up2(p * p);         // inv(S)
    double* sx = sInv + up2(p * p);          // n
    double* sr = sx + up2(n);                // innovation p
    int* piv = reinterpret_cast<int*>(sr + up2(p));
    const int tid = threadIdx.x;
    for (int e = tid; e < p * n; e += NT) sC[e] = C[e];
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        for (int e = tid; e < n * n; e += NT) sS[e] = Sigma[b * (long long)n * n + e];
        for (int i = tid; i < n; i += NT) sx[i] = x[b * n + i];
        __syncthreads();
        mm<NT, false, false>(sCS, n, sC, n, sS, n, p, n, n);             // C Sigma
        mm<NT, false, true>(sSC, p, sS, n, sC, n, n, p, n);              // Sigma C^T
        for (int i = tid; i < p; i += NT) {
            double cx = 0.0;
            for (int k = 0; k < n; ++k) cx = fma(sC[i * n + k], sx[k], cx);
            sr[i] = __dsub_rn(__dsub_rn(y[b * p + i], yref ? yref[i] : 0.0), cx);
        }
        __syncthreads();
        mm<NT, false, true>(sIn, p, sCS, n, sC, n, p, p, n, V, p);       // S = V + (C Sigma) C^T
        __syncthreads();
        lu_inverse<NT>(sIn, sInv, piv, p);
        mm<NT, false, false>(sK, p, sSC, p, sInv, p, n, p, p);           // K = (Sigma C^T) inv(S)
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            double acc = 0.0;
            for (int k = 0; k < p; ++k) acc = fma(sK[i * p + k], sr[k], acc);
            x[b * n + i] = __dadd_rn(sx[i], acc);
        }
        for (int e = tid; e < n * n; e += NT) {                          // I - K C
            const int i = e / n, j = e - i * n;
            double acc = 0.0;
            for (int k = 0; k < p; ++k) acc = fma(sK[i * p + k], sC[k * n + j], acc);
            sM[e] = __dsub_rn(i == j ? 1.0 : 0.0, acc);
        }
        __syncthreads();
        mmx<false, false>(sO, n, sM, n, sS, n, n, n, n);                 // (I - K C) Sigma
        __syncthreads();
        for (int e = tid; e < n * n; e += NT) Sigma[b * (long long)n * n + e] = sO[e];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Discrete algebraic Riccati equation.
// mode 0 -- lqr.py:6-21 `solve_riccati`, literally: value iteration from P = 0 until ||L - L_old||_F <= tol
//           (L = -solve(R + B^T P B, B^T P A)); returns the L and P of the last pass and the number of passes.
// mode 1 -- lqr.py:24-31 `dare`: the stabilising solution to working precision by the structure-preserving
//           doubling algorithm (A_{k+1} = A_k (I + G_k H_k)^-1 A_k, ...; quadratic convergence), then
//           K = -inv(B^T P B + R) (B^T P A).  scipy's solve_discrete_are reaches the same P by a QZ method.
// ---------------------------------------------------------------------------------------------------------------
struct DareSmem { int A, B, Q, R, P, T1, T2, T3, BtP, S, Sinv, L, Lold, red, piv, end; };
__host__ __device__ inline DareSmem dare_plan(int n, int m) {
    DareSmem s; int o = 0;
    auto take = [&o](int c) { const int at = o; o += up2(c); return at; };
    s.A = take(n * n); s.B = take(n * m); s.Q = take(n * n); s.R = take(m * m); s.P = take(n * n);
    s.T1 = take(n * n); s.T2 = take(n * n); s.T3 = take(n * n);
    s.BtP = take(m * n); s.S = take(m * m > n * n ? m * m : n * n); s.Sinv = take(m * m > n * n ? m * m : n * n);
    s.L = take(m * n); s.Lold = take(m * n); s.red = take(NT / 32 + 2); s.piv = take((n > m ? n : m) / 2 + 2);
    s.end = o;
    return s;
}

__device__ double cta_sum(double v, double* red) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(NT, 1)
dare_kernel(int n, int m, long long batch, int shared_cost, const double* __restrict__ Ag, const double* __restrict__ Bg,
            const double* __restrict__ Qg, const double* __restrict__ Rg, double tol, int max_iter, int mode,
            double* __restrict__ Kout, double* __restrict__ Pout, int* __restrict__ iters) {
    extern __shared__ __align__(16) double sm[];
    const DareSmem S = dare_plan(n, m);
    double *A = sm + S.A, *B = sm + S.B, *Q = sm + S.Q, *R = sm + S.R, *P = sm + S.P, *T1 = sm + S.T1, *T2 = sm + S.T2,
           *T3 = sm + S.T3, *BtP = sm + S.BtP, *Sm = sm + S.S, *Sinv = sm + S.Sinv, *L = sm + S.L, *Lold = sm + S.Lold,
           *red = sm + S.red;
    int* piv = reinterpret_cast<int*>(sm + S.piv);
    const int tid = threadIdx.x;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        for (int e = tid; e < n * n; e += NT) { A[e] = Ag[b * (long long)n * n + e]; Q[e] = Qg[(shared_cost ? 0 : b * (long long)n * n) + e]; }
        for (int e = tid; e < n * m; e += NT) B[e] = Bg[b * (long long)n * m + e];
        for (int e = tid; e < m * m; e += NT) R[e] = Rg[(shared_cost ? 0 : b * (long long)m * m) + e];
        __syncthreads();
        int it = 0;
        // gains from P: L = -inv(R + B^T P B) (B^T P A)  [np.linalg.solve], also leaves B^T P A in T3 (m x n), inv in Sinv
        auto gain = [&]() {
            mmx<true, false>(BtP, n, B, m, P, n, m, n, n);               // B^T P
            __syncthreads();
            mm<NT, false, false>(Sm, m, BtP, n, B, m, m, m, n, R, m);    // R + (B^T P) B
            mmx<false, false>(T3, n, BtP, n, A, n, m, n, n);             // (B^T P) A
            __syncthreads();
            lu_inverse<NT>(Sm, Sinv, piv, m);
            for (int e = tid; e < m * n; e += NT) {
                const int i = e / n, j = e - i * n;
                double acc = 0.0;
                for (int k = 0; k < m; ++k) acc = fma(Sinv[i * m + k], T3[k * n + j], acc);
                L[e] = -acc;
            }
            __syncthreads();
        };
        if (mode == 0) {
            for (int e = tid; e < n * n; e += NT) P[e] = 0.0;
            __syncthreads();
            gain();
            // the reference's first L has the opposite sign (lqr.py:14) and is only used in the first norm test
            for (int e = tid; e < m * n; e += NT) L[e] = -L[e];
            bool first = true;
            while (it < max_iter) {
                double d2 = 0.0;
                if (first) d2 = INFINITY;                                // ||L - inf|| = inf > tol (lqr.py:15-16)
                else {
                    double part = 0.0;
                    for (int e = tid; e < m * n; e += NT) { const double d = L[e] - Lold[e]; part = fma(d, d, part); }
                    d2 = cta_sum(part, red);
                }
                if (!(sqrt(d2) > tol)) break;
                first = false;
                for (int e = tid; e < m * n; e += NT) Lold[e] = L[e];
                // P = A^T P A - A^T P B inv(R + B^T P B) (B^T P A) + Q   (lqr.py:18, left to right)
                mmx<true, false>(T1, n, A, n, P, n, n, n, n);            // A^T P
                mmx<true, false>(BtP, n, B, m, P, n, m, n, n);           // B^T P
                __syncthreads();
                mmx<false, false>(T2, n, T1, n, A, n, n, n, n);          // (A^T P) A
                mm<NT, false, false>(Sm, m, BtP, n, B, m, m, m, n, R, m);
                mmx<false, false>(T3, n, BtP, n, A, n, m, n, n);         // B^T P A
                __syncthreads();
                lu_inverse<NT>(Sm, Sinv, piv, m);
                // T1B = (A^T P) B (n x m) -> reuse Sm region?  keep in L scratch: n x m fits m*n
                for (int e = tid; e < n * m; e += NT) {
                    const int i = e / m, j = e - i * m;
                    double acc = 0.0;
                    for (int k = 0; k < n; ++k) acc = fma(T1[i * n + k], B[k * m + j], acc);
                    L[e] = acc;                                          // A^T P B
                }
                __syncthreads();
                for (int e = tid; e < n * m; e += NT) {                  // (A^T P B) inv(S)
                    const int i = e / m, j = e - i * m;
                    double acc = 0.0;
                    for (int k = 0; k < m; ++k) acc = fma(L[i * m + k], Sinv[k * m + j], acc);
                    BtP[e] = acc;                                        // n x m (BtP is m*n doubles: same size)
                }
                __syncthreads();
                for (int e = tid; e < n * n; e += NT) {
                    const int i = e / n, j = e - i * n;
                    double acc = 0.0;
                    for (int k = 0; k < m; ++k) acc = fma(BtP[i * m + k], T3[k * n + j], acc);
                    P[e] = __dadd_rn(__dsub_rn(T2[e], acc), Q[e]);
                }
                __syncthreads();
                gain();
                ++it;
            }
        } else {
            // doubling: A_0 = A, G_0 = B R^-1 B^T, H_0 = Q
            double* G = T1; double* H = P; double* Ak = T2;
            for (int e = tid; e < m * m; e += NT) Sm[e] = R[e];
            __syncthreads();
            lu_inverse<NT>(Sm, Sinv, piv, m);
            for (int e = tid; e < n * m; e += NT) {                      // B R^-1 (n x m) in L
                const int i = e / m, j = e - i * m;
                double acc = 0.0;
                for (int k = 0; k < m; ++k) acc = fma(B[i * m + k], Sinv[k * m + j], acc);
                L[e] = acc;
            }
            __syncthreads();
            mm<NT, false, true>(G, n, L, m, B, m, n, n, m);              // G = (B R^-1) B^T
            for (int e = tid; e < n * n; e += NT) { H[e] = Q[e]; Ak[e] = A[e]; }
            __syncthreads();
            for (it = 0; it < max_iter; ++it) {
                // W = I + G H ; Winv ; A' = A Winv A ; G' = G + A Winv G A^T ; H' = H + A^T H Winv A
                mmx<false, false>(Sm, n, G, n, H, n, n, n, n);
                __syncthreads();
                for (int i = tid; i < n; i += NT) Sm[i * n + i] += 1.0;
                __syncthreads();
                lu_inverse<NT>(Sm, Sinv, piv, n);                         // Sinv = (I + G H)^-1
                mmx<false, false>(T3, n, Sinv, n, Ak, n, n, n, n);        // Winv A
                __syncthreads();
                mmx<false, false>(Sm, n, Sinv, n, G, n, n, n, n);         // Winv G
                __syncthreads();
                // H' = H + A^T (H (Winv A))
                mmx<false, false>(Lold == nullptr ? Sinv : Sinv, n, H, n, T3, n, n, n, n);   // Sinv <- H Winv A
                __syncthreads();
                double part = 0.0, hn = 0.0;
                for (int e = tid; e < n * n; e += NT) {
                    const int i = e / n, j = e - i * n;
                    double acc = 0.0;
                    for (int k = 0; k < n; ++k) acc = fma(Ak[k * n + i], Sinv[k * n + j], acc);
                    part = fma(acc, acc, part);
                    const double hv = H[e] + acc;
                    hn = fma(hv, hv, hn);
                    Q[e] = hv;                                           // H' staged in Q (Q itself is consumed)
                }
                __syncthreads();
                // G' = G + A (Winv G) A^T : first (Winv G) A^T into Sinv, then A * that
                mmx<false, true>(Sinv, n, Sm, n, Ak, n, n, n, n);
                __syncthreads();
                mmx<false, false>(Sm, n, Ak, n, Sinv, n, n, n, n, G, n);  // G + A (Winv G A^T)
                // A' = A (Winv A)
                mmx<false, false>(Sinv, n, Ak, n, T3, n, n, n, n);
                __syncthreads();
                for (int e = tid; e < n * n; e += NT) { G[e] = Sm[e]; Ak[e] = Sinv[e]; H[e] = Q[e]; }
                const double dn = cta_sum(part, red), hh = cta_sum(hn, red);
                __syncthreads();
                if (sqrt(dn) <= tol * sqrt(hh)) { ++it; break; }
            }
            // P = H ; K = -inv(B^T P B + R) (B^T P A)
            gain();
        }
        for (int e = tid; e < m * n; e += NT) Kout[b * (long long)m * n + e] = L[e];
        for (int e = tid; e < n * n; e += NT) Pout[b * (long long)n * n + e] = P[e];
        if (iters && tid == 0) iters[b] = it;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Time-varying LQR recursion (traj_tracking_lqr.py:31-41): P_T = Q; backwards over the `steps` linearisations
//   K_i = -solve(R + B^T P B, B^T P A) ; P <- Q + K^T R K + (A + B K)^T P (A + B K)
// A (batch, steps, n, n), B (batch, steps, n, m) in TIME order -> K (batch, steps, m, n), P (batch, steps + 1, n, n)
// in time order (the reference flips its reversed lists, traj_tracking_lqr.py:43-44).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
tvlqr_kernel(int n, int m, int steps, long long batch, const double* __restrict__ Ag, const double* __restrict__ Bg,
             const double* __restrict__ Qg, const double* __restrict__ Rg, double* __restrict__ Kout,
             double* __restrict__ Pout) {
    extern __shared__ __align__(16) double sm[];
    const DareSmem S = dare_plan(n, m);
    double *A = sm + S.A, *B = sm + S.B, *Q = sm + S.Q, *R = sm + S.R, *P = sm + S.P, *T1 = sm + S.T1, *T2 = sm + S.T2,
           *T3 = sm + S.T3, *BtP = sm + S.BtP, *Sm = sm + S.S, *Sinv = sm + S.Sinv, *L = sm + S.L, *Lold = sm + S.Lold;
    int* piv = reinterpret_cast<int*>(sm + S.piv);
    const int tid = threadIdx.x;
    for (int e = tid; e < n * n; e += NT) Q[e] = Qg[e];
    for (int e = tid; e < m * m; e += NT) R[e] = Rg[e];
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < n * n; e += NT) { P[e] = Q[e]; Pout[(b * (steps + 1) + steps) * (long long)n * n + e] = Q[e]; }
        for (int t = steps - 1; t >= 0; --t) {
            for (int e = tid; e < n * n; e += NT) A[e] = Ag[(b * steps + t) * (long long)n * n + e];
            for (int e = tid; e < n * m; e += NT) B[e] = Bg[(b * steps + t) * (long long)n * m + e];
            __syncthreads();
            mmx<true, false>(BtP, n, B, m, P, n, m, n, n);               // B^T P
            __syncthreads();
            mm<NT, false, false>(Sm, m, BtP, n, B, m, m, m, n, R, m);    // R + B^T P B
            mmx<false, false>(T3, n, BtP, n, A, n, m, n, n);             // B^T P A
            __syncthreads();
            lu_inverse<NT>(Sm, Sinv, piv, m);
            for (int e = tid; e < m * n; e += NT) {
                const int i = e / n, j = e - i * n;
                double acc = 0.0;
                for (int k = 0; k < m; ++k) acc = fma(Sinv[i * m + k], T3[k * n + j], acc);
                L[e] = -acc;
                Kout[(b * steps + t) * (long long)m * n + e] = -acc;
            }
            __syncthreads();
            mm<NT, false, false>(T1, n, B, m, L, n, n, n, m, A, n);      // A + B K
            mm<NT, false, false>(Lold, n, R, m, L, n, m, n, m);          // R K
            __syncthreads();
            mmx<true, false>(T2, n, T1, n, P, n, n, n, n);               // (A + B K)^T P
            mm<NT, true, false>(T3, n, L, n, Lold, n, n, n, m, Q, n);    // Q + K^T (R K)
            __syncthreads();
            mmx<false, false>(P, n, T2, n, T1, n, n, n, n, T3, n);       // (Q + K^T R K) + ((A+BK)^T P)(A+BK)
            __syncthreads();
            for (int e = tid; e < n * n; e += NT) Pout[(b * (steps + 1) + t) * (long long)n * n + e] = P[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One stored TPWL point from reduced second-order matrices (utils.py:251-286 dense branch; tpwl_utils.py:263-276):
//   Minv = inv(M) ; A = [[-Minv D, -Minv K], [I, 0]] ; B = [[Minv H], [0]] ; d = [solve(M, f + K q) ; 0]
// K, D, M (count, r, r), H (count, r, m), f, q (count, r) or NULL -> A (count, 2r, 2r), B (count, 2r, m), d (count, 2r)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
bank_point_kernel(int r, int m, long long count, const double* __restrict__ Kg, const double* __restrict__ Dg,
                  const double* __restrict__ Mg, const double* __restrict__ Hg, const double* __restrict__ fg,
                  const double* __restrict__ qg, double* __restrict__ Ao, double* __restrict__ Bo, double* __restrict__ dout) {
    extern __shared__ __align__(16) double sm[];
    double* sM = sm;                 // r x r (LU in place)
    double* sI = sM + r * r;         // inv(M)
    double* sK = sI + r * r;
    double* sD = sK + r * r;
    double* sv = sD + r * r;         // f + K q
    int* piv = reinterpret_cast<int*>(sv + up2(r));
    const int tid = threadIdx.x, n = 2 * r;
    for (long long b = blockIdx.x; b < count; b += gridDim.x) {
        for (int e = tid; e < r * r; e += NT) {
            sM[e] = Mg[b * (long long)r * r + e]; sK[e] = Kg[b * (long long)r * r + e]; sD[e] = Dg[b * (long long)r * r + e];
        }
        __syncthreads();
        if (dout && fg && qg) {
            for (int i = tid; i < r; i += NT) {
                double acc = 0.0;
                for (int k = 0; k < r; ++k) acc = fma(sK[i * r + k], qg[b * r + k], acc);
                sv[i] = __dadd_rn(fg[b * r + i], acc);
            }
        }
        lu_inverse<NT>(sM, sI, piv, r);
        double* A = Ao + b * (long long)n * n;
        for (int e = tid; e < r * r; e += NT) {
            const int i = e / r, j = e - i * r;
            double ad = 0.0, ak = 0.0;
            for (int k = 0; k < r; ++k) { ad = fma(sI[i * r + k], sD[k * r + j], ad); ak = fma(sI[i * r + k], sK[k * r + j], ak); }
            A[i * n + j] = -ad;
            A[i * n + r + j] = -ak;
            A[(r + i) * n + j] = (i == j) ? 1.0 : 0.0;
            A[(r + i) * n + r + j] = 0.0;
        }
        for (int e = tid; e < r * m; e += NT) {
            const int i = e / m, j = e - i * m;
            double acc = 0.0;
            for (int k = 0; k < r; ++k) acc = fma(sI[i * r + k], Hg[b * (long long)r * m + k * m + j], acc);
            Bo[b * (long long)n * m + i * m + j] = acc;
            Bo[b * (long long)n * m + (r + i) * m + j] = 0.0;
        }
        if (dout && fg && qg) {
            for (int i = tid; i < r; i += NT) {
                double acc = 0.0;
                for (int k = 0; k < r; ++k) acc = fma(sI[i * r + k], sv[k], acc);
                dout[b * n + i] = acc;
                dout[b * n + r + i] = 0.0;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GuSTO model accuracy (gusto.py:203-223) per trajectory: with (f_k, A_k, B_k) at the previous iterate (x_k, u_k) and
// f at the candidate (x, u):  error = sum_i dt || s o (f_i - fa_i) ||_2, approx = sum_i dt || s o fa_i ||_2,
// fa_i = f_k,i + A_k,i (x_i - x_k,i) + B_k,i (u_i - u_k,i);  rho = error / (J + approx).  One warp per trajectory.
// ---------------------------------------------------------------------------------------------------------------
__global__ void gusto_accuracy_kernel(int n, int m, int N, long long batch, double dt, const double* __restrict__ fk,
                                      const double* __restrict__ Ak, const double* __restrict__ Bk,
                                      const double* __restrict__ f, const double* __restrict__ x,
                                      const double* __restrict__ xk, const double* __restrict__ u,
                                      const double* __restrict__ uk, const double* __restrict__ fscale,
                                      const double* __restrict__ J, double* __restrict__ rho, double* __restrict__ err_o,
                                      double* __restrict__ approx_o) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    double error = 0.0, approx = 0.0;
    for (int i = 0; i < N; ++i) {
        const long long p = b * N + i;                   // linearisation points: N per trajectory
        const long long px = b * (long long)(N + 1) + i; // states: N + 1 per trajectory
        double e2 = 0.0, a2 = 0.0;
        for (int r = lane; r < n; r += 32) {
            double acc = fk[p * n + r];
            double ax = 0.0, bu = 0.0;
            for (int k = 0; k < n; ++k) ax = fma(Ak[p * (long long)n * n + r * n + k], __dsub_rn(x[px * n + k], xk[px * n + k]), ax);
            for (int k = 0; k < m; ++k) bu = fma(Bk[p * (long long)n * m + r * m + k], __dsub_rn(u[p * m + k], uk[p * m + k]), bu);
            acc = __dadd_rn(__dadd_rn(acc, ax), bu);
            const double s = fscale ? fscale[r] : 1.0;
            const double de = s * (f[p * n + r] - acc), da = s * acc;
            e2 = fma(de, de, e2);
            a2 = fma(da, da, a2);
        }
        for (int off = 16; off > 0; off >>= 1) { e2 += __shfl_xor_sync(0xffffffffu, e2, off); a2 += __shfl_xor_sync(0xffffffffu, a2, off); }
        error = fma(dt, sqrt(e2), error);
        approx = fma(dt, sqrt(a2), approx);
    }
    if (lane == 0) {
        if (err_o) err_o[b] = error;
        if (approx_o) approx_o[b] = approx;
        rho[b] = error / ((J ? J[b] : 0.0) + approx);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Receding-horizon glue (the warm-start hooks ilqr.py:24-25, 46-47 driven every control step):
//   u_applied = u_plan[:, 0] ; u_warm = [u_plan[:, 1:], u_plan[:, -1]] ; z_target window [k+1, k+1+N] of the
//   reference ; per-problem bookkeeping of the closed-loop record.  One thread per element.
// ---------------------------------------------------------------------------------------------------------------
__global__ void mpc_shift_kernel(long long batch, int N, int m, int nz, int T, int k, const double* __restrict__ u_plan,
                                 const double* __restrict__ z_ref, double* __restrict__ u_warm,
                                 double* __restrict__ u_applied, double* __restrict__ z_window,
                                 double* __restrict__ u_log) {
    const long long per = (long long)N * m + (long long)(N + 1) * nz;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= batch * per) return;
    const long long b = idx / per;
    const long long e = idx - b * per;
    if (e < (long long)N * m) {
        const int t = (int)(e / m), j = (int)(e - (long long)t * m);
        const int ts = t + 1 < N ? t + 1 : N - 1;
        u_warm[b * (long long)N * m + e] = u_plan[b * (long long)N * m + (long long)ts * m + j];
        if (t == 0) {
            const double u0 = u_plan[b * (long long)N * m + j];
            u_applied[b * m + j] = u0;
            if (u_log) u_log[(b * T + k) * (long long)m + j] = u0;
        }
    } else {
        const long long ez = e - (long long)N * m;
        const int t = (int)(ez / nz), j = (int)(ez - (long long)t * nz);
        z_window[b * (long long)(N + 1) * nz + ez] = z_ref[(b * (long long)(T + N + 1) + k + 1 + t) * nz + j];
    }
}

static int grid_for(long long batch, int per_sm = 1) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long g = (long long)sms * per_sm;
    return (int)(batch < g ? (batch > 0 ? batch : 1) : g);
}

}  // namespace ctl
}  // namespace srcb

using namespace srcb;

#define CTL_SMEM(kernel, bytes)                                                                                  \
    do {                                                                                                         \
        if ((bytes) > 227 * 1024) return fail(SRCB200_E_DIM, "dimensions need %zu B of shared memory (> 227 KB)", (size_t)(bytes)); \
        SRCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
    } while (0)

extern "C" int srcb200_ekf_predict_batch(int32_t n, int32_t m, int64_t batch, const double* A_d, const double* B_d,
                                         const double* d_d, const double* u, const double* W, double* x, double* Sigma,
                                         void* stream) {
    if (n <= 0 || m <= 0 || batch < 0) return fail(SRCB200_E_DIM, "ekf_predict: n, m must be positive");
    if (batch == 0) return 0;
    if (!A_d || !B_d || !d_d || !u || !W || !x || !Sigma) return fail(SRCB200_E_NULL, "ekf_predict: NULL argument");
    const size_t smem = sizeof(double) * (3 * (size_t)n * n + ctl::up2(n) + ctl::up2(m));
    CTL_SMEM(ctl::ekf_predict_kernel, smem);
    ctl::ekf_predict_kernel<<<ctl::grid_for(batch), ctl::NT, smem, (cudaStream_t)stream>>>(n, m, batch, A_d, B_d, d_d, u, W, x, Sigma);
    SRCB_LAUNCH_CHECK("ekf_predict_kernel");
    return 0;
}

extern "C" int srcb200_ekf_update_batch(int32_t n, int32_t p, int64_t batch, const double* C, const double* V,
                                        const double* y_ref, const double* y, double* x, double* Sigma, void* stream) {
    if (n <= 0 || p <= 0 || batch < 0) return fail(SRCB200_E_DIM, "ekf_update: n, p must be positive");
    if (batch == 0) return 0;
    if (!C || !V || !y || !x || !Sigma) return fail(SRCB200_E_NULL, "ekf_update: NULL argument");
    const size_t smem = sizeof(double) * (3 * (size_t)n * n + 4 * ctl::up2(p * n) + 2 * ctl::up2(p * p) + ctl::up2(n) +
                                          ctl::up2(p) + p / 2 + 4);
    CTL_SMEM(ctl::ekf_update_kernel, smem);
    ctl::ekf_update_kernel<<<ctl::grid_for(batch), ctl::NT, smem, (cudaStream_t)stream>>>(n, p, batch, C, V, y_ref, y, x, Sigma);
    SRCB_LAUNCH_CHECK("ekf_update_kernel");
    return 0;
}

extern "C" int srcb200_dlqr_riccati_batch(int32_t n, int32_t m, int64_t batch, const double* A, const double* B,
                                          const double* Q, const double* R, int32_t shared_cost, double tol,
                                          int32_t max_iter, int32_t mode, double* K, double* P, int32_t* iterations,
                                          void* stream) {
    if (n <= 0 || m <= 0 || batch < 0 || (mode != 0 && mode != 1)) return fail(SRCB200_E_DIM, "dlqr_riccati: bad dimensions / mode");
    if (batch == 0) return 0;
    if (!A || !B || !Q || !R || !K || !P) return fail(SRCB200_E_NULL, "dlqr_riccati: NULL argument");
    const size_t smem = sizeof(double) * (size_t)ctl::dare_plan(n, m).end;
    CTL_SMEM(ctl::dare_kernel, smem);
    ctl::dare_kernel<<<ctl::grid_for(batch), ctl::NT, smem, (cudaStream_t)stream>>>(n, m, batch, shared_cost, A, B, Q, R, tol,
                                                                                       max_iter, mode, K, P, iterations);
    SRCB_LAUNCH_CHECK("dare_kernel");
    return 0;
}

extern "C" int srcb200_tvlqr_batch(int32_t n, int32_t m, int32_t steps, int64_t batch, const double* A, const double* B,
                                   const double* Q, const double* R, double* K, double* P, void* stream) {
    if (n <= 0 || m <= 0 || steps < 0 || batch < 0) return fail(SRCB200_E_DIM, "tvlqr: bad dimensions");
    if (batch == 0) return 0;
    if (!Q || !R || !K || !P || (steps > 0 && (!A || !B))) return fail(SRCB200_E_NULL, "tvlqr: NULL argument");
    const size_t smem = sizeof(double) * (size_t)ctl::dare_plan(n, m).end;
    CTL_SMEM(ctl::tvlqr_kernel, smem);
    ctl::tvlqr_kernel<<<ctl::grid_for(batch), ctl::NT, smem, (cudaStream_t)stream>>>(n, m, steps, batch, A, B, Q, R, K, P);
    SRCB_LAUNCH_CHECK("tvlqr_kernel");
    return 0;
}

extern "C" int srcb200_tpwl_bank_point_batch(int32_t r, int32_t m, int64_t count, const double* K, const double* D,
                                             const double* M, const double* H, const double* f, const double* q,
                                             double* A_c, double* B_c, double* d_c, void* stream) {
    if (r <= 0 || m <= 0 || count < 0) return fail(SRCB200_E_DIM, "bank_point: bad dimensions");
    if (count == 0) return 0;
    if (!K || !D || !M || !H || !A_c || !B_c) return fail(SRCB200_E_NULL, "bank_point: NULL argument");
    const size_t smem = sizeof(double) * (4 * (size_t)r * r + ctl::up2(r) + r / 2 + 4);
    CTL_SMEM(ctl::bank_point_kernel, smem);
    ctl::bank_point_kernel<<<ctl::grid_for(count, 2), ctl::NT, smem, (cudaStream_t)stream>>>(r, m, count, K, D, M, H, f, q, A_c, B_c, d_c);
    SRCB_LAUNCH_CHECK("bank_point_kernel");
    return 0;
}

extern "C" int srcb200_gusto_accuracy_batch(int32_t n, int32_t m, int32_t N, int64_t batch, double dt, const double* f_k,
                                            const double* A_k, const double* B_k, const double* f, const double* x,
                                            const double* x_k, const double* u, const double* u_k, const double* f_scale,
                                            const double* J, double* rho, double* error, double* approx, void* stream) {
    if (n <= 0 || m <= 0 || N < 0 || batch < 0) return fail(SRCB200_E_DIM, "gusto_accuracy: bad dimensions");
    if (batch == 0) return 0;
    if (!f_k || !A_k || !B_k || !f || !x || !x_k || !u || !u_k || !rho) return fail(SRCB200_E_NULL, "gusto_accuracy: NULL argument");
    const int wpb = 4;
    ctl::gusto_accuracy_kernel<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        n, m, N, batch, dt, f_k, A_k, B_k, f, x, x_k, u, u_k, f_scale, J, rho, error, approx);
    SRCB_LAUNCH_CHECK("gusto_accuracy_kernel");
    return 0;
}

extern "C" int srcb200_mpc_shift_batch(int64_t batch, int32_t N, int32_t m, int32_t nz, int32_t T, int32_t k,
                                       const double* u_plan, const double* z_ref, double* u_warm, double* u_applied,
                                       double* z_window, double* u_log, void* stream) {
    if (batch < 0 || N <= 0 || m <= 0 || nz <= 0 || T <= 0 || k < 0 || k >= T) return fail(SRCB200_E_DIM, "mpc_shift: bad dimensions");
    if (batch == 0) return 0;
    if (!u_plan || !z_ref || !u_warm || !u_applied || !z_window) return fail(SRCB200_E_NULL, "mpc_shift: NULL argument");
    const long long total = batch * ((long long)N * m + (long long)(N + 1) * nz);
    ctl::mpc_shift_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(batch, N, m, nz, T, k, u_plan, z_ref,
                                                                                                u_warm, u_applied, z_window, u_log);
    SRCB_LAUNCH_CHECK("mpc_shift_kernel");
    return 0;
}
