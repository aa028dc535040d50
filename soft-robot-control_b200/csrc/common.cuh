// common.cuh -- shared host/device helpers for libsrcb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/srcb200.h"

namespace srcb {

// ---- host-side error plumbing -------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  fail(int code, const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define SRCB_CUDA(call)                                                   \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return ::srcb::cuda_fail(e__, #call);     \
    } while (0)

#define SRCB_LAUNCH_CHECK(name)                                           \
    do {                                                                  \
        cudaError_t e__ = cudaGetLastError();                             \
        if (e__ != cudaSuccess) return ::srcb::cuda_fail(e__, name);      \
    } while (0)

// ---- device: mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) for contiguous global -> shared rows ---------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// ---- device: CTA-cooperative small dense linear algebra ------------------------------------------------------
// All routines are called by every thread of the CTA; NT is the CTA size.  A one-warp CTA synchronises with
// __syncwarp(), larger CTAs with __syncthreads().  Operands may live in shared or global memory (generic
// pointers).  Every dot product is accumulated sequentially in ascending k with FMA: deterministic, and the
// same order on every launch.
template <int NT>
__device__ __forceinline__ void cta_sync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}

// C[i*ldc+j] = (D ? D[i*ldd+j] : 0) + sum_k opA(i,k) * opB(k,j),  i<M, j<N, k<K
// TA: A is stored K x M (use A^T).  TB: B is stored N x K (use B^T).
template <int NT, bool TA, bool TB>
__device__ __forceinline__ void mm(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda,
                                   const double* __restrict__ B, int ldb, int M, int N, int K,
                                   const double* __restrict__ D = nullptr, int ldd = 0) {
    for (int e = threadIdx.x; e < M * N; e += NT) {
        const int i = e / N, j = e - i * N;
        double acc = 0.0;
        for (int k = 0; k < K; ++k) {
            const double a = TA ? A[k * lda + i] : A[i * lda + k];
            const double b = TB ? B[j * ldb + k] : B[k * ldb + j];
            acc = fma(a, b, acc);
        }
        C[i * ldc + j] = D ? (D[i * ldd + j] + acc) : acc;
    }
}

// Same contract as mm<>, on the FP64 tensor pipe: each warp owns 8 x 8 output tiles (two at a time for ILP) and
// walks K in steps of 4 with mma.sync.m8n8k4.f64; fragment loads are masked, so any M, N, K works.  Operands may be
// in shared or global memory.  Used for the n x n x n products of the Riccati sweep when n is large (TPWL, n = 72).
__device__ __forceinline__ void dmma_m8n8k4_acc(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NT, bool TA, bool TB>
__device__ __forceinline__ void mm_dmma(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda,
                                        const double* __restrict__ B, int ldb, int M, int N, int K,
                                        const double* __restrict__ D = nullptr, int ldd = 0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, NW = NT / 32;
    const int tm = (M + 7) / 8, tn = (N + 7) / 8, tn2 = (tn + 1) / 2;
    for (int tile = warp; tile < tm * tn2; tile += NW) {
        const int i0 = (tile / tn2) * 8, j0 = (tile % tn2) * 16;
        const int row = i0 + g, colA = j0 + g, colB = j0 + 8 + g;
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
        for (int k0 = 0; k0 < K; k0 += 4) {
            const int k = k0 + q;
            const bool kok = k < K;
            const double a = (row < M && kok) ? (TA ? A[k * lda + row] : A[row * lda + k]) : 0.0;
            const double b0 = (colA < N && kok) ? (TB ? B[colA * ldb + k] : B[k * ldb + colA]) : 0.0;
            const double b1 = (colB < N && kok) ? (TB ? B[colB * ldb + k] : B[k * ldb + colB]) : 0.0;
            dmma_m8n8k4_acc(c00, c01, a, b0);
            dmma_m8n8k4_acc(c10, c11, a, b1);
        }
        if (row < M) {
            const int ca = j0 + 2 * q, cb = j0 + 8 + 2 * q;
            if (ca < N)     C[row * ldc + ca]     = D ? (D[row * ldd + ca] + c00) : c00;
            if (ca + 1 < N) C[row * ldc + ca + 1] = D ? (D[row * ldd + ca + 1] + c01) : c01;
            if (cb < N)     C[row * ldc + cb]     = D ? (D[row * ldd + cb] + c10) : c10;
            if (cb + 1 < N) C[row * ldc + cb + 1] = D ? (D[row * ldd + cb + 1] + c11) : c11;
        }
    }
}

// y[i] = (d ? d[i] : 0) + sum_k opA(i,k) x[k]
template <int NT, bool TA>
__device__ __forceinline__ void mv(double* __restrict__ y, const double* __restrict__ A, int lda,
                                   const double* __restrict__ x, int M, int K,
                                   const double* __restrict__ d = nullptr) {
    for (int i = threadIdx.x; i < M; i += NT) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc = fma(TA ? A[k * lda + i] : A[i * lda + k], x[k], acc);
        y[i] = d ? (d[i] + acc) : acc;
    }
}

// In-place LU with partial pivoting (first maximal |a_ik|, like LAPACK idamax) of the n x n matrix M (ld = n)
// followed by the inverse through forward/back substitution on the permuted identity (numpy.linalg.inv ->
// LAPACK gesv(A, I)).  piv: n ints of scratch.  inv: n x n output.  Returns nothing; a zero pivot yields inf/nan
// exactly like the singular-matrix path would (numpy raises there; callers treat non-finite results as failure).
template <int NT>
__device__ void lu_inverse(double* __restrict__ M, double* __restrict__ inv, int* __restrict__ piv, int n) {
    const int tid = threadIdx.x;
    for (int c = 0; c < n; ++c) {
        if (tid == 0) {
            int p = c;
            double best = fabs(M[c * n + c]);
            for (int r = c + 1; r < n; ++r) {
                const double v = fabs(M[r * n + c]);
                if (v > best) { best = v; p = r; }
            }
            piv[c] = p;
        }
        cta_sync<NT>();
        const int p = piv[c];
        if (p != c) {
            for (int j = tid; j < n; j += NT) {
                const double t = M[c * n + j];
                M[c * n + j] = M[p * n + j];
                M[p * n + j] = t;
            }
        }
        cta_sync<NT>();
        const double rp = 1.0 / M[c * n + c];
        for (int r = c + 1 + tid; r < n; r += NT) M[r * n + c] *= rp;   // LAPACK getf2 scales by the reciprocal
        cta_sync<NT>();
        const int rem = n - c - 1;
        for (int e = tid; e < rem * rem; e += NT) {
            const int r = c + 1 + e / rem, j = c + 1 + e % rem;
            M[r * n + j] = fma(-M[r * n + c], M[c * n + j], M[r * n + j]);
        }
        cta_sync<NT>();
    }
    // columns of the inverse: thread j solves L U x = P e_j
    for (int j = tid; j < n; j += NT) {
        // b = P e_j: the row swaps act on b; track where the single 1 ends up
        int pos = j;
        for (int c = 0; c < n; ++c) {
            const int p = piv[c];
            if (pos == c) pos = p; else if (pos == p) pos = c;
        }
        // forward substitution (unit lower)
        for (int i = 0; i < n; ++i) {
            double s = (i == pos) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s = fma(-M[i * n + k], inv[k * n + j], s);
            inv[i * n + j] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = inv[i * n + j];
            for (int k = i + 1; k < n; ++k) s = fma(-M[i * n + k], inv[k * n + j], s);
            inv[i * n + j] = s / M[i * n + i];
        }
    }
    cta_sync<NT>();
}

// Positive-definiteness test by the unblocked lower Cholesky of LAPACK dpotf2 (what numpy.linalg.cholesky runs;
// ilqr.py:276-280 uses it only as a PD test).  Reads the LOWER triangle of A (n x n, ld n), L is n x n scratch.
// Returns true iff every pivot is > 0 and finite.  All threads get the same answer.
template <int NT>
__device__ bool cholesky_pd(const double* __restrict__ A, double* __restrict__ L, int* __restrict__ flag, int n) {
    const int tid = threadIdx.x;
    if (tid == 0) *flag = 1;
    cta_sync<NT>();
    for (int j = 0; j < n; ++j) {
        for (int i = j + tid; i < n; i += NT) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s = fma(-L[i * n + k], L[j * n + k], s);
            L[i * n + j] = s;
        }
        cta_sync<NT>();
        const double ajj = L[j * n + j];
        if (!(ajj > 0.0) || isinf(ajj)) {   // uniform branch: every thread reads the same value
            cta_sync<NT>();
            if (tid == 0) *flag = 0;
            cta_sync<NT>();
            return false;
        }
        const double rj = sqrt(ajj);
        cta_sync<NT>();
        for (int i = j + tid; i < n; i += NT) L[i * n + j] = (i == j) ? rj : L[i * n + j] / rj;
        cta_sync<NT>();
    }
    return true;
}

// A_d, B_d, d_d from continuous (A, B, d) -- tpwl.py:272-297 / ssm.py:279-301 (fe, be, bil).
// A (n x n), B (n x m), d (n) are overwritten in place.  scratch: discretize_scratch_doubles(n, m) doubles.
// Elementwise steps use explicit single roundings (__dmul_rn/__dadd_rn) so they round like the numpy expressions.
__host__ __device__ inline int discretize_sq(int n, int m) { return n * n > n * (m + 1) ? n * n : n * (m + 1); }
__host__ __device__ inline int discretize_scratch_doubles(int n, int m) { return 3 * discretize_sq(n, m) + (n + 1) / 2 + 1; }
template <int NT>
__device__ void discretize_inplace(int method, double dt, double* __restrict__ A, double* __restrict__ B,
                                   double* __restrict__ d, int n, int m, double* __restrict__ scratch) {
    const int tid = threadIdx.x;
    if (method == SRCB200_DISCR_FE) {
        for (int e = tid; e < n * n; e += NT) {
            const int i = e / n, j = e - i * n;
            const double v = __dmul_rn(dt, A[e]);
            A[e] = (i == j) ? __dadd_rn(1.0, v) : v;
        }
        for (int e = tid; e < n * m; e += NT) B[e] = __dmul_rn(dt, B[e]);
        for (int e = tid; e < n; e += NT) d[e] = __dmul_rn(dt, d[e]);
        cta_sync<NT>();
        return;
    }
    if (method != SRCB200_DISCR_BE && method != SRCB200_DISCR_BIL) return;
    const int sq = discretize_sq(n, m);
    double* W   = scratch;           // matrix to invert / LU workspace / copy of (B_c, d_c)
    double* Inv = scratch + sq;      // its inverse
    double* Ad  = scratch + 2 * sq;  // A_d
    int*    piv = reinterpret_cast<int*>(scratch + 3 * sq);
    const double h = (method == SRCB200_DISCR_BE) ? dt : 0.5 * dt;
    // W = I - h A
    for (int e = tid; e < n * n; e += NT) {
        const int i = e / n, j = e - i * n;
        W[e] = __dsub_rn((i == j) ? 1.0 : 0.0, __dmul_rn(h, A[e]));
    }
    cta_sync<NT>();
    lu_inverse<NT>(W, Inv, piv, n);
    if (method == SRCB200_DISCR_BE) {
        for (int e = tid; e < n * n; e += NT) Ad[e] = Inv[e];
    } else {
        // A_d = (I + h A) @ inv(I - h A)
        for (int e = tid; e < n * n; e += NT) {
            const int i = e / n, j = e - i * n;
            W[e] = __dadd_rn((i == j) ? 1.0 : 0.0, __dmul_rn(h, A[e]));
        }
        cta_sync<NT>();
        mm<NT, false, false>(Ad, n, W, n, Inv, n, n, n, n);
    }
    cta_sync<NT>();
    // sep = inv(A) @ (A_d - I)
    for (int e = tid; e < n * n; e += NT) W[e] = A[e];
    cta_sync<NT>();
    lu_inverse<NT>(W, Inv, piv, n);       // Inv = inv(A_c)
    for (int e = tid; e < n * n; e += NT) {
        const int i = e / n, j = e - i * n;
        W[e] = __dsub_rn(Ad[e], (i == j) ? 1.0 : 0.0);
    }
    cta_sync<NT>();
    mm<NT, false, false>(A, n, Inv, n, W, n, n, n, n);   // A := sep (A_c no longer needed)
    cta_sync<NT>();
    // B_d = sep @ B_c ; d_d = sep @ d_c   (need B_c, d_c intact while reading: go through W)
    for (int e = tid; e < n * m; e += NT) W[e] = B[e];
    for (int e = tid; e < n; e += NT) W[n * m + e] = d[e];
    cta_sync<NT>();
    mm<NT, false, false>(B, m, A, n, W, m, n, m, n);
    mv<NT, false>(d, A, n, W + n * m, n, n);
    cta_sync<NT>();
    for (int e = tid; e < n * n; e += NT) A[e] = Ad[e];
    cta_sync<NT>();
}

}  // namespace srcb
