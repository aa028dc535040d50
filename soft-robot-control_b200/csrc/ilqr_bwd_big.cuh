// ilqr_bwd_big.cuh -- backward Riccati sweep (sofacontrol/lqr/ilqr.py:219-300) for a large state dimension (TPWL:
// n = 72, m = 4) on the FP64 tensor pipe.  Included by ilqr_impl.cuh; one CTA of NT threads per problem.
//
// One step of the sweep is three DMMA products over "extended" operands that live in shared memory with a
// bank-conflict-free leading dimension LD (LD mod 16 in {4, 12}):
//
//   A' = [A_t | B_t]            n x (n+m)    (cp.async-prefetched for step t-1 while step t computes; double buffer)
//   P' = [P | p]                n x (n+1)
//
//   1.  S1 = A'^T P'            (n+m) x (n+1):  [A^T P | A^T p ; B^T P | B^T p]          (ilqr.py:258-262, 266)
//       + rows  B^T (P + rho I) = B^T P + rho B^T   (state regularisation, ilqr.py:266-267)
//   2.  Q2 = [B^T P ; B^T(P+rho I)] A'   2m x (n+m):  [Q_ux | Q_uu - R ; Q_ux~ | Q_uu~ - R]   (ilqr.py:260-269)
//       Cholesky PD test of Q_uu~, explicit inverse, K = -inv Q_ux~, k = -inv Q_u     (ilqr.py:276-292)
//   3.  P' <- [c_xx | Q_x] + [A^T P | K^T Q_uu | K^T | Q_ux^T] [A | 0 ; K | k ; Q_ux | Q_u ; K | k]   (ilqr.py:294-295)
//
// so every matrix-matrix and matrix-vector product of the step runs as m8n8k4 DMMAs with 2 x 2 register blocking and
// no operand ever comes from global memory inside a product.
#pragma once

namespace srcb {

__host__ __device__ constexpr int cf_ld(int v) {           // smallest leading dimension >= v without LDS.64 bank conflicts
    while ((v & 15) != 4 && (v & 15) != 12) ++v;
    return v;
}

struct BigPlan {
    int LD, L3;
    int Pp, Ap0, Ap1, S1, Q2, L3b, R3b, T1, T1f, se, pf0, pf1, Qx, Qu, Quu, Quut, Lc, LU, inv, end;
};

__host__ __device__ inline BigPlan make_big(int n, int m, int nz, int base) {
    BigPlan B;
    int o = (base + 1) & ~1;
    auto take = [&o](int cnt) { const int at = o; o += (cnt + 1) & ~1; return at; };
    B.LD = cf_ld(n + m);
    B.L3 = cf_ld(3 * m);
    B.Pp = take(n * B.LD);
    B.Ap0 = take(n * B.LD);
    B.Ap1 = take(n * B.LD);
    B.S1 = take((n + 2 * m) * B.LD);
    B.Q2 = take(2 * m * B.LD);
    B.L3b = take(n * B.L3);
    B.R3b = take(3 * m * B.LD);
    B.T1 = take(n * nz);
    B.T1f = take(n * nz);
    B.se = take(nz);
    B.pf0 = take(nz + 2 * m);  B.pf1 = take(nz + 2 * m);          // prefetched e_t, u_t, u_{t-1}
    B.Qx = take(n);   B.Qu = take(m);
    B.Quu = take(m * m);  B.Quut = take(m * m);  B.Lc = take(m * m);  B.LU = take(m * m);  B.inv = take(m * m);
    B.end = o;
    return B;
}

__host__ __device__ inline int big_plan_end(int n, int m, int nz, int base) { return make_big(n, m, nz, base).end; }

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src));
}
__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// A' = [A | B] of one step into shared memory (asynchronous; completes at the next cp_async_wait_all + barrier)
template <int NT>
__device__ __forceinline__ void load_lin_async(double* __restrict__ Ap, int LD, const double* __restrict__ A,
                                               const double* __restrict__ B, int n, int m) {
    const int tid = threadIdx.x;
    const bool v16 = (((n | m | LD) & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
    if (v16) {
        const int hn = n >> 1, hm = m >> 1;
        for (int e = tid; e < n * hn; e += NT) {
            const int i = e / hn, j = (e - i * hn) * 2;
            cp_async16(Ap + i * LD + j, A + i * n + j);
        }
        for (int e = tid; e < n * hm; e += NT) {
            const int i = e / hm, j = (e - i * hm) * 2;
            cp_async16(Ap + i * LD + n + j, B + i * m + j);
        }
    } else {
        for (int e = tid; e < n * n; e += NT) {
            const int i = e / n, j = e - i * n;
            cp_async8(Ap + i * LD + j, A + e);
        }
        for (int e = tid; e < n * m; e += NT) {
            const int i = e / m, j = e - i * m;
            cp_async8(Ap + i * LD + n + j, B + e);
        }
    }
}

// C = op(A) op(B) by 16 x 16 blocks (2 x 2 DMMA tiles) handed to the warps round-robin.  a(r, k) / b(k, c) fetch one
// operand element (masking is theirs), store(r, c, v0, v1) receives the elements (r, c) and (r, c + 1).
template <class AF, class BF, class ST>
__device__ __forceinline__ void dmma_blocks(int M, int N, int K, AF a, BF b, ST store, int nwarps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int tm = (M + 7) >> 3, tn = (N + 7) >> 3, bm = (tm + 1) >> 1, bn = (tn + 1) >> 1;
    for (int blk = warp; blk < bm * bn; blk += nwarps) {
        const int bi = blk / bn;
        const int i0 = bi * 16, j0 = (blk - bi * bn) * 16;
        const bool r1 = i0 + 8 < M, c1 = j0 + 8 < N;
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0, c20 = 0.0, c21 = 0.0, c30 = 0.0, c31 = 0.0;
#pragma unroll 2
        for (int k0 = 0; k0 < K; k0 += 4) {
            const int k = k0 + q;
            const double a0 = a(i0 + g, k);
            const double b0 = b(k, j0 + g);
            const double a1 = r1 ? a(i0 + 8 + g, k) : 0.0;
            const double b1 = c1 ? b(k, j0 + 8 + g) : 0.0;
            dmma_m8n8k4_acc(c00, c01, a0, b0);
            if (c1) dmma_m8n8k4_acc(c10, c11, a0, b1);
            if (r1) dmma_m8n8k4_acc(c20, c21, a1, b0);
            if (r1 && c1) dmma_m8n8k4_acc(c30, c31, a1, b1);
        }
        store(i0 + g, j0 + 2 * q, c00, c01);
        if (c1) store(i0 + g, j0 + 8 + 2 * q, c10, c11);
        if (r1) store(i0 + 8 + g, j0 + 2 * q, c20, c21);
        if (r1 && c1) store(i0 + 8 + g, j0 + 8 + 2 * q, c30, c31);
    }
}

// The same product with strided operand views instead of element functors -- no per-element masks or index
// arithmetic inside the k loop, which unrolls completely when the segment lengths are compile-time constants.
// A view describes op(X)(i, k) = p[i * is + k * ks] for i < valid (zero rows beyond); the K range is the
// concatenation of NSEG segments (len[s] a multiple of 4), each with its own pair of views.
struct OpView { const double* p; int is, ks, valid; };

struct NoAddend { __device__ __forceinline__ double operator()(int, int) const { return 0.0; } };

// `addend(r, c)` is evaluated for the 16 elements a thread owns BEFORE the k loop (its latency -- a global-memory
// read for the c_xx term -- hides behind the DMMAs) and handed to store(r, c, v0, v1, d0, d1) afterwards.
template <int NSEG, class ST, class AD>
__device__ __forceinline__ void dmma_blocks_v(int M, int N, const int (&len)[NSEG], const OpView (&A)[NSEG],
                                              const OpView (&B)[NSEG], ST store, AD addend, int nwarps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int tm = (M + 7) >> 3, tn = (N + 7) >> 3, bm = (tm + 1) >> 1, bn = (tn + 1) >> 1;
    for (int blk = warp; blk < bm * bn; blk += nwarps) {
        const int bi = blk / bn;
        const int i0 = bi * 16, j0 = (blk - bi * bn) * 16;
        const bool r1 = i0 + 8 < M, c1 = j0 + 8 < N;
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0, c20 = 0.0, c21 = 0.0, c30 = 0.0, c31 = 0.0;
        const int ra = i0 + g, rb = i0 + 8 + g, ca = j0 + 2 * q, cb = j0 + 8 + 2 * q;
        const double d00 = addend(ra, ca), d01 = addend(ra, ca + 1);
        const double d10 = c1 ? addend(ra, cb) : 0.0, d11 = c1 ? addend(ra, cb + 1) : 0.0;
        const double d20 = r1 ? addend(rb, ca) : 0.0, d21 = r1 ? addend(rb, ca + 1) : 0.0;
        const double d30 = (r1 && c1) ? addend(rb, cb) : 0.0, d31 = (r1 && c1) ? addend(rb, cb + 1) : 0.0;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
            const bool va0 = i0 + g < A[s].valid, va1 = r1 && (i0 + 8 + g < A[s].valid);
            const bool vb0 = j0 + g < B[s].valid, vb1 = c1 && (j0 + 8 + g < B[s].valid);
            const double* pa0 = A[s].p + (va0 ? (i0 + g) * A[s].is : 0) + q * A[s].ks;
            const double* pa1 = A[s].p + (va1 ? (i0 + 8 + g) * A[s].is : 0) + q * A[s].ks;
            const double* pb0 = B[s].p + (vb0 ? (j0 + g) * B[s].is : 0) + q * B[s].ks;
            const double* pb1 = B[s].p + (vb1 ? (j0 + 8 + g) * B[s].is : 0) + q * B[s].ks;
            const int aks = 4 * A[s].ks, bks = 4 * B[s].ks;
#pragma unroll
            for (int k0 = 0; k0 < len[s]; k0 += 4) {
                const int kk = k0 >> 2;
                double a0 = pa0[kk * aks], b0 = pb0[kk * bks], a1 = pa1[kk * aks], b1 = pb1[kk * bks];
                if (!va0) a0 = 0.0;
                if (!va1) a1 = 0.0;
                if (!vb0) b0 = 0.0;
                if (!vb1) b1 = 0.0;
                dmma_m8n8k4_acc(c00, c01, a0, b0);
                if (c1) dmma_m8n8k4_acc(c10, c11, a0, b1);
                if (r1) dmma_m8n8k4_acc(c20, c21, a1, b0);
                if (r1 && c1) dmma_m8n8k4_acc(c30, c31, a1, b1);
            }
        }
        store(ra, ca, c00, c01, d00, d01);
        if (c1) store(ra, cb, c10, c11, d10, d11);
        if (r1) store(rb, ca, c20, c21, d20, d21);
        if (r1 && c1) store(rb, cb, c30, c31, d30, d31);
    }
}

// One 8 x 8 tile per warp at a time with the K range dealt round-robin to four accumulator chains (short products: 2m x (n+m)).
template <class ST>
__device__ __forceinline__ void dmma_tiles_v(int M, int N, int len, const OpView& A, const OpView& B, ST store, int nwarps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int tm = (M + 7) >> 3, tn = (N + 7) >> 3;
    for (int tile = warp; tile < tm * tn; tile += nwarps) {
        const int ti = tile / tn;
        const int i0 = ti * 8, j0 = (tile - ti * tn) * 8;
        const bool va = i0 + g < A.valid, vb = j0 + g < B.valid;
        const double* pa = A.p + (va ? (i0 + g) * A.is : 0) + q * A.ks;
        const double* pb = B.p + (vb ? (j0 + g) * B.is : 0) + q * B.ks;
        const int aks = 4 * A.ks, bks = 4 * B.ks;
        double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};     // four chains: DMMA latency, not rate, bounds this
#pragma unroll
        for (int k0 = 0; k0 < len; k0 += 16) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                if (k0 + 4 * ch < len) {
                    const int kk = (k0 >> 2) + ch;
                    double a0 = pa[kk * aks], b0 = pb[kk * bks];
                    if (!va) a0 = 0.0;
                    if (!vb) b0 = 0.0;
                    dmma_m8n8k4_acc(acc[ch][0], acc[ch][1], a0, b0);
                }
            }
        }
        const double c0 = acc[0][0] + acc[1][0], d0 = acc[2][0] + acc[3][0];
        const double c1 = acc[0][1] + acc[1][1], d1 = acc[2][1] + acc[3][1];
        store(i0 + g, j0 + 2 * q, c0 + d0, c1 + d1);
    }
}

// Cholesky PD test (dpotf2 order, as cholesky_pd<>) of an MM x MM matrix by ONE thread in registers: the same arithmetic
// as the cooperative routine without its barriers.
template <int MM>
__device__ __forceinline__ bool small_chol(const double* __restrict__ A) {
    double L[MM][MM];
#pragma unroll
    for (int j = 0; j < MM; ++j) {
#pragma unroll
        for (int i = j; i < MM; ++i) {
            double sacc = A[i * MM + j];
#pragma unroll
            for (int k = 0; k < j; ++k) sacc = fma(-L[i][k], L[j][k], sacc);
            L[i][j] = sacc;
        }
        const double ajj = L[j][j];
        if (!(ajj > 0.0) || isinf(ajj)) return false;
        const double rj = sqrt(ajj);
#pragma unroll
        for (int i = j; i < MM; ++i) L[i][j] = (i == j) ? rj : L[i][j] / rj;
    }
    return true;
}

// Column `jcol` of the explicit inverse by LU with partial pivoting (the arithmetic of lu_inverse<>): every calling
// thread factors its own register copy and solves for one column, so the MM columns come out in parallel.
template <int MM>
__device__ __forceinline__ void small_lu_inv_col(const double* __restrict__ A, double* __restrict__ inv, int jcol) {
    double W[MM][MM];
    int piv[MM];
#pragma unroll
    for (int i = 0; i < MM; ++i)
#pragma unroll
        for (int j = 0; j < MM; ++j) W[i][j] = A[i * MM + j];
#pragma unroll
    for (int cidx = 0; cidx < MM; ++cidx) {
        int pr = cidx;
        double best = fabs(W[cidx][cidx]);
#pragma unroll
        for (int r = cidx + 1; r < MM; ++r) {
            const double v = fabs(W[r][cidx]);
            if (v > best) { best = v; pr = r; }
        }
        piv[cidx] = pr;
#pragma unroll
        for (int r = cidx + 1; r < MM; ++r) {
            if (pr == r) {
#pragma unroll
                for (int j = 0; j < MM; ++j) { const double tmp = W[cidx][j]; W[cidx][j] = W[r][j]; W[r][j] = tmp; }
            }
        }
        const double rp = 1.0 / W[cidx][cidx];
#pragma unroll
        for (int r = cidx + 1; r < MM; ++r) W[r][cidx] *= rp;
#pragma unroll
        for (int r = cidx + 1; r < MM; ++r)
#pragma unroll
            for (int j = cidx + 1; j < MM; ++j) W[r][j] = fma(-W[r][cidx], W[cidx][j], W[r][j]);
    }
    int pos = jcol;
#pragma unroll
    for (int cidx = 0; cidx < MM; ++cidx) {
        const int pr = piv[cidx];
        if (pos == cidx) pos = pr; else if (pos == pr) pos = cidx;
    }
    double x[MM];
#pragma unroll
    for (int i = 0; i < MM; ++i) {
        double sacc = (i == pos) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) sacc = fma(-W[i][k], x[k], sacc);
        x[i] = sacc;
    }
#pragma unroll
    for (int i = MM - 1; i >= 0; --i) {
        double sacc = x[i];
#pragma unroll
        for (int k = i + 1; k < MM; ++k) sacc = fma(-W[i][k], x[k], sacc);
        x[i] = sacc / W[i][i];
    }
#pragma unroll
    for (int i = 0; i < MM; ++i) inv[i * MM + jcol] = x[i];
}

template <class MP>
__device__ int bwd_pass_big(const typename MP::Dev& M, const IlqrArgs& a, const Smem& S, double* sm, const Rec& rc,
                            const double* __restrict__ Adense, const double* __restrict__ Bdense,
                            const double* __restrict__ ulast, double* __restrict__ Kout, double* __restrict__ kout,
                            double* __restrict__ ab, double* __restrict__ Quout, double* __restrict__ Quuout,
                            double& rho, double& drho, double* __restrict__ cxx) {
    constexpr int NT = MP::NT;
    constexpr int NW = NT / 32;
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, N = a.N, tid = threadIdx.x;
    const BigPlan G = make_big(n, m, nz, S.P);
    const int LD = G.LD, L3 = G.L3;
    double* Pp = sm + G.Pp;    double* S1 = sm + G.S1;    double* Q2 = sm + G.Q2;    double* L3b = sm + G.L3b;
    double* R3b = sm + G.R3b;  double* T1 = sm + G.T1;    double* T1f = sm + G.T1f;  double* se = sm + G.se;
    double* Qx = sm + G.Qx;    double* Qu = sm + G.Qu;    double* Quu = sm + G.Quu;  double* Quut = sm + G.Quut;
    double* Lc = sm + G.Lc;    double* LU = sm + G.LU;    double* inv = sm + G.inv;
    const double* sQ = sm + S.Qs; const double* sR = sm + S.Rs; const double* sQf = sm + S.Qfs;
    const double* sHc = sm + S.Hcs;
    int* piv = reinterpret_cast<int*>(sm + S.ints);
    int* flag = piv + m + 2;
    const srcb200_ilqr_config& c = a.cfg;
    const bool sreg = c.regularize && c.state_regularization;
    const bool aligned4 = (n % 4 == 0) && (m % 4 == 0);      // K segments are whole m8n8k4 steps: strided views
    int pd_fail = -1;

    auto lin_of = [&](int t) {
        if (a.index_lin) return MP::bank(M, rc.idx[t]);
        if (Adense) return LinRef{Adense + (long long)t * n * n, Bdense + (long long)t * n * m, nullptr};
        return LinRef{rc.A + (long long)t * n * n, rc.B + (long long)t * n * m, nullptr};
    };

    // constant-H cost Hessian, once per pass: T1 = H^T Q, c_xx = T1 H (global scratch, read back through L1)
    mm<NT, true, false>(T1, nz, sHc, n, sQ, nz, n, nz, nz);
    mm<NT, true, false>(T1f, nz, sHc, n, sQf, nz, n, nz, nz);
    cta_sync<NT>();
    mm<NT, false, false>(cxx, n, T1, nz, sHc, n, n, n, nz);
    __threadfence_block();

    {
        // terminal_cost_vectors (ilqr.py:177-182): P_N = (H^T Qf) H, p_N = (H^T Qf) e_N
        for (int i = tid; i < nz; i += NT) se[i] = rc.e[N * nz + i];
        for (int e = tid; e < n * LD; e += NT) Pp[e] = 0.0;
        if (N > 0) {
            const LinRef l0 = lin_of(N - 1);
            load_lin_async<NT>(sm + (((N - 1) & 1) ? G.Ap1 : G.Ap0), LD, l0.A, l0.B, n, m);
        }
        cta_sync<NT>();
        mm<NT, false, false>(Pp, LD, T1f, nz, sHc, n, n, n, nz);
        for (int i = tid; i < n; i += NT) {
            double acc = 0.0;
            for (int k2 = 0; k2 < nz; ++k2) acc = fma(T1f[i * nz + k2], se[k2], acc);
            Pp[i * LD + n] = acc;
        }
        cta_sync<NT>();

        auto prefetch_small = [&](int t) {
            double* dst = sm + ((t & 1) ? G.pf1 : G.pf0);
            for (int i = tid; i < nz; i += NT) cp_async8(dst + i, rc.e + t * nz + i);
            for (int i = tid; i < m; i += NT) {
                cp_async8(dst + nz + i, rc.u + t * m + i);
                if (t > 0) cp_async8(dst + nz + m + i, rc.u + (t - 1) * m + i);
            }
        };
#ifdef SRCB_PHASE_TIMING
        long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tk = clock64();
#define PH(i) do { const long long now_ = clock64(); ph[i] += now_ - tk; tk = now_; } while (0)
#else
#define PH(i) do { } while (0)
#endif
        if (N > 0) prefetch_small(N - 1);
        for (int t = N - 1; t >= 0; --t) {
            // ---- A'_t, e_t, u_t, u_{t-1} have landed and (P | p) of step t+1 is complete; fetch step t-1
            PH(7);
            cp_async_wait_all();
            cta_sync<NT>();
            PH(0);
            const double* Ap = sm + ((t & 1) ? G.Ap1 : G.Ap0);
            const double* pf = sm + ((t & 1) ? G.pf1 : G.pf0);
            if (t > 0) {
                const LinRef ln = lin_of(t - 1);
                load_lin_async<NT>(sm + (((t - 1) & 1) ? G.Ap1 : G.Ap0), LD, ln.A, ln.B, n, m);
                prefetch_small(t - 1);
            }
            auto du_of = [&](int i) {       // du_t (ilqr.py:188-190)
                const double ut = pf[nz + i];
                if (!c.include_input_var_constraint) return ut;
                return __dsub_rn(ut, t == 0 ? (ulast ? ulast[i] : 0.0) : pf[nz + m + i]);
            };
            // ---- product 1: S1 = A'^T P'; its epilogue also forms B^T (P + rho I) = B^T P + rho B^T,
            //      Q_x = c_x + A^T p with c_x = (H^T Q) e, Q_u = c_u + B^T p with c_u = R du   (ilqr.py:186-196, 258-267)
            auto put1 = [&](int r, int cc, double v) {
                if (cc < n) {
                    S1[r * LD + cc] = v;
                    if (r >= n) S1[(r + m) * LD + cc] = sreg ? fma(rho, Ap[cc * LD + r], v) : v;
                } else if (cc == n) {
                    S1[r * LD + n] = v;
                    if (r < n) {
                        double acc = 0.0;
                        for (int k2 = 0; k2 < nz; ++k2) acc = fma(T1[r * nz + k2], pf[k2], acc);
                        Qx[r] = __dadd_rn(acc, v);
                    } else {
                        double acc = 0.0;
                        for (int k2 = 0; k2 < m; ++k2) acc = fma(sR[(r - n) * m + k2], du_of(k2), acc);
                        Qu[r - n] = __dadd_rn(acc, v);
                    }
                }
            };
            auto store1 = [&](int r, int cc, double v0, double v1) {
                if (r < n + m) { put1(r, cc, v0); put1(r, cc + 1, v1); }
            };
            if (aligned4) {
                const int len1[1] = {n};
                const OpView A1[1] = {{Ap, 1, LD, n + m}};
                const OpView B1[1] = {{Pp, 1, LD, n + 1}};
                dmma_blocks_v<1>(n + m, n + 1, len1, A1, B1,
                                 [&](int r, int cc, double v0, double v1, double, double) { store1(r, cc, v0, v1); }, NoAddend(), NW);
            } else {
                dmma_blocks(n + m, n + 1, n,
                            [&](int r, int k) { return (r < n + m && k < n) ? Ap[k * LD + r] : 0.0; },
                            [&](int k, int cc) { return (cc <= n && k < n) ? Pp[k * LD + cc] : 0.0; }, store1, NW);
            }
            PH(1);
            cta_sync<NT>();
            PH(2);
            // ---- product 2: [B^T P ; B^T (P + rho I)] A' = [Q_ux | Q_uu - R ; Q_ux~ | Q_uu~ - R]; the epilogue adds
            //      c_uu = R (ilqr.py:260-261, 268-271)
            auto put2 = [&](int r, int cc, double v) {
                if (cc < n) {
                    Q2[r * LD + cc] = v;
                } else if (cc < n + m) {
                    const int j = cc - n;
                    if (r < m) {
                        const double quu = __dadd_rn(sR[r * m + j], v);
                        Quu[r * m + j] = quu;
                        if (!sreg) Quut[r * m + j] = (c.regularize && r == j) ? __dadd_rn(quu, rho) : quu;
                    } else if (sreg) {
                        Quut[(r - m) * m + j] = __dadd_rn(sR[(r - m) * m + j], v);
                    }
                }
            };
            auto store2 = [&](int r, int cc, double v0, double v1) {
                if (r < 2 * m) { put2(r, cc, v0); put2(r, cc + 1, v1); }
            };
            if (aligned4) {
                const OpView A2 = {S1 + n * LD, LD, 1, 2 * m};
                const OpView B2 = {Ap, 1, LD, n + m};
                dmma_tiles_v(2 * m, n + m, n, A2, B2, store2, NW);
            } else {
                dmma_blocks(2 * m, n + m, n,
                            [&](int r, int k) { return (r < 2 * m && k < n) ? S1[(n + r) * LD + k] : 0.0; },
                            [&](int k, int cc) { return (cc < n + m && k < n) ? Ap[k * LD + cc] : 0.0; }, store2, NW);
            }
            cta_sync<NT>();
            PH(3);
            // ---- PD test by Cholesky (ilqr.py:276-287) and the explicit inverse (ilqr.py:289): m x m, so one thread
            //      (compile-time m) or one warp does it while the others lay out the operands that do not need it:
            //      rows (Q_ux | Q_u) of the right factor and the Q_ux^T block of the left factor
            constexpr bool small_m = MP::CM > 0 && MP::CM <= 8;
            constexpr int FILL0 = small_m ? 64 : 32;            // first thread of the operand-layout crew
            if (tid < FILL0) {
                if constexpr (small_m) {
                    // thread 0: Cholesky verdict; lanes 0..m-1 of warp 1: the LU (each its own register copy) and one
                    // inverse column each.  Two warps, so the two run concurrently; the inverse is only used when
                    // the verdict allows it, computing it regardless is harmless.
                    constexpr int MM = MP::CM > 0 ? MP::CM : 1;
                    if (tid == 0) *flag = small_chol<MM>(Quut) ? 1 : 0;
                    else if (tid >= 32 && tid < 32 + MM) small_lu_inv_col<MM>(Quut, inv, tid - 32);
                } else {
                    const bool pd1 = cholesky_pd<32>(Quut, Lc, flag, m);
                    if (pd1 || !c.regularize) {
                        for (int e = tid; e < m * m; e += 32) LU[e] = Quut[e];
                        __syncwarp();
                        lu_inverse<32>(LU, inv, piv, m);
                    }
                }
            } else {
                for (int e = tid - FILL0; e < m * (n + 1); e += NT - FILL0) {
                    const int i = e / (n + 1), j = e - i * (n + 1);
                    R3b[(m + i) * LD + j] = (j < n) ? Q2[i * LD + j] : Qu[i];
                }
                for (int e = tid - FILL0; e < n * m; e += NT - FILL0) {
                    const int i = e / m, k2 = e - i * m;
                    L3b[i * L3 + 2 * m + k2] = Q2[k2 * LD + i];
                }
            }
            cta_sync<NT>();
            PH(4);
            const bool pd = (*flag != 0);
            if (!pd && pd_fail < 0) pd_fail = t;
            if (!pd && c.regularize) {
                // ilqr.py:282-287: raise rho and leave the sweep (no restart -- the code falls through to the decrease
                // at 298).  Q_u[t], Q_uu[t] are assigned; K, k and the line-search scalars of every s <= t stay zero.
                rho_update(c, true, rho, drho);
                if (Quout) for (int i = tid; i < m; i += NT) Quout[t * m + i] = Qu[i];
                if (Quuout) for (int e = tid; e < m * m; e += NT) Quuout[(long long)t * m * m + e] = Quu[e];
                for (long long e = tid; e < (long long)(t + 1) * m * n; e += NT) Kout[e] = 0.0;
                for (int e = tid; e < (t + 1) * m; e += NT) kout[e] = 0.0;
                if (Quout) for (int e = tid; e < t * m; e += NT) Quout[e] = 0.0;
                if (Quuout) for (long long e = tid; e < (long long)t * m * m; e += NT) Quuout[e] = 0.0;
                for (int e = tid; e < 2 * t; e += NT) ab[e] = 0.0;
                if (tid == 0) {
                    double sacc = 0.0, qq = 0.0;      // k = 0 evaluated literally (0 * inf = nan like numpy)
                    for (int i = 0; i < m; ++i) sacc = __dadd_rn(sacc, __dmul_rn(0.0, Qu[i]));
                    for (int j = 0; j < m; ++j) {
                        double v = 0.0;
                        for (int i = 0; i < m; ++i) v = __dadd_rn(v, __dmul_rn(0.0, Quu[i * m + j]));
                        qq = __dadd_rn(qq, __dmul_rn(v, 0.0));
                    }
                    ab[2 * t] = sacc;
                    ab[2 * t + 1] = qq;
                }
                break;
            }
            // ---- gains (ilqr.py:289-292): K = -inv Q_ux~, k = -inv Q_u, one column per thread, and with it that
            //      column's part of both factors of the value update: K^T Q_uu, K^T, (K | k) twice; outputs
            for (int i = tid; i <= n; i += NT) {
                constexpr int MMAX = MP::CM > 0 ? MP::CM : 32;
                constexpr int UNR = MP::CM > 0 ? MP::CM : 1;       // unroll only when m is a compile-time constant
                double col[MMAX];
#pragma unroll UNR
                for (int j = 0; j < MMAX; ++j) {
                    if (j < m) {
                        double acc = 0.0;
                        if (i < n) { for (int k2 = 0; k2 < m; ++k2) acc = fma(inv[j * m + k2], Q2[(m + k2) * LD + i], acc); }
                        else       { for (int k2 = 0; k2 < m; ++k2) acc = fma(inv[j * m + k2], Qu[k2], acc); }
                        col[j] = -acc;
                        R3b[j * LD + i] = -acc;
                        R3b[(2 * m + j) * LD + i] = -acc;
                        if (i < n) Kout[(long long)t * m * n + j * n + i] = -acc;
                        else kout[t * m + j] = -acc;
                    }
                }
                if (i < n) {
#pragma unroll UNR
                    for (int k2 = 0; k2 < MMAX; ++k2) {
                        if (k2 < m) {
                            double acc = 0.0;
#pragma unroll UNR
                            for (int j = 0; j < MMAX; ++j) if (j < m) acc = fma(col[j], Quu[j * m + k2], acc);
                            L3b[i * L3 + k2] = acc;
                            L3b[i * L3 + m + k2] = col[k2];
                        }
                    }
                } else {
                    // line-search scalars a_t = k . Q_u, b_t = (k^T Q_uu) . k   (ilqr.py:69-71)
                    double sacc = 0.0, qq = 0.0;
#pragma unroll UNR
                    for (int j = 0; j < MMAX; ++j) if (j < m) sacc = fma(col[j], Qu[j], sacc);
#pragma unroll UNR
                    for (int j = 0; j < MMAX; ++j) {
                        if (j < m) {
                            double v = 0.0;
#pragma unroll UNR
                            for (int i2 = 0; i2 < MMAX; ++i2) if (i2 < m) v = fma(col[i2], Quu[i2 * m + j], v);
                            qq = fma(v, col[j], qq);
                        }
                    }
                    ab[2 * t] = sacc;
                    ab[2 * t + 1] = qq;
                    if (Quout) for (int j = 0; j < m; ++j) Quout[t * m + j] = Qu[j];
                    if (Quuout) for (int e = 0; e < m * m; ++e) Quuout[(long long)t * m * m + e] = Quu[e];
                }
            }
            cta_sync<NT>();
            PH(5);
            // ---- product 3: (P | p) = (c_xx | Q_x) + [A^T P | K^T Q_uu | K^T | Q_ux^T] [A|0 ; K|k ; Q_ux|Q_u ; K|k]
            auto add3 = [&](int r, int cc) {            // (c_xx | Q_x) element
                if (r >= n || cc > n) return 0.0;
                return (cc < n) ? cxx[r * n + cc] : Qx[r];
            };
            auto store3d = [&](int r, int cc, double v0, double v1, double d0, double d1) {
                if (r < n) {
                    if (cc <= n) Pp[r * LD + cc] = __dadd_rn(d0, v0);
                    if (cc + 1 <= n) Pp[r * LD + cc + 1] = __dadd_rn(d1, v1);
                }
            };
            auto store3 = [&](int r, int cc, double v0, double v1) { store3d(r, cc, v0, v1, add3(r, cc), add3(r, cc + 1)); };
            if (aligned4) {
                const int len3[2] = {n, 3 * m};
                const OpView A3[2] = {{S1, LD, 1, n}, {L3b, L3, 1, n}};
                const OpView B3[2] = {{Ap, 1, LD, n}, {R3b, 1, LD, n + 1}};
                dmma_blocks_v<2>(n, n + 1, len3, A3, B3, store3d, add3, NW);
            } else {
                dmma_blocks(n, n + 1, n + 3 * m,
                            [&](int r, int k) {
                                if (r >= n) return 0.0;
                                if (k < n) return S1[r * LD + k];
                                return (k < n + 3 * m) ? L3b[r * L3 + (k - n)] : 0.0;
                            },
                            [&](int k, int cc) {
                                if (cc > n) return 0.0;
                                if (k < n) return (cc < n) ? Ap[k * LD + cc] : 0.0;
                                return (k < n + 3 * m) ? R3b[(k - n) * LD + cc] : 0.0;
                            }, store3, NW);
            }
        }
        cta_sync<NT>();
#ifdef SRCB_PHASE_TIMING
        if (tid == 0 && Quout) for (int i = 0; i < 8; ++i) Quout[i] = (double)ph[i];
#endif
        cp_async_wait_all();        // an interrupted sweep may leave a prefetch in flight into buffers the caller reuses
        cta_sync<NT>();
        rho_update(c, false, rho, drho);      // ilqr.py:298 -- after a complete AND after an interrupted sweep
    }
    return pd_fail;
}

}  // namespace srcb
