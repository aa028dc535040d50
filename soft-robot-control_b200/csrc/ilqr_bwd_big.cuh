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

__host__ __device__ inline int cf_ld(int v) {           // smallest leading dimension >= v without LDS.64 bank conflicts
    while ((v & 15) != 4 && (v & 15) != 12) ++v;
    return v;
}

struct BigPlan {
    int LD, L3;
    int Pp, Ap0, Ap1, S1, Q2, L3b, R3b, T1, T1f, se, sdu, sut, cx, cu, Qx, Qu, Quu, Quut, Lc, LU, inv, end;
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
    B.se = take(nz);  B.sdu = take(m);  B.sut = take(m);
    B.cx = take(n);   B.cu = take(m);   B.Qx = take(n);   B.Qu = take(m);
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

template <class MP>
__device__ int bwd_pass_big(const typename MP::Dev& M, const IlqrArgs& a, const Smem& S, double* sm, const Rec& rc,
                            const double* __restrict__ Adense, const double* __restrict__ Bdense,
                            const double* __restrict__ ulast, double* __restrict__ Kout, double* __restrict__ kout,
                            double* __restrict__ ab, double* __restrict__ Quout, double* __restrict__ Quuout,
                            double& rho, double& drho, bool& give_up, double* __restrict__ cxx) {
    constexpr int NT = MP::NT;
    constexpr int NW = NT / 32;
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, N = a.N, tid = threadIdx.x;
    const BigPlan G = make_big(n, m, nz, S.P);
    const int LD = G.LD, L3 = G.L3;
    double* Pp = sm + G.Pp;    double* S1 = sm + G.S1;    double* Q2 = sm + G.Q2;    double* L3b = sm + G.L3b;
    double* R3b = sm + G.R3b;  double* T1 = sm + G.T1;    double* T1f = sm + G.T1f;  double* se = sm + G.se;
    double* sdu = sm + G.sdu;  double* sut = sm + G.sut;  double* cx = sm + G.cx;    double* cu = sm + G.cu;
    double* Qx = sm + G.Qx;    double* Qu = sm + G.Qu;    double* Quu = sm + G.Quu;  double* Quut = sm + G.Quut;
    double* Lc = sm + G.Lc;    double* LU = sm + G.LU;    double* inv = sm + G.inv;
    const double* sQ = sm + S.Qs; const double* sR = sm + S.Rs; const double* sQf = sm + S.Qfs;
    const double* sHc = sm + S.Hcs;
    int* piv = reinterpret_cast<int*>(sm + S.ints);
    int* flag = piv + m + 2;
    const srcb200_ilqr_config& c = a.cfg;
    const bool sreg = c.regularize && c.state_regularization;
    int restarts = 0;
    give_up = false;

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

    while (true) {
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

        bool ok = true;
        for (int t = N - 1; t >= 0; --t) {
            // ---- A'_t has landed; stage e_t, du_t; start fetching A'_{t-1} into the other buffer
            cp_async_wait_all();
            for (int i = tid; i < nz; i += NT) se[i] = rc.e[t * nz + i];
            for (int i = tid; i < m; i += NT) {
                const double ut = rc.u[t * m + i];
                double du = ut;
                if (c.include_input_var_constraint)
                    du = __dsub_rn(ut, t == 0 ? (ulast ? ulast[i] : 0.0) : rc.u[(t - 1) * m + i]);
                sdu[i] = du;
                sut[i] = ut;
            }
            cta_sync<NT>();
            const double* Ap = sm + ((t & 1) ? G.Ap1 : G.Ap0);
            if (t > 0) {
                const LinRef ln = lin_of(t - 1);
                load_lin_async<NT>(sm + (((t - 1) & 1) ? G.Ap1 : G.Ap0), LD, ln.A, ln.B, n, m);
            }
            // ---- step_cost_vectors (ilqr.py:186-196): c_x = (H^T Q) e, c_u = R du
            mv<NT, false>(cx, T1, nz, se, n, nz);
            mv<NT, false>(cu, sR, m, sdu, m, m);
            // ---- product 1: S1 = A'^T P'
            dmma_blocks(n + m, n + 1, n,
                        [&](int r, int k) { return (r < n + m && k < n) ? Ap[k * LD + r] : 0.0; },
                        [&](int k, int cc) { return (cc <= n && k < n) ? Pp[k * LD + cc] : 0.0; },
                        [&](int r, int cc, double v0, double v1) {
                            if (r < n + m) {
                                if (cc <= n) S1[r * LD + cc] = v0;
                                if (cc + 1 <= n) S1[r * LD + cc + 1] = v1;
                            }
                        }, NW);
            cta_sync<NT>();
            // ---- B^T (P + rho I) = B^T P + rho B^T;  Q_x = c_x + A^T p;  Q_u = c_u + B^T p   (ilqr.py:258-267)
            for (int e = tid; e < m * n; e += NT) {
                const int i = e / n, j = e - i * n;
                const double v = S1[(n + i) * LD + j];
                S1[(n + m + i) * LD + j] = sreg ? fma(rho, Ap[j * LD + n + i], v) : v;
            }
            for (int i = tid; i < n; i += NT) Qx[i] = __dadd_rn(cx[i], S1[i * LD + n]);
            for (int i = tid; i < m; i += NT) Qu[i] = __dadd_rn(cu[i], S1[(n + i) * LD + n]);
            cta_sync<NT>();
            // ---- product 2: Q2 = [B^T P ; B^T (P + rho I)] A'
            dmma_blocks(2 * m, n + m, n,
                        [&](int r, int k) { return (r < 2 * m && k < n) ? S1[(n + r) * LD + k] : 0.0; },
                        [&](int k, int cc) { return (cc < n + m && k < n) ? Ap[k * LD + cc] : 0.0; },
                        [&](int r, int cc, double v0, double v1) {
                            if (r < 2 * m) {
                                if (cc < n + m) Q2[r * LD + cc] = v0;
                                if (cc + 1 < n + m) Q2[r * LD + cc + 1] = v1;
                            }
                        }, NW);
            cta_sync<NT>();
            // ---- Q_uu = c_uu + B^T P B,  Q_uu~ (ilqr.py:261, 268-271)
            for (int e = tid; e < m * m; e += NT) {
                const int i = e / m, j = e - i * m;
                const double quu = __dadd_rn(sR[e], Q2[i * LD + n + j]);
                Quu[e] = quu;
                if (sreg) Quut[e] = __dadd_rn(sR[e], Q2[(m + i) * LD + n + j]);
                else Quut[e] = (c.regularize && i == j) ? __dadd_rn(quu, rho) : quu;
            }
            cta_sync<NT>();
            // ---- PD test by Cholesky (ilqr.py:276-287)
            const bool pd = cholesky_pd<NT>(Quut, Lc, flag, m);
            if (!pd && c.regularize) {
                rho_update(c, true, rho, drho);
                ok = false;
                break;
            }
            // ---- gains (ilqr.py:289-292): explicit inverse, K = -inv Q_ux~, k = -inv Q_u
            for (int e = tid; e < m * m; e += NT) LU[e] = Quut[e];
            cta_sync<NT>();
            lu_inverse<NT>(LU, inv, piv, m);
            for (int e = tid; e < m * (n + 1); e += NT) {
                const int i = e / (n + 1), j = e - i * (n + 1);
                double acc = 0.0;
                if (j < n) {
                    for (int k2 = 0; k2 < m; ++k2) acc = fma(inv[i * m + k2], Q2[(m + k2) * LD + j], acc);
                } else {
                    for (int k2 = 0; k2 < m; ++k2) acc = fma(inv[i * m + k2], Qu[k2], acc);
                }
                R3b[i * LD + j] = -acc;                                    // (K | k)
                R3b[(2 * m + i) * LD + j] = -acc;
                R3b[(m + i) * LD + j] = (j < n) ? Q2[i * LD + j] : Qu[i];      // (Q_ux | Q_u)
            }
            cta_sync<NT>();
            // ---- left factor of the value update: [K^T Q_uu | K^T | Q_ux^T]; outputs of this step
            for (int e = tid; e < n * m; e += NT) {
                const int i = e / m, k2 = e - i * m;
                double acc = 0.0;
                for (int j = 0; j < m; ++j) acc = fma(R3b[j * LD + i], Quu[j * m + k2], acc);
                L3b[i * L3 + k2] = acc;
                L3b[i * L3 + m + k2] = R3b[k2 * LD + i];
                L3b[i * L3 + 2 * m + k2] = Q2[k2 * LD + i];
            }
            for (int e = tid; e < m * n; e += NT) {
                const int i = e / n, j = e - i * n;
                Kout[(long long)t * m * n + e] = R3b[i * LD + j];
            }
            for (int i = tid; i < m; i += NT) kout[t * m + i] = R3b[i * LD + n];
            if (Quout) for (int i = tid; i < m; i += NT) Quout[t * m + i] = Qu[i];
            if (Quuout) for (int e = tid; e < m * m; e += NT) Quuout[(long long)t * m * m + e] = Quu[e];
            if (tid == NT - 1) {
                double s = 0.0;
                for (int i = 0; i < m; ++i) s = fma(R3b[i * LD + n], Qu[i], s);
                double qq = 0.0;
                for (int j = 0; j < m; ++j) {
                    double v = 0.0;
                    for (int i = 0; i < m; ++i) v = fma(R3b[i * LD + n], Quu[i * m + j], v);
                    qq = fma(v, R3b[j * LD + n], qq);
                }
                ab[2 * t] = s;
                ab[2 * t + 1] = qq;
            }
            cta_sync<NT>();
            // ---- product 3: (P | p) = (c_xx | Q_x) + [A^T P | K^T Q_uu | K^T | Q_ux^T] [A|0 ; K|k ; Q_ux|Q_u ; K|k]
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
                        },
                        [&](int r, int cc, double v0, double v1) {
                            if (r < n) {
                                if (cc < n) Pp[r * LD + cc] = __dadd_rn(cxx[r * n + cc], v0);
                                else if (cc == n) Pp[r * LD + n] = __dadd_rn(Qx[r], v0);
                                if (cc + 1 < n) Pp[r * LD + cc + 1] = __dadd_rn(cxx[r * n + cc + 1], v1);
                                else if (cc + 1 == n) Pp[r * LD + n] = __dadd_rn(Qx[r], v1);
                            }
                        }, NW);
            cta_sync<NT>();
        }
        if (ok) {
            rho_update(c, false, rho, drho);
            break;
        }
        cp_async_wait_all();        // a prefetch may still be in flight into the buffers the restart reuses
        cta_sync<NT>();
        ++restarts;
        if (restarts >= c.max_pd_restarts) { give_up = true; break; }
    }
    return restarts;
}

}  // namespace srcb
