// ilqr_impl.cuh -- batched iLQR (templates, instantiated per model policy in ilqr_ssm.cu / ilqr_tpwl.cu): one CTA per problem walks the reference algorithm (sofacontrol/lqr/ilqr.py:27-300)
// branch for branch; every per-problem decision (line search, regularisation schedule, PD test, convergence)
// is taken inside the kernel and reported per problem.
//
// Structure (Appendix B of SURVEY.md):
//   forward_pass  (ilqr.py:117-162)  -> fwd_pass<>()      : u_t = u_t + alpha k_t + K_t (x_t - x_t), cost, re-linearise, step
//   dlqr_recursion(ilqr.py:219-300)  -> bwd_pass<>()      : Riccati sweep, Cholesky PD test, explicit inverse of Q_uu~
//   ilqr_computation (ilqr.py:27-107)-> ilqr_solve_kernel : outer loop + line search
// The model is a policy (SSM polynomial model / TPWL bank) providing linearise+observe at one state.
#pragma once
#include <cstdlib>
#include <type_traits>
#include "ilqr.cuh"

namespace srcb {

struct LinRef {            // where the linearisation of one step lives (shared memory or the global bank)
    const double* A;
    const double* B;
    const double* d;
};

// ---------------------------------------------------------------------------------------------------------------
// Model policies
// ---------------------------------------------------------------------------------------------------------------
struct SsmPolicy {
    using Dev = SsmDev;
    static constexpr int NT = 32;
    static constexpr int CN = 0, CM = 0, CNZ = 0;      // compile-time dimensions (0 = run time)
    __host__ __device__ static int n(const Dev& M) { return M.n; }
    __host__ __device__ static int m(const Dev& M) { return M.m; }
    __host__ __device__ static int nz(const Dev& M) { return M.nz; }
    __host__ __device__ static bool index_lin(const Dev&, double) { return false; }
    __host__ __device__ static int scratch_doubles(const Dev& M, double) { return ssm_eval_scratch_doubles(M.n, M.m, M.nfeat); }
    // linearise (A,B,d discretised with dt) and observe (z, H) at (x,u); everything lands in shared memory
    __device__ static void eval(const Dev& M, const double* sx, const double* su, double dt, double* sA, double* sB,
                                double* sd, double* sz, double* sH, double* scr, LinRef& lin, int& idx) {
        ssm_eval<NT>(M, sx, su, dt, sA, sB, sd, sz, sH, scr);
        lin.A = sA; lin.B = sB; lin.d = sd;
        idx = 0;
    }
    __device__ static void observe(const Dev& M, const double* sx, double* sz, double* sH, double* scr) {
        ssm_eval<NT>(M, sx, nullptr, -1.0, nullptr, nullptr, nullptr, sz, sH, scr);
        cta_sync<NT>();
    }
    __device__ static LinRef bank(const Dev&, int) { return LinRef{nullptr, nullptr, nullptr}; }
};

// CN/CM/CNZ > 0 fix n/m/nz at compile time (the Diamond shape gets its own instantiation: constant strides, unrolled
// m-loops, divisions by constants); 0 keeps them run-time values.
template <int CN_, int CM_, int CNZ_>
struct TpwlPolicyT {
    using Dev = TpwlDev;
    static constexpr int NT = 512;
    static constexpr int CN = CN_, CM = CM_, CNZ = CNZ_;
    __host__ __device__ static int n(const Dev& M) { return M.n; }
    __host__ __device__ static int m(const Dev& M) { return M.m; }
    __host__ __device__ static int nz(const Dev& M) { return M.nz; }
    // nn on a bank that needs no per-step discretisation: a step's linearisation is just an index into the bank
    __host__ __device__ static bool index_lin(const Dev& M, double dt) {
        return M.method == SRCB200_TPWL_NN && (M.discr == SRCB200_DISCR_NONE || dt < 0.0);
    }
    __host__ __device__ static int scratch_doubles(const Dev& M, double dt) {
        if (index_lin(M, dt))       // the specialised forward pass (ilqr_fwd_tpwl.cuh) lays its buffers out here
            return make_fwdnn(M.n, M.m, M.nz, M.P, M.n / 2, M.wq != 0.0, M.wv != 0.0, NT).end + 8;
        int s = 0;
        if (M.method == SRCB200_TPWL_WEIGHTING) s += M.P;
        if (!index_lin(M, dt)) s += discretize_scratch_doubles(M.n, M.m);
        return s + 8;
    }
    __device__ static void eval(const Dev& M, const double* sx, const double* su, double dt, double* sA, double* sB,
                                double* sd, double* sz, double* sH, double* scr, LinRef& lin, int& idx) {
        __shared__ double red_d[NT / 32];
        __shared__ int red_i[NT / 32];
        const int n = CN_ ? CN_ : M.n, m = CM_ ? CM_ : M.m, tid = threadIdx.x;
        (void)su; (void)sH;
        // z = H x + z_ref (tpwl.py:121-122)
        for (int i = tid; i < (CNZ_ ? CNZ_ : M.nz); i += NT) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc = fma(M.H[i * n + k], sx[k], acc);
            sz[i] = __dadd_rn(acc, M.zref[i]);
        }
        if (M.method == SRCB200_TPWL_NN) {
            idx = tpwl_nearest<NT>(M, sx, nullptr, red_d, red_i, nullptr);
            if (index_lin(M, dt)) {
                lin = bank(M, idx);
                __syncthreads();
                return;
            }
            const LinRef b = bank(M, idx);
            for (int e = tid; e < n * n; e += NT) sA[e] = b.A[e];
            for (int e = tid; e < n * m; e += NT) sB[e] = b.B[e];
            for (int e = tid; e < n; e += NT) sd[e] = b.d[e];
            __syncthreads();
        } else {
            // exponential weights, then stream the whole bank once: A = sum_p w_p A_p (tpwl.py:245-248)
            double* sw = scr;
            double dmin;
            const int bi = tpwl_nearest<NT>(M, sx, sw, red_d, red_i, &dmin);
            idx = bi;
            __syncthreads();
            if (dmin == 0.0) {
                for (int p = tid; p < M.P; p += NT) sw[p] = (p == bi) ? 1.0 : 0.0;
                __syncthreads();
            } else {
                double part = 0.0;
                for (int p = tid; p < M.P; p += NT) {
                    const double e = exp(__ddiv_rn(__dmul_rn(-M.beta, sw[p]), dmin));
                    sw[p] = e;
                    part += e;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                __syncthreads();
                if ((tid & 31) == 0) red_d[tid >> 5] = part;
                __syncthreads();
                double tot = 0.0;
                for (int k = 0; k < NT / 32; ++k) tot += red_d[k];
                for (int p = tid; p < M.P; p += NT) sw[p] = __ddiv_rn(sw[p], tot);
                __syncthreads();
            }
            const long long nn = (long long)n * n, nm = (long long)n * m;
            for (int e = tid; e < nn; e += NT) {
                double acc = 0.0;
                for (int p = 0; p < M.P; ++p) acc = fma(sw[p], M.A[p * nn + e], acc);
                sA[e] = acc;
            }
            for (int e = tid; e < nm; e += NT) {
                double acc = 0.0;
                for (int p = 0; p < M.P; ++p) acc = fma(sw[p], M.B[p * nm + e], acc);
                sB[e] = acc;
            }
            for (int e = tid; e < n; e += NT) {
                double acc = 0.0;
                for (int p = 0; p < M.P; ++p) acc = fma(sw[p], M.d[(long long)p * n + e], acc);
                sd[e] = acc;
            }
            __syncthreads();
            scr += M.P;
        }
        if (dt >= 0.0 && M.discr != SRCB200_DISCR_NONE) discretize_inplace<NT>(M.discr, dt, sA, sB, sd, n, m, scr);
        lin.A = sA; lin.B = sB; lin.d = sd;
    }
    __device__ static void observe(const Dev& M, const double* sx, double* sz, double*, double*) {
        const int n = CN_ ? CN_ : M.n;
        for (int i = threadIdx.x; i < (CNZ_ ? CNZ_ : M.nz); i += NT) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc = fma(M.H[i * n + k], sx[k], acc);
            sz[i] = __dadd_rn(acc, M.zref[i]);
        }
        __syncthreads();
    }
    __device__ static LinRef bank(const Dev& M, int p) {
        const int n = CN_ ? CN_ : M.n, m = CM_ ? CM_ : M.m;
        return LinRef{M.A + (long long)p * n * n, M.B + (long long)p * n * m, M.d + (long long)p * n};
    }
};
using TpwlPolicy = TpwlPolicyT<0, 0, 0>;
using TpwlPolicyDiamond = TpwlPolicyT<72, 4, 6>;        // n = 72 (r = 36), m = 4, n_z = 6

// shared-memory plan (doubles); forward and backward phases alias the same region after the common header
struct Smem {
    int x, xn, u, uprev, dx, z, e, A, B, d, H, Qe, Rdu, mscr, fwd_end;
    int P, p, AtP, BtP, BtPr, Qux, Quxt, Quu, Quut, Lc, LU, inv, K, k, KQ, Qx, Qu, cx, cu, T1, cxx, T1f, bwd_end;
    int Qs, Rs, Qfs, Hcs, scal, ints, total;
};
__host__ __device__ inline int big_plan_end(int n, int m, int nz, int base);     // ilqr_bwd_big.cuh
// big: the backward pass runs bwd_pass_big (ilqr_bwd_big.cuh), whose shared-memory plan replaces the one below
__host__ __device__ inline bool use_big_bwd(int nt, int n, bool gn) { return nt > 32 && n >= 16 && !gn; }
__host__ __device__ inline Smem make_smem(int n, int m, int nz, int mscr, bool gn, bool big = false, bool index_lin = false) {
    Smem S;
    int o = 0;
    // persistent header: cost matrices + constant H + scalars
    S.Qs = o; o += nz * nz;
    S.Rs = o; o += m * m;
    S.Qfs = o; o += nz * nz;
    S.Hcs = o; o += nz * n;
    S.scal = o; o += 8;
    S.ints = o; o += (m + 8) / 2 + 1;      // pivots + flags
    const int base = o;
    // forward
    S.x = o; o += n;  S.xn = o; o += n;  S.u = o; o += m;  S.uprev = o; o += m;  S.dx = o; o += n;
    S.z = o; o += nz; S.e = o; o += nz;
    // index_lin: a step's linearisation is a pointer into the bank, no shared-memory copy
    S.A = o; o += index_lin ? 0 : n * n; S.B = o; o += index_lin ? 0 : n * m; S.d = o; o += index_lin ? 0 : n;
    S.H = o; o += nz * n; S.Qe = o; o += nz; S.Rdu = o; o += m; S.mscr = o; o += mscr;
    S.fwd_end = o;
    // backward (aliases the forward region)
    o = base;
    S.P = o; o += n * n;  S.p = o; o += n;  S.AtP = o; o += n * n;  S.BtP = o; o += m * n;  S.BtPr = o; o += m * n;
    S.Qux = o; o += m * n; S.Quxt = o; o += m * n; S.Quu = o; o += m * m; S.Quut = o; o += m * m;
    S.Lc = o; o += m * m;  S.LU = o; o += m * m;  S.inv = o; o += m * m;  S.K = o; o += m * n;  S.k = o; o += m;
    S.KQ = o; o += n * m;  S.Qx = o; o += n;  S.Qu = o; o += m;  S.cx = o; o += n;  S.cu = o; o += m;
    S.T1 = o; o += n * nz; S.cxx = o; o += gn ? n * n : 0; S.T1f = o; o += n * nz;
    S.bwd_end = big ? big_plan_end(n, m, nz, base) : o;
    // the backward pass also needs H_t / e_t / u rows staged: reuse tail
    S.total = (S.fwd_end > S.bwd_end ? S.fwd_end : S.bwd_end) + nz * n + nz + 2 * m + 4;
    return S;
}

// x+ = (A x + B u) + d
template <int NT>
__device__ __forceinline__ void affine_step(const LinRef& lin, const double* __restrict__ x, const double* __restrict__ u,
                                            double* __restrict__ xn, int n, int m) {
    if (NT == 32) {
        for (int i = threadIdx.x; i < n; i += NT) {
            double ax = 0.0, bu = 0.0;
            for (int k = 0; k < n; ++k) ax = fma(lin.A[i * n + k], x[k], ax);
            for (int k = 0; k < m; ++k) bu = fma(lin.B[i * m + k], u[k], bu);
            xn[i] = __dadd_rn(__dadd_rn(ax, bu), lin.d[i]);
        }
    } else {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int i = warp; i < n; i += NT / 32) {
            double ax = 0.0, bu = 0.0;
            for (int k = lane; k < n; k += 32) ax = fma(lin.A[i * n + k], x[k], ax);
            for (int k = lane; k < m; k += 32) bu = fma(lin.B[i * m + k], u[k], bu);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, off);
                bu += __shfl_xor_sync(0xffffffffu, bu, off);
            }
            if (lane == 0) xn[i] = __dadd_rn(__dadd_rn(ax, bu), lin.d[i]);
        }
    }
    cta_sync<NT>();
}

}  // namespace srcb
#include "ilqr_bwd_big.cuh"
#include "ilqr_fwd_tpwl.cuh"
namespace srcb {

// ---------------------------------------------------------------------------------------------------------------
// Forward pass (ilqr.py:117-162).  nom: nominal trajectory (x_prev, u_prev); K/k may be nullptr (zeros).
// Writes the trial record `tr` (x, u, e = z - z*, H_t, A_t/B_t or idx_t) and returns the cost to every thread.
// Aout/Bout/dout: optional dense outputs (N x n x n, N x n x m, N x n) for the standalone entry point.
// ---------------------------------------------------------------------------------------------------------------
template <class MP>
__device__ double fwd_pass(const typename MP::Dev& M, const IlqrArgs& a, const Smem& S, double* sm,
                           const double* __restrict__ nx, const double* __restrict__ nu, double alpha,
                           const double* __restrict__ K, const double* __restrict__ k, const Rec& tr,
                           const double* __restrict__ ztar, const double* __restrict__ ulast,
                           double* __restrict__ dout) {
    constexpr int NT = MP::NT;
    if constexpr (MP::NT > 32) {
        if (a.index_lin && !a.gn) return fwd_pass_tpwl_nn<MP>(M, a, S, sm, nx, nu, alpha, K, k, tr, ztar, ulast, dout);
    }
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, N = a.N, tid = threadIdx.x;
    double* sx = sm + S.x;  double* sxn = sm + S.xn;  double* su = sm + S.u;  double* sup = sm + S.uprev;
    double* sdx = sm + S.dx; double* sz = sm + S.z;   double* se = sm + S.e;  double* sA = sm + S.A;
    double* sB = sm + S.B;  double* sd = sm + S.d;    double* sH = sm + S.H;  double* sQe = sm + S.Qe;
    double* sRdu = sm + S.Rdu; double* mscr = sm + S.mscr;
    const double* sQ = sm + S.Qs; const double* sR = sm + S.Rs; const double* sQf = sm + S.Qfs;
    double* scal = sm + S.scal;
    double cost = 0.0;   // meaningful on thread 0

    for (int i = tid; i < n; i += NT) { sx[i] = nx[i]; tr.x[i] = nx[i]; }
    for (int i = tid; i < m; i += NT) sup[i] = ulast ? ulast[i] : 0.0;
    cta_sync<NT>();

    for (int t = 0; t < N; ++t) {
        // u_t = u_prev[t] + alpha * k[t] + K[t] @ (x[t] - x_prev[t])   (ilqr.py:140)
        for (int i = tid; i < n; i += NT) sdx[i] = __dsub_rn(sx[i], nx[t * n + i]);
        cta_sync<NT>();
        for (int i = tid; i < m; i += NT) {
            double v = nu[t * m + i];
            if (k) v = __dadd_rn(v, __dmul_rn(alpha, k[t * m + i]));
            if (K) {
                double acc = 0.0;
                for (int j = 0; j < n; ++j) acc = fma(K[((long long)t * m + i) * n + j], sdx[j], acc);
                v = __dadd_rn(v, acc);
            }
            su[i] = v;
            tr.u[t * m + i] = v;
        }
        cta_sync<NT>();
        // model: linearise at (x_t, u_t) and observe z_t (+ H_t)
        LinRef lin;
        int idx = 0;
        MP::eval(M, sx, su, a.dt, sA, sB, sd, sz, a.gn ? sH : nullptr, mscr, lin, idx);
        cta_sync<NT>();
        // e_t = z_t - z*_t ; du = u_t - u_{t-1}
        for (int i = tid; i < nz; i += NT) { se[i] = __dsub_rn(sz[i], ztar[t * nz + i]); tr.e[t * nz + i] = se[i]; }
        for (int i = tid; i < m; i += NT)
            sup[i] = a.cfg.include_input_var_constraint ? __dsub_rn(su[i], sup[i]) : su[i];   // sup now holds du
        cta_sync<NT>();
        // step cost (ilqr.py:168-175): .5 e^T Q e + .5 du^T R du, row-vector-times-matrix first
        for (int j = tid; j < nz; j += NT) {
            double acc = 0.0;
            for (int i = 0; i < nz; ++i) acc = fma(se[i], sQ[i * nz + j], acc);
            sQe[j] = acc;
        }
        for (int j = tid; j < m; j += NT) {
            double acc = 0.0;
            for (int i = 0; i < m; ++i) acc = fma(sup[i], sR[i * m + j], acc);
            sRdu[j] = acc;
        }
        // persist the linearisation of this step
        if (a.index_lin) {
            if (tid == 0) tr.idx[t] = idx;
        } else {
            for (int e = tid; e < n * n; e += NT) tr.A[(long long)t * n * n + e] = lin.A[e];
            for (int e = tid; e < n * m; e += NT) tr.B[(long long)t * n * m + e] = lin.B[e];
        }
        if (dout) for (int e = tid; e < n; e += NT) dout[(long long)t * n + e] = lin.d[e];
        if (a.gn) for (int e = tid; e < nz * n; e += NT) tr.H[(long long)t * nz * n + e] = sH[e];
        cta_sync<NT>();
        if (tid == 0) {
            double s1 = 0.0, s2 = 0.0;
            for (int j = 0; j < nz; ++j) s1 = fma(sQe[j], se[j], s1);
            for (int j = 0; j < m; ++j) s2 = fma(sRdu[j], sup[j], s2);
            cost = __dadd_rn(cost, __dadd_rn(__dmul_rn(0.5, s1), __dmul_rn(0.5, s2)));
        }
        // x_{t+1} = A x + B u + d
        affine_step<NT>(lin, sx, su, sxn, n, m);
        for (int i = tid; i < n; i += NT) { sx[i] = sxn[i]; tr.x[(long long)(t + 1) * n + i] = sxn[i]; }
        for (int i = tid; i < m; i += NT) sup[i] = su[i];
        cta_sync<NT>();
    }
    // terminal cost (ilqr.py:164-166)
    MP::observe(M, sx, sz, a.gn ? sH : nullptr, mscr);
    for (int i = tid; i < nz; i += NT) { se[i] = __dsub_rn(sz[i], ztar[N * nz + i]); tr.e[N * nz + i] = se[i]; }
    if (a.gn) for (int e = tid; e < nz * n; e += NT) tr.H[(long long)N * nz * n + e] = sH[e];
    cta_sync<NT>();
    for (int j = tid; j < nz; j += NT) {
        double acc = 0.0;
        for (int i = 0; i < nz; ++i) acc = fma(se[i], sQf[i * nz + j], acc);
        sQe[j] = acc;
    }
    cta_sync<NT>();
    if (tid == 0) {
        double s1 = 0.0;
        for (int j = 0; j < nz; ++j) s1 = fma(sQe[j], se[j], s1);
        cost = __dadd_rn(cost, __dmul_rn(0.5, s1));
        scal[0] = cost;
    }
    cta_sync<NT>();
    cost = scal[0];
    cta_sync<NT>();
    return cost;
}

// ---------------------------------------------------------------------------------------------------------------
// Backward pass (ilqr.py:219-300).  Reads the accepted record `rc`, writes K (N x m x n), k (N x m) and the two
// line-search scalars per step ab[2t] = k_t . Q_u,t, ab[2t+1] = (k_t^T Q_uu,t) . k_t.  Optional dense Q_u / Q_uu.
// Returns the horizon index of the (first) failed PD test, -1 if every Q_uu~ was PD; rho/drho are updated in place.
// Control flow of the failed test = the reference's literal code (ilqr.py:282-299): with `regularize` rho is raised,
// the sweep STOPS at that step (K_t = k_t = 0 for every t <= t_fail, Q_u / Q_uu of t_fail already written), then rho
// is lowered once like after a complete sweep -- there is no restart.  Without `regularize` the sweep goes on.
// ---------------------------------------------------------------------------------------------------------------
template <class MP>
__device__ int bwd_pass(const typename MP::Dev& M, const IlqrArgs& a, const Smem& S, double* sm, const Rec& rc,
                        const double* __restrict__ Adense, const double* __restrict__ Bdense,
                        const double* __restrict__ ulast, double* __restrict__ Kout, double* __restrict__ kout,
                        double* __restrict__ ab, double* __restrict__ Quout, double* __restrict__ Quuout,
                        double& rho, double& drho, double* __restrict__ cxx_global) {
    constexpr int NT = MP::NT;
    if constexpr (MP::NT > 32) {
        if (use_big_bwd(NT, MP::CN ? MP::CN : a.n, a.gn))
            return bwd_pass_big<MP>(M, a, S, sm, rc, Adense, Bdense, ulast, Kout, kout, ab, Quout, Quuout, rho, drho,
                                    cxx_global);
    }
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, N = a.N, tid = threadIdx.x;
    double* P = sm + S.P;      double* p = sm + S.p;      double* AtP = sm + S.AtP;   double* BtP = sm + S.BtP;
    double* BtPr = sm + S.BtPr; double* Qux = sm + S.Qux; double* Quxt = sm + S.Quxt; double* Quu = sm + S.Quu;
    double* Quut = sm + S.Quut; double* Lc = sm + S.Lc;   double* LU = sm + S.LU;     double* inv = sm + S.inv;
    double* Kt = sm + S.K;     double* kt = sm + S.k;     double* KQ = sm + S.KQ;     double* Qx = sm + S.Qx;
    double* Qu = sm + S.Qu;    double* cx = sm + S.cx;    double* cu = sm + S.cu;     double* T1 = sm + S.T1;
    double* cxx = a.gn ? sm + S.cxx : cxx_global;  double* T1f = sm + S.T1f;
    const int tail = (S.fwd_end > S.bwd_end ? S.fwd_end : S.bwd_end);
    double* sH = sm + tail;    double* se = sH + nz * n;  double* sdu = se + nz;      double* sut = sdu + m;
    const double* sQ = sm + S.Qs; const double* sR = sm + S.Rs; const double* sQf = sm + S.Qfs;
    const double* sHc = sm + S.Hcs;
    int* piv = reinterpret_cast<int*>(sm + S.ints);
    int* flag = piv + m + 2;
    const srcb200_ilqr_config& c = a.cfg;
    int pd_fail = -1;

    // constant-H cost Hessians are computed once per pass (cheap): T1 = H^T Q, cxx = T1 H
    if (!a.gn) {
        mm<NT, true, false>(T1, nz, sHc, n, sQ, nz, n, nz, nz);
        cta_sync<NT>();
        mm<NT, false, false>(cxx, n, T1, nz, sHc, n, n, n, nz);
        cta_sync<NT>();
    }

    {
        // terminal_cost_vectors (ilqr.py:177-182): P_N = (H^T Qf) H, p_N = (H^T Qf) e_N
        const double* HN = sHc;
        if (a.gn) {
            for (int e = tid; e < nz * n; e += NT) sH[e] = rc.H[(long long)N * nz * n + e];
            HN = sH;
        }
        for (int i = tid; i < nz; i += NT) se[i] = rc.e[N * nz + i];
        cta_sync<NT>();
        mm<NT, true, false>(T1f, nz, HN, n, sQf, nz, n, nz, nz);
        cta_sync<NT>();
        mm<NT, false, false>(P, n, T1f, nz, HN, n, n, n, nz);
        mv<NT, false>(p, T1f, nz, se, n, nz);
        cta_sync<NT>();

        for (int t = N - 1; t >= 0; --t) {
            // ---- stage step data
            LinRef lin;
            if (a.index_lin) lin = MP::bank(M, rc.idx[t]);
            else if (Adense) lin = LinRef{Adense + (long long)t * n * n, Bdense + (long long)t * n * m, nullptr};
            else lin = LinRef{rc.A + (long long)t * n * n, rc.B + (long long)t * n * m, nullptr};
            const double* Ht = sHc;
            if (a.gn) {
                for (int e = tid; e < nz * n; e += NT) sH[e] = rc.H[(long long)t * nz * n + e];
                Ht = sH;
            }
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
            // ---- step_cost_vectors (ilqr.py:186-196)
            if (a.gn) {
                mm<NT, true, false>(T1, nz, Ht, n, sQ, nz, n, nz, nz);
                cta_sync<NT>();
                mm<NT, false, false>(cxx, n, T1, nz, Ht, n, n, n, nz);
            }
            mv<NT, false>(cx, T1, nz, se, n, nz);
            mv<NT, false>(cu, sR, m, sdu, m, m);
            cta_sync<NT>();
            // ---- Q terms (ilqr.py:258-262)
            mv<NT, true>(Qx, lin.A, n, p, n, n, cx);            // Q_x = c_x + A^T p
            mv<NT, true>(Qu, lin.B, m, p, m, n, cu);            // Q_u = c_u + B^T p
            const bool big = (NT > 32) && (n >= 16);                // n x n x n products on the DMMA pipe (TPWL)
            if (big) mm_dmma<NT, true, false>(AtP, n, lin.A, n, P, n, n, n, n);
            else     mm<NT, true, false>(AtP, n, lin.A, n, P, n, n, n, n);   // A^T P
            mm<NT, true, false>(BtP, n, lin.B, m, P, n, m, n, n);   // B^T P
            if (c.regularize && c.state_regularization) {
                // B^T (P + rho I): the diagonal is P_jj + rho, off-diagonals P_kj + 0 (ilqr.py:266-267)
                for (int e = tid; e < m * n; e += NT) {
                    const int i = e / n, j = e - i * n;
                    double acc = 0.0;
                    for (int k2 = 0; k2 < n; ++k2) {
                        const double pv = (k2 == j) ? __dadd_rn(P[k2 * n + j], rho) : P[k2 * n + j];
                        acc = fma(lin.B[k2 * m + i], pv, acc);
                    }
                    BtPr[e] = acc;
                }
            }
            cta_sync<NT>();
            mm<NT, false, false>(Quu, m, BtP, n, lin.B, m, m, m, n, sR, m);   // Q_uu = c_uu + (B^T P) B
            mm<NT, false, false>(Qux, n, BtP, n, lin.A, n, m, n, n);          // Q_ux = (B^T P) A
            if (c.regularize && c.state_regularization) {
                mm<NT, false, false>(Quut, m, BtPr, n, lin.B, m, m, m, n, sR, m);
                mm<NT, false, false>(Quxt, n, BtPr, n, lin.A, n, m, n, n);
            }
            cta_sync<NT>();
            if (!(c.regularize && c.state_regularization)) {
                for (int e = tid; e < m * m; e += NT) {
                    const int i = e / m, j = e - i * m;
                    Quut[e] = (c.regularize && i == j) ? __dadd_rn(Quu[e], rho) : Quu[e];
                }
                for (int e = tid; e < m * n; e += NT) Quxt[e] = Qux[e];
                cta_sync<NT>();
            }
            // ---- PD test by Cholesky (ilqr.py:276-287)
            const bool pd = cholesky_pd<NT>(Quut, Lc, flag, m);
            if (!pd && pd_fail < 0) pd_fail = t;
            if (!pd && c.regularize) {
                // ilqr.py:282-287: raise rho and leave the sweep.  Q_u[t], Q_uu[t] are already assigned (258-261);
                // K, k (and with them the line-search scalars) stay at their zero initialisation for every s <= t.
                rho_update(c, true, rho, drho);
                if (Quout) for (int i = tid; i < m; i += NT) Quout[t * m + i] = Qu[i];
                if (Quuout) for (int e = tid; e < m * m; e += NT) Quuout[(long long)t * m * m + e] = Quu[e];
                for (long long e = tid; e < (long long)(t + 1) * m * n; e += NT) Kout[e] = 0.0;
                for (int e = tid; e < (t + 1) * m; e += NT) kout[e] = 0.0;
                if (Quout) for (int e = tid; e < t * m; e += NT) Quout[e] = 0.0;
                if (Quuout) for (long long e = tid; e < (long long)t * m * m; e += NT) Quuout[e] = 0.0;
                for (int e = tid; e < 2 * t; e += NT) ab[e] = 0.0;
                if (tid == 0) {
                    // alpha * k^T Q_u + alpha^2/2 * k^T Q_uu k with k = 0, evaluated literally (0 * inf = nan)
                    double s = 0.0, q = 0.0;
                    for (int i = 0; i < m; ++i) s = __dadd_rn(s, __dmul_rn(0.0, Qu[i]));
                    for (int j = 0; j < m; ++j) {
                        double v = 0.0;
                        for (int i = 0; i < m; ++i) v = __dadd_rn(v, __dmul_rn(0.0, Quu[i * m + j]));
                        q = __dadd_rn(q, __dmul_rn(v, 0.0));
                    }
                    ab[2 * t] = s;
                    ab[2 * t + 1] = q;
                }
                cta_sync<NT>();
                break;
            }
            // ---- gains (ilqr.py:289-292): explicit inverse, K = -inv Q_ux~, k = -inv Q_u
            for (int e = tid; e < m * m; e += NT) LU[e] = Quut[e];
            cta_sync<NT>();
            lu_inverse<NT>(LU, inv, piv, m);
            for (int e = tid; e < m * n; e += NT) {
                const int i = e / n, j = e - i * n;
                double acc = 0.0;
                for (int k2 = 0; k2 < m; ++k2) acc = fma(inv[i * m + k2], Quxt[k2 * n + j], acc);
                Kt[e] = -acc;
            }
            for (int i = tid; i < m; i += NT) {
                double acc = 0.0;
                for (int k2 = 0; k2 < m; ++k2) acc = fma(inv[i * m + k2], Qu[k2], acc);
                kt[i] = -acc;
            }
            cta_sync<NT>();
            // ---- value function (ilqr.py:294-295)
            mm<NT, true, false>(KQ, m, Kt, n, Quu, m, n, m, m);     // K^T Q_uu  (n x m)
            cta_sync<NT>();
            // p = ((Q_x + KQ k) + K^T Q_u) + Q_ux^T k
            for (int i = tid; i < n; i += NT) {
                double s1 = 0.0, s2 = 0.0, s3 = 0.0;
                for (int k2 = 0; k2 < m; ++k2) {
                    s1 = fma(KQ[i * m + k2], kt[k2], s1);
                    s2 = fma(Kt[k2 * n + i], Qu[k2], s2);
                    s3 = fma(Qux[k2 * n + i], kt[k2], s3);
                }
                p[i] = __dadd_rn(__dadd_rn(__dadd_rn(Qx[i], s1), s2), s3);
            }
            // P = (((c_xx + AtP A) + KQ K) + K^T Q_ux) + Q_ux^T K      (P is dead: overwrite in place)
            if (big) {
                mm_dmma<NT, false, false>(P, n, AtP, n, lin.A, n, n, n, n, cxx, n);     // c_xx + (A^T P) A
                cta_sync<NT>();
            }
            for (int e = tid; e < n * n; e += NT) {
                const int i = e / n, j = e - i * n;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                if (!big) for (int k2 = 0; k2 < n; ++k2) s0 = fma(AtP[i * n + k2], lin.A[k2 * n + j], s0);
                for (int k2 = 0; k2 < m; ++k2) {
                    s1 = fma(KQ[i * m + k2], Kt[k2 * n + j], s1);
                    s2 = fma(Kt[k2 * n + i], Qux[k2 * n + j], s2);
                    s3 = fma(Qux[k2 * n + i], Kt[k2 * n + j], s3);
                }
                const double base = big ? P[e] : __dadd_rn(cxx[e], s0);
                P[e] = __dadd_rn(__dadd_rn(__dadd_rn(base, s1), s2), s3);
            }
            // ---- outputs of this step
            for (int e = tid; e < m * n; e += NT) Kout[(long long)t * m * n + e] = Kt[e];
            for (int i = tid; i < m; i += NT) kout[t * m + i] = kt[i];
            if (Quout) for (int i = tid; i < m; i += NT) Quout[t * m + i] = Qu[i];
            if (Quuout) for (int e = tid; e < m * m; e += NT) Quuout[(long long)t * m * m + e] = Quu[e];
            if (tid == 0) {
                double s = 0.0;
                for (int i = 0; i < m; ++i) s = fma(kt[i], Qu[i], s);
                double q = 0.0;
                for (int j = 0; j < m; ++j) {
                    double v = 0.0;
                    for (int i = 0; i < m; ++i) v = fma(kt[i], Quu[i * m + j], v);
                    q = fma(v, kt[j], q);
                }
                ab[2 * t] = s;
                ab[2 * t + 1] = q;
            }
            cta_sync<NT>();
        }
        rho_update(c, false, rho, drho);      // ilqr.py:298 -- reached after a complete AND after an interrupted sweep
    }
    return pd_fail;
}

// ---------------------------------------------------------------------------------------------------------------
// Solve kernel: ilqr_computation (ilqr.py:27-107)
// ---------------------------------------------------------------------------------------------------------------
template <class MP>
__device__ __forceinline__ void load_costs(const IlqrArgs& a, const Smem& S, double* sm) {
    constexpr int NT = MP::NT;
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, tid = threadIdx.x;
    for (int e = tid; e < nz * nz; e += NT) { sm[S.Qs + e] = a.Q[e]; sm[S.Qfs + e] = a.Qf[e]; }
    for (int e = tid; e < m * m; e += NT) sm[S.Rs + e] = a.R[e];
    for (int e = tid; e < nz * n; e += NT) sm[S.Hcs + e] = a.Hc ? a.Hc[e] : 0.0;
    cta_sync<NT>();
}

template <class MP>
__global__ void __launch_bounds__(MP::NT, 1)
ilqr_solve_kernel(typename MP::Dev M, IlqrArgs a) {
    constexpr int NT = MP::NT;
    extern __shared__ __align__(16) double sm[];
    const Smem S = make_smem(MP::CN ? MP::CN : a.n, MP::CM ? MP::CM : a.m, MP::CNZ ? MP::CNZ : a.nz, a.model_scratch, a.gn,
                              use_big_bwd(MP::NT, MP::CN ? MP::CN : a.n, a.gn), a.index_lin != 0);
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, N = a.N, tid = threadIdx.x;
    const srcb200_ilqr_config& c = a.cfg;
    load_costs<MP>(a, S, sm);

    // Problems differ a lot in iteration count (TPWL: 10 .. 51), so the unit of work is ONE iteration: everything a
    // solve needs between iterations lives in global memory (records, gains) plus 8 doubles of solver state, and a
    // CTA that finishes an iteration puts the problem back on the task queue (ilqr.cuh, namespace ilqrq; FIFO here) and takes the next task.  Results do not depend on which CTA runs which iteration.
    __shared__ int s_next, s_cls;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_next = ilqrq::pop_one(a.work_counter, a.queue_cap);
        __syncthreads();
        const long long b = s_next;
        if (b < 0) break;
        __threadfence();                // acquire: what the previous owner of this problem wrote
        double* wsb = a.ws + b * a.L.total;
        Rec rec[2] = {rec_at(wsb, a.L), rec_at(wsb + a.L.rec, a.L)};
        double* kbuf = wsb + a.L.k;
        double* ab = wsb + a.L.ab;
        double* Kbuf = a.oK + b * (long long)N * m * n;
        const double* ztar = a.z_target + (a.shared_target ? 0 : b * (long long)(N + 1) * nz);
        const double* ulast = a.u_last ? a.u_last + b * m : nullptr;
        double* trace = a.otrace ? a.otrace + b * (long long)(c.max_iter + 1) * 4 : nullptr;
        double* sv = wsb + a.L.state;   // [rho, drho, cost, (fails, cur), (status, trials), (it, started), (class, -)]
        int* svi = reinterpret_cast<int*>(sv + 3);

        double rho, drho, cost;
        int fails, cur, status, trials, it, cls;
        if (svi[5] != ilqrq::kStarted) {
            rho = c.rho0; drho = c.drho0;
            fails = 0; cur = 0; status = 0; trials = 0; it = 0;
            // initial rollout: x_prev = [x0, 0, ...], u = warm start or zeros, K = k = 0 (ilqr.py:41-49).
            // With K = 0 the x_prev rows are never used beyond row 0, so the nominal can alias the trial record.
            Rec& nom = rec[1];
            for (int i = tid; i < n; i += NT) nom.x[i] = a.x0[b * n + i];
            for (long long e = tid; e < (long long)N * m; e += NT) nom.u[e] = a.u_init ? a.u_init[b * (long long)N * m + e] : 0.0;
            cta_sync<NT>();
            cost = fwd_pass<MP>(M, a, S, sm, rec[1].x, rec[1].u, 1.0, nullptr, nullptr, rec[0], ztar, ulast, nullptr);
            if (tid == 0) {
                if (a.ocost0) a.ocost0[b] = cost;
                // one class (FIFO): on the TPWL workload the initial cost does not predict the iteration count, measured
                // 2267 solves/s FIFO vs 2206 with the two-class rule that helps the Trunk-SSM kernel
                s_cls = 1;
            }
            __syncthreads();
            cls = s_cls;
        } else {
            rho = sv[0]; drho = sv[1]; cost = sv[2];
            fails = svi[0]; cur = svi[1]; status = svi[2]; trials = svi[3]; it = svi[4]; cls = svi[6];
        }

        bool conv = false, stop = false;
        if (it <= c.max_iter) {         // one pass of the `while not converged and nbr_iter <= max_iter` loop
            const int pd_fail = bwd_pass<MP>(M, a, S, sm, rec[cur], nullptr, nullptr, ulast, Kbuf, kbuf, ab, nullptr,
                                             nullptr, rho, drho, wsb + a.L.cxx);
            const double rho_bwd = rho;
            if (pd_fail >= 0) status |= SRCB200_ILQR_ST_NONPD;
            {
                const double prev_cost = cost;
                double alpha = c.alpha0;
                bool improved = false, failed = false;
                double cost_t = cost, alpha_acc = 0.0;
                while (!improved && !failed) {
                    improved = true;
                    cost_t = fwd_pass<MP>(M, a, S, sm, rec[cur].x, rec[cur].u, alpha, Kbuf, kbuf, rec[cur ^ 1], ztar, ulast, nullptr);
                    ++trials;
                    // delta_cost = sum_t alpha k^T Q_u + alpha^2/2 k^T Q_uu k, accumulated in t order (ilqr.py:69-71)
                    double dc = 0.0;
                    const double a2 = __dmul_rn(__dmul_rn(alpha, alpha), 0.5);
                    for (int t = 0; t < N; ++t)
                        dc = __dadd_rn(dc, __dadd_rn(__dmul_rn(alpha, ab[2 * t]), __dmul_rn(a2, ab[2 * t + 1])));
                    alpha_acc = alpha;
                    if (c.do_linesearch) {
                        const double ratio = __ddiv_rn(__dsub_rn(cost_t, prev_cost), dc);
                        if (ratio <= c.improv_lb || ratio > c.improv_ub) {
                            alpha = __dmul_rn(c.alpha_scaling, alpha);
                            improved = false;
                            if (alpha < c.alpha_min) {
                                rho_update(c, true, rho, drho);
                                rho = __dadd_rn(rho, c.rho_increase_fp);
                                failed = true;
                            }
                        }
                    }
                }
                if (!failed) {
                    cur ^= 1;
                    cost = cost_t;
                    const double dJ = __dsub_rn(prev_cost, cost);
                    conv = (dJ < c.epsilon) && (dJ >= 0.0);
                    if (conv) status |= SRCB200_ILQR_ST_CONVERGED;
                    fails = 0;
                } else {
                    ++fails;
                    if (fails >= c.counter_limit) { conv = true; status |= SRCB200_ILQR_ST_ABANDONED; }
                }
                if (trace && tid == 0) {
                    trace[it * 4 + 0] = cost;
                    trace[it * 4 + 1] = failed ? 0.0 : alpha_acc;
                    trace[it * 4 + 2] = rho_bwd;
                    trace[it * 4 + 3] = (double)pd_fail;
                }
                ++it;
                if (!isfinite(cost)) { status |= SRCB200_ILQR_ST_NONFINITE; stop = true; }
            }
        }
        if (!conv && !stop && it <= c.max_iter) {
            // not finished: save the state, hand the problem to whichever CTA is free next
            if (tid == 0) {
                sv[0] = rho; sv[1] = drho; sv[2] = cost;
                svi[0] = fails; svi[1] = cur; svi[2] = status; svi[3] = trials; svi[4] = it; svi[5] = ilqrq::kStarted; svi[6] = cls;
            }
            __threadfence();            // release: records / gains / state of every thread
            __syncthreads();
            if (tid == 0) ilqrq::push_one(a.work_counter, a.queue_cap, (int)b, cls);
            continue;
        }
        if (!conv && !stop && it > c.max_iter) status |= SRCB200_ILQR_ST_MAXITER;

        // results
        const Rec& fin = rec[cur];
        for (long long e = tid; e < (long long)(N + 1) * n; e += NT) a.ox[b * (long long)(N + 1) * n + e] = fin.x[e];
        for (long long e = tid; e < (long long)N * m; e += NT) a.ou[b * (long long)N * m + e] = fin.u[e];
        if (tid == 0) {
            a.ocost[b] = cost;
            if (a.orho) a.orho[b] = rho;
            a.oiter[b] = it;
            a.ostatus[b] = status;
            if (a.otrials) a.otrials[b] = trials;
            __threadfence();
            atomicSub(a.work_counter + ilqrq::Q_REMAINING, 1);
        }
        cta_sync<NT>();
    }
}

// standalone forward pass (srcb200_ilqr_forward_pass)
template <class MP>
__global__ void __launch_bounds__(MP::NT, 1)
ilqr_forward_kernel(typename MP::Dev M, IlqrArgs a, const double* __restrict__ xprev, const double* __restrict__ uprev,
                    double alpha, const double* __restrict__ K, const double* __restrict__ k, double* __restrict__ xo,
                    double* __restrict__ uo, double* __restrict__ costo, double* __restrict__ Ao, double* __restrict__ Bo,
                    double* __restrict__ dout) {
    constexpr int NT = MP::NT;
    extern __shared__ __align__(16) double sm[];
    const Smem S = make_smem(MP::CN ? MP::CN : a.n, MP::CM ? MP::CM : a.m, MP::CNZ ? MP::CNZ : a.nz, a.model_scratch, a.gn,
                              use_big_bwd(MP::NT, MP::CN ? MP::CN : a.n, a.gn), a.index_lin != 0);
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, N = a.N, tid = threadIdx.x;
    load_costs<MP>(a, S, sm);
    for (long long b = blockIdx.x; b < a.batch; b += gridDim.x) {
        double* wsb = a.ws + b * a.L.total;
        Rec tr = rec_at(wsb, a.L);
        const double* ztar = a.z_target + (a.shared_target ? 0 : b * (long long)(N + 1) * nz);
        const double* ulast = a.u_last ? a.u_last + b * m : nullptr;
        const double cost = fwd_pass<MP>(M, a, S, sm, xprev + b * (long long)(N + 1) * n, uprev + b * (long long)N * m, alpha,
                                         K ? K + b * (long long)N * m * n : nullptr, k ? k + b * (long long)N * m : nullptr, tr, ztar,
                                         ulast, dout ? dout + b * (long long)N * n : nullptr);
        for (long long e = tid; e < (long long)(N + 1) * n; e += NT) xo[b * (long long)(N + 1) * n + e] = tr.x[e];
        for (long long e = tid; e < (long long)N * m; e += NT) uo[b * (long long)N * m + e] = tr.u[e];
        if (tid == 0) costo[b] = cost;
        if (Ao || Bo) {
            for (int t = 0; t < N; ++t) {
                LinRef lin = a.index_lin ? MP::bank(M, tr.idx[t])
                                         : LinRef{tr.A + (long long)t * n * n, tr.B + (long long)t * n * m, nullptr};
                if (Ao) for (int e = tid; e < n * n; e += NT) Ao[(b * N + t) * (long long)n * n + e] = lin.A[e];
                if (Bo) for (int e = tid; e < n * m; e += NT) Bo[(b * N + t) * (long long)n * m + e] = lin.B[e];
            }
        }
        cta_sync<NT>();
    }
}

// standalone backward pass (srcb200_ilqr_backward_pass): e_t / H_t are rebuilt from x first
template <class MP>
__global__ void __launch_bounds__(MP::NT, 1)
ilqr_backward_kernel(typename MP::Dev M, IlqrArgs a, const double* __restrict__ x, const double* __restrict__ u,
                     const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ K,
                     double* __restrict__ k, double* __restrict__ Qu, double* __restrict__ Quu, double* __restrict__ rho_io,
                     double* __restrict__ drho_io, int* __restrict__ restarts_o) {
    constexpr int NT = MP::NT;
    extern __shared__ __align__(16) double sm[];
    const Smem S = make_smem(MP::CN ? MP::CN : a.n, MP::CM ? MP::CM : a.m, MP::CNZ ? MP::CNZ : a.nz, a.model_scratch, a.gn,
                              use_big_bwd(MP::NT, MP::CN ? MP::CN : a.n, a.gn), a.index_lin != 0);
    const int n = MP::CN ? MP::CN : a.n, m = MP::CM ? MP::CM : a.m, nz = MP::CNZ ? MP::CNZ : a.nz, N = a.N, tid = threadIdx.x;
    load_costs<MP>(a, S, sm);
    for (long long b = blockIdx.x; b < a.batch; b += gridDim.x) {
        double* wsb = a.ws + b * a.L.total;
        Rec rc = rec_at(wsb, a.L);
        const double* ztar = a.z_target + (a.shared_target ? 0 : b * (long long)(N + 1) * nz);
        const double* ulast = a.u_last ? a.u_last + b * m : nullptr;
        double* sx = sm + S.x; double* sz = sm + S.z; double* sH = sm + S.H; double* mscr = sm + S.mscr;
        for (int t = 0; t <= N; ++t) {
            for (int i = tid; i < n; i += NT) sx[i] = x[(b * (N + 1) + t) * (long long)n + i];
            cta_sync<NT>();
            MP::observe(M, sx, sz, a.gn ? sH : nullptr, mscr);
            for (int i = tid; i < nz; i += NT) rc.e[t * nz + i] = __dsub_rn(sz[i], ztar[t * nz + i]);
            if (a.gn) for (int e = tid; e < nz * n; e += NT) rc.H[(long long)t * nz * n + e] = sH[e];
            cta_sync<NT>();
        }
        for (long long e = tid; e < (long long)N * m; e += NT) rc.u[e] = u[b * (long long)N * m + e];
        cta_sync<NT>();
        __threadfence_block();
        double rho = rho_io[b], drho = drho_io[b];
        IlqrArgs a2 = a;
        a2.index_lin = 0;     // dense A/B are supplied by the caller
        const int r = bwd_pass<MP>(M, a2, S, sm, rc, A + b * (long long)N * n * n, B + b * (long long)N * n * m, ulast,
                                   K + b * (long long)N * m * n, k + b * (long long)N * m, wsb + a.L.ab,
                                   Qu ? Qu + b * (long long)N * m : nullptr, Quu ? Quu + b * (long long)N * m * m : nullptr, rho, drho,
                                   wsb + a.L.cxx);
        if (tid == 0) {
            rho_io[b] = rho;
            drho_io[b] = drho;
            if (restarts_o) restarts_o[b] = r;          // horizon index of the failed PD test, -1: none
        }
        cta_sync<NT>();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
template <class MP>
static int fill_args(const typename MP::Dev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr,
                     IlqrArgs& a, size_t& smem) {
    memset(&a, 0, sizeof(a));
    a.n = MP::n(M); a.m = MP::m(M); a.nz = MP::nz(M); a.N = pr->N;
    if (a.nz <= 0) return fail(SRCB200_E_DIM, "ilqr: the model has no output map (Need to set output or meas. model)");
    if (a.m > 32) return fail(SRCB200_E_DIM, "ilqr: m=%d > 32", a.m);
    a.gn = pr->gauss_newton ? 1 : 0;
    a.index_lin = MP::index_lin(M, pr->dt) ? 1 : 0;
    a.shared_target = pr->shared_target;
    a.batch = pr->batch;
    a.dt = pr->dt;
    if (cfg) a.cfg = *cfg;
    a.x0 = pr->x0; a.u_init = pr->u_init; a.z_target = pr->z_target; a.u_last = pr->u_last;
    a.Q = pr->Q; a.R = pr->R; a.Qf = pr->Qf; a.Hc = pr->H_const;
    a.L = make_layout(a.n, a.m, a.nz, a.N, a.gn, a.index_lin);
    a.model_scratch = MP::scratch_doubles(M, pr->dt);
    const Smem S = make_smem(a.n, a.m, a.nz, a.model_scratch, a.gn, use_big_bwd(MP::NT, a.n, a.gn), a.index_lin != 0);
    smem = sizeof(double) * (size_t)S.total;
    if (smem > 227 * 1024) return fail(SRCB200_E_DIM, "ilqr: n=%d m=%d needs %zu B of shared memory per CTA", a.n, a.m, smem);
    return 0;
}

template <class MP>
static int grid_size(long long batch, size_t smem, int waves = 4) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long per_sm = (long long)(227 * 1024) / (long long)(smem + 1024);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 32) per_sm = 32;
    const long long cap = per_sm * sms * waves;      // waves = 1: persistent CTAs (the solve kernel's work queue)
    return (int)(batch < cap ? batch : cap);
}

template <class MP>
static int solve_impl(const typename MP::Dev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr,
                      const srcb200_ilqr_result* res, void* ws, size_t ws_bytes, cudaStream_t st) {
    IlqrArgs a;
    size_t smem;
    if (int e = fill_args<MP>(M, cfg, pr, a, smem)) return e;
    if (!res || !res->x || !res->u || !res->K || !res->cost || !res->iterations || !res->status)
        return fail(SRCB200_E_NULL, "ilqr: result x/u/K/cost/iterations/status must be provided");
    if (!ws || ws_bytes < sizeof(double) * (size_t)a.L.total * (size_t)a.batch + ilqr_queue_bytes(a.batch))
        return fail(SRCB200_E_WORKSPACE, "ilqr: workspace too small (%zu < %zu)", ws_bytes,
                    sizeof(double) * (size_t)a.L.total * (size_t)a.batch + ilqr_queue_bytes(a.batch));
    a.work_counter = reinterpret_cast<int*>((double*)ws + (size_t)a.L.total * (size_t)a.batch);
    a.queue_cap = (int)ilqr_queue_cap(a.batch);
    a.prio_frac = 1.0;          // fast kernel: HIGH class = initial cost above the running mean (0.75 .. 1.25 measured equal)
    a.ox = res->x; a.ou = res->u; a.oK = res->K; a.ocost = res->cost; a.ocost0 = res->cost0; a.orho = res->rho;
    a.otrace = res->trace; a.oiter = res->iterations; a.ostatus = res->status; a.otrials = res->trials;
    a.ws = (double*)ws;
    if constexpr (std::is_same<MP, SsmPolicy>::value) {
        bool handled = false;
        if (int e = ilqr_ssm_fast_launch(M, a, st, &handled)) return e;
        if (handled) return 0;
    }
    auto kern = ilqr_solve_kernel<MP>;
    SRCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ilqrq::queue_init_kernel<<<(a.queue_cap + 255) / 256, 256, 0, st>>>(a.work_counter, a.queue_cap, (int)a.batch, a.ws,
                                                                        a.L.total, a.L.state);
    SRCB_LAUNCH_CHECK("queue_init_kernel");
    kern<<<grid_size<MP>(a.batch, smem, 1), MP::NT, smem, st>>>(M, a);
    SRCB_LAUNCH_CHECK("ilqr_solve_kernel");
    return 0;
}

template <class MP>
static int forward_impl(const typename MP::Dev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr,
                        const double* xp, const double* up, double alpha, const double* K, const double* k, double* x,
                        double* u, double* cost, double* A, double* B, double* d, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
    IlqrArgs a;
    size_t smem;
    if (int e = fill_args<MP>(M, cfg, pr, a, smem)) return e;
    if (!xp || !up || !x || !u || !cost) return fail(SRCB200_E_NULL, "forward_pass: NULL argument");
    if (!ws || ws_bytes < sizeof(double) * (size_t)a.L.total * (size_t)a.batch)
        return fail(SRCB200_E_WORKSPACE, "forward_pass: workspace too small");
    a.ws = (double*)ws;
    auto kern = ilqr_forward_kernel<MP>;
    SRCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_size<MP>(a.batch, smem), MP::NT, smem, st>>>(M, a, xp, up, alpha, K, k, x, u, cost, A, B, d);
    SRCB_LAUNCH_CHECK("ilqr_forward_kernel");
    return 0;
}

template <class MP>
static int backward_impl(const typename MP::Dev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr,
                         const double* x, const double* u, const double* A, const double* B, double* K, double* k,
                         double* Qu, double* Quu, double* rho, double* drho, int32_t* restarts, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
    IlqrArgs a;
    size_t smem;
    if (int e = fill_args<MP>(M, cfg, pr, a, smem)) return e;
    if (!x || !u || !A || !B || !K || !k || !rho || !drho) return fail(SRCB200_E_NULL, "backward_pass: NULL argument");
    if (!ws || ws_bytes < sizeof(double) * (size_t)a.L.total * (size_t)a.batch)
        return fail(SRCB200_E_WORKSPACE, "backward_pass: workspace too small");
    a.ws = (double*)ws;
    auto kern = ilqr_backward_kernel<MP>;
    SRCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_size<MP>(a.batch, smem), MP::NT, smem, st>>>(M, a, x, u, A, B, K, k, Qu, Quu, rho, drho, restarts);
    SRCB_LAUNCH_CHECK("ilqr_backward_kernel");
    return 0;
}


}  // namespace srcb
