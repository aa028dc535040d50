// ssm.cuh -- device-side SSM polynomial model evaluation (generic CTA-cooperative path).
// Reference math: sofacontrol/SSM/ssm.py:158-178 (basis + maps), 198-235 (Jacobians), 279-301 (discretisation).
#pragma once
#include "common.cuh"

namespace srcb {

struct SsmDev {
    int n, m, nz, order, nfeat, discr;
    const double* r;
    const double* w;
    const double* v;
    const double* B;
    const double* zref;
    const uint8_t* mono;   // nfeat x SRCB200_SSM_MAX_ORDER, 0xFF padded
};

inline SsmDev to_dev(const srcb200_ssm_model& s) {
    SsmDev d;
    d.n = s.n; d.m = s.m; d.nz = s.nz; d.order = s.order; d.nfeat = s.nfeat; d.discr = s.discr_method;
    d.r = s.r_coeff; d.w = s.w_coeff; d.v = s.v_coeff; d.B = s.B_r; d.zref = s.z_ref; d.mono = s.mono;
    return d;
}

int check_ssm_model(const srcb200_ssm_model* mdl);

// phi[k] (and dphi[k*n + j] when dphi != nullptr) for k < nfeat.  Products left to right over the sorted variable
// indices; derivative = multiplicity * (product of the remaining factors, left to right) -- the same order as the
// oracle (oracle/ssm_np.py poly_features / poly_features_jac).
template <int NT>
__device__ __forceinline__ void ssm_features(const SsmDev& M, const double* __restrict__ x,
                                             double* __restrict__ phi, double* __restrict__ dphi) {
    const int n = M.n;
    for (int k = threadIdx.x; k < M.nfeat; k += NT) {
        int idx[SRCB200_SSM_MAX_ORDER];
        int deg = 0;
#pragma unroll
        for (int q = 0; q < SRCB200_SSM_MAX_ORDER; ++q) {
            const int v = M.mono[k * SRCB200_SSM_MAX_ORDER + q];
            idx[q] = v;
            if (v != 0xFF) deg = q + 1;
        }
        double p = x[idx[0]];
        for (int q = 1; q < deg; ++q) p = __dmul_rn(p, x[idx[q]]);
        phi[k] = p;
        if (dphi) {
            for (int j = 0; j < n; ++j) dphi[k * n + j] = 0.0;
            for (int q = 0; q < deg; ++q) {
                const int j = idx[q];
                if (q > 0 && idx[q - 1] == j) continue;   // handle each distinct variable once
                int mult = 0;
                for (int s = 0; s < deg; ++s) mult += (idx[s] == j);
                double rest = 1.0;
                bool skipped = false;
                for (int s = 0; s < deg; ++s) {
                    if (!skipped && idx[s] == j) { skipped = true; continue; }
                    rest = __dmul_rn(rest, x[idx[s]]);
                }
                dphi[k * n + j] = __dmul_rn((double)mult, rest);
            }
        }
    }
    cta_sync<NT>();
}

// out[i] = sum_k C[i*nfeat + k] phi[k]   (i < rows)
template <int NT>
__device__ __forceinline__ void ssm_contract_vec(double* __restrict__ out, const double* __restrict__ C,
                                                 const double* __restrict__ phi, int rows, int nfeat) {
    for (int i = threadIdx.x; i < rows; i += NT) {
        double acc = 0.0;
        for (int k = 0; k < nfeat; ++k) acc = fma(C[i * nfeat + k], phi[k], acc);
        out[i] = acc;
    }
}

// out[i*n + j] = sum_k C[i*nfeat + k] dphi[k*n + j]
template <int NT>
__device__ __forceinline__ void ssm_contract_jac(double* __restrict__ out, const double* __restrict__ C,
                                                 const double* __restrict__ dphi, int rows, int n, int nfeat) {
    for (int e = threadIdx.x; e < rows * n; e += NT) {
        const int i = e / n, j = e - i * n;
        double acc = 0.0;
        for (int k = 0; k < nfeat; ++k) acc = fma(C[i * nfeat + k], dphi[k * n + j], acc);
        out[e] = acc;
    }
}

__host__ __device__ inline int ssm_eval_scratch_doubles(int n, int m, int nfeat) {
    return nfeat + nfeat * n + n + discretize_scratch_doubles(n, m);
}

// Full evaluation at (x, u), all operands in shared memory:
//   dynamics (if A != nullptr): A (n x n), B (n x m), d (n) discretised with dt (dt < 0: continuous)
//   observation (if z != nullptr): z = C(x) + z_ref; H (nz x n) = dC/dx if H != nullptr
// scratch: ssm_eval_scratch_doubles() doubles.
template <int NT>
__device__ void ssm_eval(const SsmDev& M, const double* __restrict__ x, const double* __restrict__ u, double dt,
                         double* __restrict__ A, double* __restrict__ B, double* __restrict__ d,
                         double* __restrict__ z, double* __restrict__ H, double* __restrict__ scratch,
                         double* __restrict__ zraw = nullptr) {
    const int n = M.n, m = M.m, nz = M.nz, nf = M.nfeat, tid = threadIdx.x;
    double* phi  = scratch;
    double* dphi = phi + nf;
    double* f    = dphi + nf * n;
    double* dscr = f + n;
    const bool need_jac = (A != nullptr) || (H != nullptr);
    ssm_features<NT>(M, x, phi, need_jac ? dphi : nullptr);
    if (A) {
        ssm_contract_jac<NT>(A, M.r, dphi, n, n, nf);
        // f = r_coeff phi + B_r u   (ssm.py:168): two dots, then one add
        for (int i = tid; i < n; i += NT) {
            double a = 0.0, b = 0.0;
            for (int k = 0; k < nf; ++k) a = fma(M.r[i * nf + k], phi[k], a);
            for (int k = 0; k < m; ++k) b = fma(M.B[i * m + k], u[k], b);
            f[i] = __dadd_rn(a, b);
        }
        for (int e = tid; e < n * m; e += NT) B[e] = M.B[e];
    }
    if (z) {
        for (int i = tid; i < nz; i += NT) {
            double a = 0.0;
            for (int k = 0; k < nf; ++k) a = fma(M.w[i * nf + k], phi[k], a);
            z[i] = __dadd_rn(a, M.zref[i]);
            if (zraw) zraw[i] = a;
        }
    }
    if (H) ssm_contract_jac<NT>(H, M.w, dphi, nz, n, nf);
    cta_sync<NT>();
    if (A) {
        // d = f - A x - B u   (ssm.py:203 / 211)
        for (int i = tid; i < n; i += NT) {
            double ax = 0.0, bu = 0.0;
            for (int k = 0; k < n; ++k) ax = fma(A[i * n + k], x[k], ax);
            for (int k = 0; k < m; ++k) bu = fma(B[i * m + k], u[k], bu);
            d[i] = __dsub_rn(__dsub_rn(f[i], ax), bu);
        }
        cta_sync<NT>();
        if (dt >= 0.0 && M.discr != SRCB200_DISCR_NONE) discretize_inplace<NT>(M.discr, dt, A, B, d, n, m, dscr);
    }
}

}  // namespace srcb
