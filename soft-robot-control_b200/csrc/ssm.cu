// ssm.cu -- SSM polynomial ROM: batched evaluation + linearisation, polynomial maps, batched rollout.
// Reference: sofacontrol/SSM/ssm.py (see include/srcb200.h for the per-entry mapping).
#include "ssm.cuh"

namespace srcb {

int ssm_rollout_fast_launch(const SsmDev& M, long long batch, int N, const double* x0, const double* u, double dt,
                            double* x, double* z, cudaStream_t st, bool* handled);   // ilqr_fast.cu

int ssm_eval_dmma_launch(const SsmDev& M, long long count, const double* x, const double* u, double dt, double* A,
                         double* B, double* d, double* H, double* c, double* z, cudaStream_t st, bool* handled);   // ilqr_fast.cu

int check_ssm_model(const srcb200_ssm_model* s) {
    if (!s) return fail(SRCB200_E_NULL, "ssm model is NULL");
    if (s->n < 1 || s->n > SRCB200_SSM_MAX_N || s->m < 1 || s->m > SRCB200_SSM_MAX_M ||
        s->order < 1 || s->order > SRCB200_SSM_MAX_ORDER || s->nfeat < 1 || s->nfeat > SRCB200_SSM_MAX_FEAT)
        return fail(SRCB200_E_DIM, "ssm dims out of range: n=%d m=%d order=%d nfeat=%d", s->n, s->m, s->order, s->nfeat);
    if (s->nz != s->n)
        return fail(SRCB200_E_DIM, "ssm output_dim (%d) must equal state_dim (%d): the reference applies the output "
                                   "basis to x (ssm.py:39-40,170-174)", s->nz, s->n);
    if (!s->r_coeff || !s->w_coeff || !s->B_r || !s->z_ref || !s->mono)
        return fail(SRCB200_E_NULL, "ssm model has NULL coefficient pointers");
    if (s->discr_method != SRCB200_DISCR_FE && s->discr_method != SRCB200_DISCR_BE &&
        s->discr_method != SRCB200_DISCR_BIL && s->discr_method != SRCB200_DISCR_NONE)
        return fail(SRCB200_E_METHOD, "self.discr_method must be in [fe, be, bil, zoh]");   // ssm.py:299 (zoh raises too)
    return 0;
}

constexpr int kNT = 32;

__global__ void __launch_bounds__(kNT)
ssm_eval_kernel(SsmDev M, long long count, const double* __restrict__ x, const double* __restrict__ u, double dt,
                double* __restrict__ A, double* __restrict__ B, double* __restrict__ d, double* __restrict__ H,
                double* __restrict__ c, double* __restrict__ z) {
    extern __shared__ double sm[];
    const int n = M.n, m = M.m, nz = M.nz, tid = threadIdx.x;
    double* sx = sm;
    double* su = sx + n;
    double* sA = su + m;
    double* sB = sA + n * n;
    double* sd = sB + n * m;
    double* sz = sd + n;
    double* szr = sz + nz;
    double* sH = szr + nz;
    double* scr = sH + nz * n;
    const bool dyn = (A || B || d), obs = (z || H || c);
    for (long long s = blockIdx.x; s < count; s += gridDim.x) {
        for (int i = tid; i < n; i += kNT) sx[i] = x[s * n + i];
        for (int i = tid; i < m; i += kNT) su[i] = u ? u[s * m + i] : 0.0;
        cta_sync<kNT>();
        ssm_eval<kNT>(M, sx, su, dt, dyn ? sA : nullptr, sB, sd, obs ? sz : nullptr, (H || c) ? sH : nullptr, scr, szr);
        cta_sync<kNT>();
        if (A) for (int e = tid; e < n * n; e += kNT) A[s * n * n + e] = sA[e];
        if (B) for (int e = tid; e < n * m; e += kNT) B[s * n * m + e] = sB[e];
        if (d) for (int e = tid; e < n; e += kNT) d[s * n + e] = sd[e];
        if (z) for (int e = tid; e < nz; e += kNT) z[s * nz + e] = sz[e];
        if (H) for (int e = tid; e < nz * n; e += kNT) H[s * nz * n + e] = sH[e];
        if (c) {   // c_res = C(x) - H x  (ssm.py:234)
            for (int i = tid; i < nz; i += kNT) {
                double hx = 0.0;
                for (int k = 0; k < n; ++k) hx = fma(sH[i * n + k], sx[k], hx);
                c[s * nz + i] = __dsub_rn(szr[i], hx);
            }
        }
        cta_sync<kNT>();
    }
}

__global__ void __launch_bounds__(kNT)
ssm_map_kernel(SsmDev M, int which, int add_ref, long long count, const double* __restrict__ in,
               const double* __restrict__ u, double* __restrict__ out) {
    extern __shared__ double sm[];
    const int n = M.n, nf = M.nfeat, tid = threadIdx.x;
    double* sx = sm;
    double* phi = sx + n;
    const double* C = which == 0 ? M.w : (which == 1 ? M.v : M.r);
    for (long long s = blockIdx.x; s < count; s += gridDim.x) {
        for (int i = tid; i < n; i += kNT) {
            double v = in[s * n + i];
            if (which == 1 && add_ref) v = __dsub_rn(v, M.zref[i]);
            sx[i] = v;
        }
        cta_sync<kNT>();
        ssm_features<kNT>(M, sx, phi, nullptr);
        for (int i = tid; i < n; i += kNT) {
            double a = 0.0;
            for (int k = 0; k < nf; ++k) a = fma(C[i * nf + k], phi[k], a);
            if (which == 0 && add_ref) a = __dadd_rn(a, M.zref[i]);
            if (which == 2 && u) {
                double b = 0.0;
                for (int k = 0; k < M.m; ++k) b = fma(M.B[i * M.m + k], u[s * M.m + k], b);
                a = __dadd_rn(a, b);
            }
            out[s * n + i] = a;
        }
        cta_sync<kNT>();
    }
}

// One CTA (one warp) walks one trajectory: x+ = (A_d x + B_d u) + d_d with (A_d,B_d,d_d) re-linearised at every
// step (ssm.py:134-156, 187-195, 330-333); z = C_map(x) + z_ref for all N+1 states.
__global__ void __launch_bounds__(kNT)
ssm_rollout_kernel(SsmDev M, long long batch, int N, const double* __restrict__ x0, const double* __restrict__ u,
                   double dt, double* __restrict__ xo, double* __restrict__ zo) {
    extern __shared__ double sm[];
    const int n = M.n, m = M.m, nz = M.nz, tid = threadIdx.x;
    double* sx = sm;
    double* su = sx + n;
    double* sA = su + m;
    double* sB = sA + n * n;
    double* sd = sB + n * m;
    double* sz = sd + n;
    double* sxn = sz + nz;
    double* scr = sxn + n;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        double* xb = xo + b * (long long)(N + 1) * n;
        double* zb = zo ? zo + b * (long long)(N + 1) * nz : nullptr;
        const double* ub = u + b * (long long)N * m;
        for (int i = tid; i < n; i += kNT) { sx[i] = x0[b * n + i]; xb[i] = sx[i]; }
        cta_sync<kNT>();
        for (int t = 0; t < N; ++t) {
            for (int i = tid; i < m; i += kNT) su[i] = ub[t * m + i];
            cta_sync<kNT>();
            ssm_eval<kNT>(M, sx, su, dt, sA, sB, sd, zb ? sz : nullptr, nullptr, scr);
            cta_sync<kNT>();
            for (int i = tid; i < n; i += kNT) {
                double ax = 0.0, bu = 0.0;
                for (int k = 0; k < n; ++k) ax = fma(sA[i * n + k], sx[k], ax);
                for (int k = 0; k < m; ++k) bu = fma(sB[i * m + k], su[k], bu);
                sxn[i] = __dadd_rn(__dadd_rn(ax, bu), sd[i]);
            }
            if (zb) for (int i = tid; i < nz; i += kNT) zb[t * nz + i] = sz[i];
            cta_sync<kNT>();
            for (int i = tid; i < n; i += kNT) { sx[i] = sxn[i]; xb[(t + 1) * n + i] = sxn[i]; }
            cta_sync<kNT>();
        }
        if (zb) {
            ssm_eval<kNT>(M, sx, su, dt, nullptr, nullptr, nullptr, sz, nullptr, scr);
            cta_sync<kNT>();
            for (int i = tid; i < nz; i += kNT) zb[N * nz + i] = sz[i];
        }
        cta_sync<kNT>();
    }
}

static int grid_for(long long count) {
    const long long cap = 148LL * 32 * 8;   // persistent-ish: at most 8 waves of 32 one-warp CTAs per SM
    return (int)(count < cap ? (count < 1 ? 1 : count) : cap);
}

}  // namespace srcb

using namespace srcb;

extern "C" int srcb200_ssm_eval_linearize_batch(const srcb200_ssm_model* mdl, int64_t count, const double* x,
                                                const double* u, double dt, double* A, double* B, double* d,
                                                double* H, double* c, double* z, void* stream) {
    if (int e = check_ssm_model(mdl)) return e;
    if (count < 0) return fail(SRCB200_E_DIM, "count < 0");
    if (count == 0) return 0;
    if (!x) return fail(SRCB200_E_NULL, "x is NULL");
    if ((A || B || d) && !u) return fail(SRCB200_E_NULL, "u is NULL (Need to supply current input)");
    SsmDev M = to_dev(*mdl);
    {   // Trunk / Diamond shape: dense contraction on the FP64 tensor pipe (DMMA)
        bool handled = false;
        if (int e = ssm_eval_dmma_launch(M, count, x, u, dt, A, B, d, H, c, z, (cudaStream_t)stream, &handled)) return e;
        if (handled) return 0;
    }
    const int n = M.n, m = M.m, nz = M.nz;
    const size_t smem = sizeof(double) * (n + m + n * n + n * m + n + 2 * nz + nz * n + ssm_eval_scratch_doubles(n, m, M.nfeat));
    SRCB_CUDA(cudaFuncSetAttribute(ssm_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssm_eval_kernel<<<grid_for(count), kNT, smem, (cudaStream_t)stream>>>(M, count, x, u, dt, A, B, d, H, c, z);
    SRCB_LAUNCH_CHECK("ssm_eval_kernel");
    return 0;
}

extern "C" int srcb200_ssm_map_batch(const srcb200_ssm_model* mdl, int32_t which, int32_t add_ref, int64_t count,
                                     const double* in, const double* u, double* out, void* stream) {
    if (int e = check_ssm_model(mdl)) return e;
    if (which < 0 || which > 2) return fail(SRCB200_E_DIM, "which must be 0 (C_map), 1 (W_map) or 2 (r_coeff)");
    if (which == 1 && !mdl->v_coeff) return fail(SRCB200_E_NULL, "v_coeff is NULL");
    if (count < 0) return fail(SRCB200_E_DIM, "count < 0");
    if (count == 0) return 0;
    if (!in || !out) return fail(SRCB200_E_NULL, "in/out is NULL");
    SsmDev M = to_dev(*mdl);
    const size_t smem = sizeof(double) * (M.n + M.nfeat);
    ssm_map_kernel<<<grid_for(count), kNT, smem, (cudaStream_t)stream>>>(M, which, add_ref, count, in, u, out);
    SRCB_LAUNCH_CHECK("ssm_map_kernel");
    return 0;
}

extern "C" int srcb200_ssm_rollout_batch(const srcb200_ssm_model* mdl, int64_t batch, int32_t N, const double* x0,
                                         const double* u, double dt, double* x, double* z, void* stream) {
    if (int e = check_ssm_model(mdl)) return e;
    if (batch < 0 || N < 0) return fail(SRCB200_E_DIM, "batch/N < 0");
    if (batch == 0) return 0;
    if (!x0 || !x || (N > 0 && !u)) return fail(SRCB200_E_NULL, "x0/u/x is NULL");
    SsmDev M = to_dev(*mdl);
    {   // Trunk / Diamond shape: warp-per-trajectory kernel sharing the specialised iLQR device code
        bool handled = false;
        if (int e = ssm_rollout_fast_launch(M, batch, N, x0, u, dt, x, z, (cudaStream_t)stream, &handled)) return e;
        if (handled) return 0;
    }
    const int n = M.n, m = M.m, nz = M.nz;
    const size_t smem = sizeof(double) * (n + m + n * n + n * m + n + nz + n + ssm_eval_scratch_doubles(n, m, M.nfeat));
    SRCB_CUDA(cudaFuncSetAttribute(ssm_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssm_rollout_kernel<<<grid_for(batch), kNT, smem, (cudaStream_t)stream>>>(M, batch, N, x0, u, dt, x, z);
    SRCB_LAUNCH_CHECK("ssm_rollout_kernel");
    return 0;
}
