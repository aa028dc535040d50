// expm.cu -- zero-order-hold discretisation of a batch of affine systems (A, B, d):
//     expm([[A, B, d], [0, 0, 0]] * dt)  ->  A_d = E[:n,:n], B_d = E[:n,n:n+m], d_d = E[:n,n+m]
// Replaces sofacontrol/utils.py:302-335 (zoh_linear / zoh_affine, scipy expm) as used by
// TPWLATV.discretize_dynamics('zoh') and pre_discretize (sofacontrol/tpwl/tpwl.py:272-322).
//
// Scaling and squaring with the degree-13 Pade approximant (Higham 2005 / Al-Mohy & Higham 2009, the family
// scipy's expm implements): s = max(0, ceil(log2(||M||_1 / theta_13))), theta_13 = 4.25; U, V from M^2, M^4, M^6;
// (V - U) X = V + U by LU with partial pivoting; s squarings.  One CTA per matrix, all matrices in a
// caller-provided global workspace (L2 resident), products by the cooperative routines of common.cuh.
#include "common.cuh"

namespace srcb {

constexpr int kExpmThreads = 256;
constexpr int kExpmMats = 9;   // M, M2, M4, M6, T1, T2, U, V, X

__global__ void __launch_bounds__(kExpmThreads)
zoh_expm_kernel(int n, int m, long long count, double dt, const double* A, const double* B, const double* d,
                double* Ad, double* Bd, double* dd, double* __restrict__ ws) {   // in-place (A == Ad ...) allowed
    constexpr int NT = kExpmThreads;
    const int s = n + m + 1, ss = s * s, tid = threadIdx.x;
    __shared__ double red[NT / 32];
    __shared__ int piv[192];
    __shared__ int s_squarings;
    double* base = ws + (size_t)blockIdx.x * kExpmMats * ss;
    double* Mx = base;            double* M2 = base + ss;      double* M4 = base + 2 * ss;  double* M6 = base + 3 * ss;
    double* T1 = base + 4 * ss;   double* T2 = base + 5 * ss;  double* U = base + 6 * ss;   double* V = base + 7 * ss;
    double* X = base + 8 * ss;
    const double b[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
                          10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.};
    for (long long it = blockIdx.x; it < count; it += gridDim.x) {
        // ---- M = [[A, B, d], [0]] * dt
        for (int e = tid; e < ss; e += NT) {
            const int i = e / s, j = e - i * s;
            double v = 0.0;
            if (i < n) {
                if (j < n) v = A[it * n * n + i * n + j];
                else if (j < n + m) v = B[it * n * m + i * m + (j - n)];
                else v = d[it * n + i];
            }
            Mx[e] = __dmul_rn(v, dt);
        }
        __syncthreads();
        // ---- 1-norm (max column sum) -> number of squarings
        double cmax = 0.0;
        for (int j = tid; j < s; j += NT) {
            double cs = 0.0;
            for (int i = 0; i < s; ++i) cs += fabs(Mx[i * s + j]);
            cmax = fmax(cmax, cs);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, off));
        if ((tid & 31) == 0) red[tid >> 5] = cmax;
        __syncthreads();
        if (tid == 0) {
            double nrm = 0.0;
            for (int k = 0; k < NT / 32; ++k) nrm = fmax(nrm, red[k]);
            int sq = 0;
            if (nrm > 4.25) sq = (int)ceil(log2(nrm / 4.25));
            if (sq < 0) sq = 0;
            if (sq > 60) sq = 60;
            s_squarings = sq;
        }
        __syncthreads();
        const int sq = s_squarings;
        const double scale = ldexp(1.0, -sq);
        for (int e = tid; e < ss; e += NT) Mx[e] *= scale;     // exact (power of two)
        __syncthreads();
        // ---- powers
        mm<NT, false, false>(M2, s, Mx, s, Mx, s, s, s, s);
        __syncthreads();
        mm<NT, false, false>(M4, s, M2, s, M2, s, s, s, s);
        __syncthreads();
        mm<NT, false, false>(M6, s, M4, s, M2, s, s, s, s);
        __syncthreads();
        // ---- U = M (M6 (b13 M6 + b11 M4 + b9 M2) + b7 M6 + b5 M4 + b3 M2 + b1 I)
        //      V =     M6 (b12 M6 + b10 M4 + b8 M2) + b6 M6 + b4 M4 + b2 M2 + b0 I
        for (int e = tid; e < ss; e += NT) {
            T1[e] = b[13] * M6[e] + b[11] * M4[e] + b[9] * M2[e];
            T2[e] = b[12] * M6[e] + b[10] * M4[e] + b[8] * M2[e];
        }
        __syncthreads();
        mm<NT, false, false>(X, s, M6, s, T1, s, s, s, s);     // X = M6 T1
        mm<NT, false, false>(V, s, M6, s, T2, s, s, s, s);     // V = M6 T2
        __syncthreads();
        for (int e = tid; e < ss; e += NT) {
            const int i = e / s, j = e - i * s;
            const double id = (i == j) ? 1.0 : 0.0;
            T1[e] = X[e] + b[7] * M6[e] + b[5] * M4[e] + b[3] * M2[e] + b[1] * id;
            V[e] = V[e] + b[6] * M6[e] + b[4] * M4[e] + b[2] * M2[e] + b[0] * id;
        }
        __syncthreads();
        mm<NT, false, false>(U, s, Mx, s, T1, s, s, s, s);     // U = M T1
        __syncthreads();
        // ---- (V - U) X = V + U
        for (int e = tid; e < ss; e += NT) {
            T1[e] = V[e] - U[e];
            T2[e] = V[e] + U[e];
        }
        __syncthreads();
        lu_inverse<NT>(T1, M2, piv, s);                         // M2 = inv(V - U)   (M2 is free now)
        mm<NT, false, false>(X, s, M2, s, T2, s, s, s, s);
        __syncthreads();
        // ---- squarings
        double* cur = X;
        double* nxt = M4;
        for (int k = 0; k < sq; ++k) {
            mm<NT, false, false>(nxt, s, cur, s, cur, s, s, s, s);
            __syncthreads();
            double* t = cur; cur = nxt; nxt = t;
        }
        // ---- slice
        for (int e = tid; e < n * n; e += NT) Ad[it * n * n + e] = cur[(e / n) * s + e % n];
        for (int e = tid; e < n * m; e += NT) Bd[it * n * m + e] = cur[(e / m) * s + n + e % m];
        for (int e = tid; e < n; e += NT) dd[it * n + e] = cur[e * s + n + m];
        __syncthreads();
    }
}

static int expm_grid(long long count) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long cap = 2LL * sms;
    return (int)(count < cap ? count : cap);
}

}  // namespace srcb

using namespace srcb;

extern "C" size_t srcb200_zoh_workspace(int32_t n, int32_t m, int64_t count) {
    if (n < 1 || m < 0 || count <= 0) return 0;
    const size_t s = (size_t)n + m + 1;
    return sizeof(double) * (size_t)expm_grid(count) * kExpmMats * s * s;
}

extern "C" int srcb200_zoh_batch(int32_t n, int32_t m, int64_t count, double dt, const double* A_c, const double* B_c,
                                 const double* d_c, double* A_d, double* B_d, double* d_d, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    if (n < 1 || m < 1 || n + m + 1 > 192 || count < 0) return fail(SRCB200_E_DIM, "zoh: bad dims n=%d m=%d", n, m);
    if (count == 0) return 0;
    if (!A_c || !B_c || !d_c || !A_d || !B_d || !d_d) return fail(SRCB200_E_NULL, "zoh: NULL operand");
    if (!workspace || workspace_bytes < srcb200_zoh_workspace(n, m, count)) return fail(SRCB200_E_WORKSPACE, "zoh: workspace too small");
    zoh_expm_kernel<<<expm_grid(count), kExpmThreads, 0, (cudaStream_t)stream>>>(n, m, count, dt, A_c, B_c, d_c, A_d, B_d, d_d,
                                                                                   (double*)workspace);
    SRCB_LAUNCH_CHECK("zoh_expm_kernel");
    return 0;
}
