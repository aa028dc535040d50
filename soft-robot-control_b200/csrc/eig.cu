// eig.cu -- eigen-decomposition of a small symmetric positive semi-definite matrix (n <= 160) in ONE CTA: the
// Rayleigh-Ritz / orthonormalisation step of the leading-eigenpair solver that replaces the full SVD of
// sofacontrol/mor/pod.py:191 (mor/eig.py: block subspace iteration on the snapshot Gram matrix).
//
// One-sided (Hestenes) Jacobi on the columns of W = T: plane rotations V make the columns of W V orthogonal; for a
// symmetric PSD T = Q L Q^T the limit is W V = Q L, so the column norms are the eigenvalues and the normalised
// columns the eigenvectors -- no accumulated rotation matrix is needed, the whole state is the n x n array in
// shared memory (column-major: a lane walks a column with stride 1).  A round of the round-robin tournament
// holds n / 2 independent pairs, spread over the 32 warps; dot products are warp-shuffle trees in FP64.
#include <math.h>
#include "common.cuh"

namespace srcb {
namespace eig {

constexpr int NT = 1024;
constexpr int MAXN = 160;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}

__global__ void __launch_bounds__(NT, 1)
sym_eig_psd_kernel(int n, const double* __restrict__ A, long long lda, double* __restrict__ evals,
                   double* __restrict__ V, long long ldv, int max_sweeps, double tol, int* __restrict__ info) {
    extern __shared__ __align__(16) double W[];          // n x n, column-major (+ n doubles of norms, n ints of ranks)
    __shared__ int s_rot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int np = (n + 1) & ~1;                          // players of the tournament (a dummy when n is odd)
    const int m = np - 1;
    double* nrm = W + (size_t)n * n;
    int* rank = reinterpret_cast<int*>(nrm + n);

    // column i of W = row i of the symmetrised input
    for (int e = tid; e < n * n; e += NT) {
        const int i = e / n, r = e - i * n;
        W[(size_t)i * n + r] = 0.5 * (A[(long long)i * lda + r] + A[(long long)r * lda + i]);
    }
    __syncthreads();

    int sweeps = 0;
    for (; sweeps < max_sweeps; ++sweeps) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int round = 0; round < m; ++round) {
            for (int k = warp; k < np / 2; k += NT / 32) {
                int a, b;
                if (k == 0) { a = round % m; b = np - 1; }
                else { a = (round + k) % m; b = (round - k + m) % m; }
                if (a >= n || b >= n) continue;           // the dummy player
                if (a > b) { const int t = a; a = b; b = t; }
                double* wa = W + (size_t)a * n;
                double* wb = W + (size_t)b * n;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int r = lane; r < n; r += 32) {
                    const double x = wa[r], y = wb[r];
                    al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
                }
                al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                if (fabs(ga) > tol * sqrt(al * be) && ga != 0.0) {
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    for (int r = lane; r < n; r += 32) {
                        const double x = wa[r], y = wb[r];
                        wa[r] = c * x - s * y;
                        wb[r] = s * x + c * y;
                    }
                    if (lane == 0) s_rot = 1;
                }
            }
            __syncthreads();
        }
        const int any = s_rot;
        __syncthreads();
        if (!any) break;
    }

    // eigenvalues = column norms, sorted descending (rank by counting; ties by index)
    for (int i = warp; i < n; i += NT / 32) {
        const double* wi = W + (size_t)i * n;
        double s2 = 0.0;
        for (int r = lane; r < n; r += 32) s2 = fma(wi[r], wi[r], s2);
        s2 = warp_sum(s2);
        if (lane == 0) nrm[i] = sqrt(s2);
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        int rk = 0;
        const double li = nrm[i];
        for (int j = 0; j < n; ++j) rk += (nrm[j] > li) || (nrm[j] == li && j < i);
        rank[i] = rk;
        evals[rk] = li;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += NT) {
        const int i = e / n, r = e - i * n;
        const double li = nrm[i];
        V[(long long)r * ldv + rank[i]] = li > 0.0 ? W[(size_t)i * n + r] / li : (r == i ? 1.0 : 0.0);
    }
    if (tid == 0 && info) *info = (sweeps >= max_sweeps) ? -1 : sweeps + 1;
}

}  // namespace eig
}  // namespace srcb

using namespace srcb;

extern "C" int srcb200_sym_eig_psd(int32_t n, const double* A, int64_t lda, double* evals, double* V, int64_t ldv,
                                   int32_t* info, void* stream) {
    if (n < 0 || n > eig::MAXN) return fail(SRCB200_E_DIM, "sym_eig_psd: n=%d outside 0..%d", n, eig::MAXN);
    if (n == 0) return 0;
    if (!A || !evals || !V) return fail(SRCB200_E_NULL, "sym_eig_psd: NULL operand");
    if (lda < n || ldv < n) return fail(SRCB200_E_DIM, "sym_eig_psd: leading dimension too small");
    const size_t smem = sizeof(double) * ((size_t)n * n + n) + sizeof(int) * n + 16;
    SRCB_CUDA(cudaFuncSetAttribute(eig::sym_eig_psd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eig::sym_eig_psd_kernel<<<1, eig::NT, smem, (cudaStream_t)stream>>>(n, A, lda, evals, V, ldv, 40, 1e-15 * sqrt((double)n), info);
    SRCB_LAUNCH_CHECK("sym_eig_psd_kernel");
    return 0;
}
