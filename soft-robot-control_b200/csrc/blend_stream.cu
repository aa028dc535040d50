// blend_stream.cu -- TPWL weighted bank blend for SMALL batches (kernel a3 of BASELINE.json):
//     out[b][e] = sum_p W[b][p] * bank[p][e],   b < nb <= 8,   bank = A_c (P x n^2), B_c (P x n m) or d_c (P x n)
// (np.einsum("i,ijk->jk", weights, bank), sofacontrol/tpwl/tpwl.py:246-248).  For a handful of trajectories the
// blend is a pure stream of the bank (44.4 MB at the Diamond size): every CTA owns a 128-column slab and a chunk
// of the stored points, pulls its rows through a 4-stage ring of TMA bulk copies (cp.async.bulk + mbarrier, no
// tensor map needed for contiguous rows) and accumulates the <= 8 weighted sums in registers; chunk partials are
// combined in a fixed order (deterministic).  Large batches use the DMMA GEMM instead (gemm.cu), where the bank
// tile is reused across the batch.
#include "common.cuh"

namespace srcb {

constexpr int BS_COLS = 128;     // threads per CTA = columns of a slab
constexpr int BS_ROWS = 16;      // bank rows per stage
constexpr int BS_STAGES = 4;
constexpr int BS_MAXB = 8;       // trajectories per pass
constexpr int BS_CHUNKS = 8;     // split of the stored points across CTAs (41 slabs x 8 chunks = 328 CTAs for A)

// grid = (slabs, BS_CHUNKS).  W is (nb x P) row-major; partial is (BS_CHUNKS x BS_MAXB x width).
__global__ void __launch_bounds__(BS_COLS)
blend_stream_kernel(const double* __restrict__ bank, int P, int width, const double* __restrict__ W, int nb,
                    double* __restrict__ partial) {
    extern __shared__ __align__(128) double bsm[];
    __shared__ __align__(8) uint64_t full[BS_STAGES];
    double* ring = bsm;                                        // BS_STAGES x BS_ROWS x BS_COLS
    double* sw = bsm + BS_STAGES * BS_ROWS * BS_COLS;          // rows_in_chunk x BS_MAXB weights (transposed)
    const int tid = threadIdx.x;
    const int e0 = blockIdx.x * BS_COLS;
    const int cols = min(BS_COLS, width - e0);
    const int per = (P + BS_CHUNKS - 1) / BS_CHUNKS;
    const int p0 = blockIdx.y * per, p1 = min(P, p0 + per), rows = max(0, p1 - p0);
    const int iters = (rows + BS_ROWS - 1) / BS_ROWS;
    for (int e = tid; e < rows * BS_MAXB; e += BS_COLS) {
        const int r = e / BS_MAXB, b = e - r * BS_MAXB;
        sw[e] = (b < nb) ? W[(size_t)b * P + p0 + r] : 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < BS_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // the first BS_ROWS threads issue one row copy each (parallel issue); thread 0 arms the barrier with the byte count
    auto issue = [&](int it) {
        const int s = it % BS_STAGES;
        const int r0 = it * BS_ROWS, nr = min(BS_ROWS, rows - r0);
        if (tid == 0) mbar_expect_tx(&full[s], (unsigned)(nr * cols * sizeof(double)));
        if (tid < nr)
            tma_bulk_g2s(ring + (s * BS_ROWS + tid) * BS_COLS, bank + (size_t)(p0 + r0 + tid) * width + e0,
                         (unsigned)(cols * sizeof(double)), &full[s]);
    };
    if (tid < BS_ROWS)
        for (int it = 0; it < BS_STAGES && it < iters; ++it) issue(it);
    double acc[BS_MAXB];
#pragma unroll
    for (int b = 0; b < BS_MAXB; ++b) acc[b] = 0.0;
    for (int it = 0; it < iters; ++it) {
        const int s = it % BS_STAGES;
        mbar_wait(&full[s], (unsigned)((it / BS_STAGES) & 1));
        const int r0 = it * BS_ROWS, nr = min(BS_ROWS, rows - r0);
        if (tid < cols) {
            for (int r = 0; r < nr; ++r) {
                const double v = ring[(s * BS_ROWS + r) * BS_COLS + tid];
                const double2* wr = reinterpret_cast<const double2*>(sw + (size_t)(r0 + r) * BS_MAXB);
#pragma unroll
                for (int b2 = 0; b2 < BS_MAXB / 2; ++b2) {
                    const double2 w2 = wr[b2];
                    acc[2 * b2] = fma(w2.x, v, acc[2 * b2]);
                    acc[2 * b2 + 1] = fma(w2.y, v, acc[2 * b2 + 1]);
                }
            }
        }
        __syncthreads();                                        // everyone is done with slot s
        if (tid < BS_ROWS && it + BS_STAGES < iters) issue(it + BS_STAGES);
    }
    if (tid < cols) {
#pragma unroll
        for (int b = 0; b < BS_MAXB; ++b)
            if (b < nb) partial[((size_t)blockIdx.y * BS_MAXB + b) * width + e0 + tid] = acc[b];
    }
}

// out[b][e] = sum over chunks in fixed order
__global__ void blend_reduce_kernel(const double* __restrict__ partial, int width, int nb, double* __restrict__ out,
                                    long long out_stride) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= width) return;
    for (int b = 0; b < nb; ++b) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < BS_CHUNKS; ++c) s += partial[((size_t)c * BS_MAXB + b) * width + e];
        out[(size_t)b * out_stride + e] = s;
    }
}

size_t blend_stream_workspace(int width) { return sizeof(double) * (size_t)BS_CHUNKS * BS_MAXB * (size_t)width; }

// out (count x width, row stride = width) = W (count x P) @ bank (P x width); count processed 8 at a time.
int blend_stream(const double* bank, int P, int width, const double* W, long long count, double* out, void* ws,
                 cudaStream_t st) {
    if ((width & 1) || ((uintptr_t)bank % 16) != 0) return -1;   // bulk copies need 16-byte aligned rows: caller falls back
    const int per = (P + BS_CHUNKS - 1) / BS_CHUNKS;
    const size_t smem = sizeof(double) * ((size_t)BS_STAGES * BS_ROWS * BS_COLS + (size_t)per * BS_MAXB);
    if (smem > 200 * 1024) return -1;
    cudaError_t e = cudaFuncSetAttribute(blend_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "blend_stream attr");
    const int slabs = (width + BS_COLS - 1) / BS_COLS;
    for (long long b0 = 0; b0 < count; b0 += BS_MAXB) {
        const int nb = (int)((count - b0) < BS_MAXB ? (count - b0) : BS_MAXB);
        blend_stream_kernel<<<dim3(slabs, BS_CHUNKS), BS_COLS, smem, st>>>(bank, P, width, W + b0 * P, nb, (double*)ws);
        blend_reduce_kernel<<<(width + 255) / 256, 256, 0, st>>>((const double*)ws, width, nb, out + b0 * width, width);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "blend_stream_kernel");
    return 0;
}

}  // namespace srcb
