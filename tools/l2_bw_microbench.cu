// l2_bw_microbench.cu -- L2 -> SM read bandwidth of a B200 on an L2-resident buffer, the denominator of the TPWL
// nearest-neighbour rollout's roofline (its 44 MB bank of [A_i | B_i | d_i] entries is L2 resident and every
// trajectory-step gathers one 44 KB entry).  Two access patterns, both with 16-byte loads:
//   stream : every CTA sweeps the whole buffer (coalesced, all SMs pulling at once)
//   gather : every 256-thread half-CTA reads a random 44 352 B entry per iteration, 18 loads in flight per thread --
//            the access pattern of tpwl_rollout_nn_screen_kernel
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/l2_bw_microbench tools/l2_bw_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) stream_kernel(const double2* __restrict__ buf, size_t n16, int reps, double* sink) {
    double acc = 0.0;
    for (int rp = 0; rp < reps; ++rp) {
        // CTAs start at different offsets so that the L2 slices are hit evenly
        const size_t start = ((size_t)blockIdx.x * 7919u * 512u) % n16;
        for (size_t i = threadIdx.x; i < n16; i += 512 * 4) {
            size_t a = start + i, b = a + 512, c = a + 1024, d = a + 1536;
            a = a >= n16 ? a - n16 : a; b = b >= n16 ? b - n16 : b; c = c >= n16 ? c - n16 : c; d = d >= n16 ? d - n16 : d;
            const double2 v0 = buf[a], v1 = buf[b], v2 = buf[c], v3 = buf[d];
            acc += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
        }
    }
    if (acc == 1.2345) sink[0] = acc;
}

__global__ void __launch_bounds__(512, 1) gather_kernel(const double2* __restrict__ buf, int entries, int e16, int iters,
                                                        unsigned seed, double* sink) {
    const int half = threadIdx.x >> 8, ht = threadIdx.x & 255;
    unsigned s = seed ^ (blockIdx.x * 2654435761u + half * 40503u);
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        const double2* e = buf + (size_t)((s >> 8) % (unsigned)entries) * e16;
        for (int base = 0; base < e16; base += 256 * 9) {
            double2 v[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int i = base + k * 256 + ht;
                v[k] = i < e16 ? e[i] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) acc += v[k].x + v[k].y;
        }
    }
    if (acc == 1.2345) sink[0] = acc;
}

int main() {
    const int entries = 1000, e16 = 44352 / 16;
    const size_t n16 = (size_t)entries * e16;
    double2* buf; double* sink;
    cudaMalloc(&buf, n16 * 16); cudaMalloc(&sink, 8);
    cudaMemset(buf, 0, n16 * 16);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        const int reps = 8;
        cudaEventRecord(a);
        stream_kernel<<<sms, 512>>>(buf, n16, reps, sink);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        const double bytes = (double)n16 * 16 * reps * sms;
        if (rep) printf("stream : %d CTAs x 512 threads, %.1f MB buffer x %d sweeps each: %.3f ms, %.0f GB/s L2 -> SM (%.1f B/clk/SM at %d MHz)\n",
                        sms, n16 * 16 / 1e6, reps, ms, bytes / ms / 1e6, bytes / ms / 1e6 * 1e9 / sms / (clk * 1e3), clk / 1000);
    }
    for (int rep = 0; rep < 2; ++rep) {
        const int iters = 2000;
        cudaEventRecord(a);
        gather_kernel<<<sms, 512>>>(buf, entries, e16, iters, 12345u, sink);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        const double bytes = (double)e16 * 16 * iters * 2 * sms;
        if (rep) printf("gather : %d CTAs x 2 halves, random 44352 B entries, %d each: %.3f ms, %.0f GB/s L2 -> SM (%.1f B/clk/SM)\n",
                        sms, iters, ms, bytes / ms / 1e6, bytes / ms / 1e6 * 1e9 / sms / (clk * 1e3));
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
