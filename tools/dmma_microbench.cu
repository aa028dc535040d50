// dmma_microbench.cu -- latency and issue rate of mma.sync.m8n8k4.f64 (SASS DMMA) on one SM of a B200:
//   ./dmma_microbench  prints, for W warps on the SM (all in one CTA) and C independent accumulator chains per warp,
//   the cycles per DMMA per warp and the SM-wide DMMAs per cycle.  Used to size the register blocking of the
//   TPWL Riccati products (ilqr_bwd_big.cuh) and the SSM kernel's tile chains.
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void k(long long* out, double* sink, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c0[C], c1[C];
#pragma unroll
    for (int i = 0; i < C; ++i) { c0[i] = i; c1[i] = -i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int C>
void run(int warps, long long* dout, double* sink) {
    const int iters = 2000;
    k<C><<<1, warps * 32>>>(dout, sink, iters);
    k<C><<<1, warps * 32>>>(dout, sink, iters);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, dout, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double per = (double)cyc / (iters * C);
    printf("warps %2d chains %d : %.1f cycles per DMMA per warp, %.3f DMMA/cycle/SM (%.1f FMA/cycle/SM)\n", warps, C, per,
           warps / per, 256.0 * warps / per);
}

int main() {
    long long* dout; double* sink;
    cudaMalloc(&dout, 8); cudaMalloc(&sink, 8);
    for (int w : {1, 2, 4, 8, 16}) { run<1>(w, dout, sink); run<2>(w, dout, sink); run<4>(w, dout, sink); run<8>(w, dout, sink); }
    return 0;
}
