// lat_microbench.cu -- dependent-issue latencies (cycles) of the instruction kinds the Trunk-SSM iLQR step is made of,
// one warp on one SM of a B200: DFMA / DADD / DMUL chains, __drcp_rn, 64-bit shuffle, LDS pointer chase, STS ->
// __syncwarp -> LDS round trip, REDUX, ballot, DMMA.  Used to budget the per-step dependency chain (profiles/).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define TIME(name, body)                                                     \
    {                                                                        \
        __syncwarp();                                                        \
        const long long t0 = clock64();                                      \
        _Pragma("unroll 1") for (int it = 0; it < ITERS / 8; ++it) {         \
            body body body body body body body body                          \
        }                                                                    \
        const long long t1 = clock64();                                      \
        if (threadIdx.x == 0) out[idx] = (double)(t1 - t0) / ITERS;          \
        if (threadIdx.x == 0 && blockIdx.x == 0 && names) names[idx] = name; \
        ++idx;                                                               \
    }

__global__ void k(double* out, const char** names, double seed, double* sink) {
    __shared__ double sh[64];
    __shared__ int chase[32];
    const int lane = threadIdx.x;
    sh[lane] = 1.0 + lane; sh[32 + lane] = 0.5;
    chase[lane] = (lane + 1) & 31;
    __syncwarp();
    int idx = 0;
    double x = seed, y = 1.0 + 1e-9 * lane, z = 0.999999;
    unsigned u = lane + 1;
    int p = lane;
    TIME("DFMA chain", x = fma(x, y, z);)
    TIME("DADD chain", x = __dadd_rn(x, y);)
    TIME("DMUL chain", x = __dmul_rn(x, z);)
    TIME("__drcp_rn chain", x = __drcp_rn(x) + y;)
    TIME("1.0/x (div) chain", x = 1.0 / x + y;)
    TIME("shfl 64-bit chain", x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);)
    TIME("shfl_xor 32-bit chain", u = __shfl_xor_sync(0xffffffffu, u, 1);)
    TIME("LDS chase", p = chase[p];)
    TIME("STS+syncwarp+LDS", sh[lane] = x; __syncwarp(); x = sh[(lane + 1) & 31]; __syncwarp();)
    TIME("STS+syncwarp+LDS.128 bcast", sh[lane] = x; __syncwarp(); { double2 v = *reinterpret_cast<double2*>(sh + 2 * (it & 7)); x = v.x + v.y; } __syncwarp();)
    TIME("redux.max", u = __reduce_max_sync(0xffffffffu, u) + lane;)
    TIME("ballot+ffs", u = __ffs(__ballot_sync(0xffffffffu, u & 1)) + lane;)
    TIME("syncwarp", __syncwarp();)
    {
        double c0 = 0, c1 = 0;
        TIME("DMMA chain", asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(y), "d"(z));)
        TIME("DMMA -> A operand chain", asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(c0), "d"(z));)
        x += c0 + c1;
    }
    TIME("DFMA 2 chains", x = fma(x, y, z); y = fma(y, z, z);)
    TIME("DFMA 4 chains", x = fma(x, y, z); y = fma(y, z, z); z = fma(z, z, 0.1); seed = fma(seed, seed, 0.2);)
    if (x + y + z + u + p + seed == 12345.6789) sink[0] = x;
}

int main() {
    double* out; const char** names; double* sink;
    cudaMallocManaged(&out, 64 * sizeof(double)); cudaMallocManaged(&names, 64 * sizeof(char*)); cudaMalloc(&sink, 8);
    for (int i = 0; i < 64; ++i) { out[i] = -1; names[i] = nullptr; }
    k<<<1, 32>>>(out, nullptr, 1.000001, sink);
    cudaDeviceSynchronize();
    k<<<1, 32>>>(out, nullptr, 1.000001, sink);
    cudaDeviceSynchronize();
    const char* nm[] = {"DFMA chain", "DADD chain", "DMUL chain", "__drcp_rn + DADD chain", "1.0/x + DADD chain", "shfl 64-bit chain",
                        "shfl_xor 32-bit chain", "LDS pointer chase", "STS + syncwarp + LDS + syncwarp", "STS + syncwarp + LDS.128 + DADD + syncwarp",
                        "redux.max + IADD", "ballot + ffs + IADD", "syncwarp alone", "DMMA accumulator chain", "DMMA result -> A operand chain",
                        "DFMA, 2 independent chains (per pair)", "DFMA, 4 independent chains (per quad)"};
    for (int i = 0; i < 17; ++i) printf("%-48s %.1f cycles\n", nm[i], out[i]);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
