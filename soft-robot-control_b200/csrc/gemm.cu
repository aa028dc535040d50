// gemm.cu -- FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) GEMM / SYRK for the POD Gram matrix, the POD
// projections and the TPWL weighted bank blend.
//
// tcgen05/UMMA has no FP64 kind, so on sm_100a the FP64 tensor path is the warp-level DMMA; operands are staged
// through shared memory with a 4-stage cp.async pipeline, padded so that every fragment load is conflict-free.  The
// stages are handed over through mbarriers instead of CTA barriers: a thread's copies of a stage arrive on its "full"
// barrier when they land (cp.async.mbarrier.arrive), a warp arrives on the "empty" barrier when it has read the stage,
// and a stage is refilled two K chunks after its last use -- a warp never waits for the slowest warp of the CTA on
// the chunk it is about to multiply (with a __syncthreads per chunk the tensor pipe idled 21 % of the time).
//
// CTA tile 128 x 128, K chunk 16, 8 warps arranged 2 (M) x 4 (N), warp tile 64 x 32 = 8 x 4 DMMA tiles
// (64 FP64 accumulators per thread).  Per k4 step a warp issues 12 LDS.64 for 32 DMMAs.
//
// Reference math sites: sofacontrol/mor/pod.py:181-200 (SVD of the snapshot matrix -> eig of X^T X),
// pod.py:22-72 (projections), sofacontrol/tpwl/tpwl.py:246-248 (einsum bank blend).
#include <cstdlib>
#include "common.cuh"

namespace srcb {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, PREFETCH = 2, GEMM_THREADS = 256;   // BK 32 / 3 stages / prefetch 1 measured: no gain (Gram 58.6 vs 61.1)
constexpr int LDT = BM + 4;      // [k][m] / [k][n] tiles: row stride 132 doubles (== 4 mod 16 -> conflict-free frags)
constexpr int LDA_NT = BK + 4;   // non-transposed A tile stored [m][k]: row stride 20 doubles
constexpr int A_TILE = (BK * LDT > BM * LDA_NT) ? BK * LDT : BM * LDA_NT;
constexpr int B_TILE = BK * LDT;
constexpr size_t GEMM_SMEM = sizeof(double) * STAGES * (A_TILE + B_TILE);
constexpr int kPaceSlots = 4, kPaceInts = 1 << 16;
__device__ int g_gemm_pace[kPaceSlots * kPaceInts];   // pace-keeper counters of up to kPaceSlots SYRK launches in flight

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
// this thread's earlier cp.async copies arrive on `bar` when they complete (counted in the barrier's expected arrivals)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Loads a [rows][cols] panel (row-major in global, leading dim ld) starting at (r0, c0) into smem with row stride
// lds, zero-filling outside (R, Cn).  cols must be even.  ALIGN16: ld even and base 16B aligned.
template <bool ALIGN16>
__device__ __forceinline__ void load_panel(double* __restrict__ s, int lds, const double* __restrict__ g, long long ld,
                                           long long r0, long long c0, long long R, long long Cn, int rows, int cols) {
    if (ALIGN16) {
        const int cpr = cols / 2;   // 16B chunks per row
        for (int c = threadIdx.x; c < rows * cpr; c += GEMM_THREADS) {
            const int rr = c / cpr, cc = (c - rr * cpr) * 2;
            const long long gr = r0 + rr, gc = c0 + cc;
            int bytes = 0;
            if (gr < R && gc < Cn) bytes = (gc + 1 < Cn) ? 16 : 8;
            const double* src = (bytes > 0) ? (g + gr * ld + gc) : g;
            cp_async16(s + rr * lds + cc, src, bytes);
        }
    } else {
        for (int c = threadIdx.x; c < rows * cols; c += GEMM_THREADS) {
            const int rr = c / cols, cc = c - rr * cols;
            const long long gr = r0 + rr, gc = c0 + cc;
            const bool ok = (gr < R && gc < Cn);
            cp_async8(s + rr * lds + cc, ok ? (g + gr * ld + gc) : g, ok ? 8 : 0);
        }
    }
}

// C = alpha * op(A) * B.   TRANSA: A is K x M (lda), else M x K.  B is K x N (ldb).  C is M x N (ldc).
// SYRK: A == B == X (K x M), only tiles with tile_n >= tile_m are computed and each is written to both triangles;
//       ACCUM adds into C instead of overwriting.
template <bool TRANSA, bool SYRK, bool ALIGN16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dgemm_kernel(long long M, long long N, long long K, double alpha, const double* __restrict__ A, long long lda,
             const double* __restrict__ B, long long ldb, double* __restrict__ C, long long ldc, int accumulate,
             int tiles_m, int tiles_n, long long num_tiles, int sb, int* __restrict__ pace, int pace_w, int pace_windows, int pace_slack) {
    extern __shared__ __align__(16) double gsm[];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    double* sA = gsm;
    double* sB = gsm + STAGES * A_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;        // 2 x 4 warps
    const int g = lane >> 2, q = lane & 3;          // groupID, threadID_in_group
    const long long KT = (K + BK - 1) / BK;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], GEMM_THREADS); mbar_init(&empty[s], GEMM_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    long long gk = 0;                               // K chunks this CTA has walked so far (all tiles): stage = gk % STAGES

    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int tm, tn;
        if (SYRK) {
            // Upper triangle (tn >= tm), enumerated SUPER-BLOCK by super-block: sb x sb tiles with sb ~ sqrt(grid), so
            // the ~148 tiles in flight share 2 sb column panels of X instead of (1 + 148): every CTA walks K at the
            // same pace, a panel chunk is fetched from DRAM once per super-block and served to its sb users by L2.
            // (Row-by-row enumeration re-read all of X once per wave: 16 x the algorithmic bytes at ns = 8192.)
            const int nsb = (tiles_m + sb - 1) / sb;
            long long t = tile;
            tm = tn = 0;
            bool found = false;
            for (int I = 0; I < nsb && !found; ++I) {
                const int r0 = I * sb, h = min(sb, tiles_m - r0);
                for (int J = I; J < nsb; ++J) {
                    const int c0 = J * sb, w = min(sb, tiles_m - c0);
                    const long long cnt = (J == I) ? (long long)h * (h + 1) / 2 : (long long)h * w;
                    if (t < cnt) {
                        if (J == I) {
                            int rr = 0, rowlen = h;
                            while (t >= rowlen) { t -= rowlen; ++rr; --rowlen; }
                            tm = r0 + rr; tn = r0 + rr + (int)t;
                        } else {
                            tm = r0 + (int)(t / w); tn = c0 + (int)(t % w);
                        }
                        found = true;
                        break;
                    }
                    t -= cnt;
                }
            }
        } else {
            // same idea for the rectangular tile grid: sb x sb super-blocks, row-major inside
            const int nsbn = (tiles_n + sb - 1) / sb;
            const long long per_row = (long long)sb * tiles_n;               // tiles of one full super-block row
            const int I = (int)(tile / per_row);
            const int r0 = I * sb, h = min(sb, tiles_m - r0);
            long long t = tile - (long long)I * per_row;                     // index inside super-block row I (h rows)
            const long long per_blk = (long long)h * sb;
            int J = (int)(t / per_blk);
            if (J >= nsbn) J = nsbn - 1;
            t -= (long long)J * per_blk;
            const int c0 = J * sb, w = min(sb, tiles_n - c0);
            tm = r0 + (int)(t / w);
            tn = c0 + (int)(t % w);
        }
        const long long m0 = (long long)tm * BM, n0 = (long long)tn * BN;

        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // K chunk kt of this tile is the CTA's chunk number gk + kt: stage (gk + kt) % STAGES, in its use number
        // (gk + kt) / STAGES.  Before the refill, the stage's previous use must have been read by all eight warps.
        auto issue = [&](long long kt) {
            const long long j = gk + kt;
            const int st = (int)(j % STAGES);
            const long long use = j / STAGES;
            if (pace && tid == 0 && ((int)kt & (pace_w - 1)) == 0) {       // pace_w is a power of two
                // Pace keeper (long-K SYRK): the CTAs that work on the same round of tiles share column panels of X
                // through L2 only while they walk K together.  Every pace_w chunks a CTA reports its window and waits
                // (bounded: this is a hint, never a dependency) until all CTAs of the round have entered the previous
                // one: the pack stays within two windows, far inside what L2 holds.
                const long long round = tile / gridDim.x, left = num_tiles - round * gridDim.x;
                const int expect = (int)(left < gridDim.x ? left : gridDim.x), w = (int)kt >> (31 - __clz(pace_w));
                int* row = pace + round * pace_windows;
                atomicAdd(row + w, 1);
                if (w >= pace_slack) {
                    const long long t0 = clock64();
                    while (*(volatile int*)(row + w - pace_slack) < expect && clock64() - t0 < 60000) { }
                }
            }
            if (use > 0) mbar_wait(&empty[st], (unsigned)((use - 1) & 1));
            double* a = sA + st * A_TILE;
            double* b = sB + st * B_TILE;
            const long long k0 = kt * BK;
            if (TRANSA) load_panel<ALIGN16>(a, LDT, A, lda, k0, m0, K, M, BK, BM);
            else        load_panel<ALIGN16>(a, LDA_NT, A, lda, m0, k0, M, K, BM, BK);
            load_panel<ALIGN16>(b, LDT, B, ldb, k0, n0, K, N, BK, BN);
            cp_async_arrive(&full[st]);
        };

#pragma unroll
        for (int s = 0; s < PREFETCH; ++s)
            if (s < KT) issue(s);
        for (long long kt = 0; kt < KT; ++kt) {
            if (kt + PREFETCH < KT) issue(kt + PREFETCH);
            const long long j = gk + kt;
            const int st = (int)(j % STAGES);
            mbar_wait(&full[st], (unsigned)((j / STAGES) & 1));
            const double* a = sA + st * A_TILE;
            const double* b = sB + st * B_TILE;
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int mrow = wm * 64 + i * 8 + g;
                    af[i] = TRANSA ? a[(kk + q) * LDT + mrow] : a[mrow * LDA_NT + kk + q];
                }
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) bf[j2] = b[(kk + q) * LDT + wn * 32 + j2 * 8 + g];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j2 = 0; j2 < 4; ++j2) dmma_m8n8k4(acc[i][j2][0], acc[i][j2][1], af[i], bf[j2]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        gk += KT;

        // epilogue
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long row = m0 + wm * 64 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const long long col = n0 + wn * 32 + j * 8 + q * 2 + e;
                    if (row < M && col < N) {
                        const double v = alpha * acc[i][j][e];
                        if (SYRK) {
                            if (tm == tn) {
                                // diagonal tile: both (row,col) and (col,row) are produced inside this tile
                                C[row * ldc + col] = accumulate ? C[row * ldc + col] + v : v;
                            } else {
                                C[row * ldc + col] = accumulate ? C[row * ldc + col] + v : v;
                                C[col * ldc + row] = accumulate ? C[col * ldc + row] + v : v;
                            }
                        } else {
                            C[row * ldc + col] = accumulate ? C[row * ldc + col] + v : v;
                        }
                    }
                }
            }
        }
    }
}

template <bool TRANSA, bool SYRK>
static int launch_gemm(long long M, long long N, long long K, double alpha, const double* A, long long lda,
                       const double* B, long long ldb, double* C, long long ldc, int accumulate, cudaStream_t st) {
    const int tiles_m = (int)((M + BM - 1) / BM), tiles_n = (int)((N + BN - 1) / BN);
    const long long num_tiles = SYRK ? (long long)tiles_m * (tiles_m + 1) / 2 : (long long)tiles_m * tiles_n;
    if (num_tiles == 0) return 0;
    const bool al = ((lda % 2) == 0) && ((ldb % 2) == 0) && (((uintptr_t)A % 16) == 0) && (((uintptr_t)B % 16) == 0);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(num_tiles < sms ? num_tiles : sms);   // persistent: one CTA per SM walks the tile list
    int sb = 1;
    while ((sb + 1) * (sb + 1) <= grid) ++sb;                    // super-block edge of the SYRK walk
    // pace keeper of the long-K SYRK (see the kernel): one counter per (round of tiles, window of K chunks).  OFF by
    // default: measured at 131072 x 8192 (profiles/ncu_gram_r2_mbar_pace.txt) free-running CTAs read 83-168 GB from
    // DRAM (8.6 GB algorithmic) in 290 ms, paced ones (SRCB200_GEMM_PACE=1: slack of one window) 60.9 GB in 309 ms --
    // the kernel is bound by the tensor pipe (88 % active), DRAM runs at 5-9 % of its peak, so the re-reads are free
    // and waiting for the slowest CTA is not.
    int* pace = nullptr;
    int pace_w = 32, pace_windows = 0, pace_slack = 0;
    const long long KT = (K + BK - 1) / BK;
    if (const char* env = getenv("SRCB200_GEMM_PACE")) pace_slack = atoi(env);     // 0: free-running CTAs
    if (SYRK && KT >= 512 && grid > 1 && pace_slack > 0) {
        static int* base = nullptr;
        static unsigned launches = 0;
        if (!base) SRCB_CUDA(cudaGetSymbolAddress((void**)&base, g_gemm_pace));
        const long long rounds = (num_tiles + grid - 1) / grid;
        while (rounds * ((KT + pace_w - 1) / pace_w) > kPaceInts) pace_w *= 2;
        pace_windows = (int)((KT + pace_w - 1) / pace_w);
        pace = base + (size_t)(launches++ % kPaceSlots) * kPaceInts;
        SRCB_CUDA(cudaMemsetAsync(pace, 0, sizeof(int) * (size_t)rounds * pace_windows, st));
    }
    if (al) {
        auto kern = dgemm_kernel<TRANSA, SYRK, true>;
        SRCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        kern<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(M, N, K, alpha, A, lda, B, ldb, C, ldc, accumulate, tiles_m,
                                                    tiles_n, num_tiles, sb, pace, pace_w, pace_windows, pace_slack);
    } else {
        auto kern = dgemm_kernel<TRANSA, SYRK, false>;
        SRCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        kern<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(M, N, K, alpha, A, lda, B, ldb, C, ldc, accumulate, tiles_m,
                                                    tiles_n, num_tiles, sb, pace, pace_w, pace_windows, pace_slack);
    }
    SRCB_LAUNCH_CHECK("dgemm_kernel");
    return 0;
}

int dgemm_device(int transA, long long M, long long N, long long K, double alpha, const double* A, long long lda,
                 const double* B, long long ldb, double* C, long long ldc, cudaStream_t st) {
    if (transA) return launch_gemm<true, false>(M, N, K, alpha, A, lda, B, ldb, C, ldc, 0, st);
    return launch_gemm<false, false>(M, N, K, alpha, A, lda, B, ldb, C, ldc, 0, st);
}

}  // namespace srcb

using namespace srcb;

extern "C" int srcb200_pod_gram(int64_t nf, int64_t ns, const double* X, int64_t ldx, double* G, int64_t ldg,
                                int32_t accumulate, void* stream) {
    if (nf < 0 || ns < 0 || ldx < ns || ldg < ns) return fail(SRCB200_E_DIM, "pod_gram: bad dims nf=%lld ns=%lld ldx=%lld ldg=%lld",
                                                              (long long)nf, (long long)ns, (long long)ldx, (long long)ldg);
    if (ns == 0) return 0;
    if (!X || !G) return fail(SRCB200_E_NULL, "pod_gram: X/G is NULL");
    return launch_gemm<true, true>(ns, ns, nf, 1.0, X, ldx, X, ldx, G, ldg, accumulate, (cudaStream_t)stream);
}

extern "C" int srcb200_dgemm(int32_t transA, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                             int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, void* stream) {
    if (M < 0 || N < 0 || K < 0) return fail(SRCB200_E_DIM, "dgemm: negative dims");
    if (M == 0 || N == 0) return 0;
    if (!C) return fail(SRCB200_E_NULL, "dgemm: NULL operand");
    if (K == 0) {   // empty contraction: C = 0
        SRCB_CUDA(cudaMemset2DAsync(C, sizeof(double) * ldc, 0, sizeof(double) * N, M, (cudaStream_t)stream));
        return 0;
    }
    if (!A || !B) return fail(SRCB200_E_NULL, "dgemm: NULL operand");
    if (ldb < N || ldc < N || lda < (transA ? M : K)) return fail(SRCB200_E_DIM, "dgemm: leading dimension too small");
    return dgemm_device(transA, M, N, K, alpha, A, lda, B, ldb, C, ldc, (cudaStream_t)stream);
}
