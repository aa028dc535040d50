// tpwl.cu -- TPWL model: nearest-point selection, exponential weights, linearisation (gather / weighted bank
// blend + discretisation) and batched rollout.  Reference: sofacontrol/tpwl/tpwl.py (see include/srcb200.h).
#include "tpwl.cuh"

extern "C" size_t srcb200_zoh_workspace(int32_t n, int32_t m, int64_t count);
extern "C" int srcb200_zoh_batch(int32_t n, int32_t m, int64_t count, double dt, const double* A_c, const double* B_c,
                                 const double* d_c, double* A_d, double* B_d, double* d_d, void* workspace,
                                 size_t workspace_bytes, void* stream);

namespace srcb {

size_t blend_stream_workspace(int width);
int blend_stream(const double* bank, int P, int width, const double* W, long long count, double* out, void* ws,
                 cudaStream_t st);   // blend_stream.cu: TMA-staged bank stream for small batches (-1: not applicable)

constexpr long long kBlendStreamMaxBatch = 32;   // above this the DMMA GEMM (bank tile reused across the batch) wins

// W (count x P) @ bank (P x width): bank stream for small batches, DMMA GEMM otherwise
static int blend(const double* bank, int P, long long width, const double* W, long long count, double* out, void* stream_ws,
                 cudaStream_t st);

int dgemm_device(int transA, long long M, long long N, long long K, double alpha, const double* A, long long lda,
                 const double* B, long long ldb, double* C, long long ldc, cudaStream_t st);

static int blend(const double* bank, int P, long long width, const double* W, long long count, double* out, void* stream_ws,
                 cudaStream_t st) {
    if (stream_ws && count <= kBlendStreamMaxBatch) {
        const int rc = blend_stream(bank, P, (int)width, W, count, out, stream_ws, st);
        if (rc >= 0) return rc;
    }
    return dgemm_device(0, count, width, P, 1.0, W, P, bank, width, out, width, st);
}

int check_tpwl_model(const srcb200_tpwl_model* s) {
    if (!s) return fail(SRCB200_E_NULL, "tpwl model is NULL");
    if (s->n < 2 || (s->n & 1) || s->m < 1 || s->P < 1 || s->nz < 0)
        return fail(SRCB200_E_DIM, "tpwl dims invalid: n=%d m=%d nz=%d P=%d", s->n, s->m, s->nz, s->P);
    if (s->n > 128 || s->m > 32 || s->nz > 32)
        return fail(SRCB200_E_DIM, "tpwl dims beyond kernel limits (n<=128, m<=32, nz<=32): n=%d m=%d nz=%d", s->n, s->m, s->nz);
    if (!s->qT || !s->vT || !s->A || !s->B || !s->d) return fail(SRCB200_E_NULL, "tpwl model has NULL bank pointers");
    if (s->nz > 0 && (!s->H || !s->z_ref)) return fail(SRCB200_E_NULL, "tpwl output model has NULL H/z_ref");
    if (s->method != SRCB200_TPWL_NN && s->method != SRCB200_TPWL_WEIGHTING)
        return fail(SRCB200_E_METHOD, "tpwl method should be nn or weighting");               // tpwl.py:268
    if (s->discr_method < SRCB200_DISCR_FE || s->discr_method > SRCB200_DISCR_NONE)
        return fail(SRCB200_E_METHOD, "self.discr_method must be in [fe, be, bil, zoh]");     // tpwl.py:295
    return 0;
}

constexpr int kSel = 128;   // threads per CTA for selection / rollout kernels

// ---- nearest point ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSel)
tpwl_nearest_kernel(TpwlDev M, long long count, const double* __restrict__ x, int* __restrict__ idx,
                    double* __restrict__ dist) {
    extern __shared__ double sm[];
    __shared__ double red_d[kSel / 32];
    __shared__ int red_i[kSel / 32];
    double* sx = sm;
    for (long long s = blockIdx.x; s < count; s += gridDim.x) {
        for (int i = threadIdx.x; i < M.n; i += kSel) sx[i] = x[s * M.n + i];
        __syncthreads();
        double dmin;
        const int bi = tpwl_nearest<kSel, true>(M, sx, nullptr, red_d, red_i, &dmin);
        if (threadIdx.x == 0) {
            idx[s] = bi;
            if (dist) dist[s] = dmin;
        }
        __syncthreads();
    }
}

// ---- exponential weights (tpwl.py:170-191) --------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ double cta_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) s += red[k];
    __syncthreads();
    return s;
}

// weights of one state sx (shared memory) into ws (global, P doubles); sd: P doubles of shared scratch
template <int NT>
__device__ __forceinline__ void tpwl_weights_one(const TpwlDev& M, const double* __restrict__ sx, double* __restrict__ sd,
                                                 double* red_d, int* red_i, double* __restrict__ ws) {
    double dmin;
    const int bi = tpwl_nearest<NT, true>(M, sx, sd, red_d, red_i, &dmin);
    __syncthreads();
    if (dmin == 0.0) {
        for (int p = threadIdx.x; p < M.P; p += NT) ws[p] = (p == bi) ? 1.0 : 0.0;
    } else {
        double part = 0.0;
        for (int p = threadIdx.x; p < M.P; p += NT) {
            // np.exp(-beta * dist / m): ((-beta) * dist) / m
            const double e = exp(__ddiv_rn(__dmul_rn(-M.beta, sd[p]), dmin));
            sd[p] = e;
            part += e;
        }
        const double tot = cta_sum<NT>(part, red_d);
        for (int p = threadIdx.x; p < M.P; p += NT) ws[p] = __ddiv_rn(sd[p], tot);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kSel)
tpwl_weights_kernel(TpwlDev M, long long count, const double* __restrict__ x, long long xstride,
                    double* __restrict__ w) {
    extern __shared__ double sm[];
    __shared__ double red_d[kSel / 32];
    __shared__ int red_i[kSel / 32];
    double* sx = sm;
    double* sd = sm + M.n;   // P distances
    for (long long s = blockIdx.x; s < count; s += gridDim.x) {
        for (int i = threadIdx.x; i < M.n; i += kSel) sx[i] = x[s * xstride + i];
        __syncthreads();
        tpwl_weights_one<kSel>(M, sx, sd, red_d, red_i, w + s * (long long)M.P);
    }
}

// ---- weighting-mode rollout, fused per time step ---------------------------------------------------------------
// The three banks are concatenated once per call into one (P x wd) matrix [A_p | B_p | d_p], so the blend of a time step
// is ONE DMMA GEMM  W (batch x P) * bank (P x wd)  (tpwl.py:246-248 are three einsums over the same weights); this
// kernel then takes each trajectory's blended row, discretises it in shared memory (tpwl.py:272-297), steps the state
// (tpwl.py:336-339) and computes the weights of the NEW state for the next time step's GEMM: two launches per step.
constexpr int kWS = 256;
__global__ void tpwl_concat_bank_kernel(TpwlDev M, long long wd, double* __restrict__ cat) {
    const long long nn = (long long)M.n * M.n, nm = (long long)M.n * M.m;
    for (long long p = blockIdx.x; p < M.P; p += gridDim.x) {
        double* row = cat + p * wd;
        for (long long e = threadIdx.x; e < nn; e += blockDim.x) row[e] = M.A[p * nn + e];
        for (long long e = threadIdx.x; e < nm; e += blockDim.x) row[nn + e] = M.B[p * nm + e];
        for (long long e = threadIdx.x; e < M.n; e += blockDim.x) row[nn + nm + e] = M.d[p * M.n + e];
        for (long long e = nn + nm + M.n + threadIdx.x; e < wd; e += blockDim.x) row[e] = 0.0;
    }
}

__global__ void __launch_bounds__(kWS)
tpwl_weighting_step_kernel(TpwlDev M, long long batch, int disc, double dt, const double* __restrict__ blended, long long wd,
                           const double* __restrict__ x, long long xstride, const double* __restrict__ u, long long ustride,
                           double* __restrict__ xn, long long xnstride, double* __restrict__ wnext) {
    extern __shared__ __align__(128) double sm[];
    __shared__ double red_d[kWS / 32];
    __shared__ int red_i[kWS / 32];
    __shared__ __align__(8) uint64_t full[2];
    const int n = M.n, m = M.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rowlen = n * n + n * m + n;                       // doubles of one blended row [A | B | d]
    const unsigned rowbytes = (unsigned)(sizeof(double) * (size_t)((rowlen + 1) & ~1));   // == wd doubles: 16-byte multiple
    double* buf0 = sm;
    double* buf1 = sm + wd;
    double* sx = sm + 2 * wd;
    double* su = sx + n;
    double* sxn = su + ((m + 1) & ~1);
    double* scr = sxn + n;                       // discretisation scratch, then the P distances of the weights
    // Each trajectory's 44 KB row is pulled by ONE TMA bulk copy (cp.async.bulk + mbarrier) into a two-slot ring: the
    // row of the CTA's next trajectory streams in while the current one is discretised, stepped and weighted.
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    long long b = blockIdx.x;
    if (tid == 0 && b < batch) {
        mbar_expect_tx(&full[0], rowbytes);
        tma_bulk_g2s(buf0, blended + b * wd, rowbytes, &full[0]);
    }
    int it = 0;
    for (; b < batch; b += gridDim.x, ++it) {
        const int slot = it & 1;
        double* sA = slot ? buf1 : buf0;
        double* sB = sA + n * n;
        double* sd = sB + n * m;
        const long long bnext = b + gridDim.x;
        if (tid == 0 && bnext < batch) {          // slot ^ 1 was released by the __syncthreads that ended the previous trajectory
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // its generic-proxy writes (in-place discretisation) come first
            mbar_expect_tx(&full[slot ^ 1], rowbytes);
            tma_bulk_g2s(slot ? buf0 : buf1, blended + bnext * wd, rowbytes, &full[slot ^ 1]);
        }
        for (int e = tid; e < n; e += kWS) sx[e] = x[b * xstride + e];
        for (int e = tid; e < m; e += kWS) su[e] = u[b * ustride + e];
        mbar_wait(&full[slot], (unsigned)((it >> 1) & 1));
        __syncthreads();
        if (disc) {
            discretize_inplace<kWS>(M.discr, dt, sA, sB, sd, n, m, scr);
            __syncthreads();
        }
        for (int i = warp; i < n; i += kWS / 32) {
            double ax = 0.0, bu = 0.0;
            for (int k = lane; k < n; k += 32) ax = fma(sA[i * n + k], sx[k], ax);
            for (int k = lane; k < m; k += 32) bu = fma(sB[i * m + k], su[k], bu);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, off);
                bu += __shfl_xor_sync(0xffffffffu, bu, off);
            }
            if (lane == 0) {
                const double v = __dadd_rn(__dadd_rn(ax, bu), sd[i]);
                sxn[i] = v;
                xn[b * xnstride + i] = v;
            }
        }
        __syncthreads();
        if (wnext) tpwl_weights_one<kWS>(M, sxn, scr, red_d, red_i, wnext + b * (long long)M.P);
        __syncthreads();
    }
}

// ---- gather the selected bank entries (nn) ---------------------------------------------------------------------
__global__ void tpwl_gather_kernel(TpwlDev M, long long count, const int* __restrict__ idx, double* __restrict__ A,
                                   double* __restrict__ B, double* __restrict__ d) {
    const long long nn = (long long)M.n * M.n, nm = (long long)M.n * M.m, n = M.n;
    for (long long s = blockIdx.x; s < count; s += gridDim.x) {
        const long long p = idx[s];
        if (A) for (int e = threadIdx.x; e < nn; e += blockDim.x) A[s * nn + e] = M.A[p * nn + e];
        if (B) for (int e = threadIdx.x; e < nm; e += blockDim.x) B[s * nm + e] = M.B[p * nm + e];
        if (d) for (int e = threadIdx.x; e < n; e += blockDim.x) d[s * n + e] = M.d[p * n + e];
    }
}

// ---- batched discretisation (fe / be / bil) ---------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
discretize_kernel(int n, int m, int method, long long count, double dt, const double* Ac,
                  const double* Bc, const double* dc, double* Ad,
                  double* Bd, double* dd) {
    extern __shared__ double sm[];
    double* sA = sm;
    double* sB = sA + n * n;
    double* sd = sB + n * m;
    double* scr = sd + n;
    const int tid = threadIdx.x;
    for (long long s = blockIdx.x; s < count; s += gridDim.x) {
        for (int e = tid; e < n * n; e += NT) sA[e] = Ac[s * n * n + e];
        for (int e = tid; e < n * m; e += NT) sB[e] = Bc[s * n * m + e];
        for (int e = tid; e < n; e += NT) sd[e] = dc[s * n + e];
        cta_sync<NT>();
        discretize_inplace<NT>(method, dt, sA, sB, sd, n, m, scr);
        cta_sync<NT>();
        for (int e = tid; e < n * n; e += NT) Ad[s * n * n + e] = sA[e];
        for (int e = tid; e < n * m; e += NT) Bd[s * n * m + e] = sB[e];
        for (int e = tid; e < n; e += NT) dd[s * n + e] = sd[e];
        cta_sync<NT>();
    }
}

int discretize_launch(int n, int m, int method, long long count, double dt, const double* Ac, const double* Bc,
                      const double* dc, double* Ad, double* Bd, double* dd, cudaStream_t st) {
    if (count == 0) return 0;
    const size_t smem = sizeof(double) * (n * n + n * m + n + discretize_scratch_doubles(n, m));
    if (smem > 227 * 1024) return fail(SRCB200_E_DIM, "discretize: n=%d m=%d needs %zu B of shared memory", n, m, smem);
    const int grid = (int)(count < 148 * 16 ? count : 148 * 16);
    if (n <= 12) {
        SRCB_CUDA(cudaFuncSetAttribute(discretize_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        discretize_kernel<32><<<grid, 32, smem, st>>>(n, m, method, count, dt, Ac, Bc, dc, Ad, Bd, dd);
    } else {
        SRCB_CUDA(cudaFuncSetAttribute(discretize_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        discretize_kernel<256><<<grid, 256, smem, st>>>(n, m, method, count, dt, Ac, Bc, dc, Ad, Bd, dd);
    }
    SRCB_LAUNCH_CHECK("discretize_kernel");
    return 0;
}

// fe / be / bil through discretize_kernel, zoh through the expm kernel (expm.cu) with caller workspace
static int discretize_any(int n, int m, int method, long long count, double dt, double* A, double* B, double* d,
                          void* zoh_ws, size_t zoh_ws_bytes, cudaStream_t st) {
    if (method == SRCB200_DISCR_ZOH)
        return srcb200_zoh_batch(n, m, count, dt, A, B, d, A, B, d, zoh_ws, zoh_ws_bytes, (void*)st);
    return discretize_launch(n, m, method, count, dt, A, B, d, A, B, d, st);
}

// ---- affine step x+ = (A x + B u) + d for a batch with per-trajectory matrices (weighting mode) -----------------
__global__ void __launch_bounds__(128)
tpwl_step_kernel(int n, int m, long long batch, const double* __restrict__ A, const double* __restrict__ B,
                 const double* __restrict__ d, const double* __restrict__ x, long long xstride,
                 const double* __restrict__ u, long long ustride, double* __restrict__ xn, long long xnstride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + b * (long long)n * n;
        const double* Bb = B + b * (long long)n * m;
        const double* xb = x + b * xstride;
        const double* ub = u + b * ustride;
        for (int i = warp; i < n; i += nw) {
            double ax = 0.0, bu = 0.0;
            for (int k = lane; k < n; k += 32) ax = fma(Ab[i * n + k], xb[k], ax);
            for (int k = lane; k < m; k += 32) bu = fma(Bb[i * m + k], ub[k], bu);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, off);
                bu += __shfl_xor_sync(0xffffffffu, bu, off);
            }
            if (lane == 0) xn[b * xnstride + i] = __dadd_rn(__dadd_rn(ax, bu), d[b * n + i]);
        }
    }
}

// z = H x + z_ref for count points (tpwl.py:115-126)
__global__ void tpwl_output_kernel(TpwlDev M, long long count, const double* __restrict__ x, double* __restrict__ z) {
    const long long tot = count * M.nz;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        const long long s = e / M.nz;
        const int i = (int)(e - s * M.nz);
        double acc = 0.0;
        for (int k = 0; k < M.n; ++k) acc = fma(M.H[i * M.n + k], x[s * M.n + k], acc);
        z[e] = __dadd_rn(acc, M.zref[i]);
    }
}

// ---- nn rollout: one CTA walks one trajectory (tpwl.py:193-216 with update_state 226-234) -----------------------
// PREDISC: the bank is already discrete (or dt < 0): the selected (A, B, d) are used straight from global memory.
// Otherwise the selected entry is copied to shared memory and discretised (fe/be/bil) every step, like the
// reference does when pre_discretize() was not called.
template <bool PREDISC>
__global__ void __launch_bounds__(kSel)
tpwl_rollout_nn_kernel(TpwlDev M, long long batch, int N, const double* __restrict__ x0,
                       const double* __restrict__ u, double dt, double* __restrict__ xo, int* __restrict__ idxo) {
    extern __shared__ double sm[];
    __shared__ double red_d[kSel / 32];
    __shared__ int red_i[kSel / 32];
    const int n = M.n, m = M.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* sx = sm;
    double* su = sx + n;
    double* sxn = su + m;
    double* sA = sxn + n;                 // only when !PREDISC
    double* sB = sA + n * n;
    double* sd = sB + n * m;
    double* scr = sd + n;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        double* xb = xo + b * (long long)(N + 1) * n;
        const double* ub = u + b * (long long)N * m;
        for (int i = tid; i < n; i += kSel) { sx[i] = x0[b * n + i]; xb[i] = sx[i]; }
        __syncthreads();
        for (int t = 0; t < N; ++t) {
            for (int i = tid; i < m; i += kSel) su[i] = ub[t * m + i];
            const int p = tpwl_nearest<kSel, true>(M, sx, nullptr, red_d, red_i, nullptr);
            if (idxo && tid == 0) idxo[b * (long long)N + t] = p;
            const double* Ap = M.A + (long long)p * n * n;
            const double* Bp = M.B + (long long)p * n * m;
            const double* dp = M.d + (long long)p * n;
            if (!PREDISC) {
                for (int e = tid; e < n * n; e += kSel) sA[e] = Ap[e];
                for (int e = tid; e < n * m; e += kSel) sB[e] = Bp[e];
                for (int e = tid; e < n; e += kSel) sd[e] = dp[e];
                __syncthreads();
                discretize_inplace<kSel>(M.discr, dt, sA, sB, sd, n, m, scr);
                Ap = sA; Bp = sB; dp = sd;
            }
            __syncthreads();
            for (int i = warp; i < n; i += kSel / 32) {
                double ax = 0.0, bu = 0.0;
                for (int k = lane; k < n; k += 32) ax = fma(Ap[i * n + k], sx[k], ax);
                for (int k = lane; k < m; k += 32) bu = fma(Bp[i * m + k], su[k], bu);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    ax += __shfl_xor_sync(0xffffffffu, ax, off);
                    bu += __shfl_xor_sync(0xffffffffu, bu, off);
                }
                if (lane == 0) sxn[i] = __dadd_rn(__dadd_rn(ax, bu), dp[i]);
            }
            __syncthreads();
            for (int i = tid; i < n; i += kSel) { sx[i] = sxn[i]; xb[(t + 1) * n + i] = sxn[i]; }
            __syncthreads();
        }
    }
}

// ---- nn rollout on a pre-discretised bank, T trajectories per CTA ----------------------------------------------
// The distance bank (r x P, L2 resident) is read ONCE per CTA-step and applied to kMT trajectories, which divides
// the L2 traffic of the selection by kMT; sums follow numpy's pairwise order (8 strided accumulators per row,
// r <= 128), so indices stay bit-exact.  RT > 0 fixes r at compile time (Diamond: 36) so the whole distance loop
// unrolls with immediate offsets.  The affine step is one thread per row against the gathered A_i (L2).
constexpr int kMT = 4;        // trajectories per CTA
constexpr int kMThreads = 256;

// sum_j (bank[j][p] - x[tr][j])^2 in numpy's pairwise order for the kMT trajectories; xs = transposed states
// xs[j * kMT + tr].  Adds w * sqrt(sum) to out[tr].
template <int RT>
__device__ __forceinline__ void multi_distances(int r_runtime, int P, const double* __restrict__ bank_p,
                                                const double* __restrict__ xs, double w, double (&out)[kMT]) {
    const int r = RT > 0 ? RT : r_runtime;
    double res[kMT];
    if (r < 8) {
#pragma unroll
        for (int tr = 0; tr < kMT; ++tr) res[tr] = 0.0;
        for (int j = 0; j < r; ++j) {
            const double qv = bank_p[(size_t)j * P];
#pragma unroll
            for (int tr = 0; tr < kMT; ++tr) {
                const double t = __dsub_rn(qv, xs[j * kMT + tr]);
                res[tr] = __dadd_rn(res[tr], __dmul_rn(t, t));
            }
        }
    } else {
        double acc[kMT][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double qv = bank_p[(size_t)k * P];
#pragma unroll
            for (int tr = 0; tr < kMT; ++tr) {
                const double t = __dsub_rn(qv, xs[k * kMT + tr]);
                acc[tr][k] = __dmul_rn(t, t);
            }
        }
        const int rfull = r - (r % 8);
#pragma unroll
        for (int j = 8; j < rfull; j += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double qv = bank_p[(size_t)(j + k) * P];
#pragma unroll
                for (int tr = 0; tr < kMT; ++tr) {
                    const double t = __dsub_rn(qv, xs[(j + k) * kMT + tr]);
                    acc[tr][k] = __dadd_rn(acc[tr][k], __dmul_rn(t, t));
                }
            }
        }
#pragma unroll
        for (int tr = 0; tr < kMT; ++tr)
            res[tr] = __dadd_rn(__dadd_rn(__dadd_rn(acc[tr][0], acc[tr][1]), __dadd_rn(acc[tr][2], acc[tr][3])),
                                __dadd_rn(__dadd_rn(acc[tr][4], acc[tr][5]), __dadd_rn(acc[tr][6], acc[tr][7])));
#pragma unroll
        for (int j = rfull; j < r; ++j) {
            const double qv = bank_p[(size_t)j * P];
#pragma unroll
            for (int tr = 0; tr < kMT; ++tr) {
                const double t = __dsub_rn(qv, xs[j * kMT + tr]);
                res[tr] = __dadd_rn(res[tr], __dmul_rn(t, t));
            }
        }
    }
#pragma unroll
    for (int tr = 0; tr < kMT; ++tr) out[tr] = __dadd_rn(out[tr], __dmul_rn(w, sqrt(res[tr])));
}

template <int RT>
__global__ void __launch_bounds__(kMThreads, 2)
tpwl_rollout_nn_multi_kernel(TpwlDev M, long long batch, int N, const double* __restrict__ x0,
                             const double* __restrict__ u, double* __restrict__ xo, int* __restrict__ idxo) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red_d[kMT][kMThreads / 32];
    __shared__ int red_i[kMT][kMThreads / 32];
    __shared__ int sel[kMT];
    const int n = M.n, m = M.m, r = RT > 0 ? RT : M.r, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* sxT = sm;                 // n x kMT  (transposed: the kMT states of one coordinate are contiguous)
    double* su = sxT + kMT * n;       // kMT x m
    double* sxn = su + kMT * m;       // kMT x n
    const long long groups = (batch + kMT - 1) / kMT;
    for (long long gidx = blockIdx.x; gidx < groups; gidx += gridDim.x) {
        const long long b0 = gidx * kMT;
        const int nt = (int)((batch - b0) < kMT ? (batch - b0) : kMT);
        for (int e = tid; e < kMT * n; e += kMThreads) {
            const int tr = e / n, i = e - tr * n;
            const double v = (tr < nt) ? x0[(b0 + tr) * n + i] : 0.0;
            sxT[i * kMT + tr] = v;
            if (tr < nt) xo[(b0 + tr) * (long long)(N + 1) * n + i] = v;
        }
        __syncthreads();
        for (int t = 0; t < N; ++t) {
            for (int e = tid; e < kMT * m; e += kMThreads) {
                const int tr = e / m, i = e - tr * m;
                su[e] = (tr < nt) ? u[((b0 + tr) * (long long)N + t) * m + i] : 0.0;
            }
            // ---- nearest stored point for the kMT states (x = [v; q]: q at offset r, v at offset 0)
            double best[kMT];
            int bi[kMT];
#pragma unroll
            for (int tr = 0; tr < kMT; ++tr) { best[tr] = INFINITY; bi[tr] = 0x7fffffff; }
            for (int p = tid; p < M.P; p += kMThreads) {
                double dd[kMT];
#pragma unroll
                for (int tr = 0; tr < kMT; ++tr) dd[tr] = 0.0;
                if (M.wq != 0.0) multi_distances<RT>(r, M.P, M.qT + p, sxT + r * kMT, M.wq, dd);
                if (M.wv != 0.0) {
                    double dv[kMT];
#pragma unroll
                    for (int tr = 0; tr < kMT; ++tr) dv[tr] = 0.0;
                    multi_distances<RT>(r, M.P, M.vT + p, sxT, M.wv, dv);
#pragma unroll
                    for (int tr = 0; tr < kMT; ++tr) dd[tr] = __dadd_rn(dd[tr], dv[tr]);
                }
#pragma unroll
                for (int tr = 0; tr < kMT; ++tr)
                    if (dd[tr] < best[tr]) { best[tr] = dd[tr]; bi[tr] = p; }
            }
#pragma unroll
            for (int tr = 0; tr < kMT; ++tr) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, best[tr], off);
                    const int op = __shfl_xor_sync(0xffffffffu, bi[tr], off);
                    if (od < best[tr] || (od == best[tr] && op < bi[tr])) { best[tr] = od; bi[tr] = op; }
                }
                if (lane == 0) { red_d[tr][warp] = best[tr]; red_i[tr][warp] = bi[tr]; }
            }
            __syncthreads();
            if (tid < kMT) {
                double d = red_d[tid][0];
                int p = red_i[tid][0];
                for (int k = 1; k < kMThreads / 32; ++k) {
                    const double od = red_d[tid][k];
                    const int op = red_i[tid][k];
                    if (od < d || (od == d && op < p)) { d = od; p = op; }
                }
                if (p == 0x7fffffff) p = 0;
                sel[tid] = p;
                if (idxo && tid < nt) idxo[(b0 + tid) * (long long)N + t] = p;
            }
            __syncthreads();
            // ---- x+ = (A_i x + B_i u) + d_i : one thread per (trajectory, row), sequential dot in ascending k
            for (int row = tid; row < kMT * n; row += kMThreads) {
                const int tr = row / n, i = row - tr * n;
                const long long p = sel[tr];
                const double* Ap = M.A + (p * n + i) * n;
                const double* Bp = M.B + (p * n + i) * m;
                double ax = 0.0, bu = 0.0;
                if ((n & 1) == 0) {
                    for (int k = 0; k < n; k += 2) {
                        const double2 a2 = *reinterpret_cast<const double2*>(Ap + k);
                        ax = fma(a2.x, sxT[k * kMT + tr], ax);
                        ax = fma(a2.y, sxT[(k + 1) * kMT + tr], ax);
                    }
                } else {
                    for (int k = 0; k < n; ++k) ax = fma(Ap[k], sxT[k * kMT + tr], ax);
                }
                for (int k = 0; k < m; ++k) bu = fma(Bp[k], su[tr * m + k], bu);
                sxn[row] = __dadd_rn(__dadd_rn(ax, bu), M.d[p * n + i]);
            }
            __syncthreads();
            for (int e = tid; e < kMT * n; e += kMThreads) {
                const int tr = e / n, i = e - tr * n;
                sxT[i * kMT + tr] = sxn[e];
                if (tr < nt) xo[((b0 + tr) * (long long)(N + 1) + t + 1) * n + i] = sxn[e];
            }
            __syncthreads();
        }
    }
}

int tpwl_rollout_nn_screen_launch(const TpwlDev& M, long long batch, int N, const double* x0, const double* u,
                                  double* x, int* idx, cudaStream_t st, bool* handled);
int tpwl_rollout_nn_resident_launch(const TpwlDev& M, long long batch, int N, const double* x0, const double* u,
                                    double* x, int* idx, cudaStream_t st, bool* handled);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace srcb

using namespace srcb;

extern "C" int srcb200_tpwl_nearest_batch(const srcb200_tpwl_model* mdl, int64_t count, const double* x,
                                          int32_t* idx, double* dist, void* stream) {
    if (int e = check_tpwl_model(mdl)) return e;
    if (count < 0) return fail(SRCB200_E_DIM, "count < 0");
    if (count == 0) return 0;
    if (!x || !idx) return fail(SRCB200_E_NULL, "x/idx is NULL");
    TpwlDev M = to_dev(*mdl);
    const int grid = (int)(count < 148 * 16 ? count : 148 * 16);
    tpwl_nearest_kernel<<<grid, kSel, sizeof(double) * M.n, (cudaStream_t)stream>>>(M, count, x, idx, dist);
    SRCB_LAUNCH_CHECK("tpwl_nearest_kernel");
    return 0;
}

extern "C" int srcb200_tpwl_weights_batch(const srcb200_tpwl_model* mdl, int64_t count, const double* x, double* w,
                                          void* stream) {
    if (int e = check_tpwl_model(mdl)) return e;
    if (count < 0) return fail(SRCB200_E_DIM, "count < 0");
    if (count == 0) return 0;
    if (!x || !w) return fail(SRCB200_E_NULL, "x/w is NULL");
    TpwlDev M = to_dev(*mdl);
    const size_t smem = sizeof(double) * (M.n + M.P);
    if (smem > 227 * 1024) return fail(SRCB200_E_DIM, "weights: P=%d too large for shared memory", M.P);
    SRCB_CUDA(cudaFuncSetAttribute(tpwl_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(count < 148 * 16 ? count : 148 * 16);
    tpwl_weights_kernel<<<grid, kSel, smem, (cudaStream_t)stream>>>(M, count, x, M.n, w);
    SRCB_LAUNCH_CHECK("tpwl_weights_kernel");
    return 0;
}

extern "C" int srcb200_tpwl_output_batch(const srcb200_tpwl_model* mdl, int64_t count, const double* x, double* z,
                                         void* stream) {
    if (int e = check_tpwl_model(mdl)) return e;
    if (mdl->nz == 0) return fail(SRCB200_E_NULL, "Need to set output or meas. model");   // tpwl.py:126
    if (count < 0) return fail(SRCB200_E_DIM, "count < 0");
    if (count == 0) return 0;
    if (!x || !z) return fail(SRCB200_E_NULL, "x/z is NULL");
    TpwlDev M = to_dev(*mdl);
    const long long blocks = (count * M.nz + 255) / 256;
    tpwl_output_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(M, count, x, z);
    SRCB_LAUNCH_CHECK("tpwl_output_kernel");
    return 0;
}

extern "C" int srcb200_discretize_batch(int32_t n, int32_t m, int32_t discr_method, int64_t count, double dt,
                                        const double* A_c, const double* B_c, const double* d_c, double* A_d,
                                        double* B_d, double* d_d, void* stream) {
    if (n < 1 || m < 1 || n > 128 || m > 32 || count < 0) return fail(SRCB200_E_DIM, "discretize: bad dims n=%d m=%d", n, m);
    if (discr_method != SRCB200_DISCR_FE && discr_method != SRCB200_DISCR_BE && discr_method != SRCB200_DISCR_BIL)
        return fail(SRCB200_E_METHOD, "self.discr_method must be in [fe, be, bil, zoh]");
    if (count == 0) return 0;
    if (!A_c || !B_c || !d_c || !A_d || !B_d || !d_d) return fail(SRCB200_E_NULL, "discretize: NULL operand");
    return discretize_launch(n, m, discr_method, count, dt, A_c, B_c, d_c, A_d, B_d, d_d, (cudaStream_t)stream);
}

// workspace layout for linearize (weighting): [ W (count x P) | blend (count x width) ] ; nn: [ idx (count) ]
extern "C" size_t srcb200_tpwl_linearize_workspace(const srcb200_tpwl_model* mdl, int64_t count) {
    if (!mdl || count <= 0) return 0;
    TpwlDev M = to_dev(*mdl);
    const size_t z = (M.discr == SRCB200_DISCR_ZOH) ? align_up(srcb200_zoh_workspace(M.n, M.m, count), 256) : 0;
    if (M.method == SRCB200_TPWL_NN) return align_up(sizeof(int32_t) * (size_t)count, 256) + z;
    return align_up(sizeof(double) * (size_t)count * M.P, 256) + z + align_up(blend_stream_workspace(M.n * M.n), 256);
}

extern "C" int srcb200_tpwl_linearize_batch(const srcb200_tpwl_model* mdl, int64_t count, const double* x, double dt,
                                            double* A, double* B, double* d, int32_t* idx, void* workspace,
                                            size_t workspace_bytes, void* stream) {
    if (int e = check_tpwl_model(mdl)) return e;
    if (count < 0) return fail(SRCB200_E_DIM, "count < 0");
    if (count == 0) return 0;
    if (!x || !A || !B || !d) return fail(SRCB200_E_NULL, "linearize: x/A/B/d is NULL");
    if (!workspace || workspace_bytes < srcb200_tpwl_linearize_workspace(mdl, count))
        return fail(SRCB200_E_WORKSPACE, "linearize: workspace too small");
    TpwlDev M = to_dev(*mdl);
    cudaStream_t st = (cudaStream_t)stream;
    const int n = M.n, m = M.m;
    const bool disc = (dt >= 0.0 && M.discr != SRCB200_DISCR_NONE);
    const int grid = (int)(count < 148 * 16 ? count : 148 * 16);
    if (M.method == SRCB200_TPWL_NN) {
        int32_t* widx = idx ? idx : (int32_t*)workspace;
        tpwl_nearest_kernel<<<grid, kSel, sizeof(double) * n, st>>>(M, count, x, widx, nullptr);
        SRCB_LAUNCH_CHECK("tpwl_nearest_kernel");
        tpwl_gather_kernel<<<grid, 256, 0, st>>>(M, count, widx, A, B, d);
        SRCB_LAUNCH_CHECK("tpwl_gather_kernel");
    } else {
        // the bank is three arrays (A, B, d) -> three GEMMs against the same weight matrix
        double* W = (double*)workspace;
        void* sws = (char*)workspace + workspace_bytes - align_up(blend_stream_workspace(n * n), 256);   // tail of the workspace
        if (int e = srcb200_tpwl_weights_batch(mdl, count, x, W, stream)) return e;
        if (int e = blend(M.A, M.P, (long long)n * n, W, count, A, sws, st)) return e;
        if (int e = blend(M.B, M.P, (long long)n * m, W, count, B, sws, st)) return e;
        if (int e = blend(M.d, M.P, n, W, count, d, sws, st)) return e;
    }
    if (disc) {
        const size_t first = (M.method == SRCB200_TPWL_NN) ? align_up(sizeof(int32_t) * (size_t)count, 256)
                                                           : align_up(sizeof(double) * (size_t)count * M.P, 256);
        const size_t zb = (M.discr == SRCB200_DISCR_ZOH) ? align_up(srcb200_zoh_workspace(n, m, count), 256) : 0;
        return discretize_any(n, m, M.discr, count, dt, A, B, d, (char*)workspace + first, zb, st);
    }
    return 0;
}

// rollout workspace (weighting only): [ W (batch x P) | A (batch x n x n) | B (batch x n x m) | d (batch x n) | zoh scratch |
//                                       concatenated bank (P x wd) | blended rows (batch x wd) | blend-stream scratch ]
static long long weighting_width(const TpwlDev& M) { return ((long long)M.n * M.n + (long long)M.n * M.m + M.n + 1) & ~1LL; }
extern "C" size_t srcb200_tpwl_rollout_workspace(const srcb200_tpwl_model* mdl, int64_t batch) {
    if (!mdl || batch <= 0) return 0;
    TpwlDev M = to_dev(*mdl);
    if (M.method == SRCB200_TPWL_NN) return 256;
    const size_t z = (M.discr == SRCB200_DISCR_ZOH) ? align_up(srcb200_zoh_workspace(M.n, M.m, batch), 256) : 0;
    return align_up(sizeof(double) * (size_t)batch * M.P, 256) + align_up(sizeof(double) * (size_t)batch * M.n * M.n, 256) +
           align_up(sizeof(double) * (size_t)batch * M.n * M.m, 256) + align_up(sizeof(double) * (size_t)batch * M.n, 256) + z +
           align_up(sizeof(double) * (size_t)M.P * weighting_width(M), 256) +
           align_up(sizeof(double) * (size_t)batch * weighting_width(M), 256) + align_up(blend_stream_workspace(M.n * M.n), 256);
}

extern "C" int srcb200_tpwl_rollout_batch(const srcb200_tpwl_model* mdl, int64_t batch, int32_t N, const double* x0,
                                          const double* u, double dt, double* x, double* z, int32_t* idx,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = check_tpwl_model(mdl)) return e;
    if (batch < 0 || N < 0) return fail(SRCB200_E_DIM, "batch/N < 0");
    if (batch == 0) return 0;
    if (!x0 || !x || (N > 0 && !u)) return fail(SRCB200_E_NULL, "rollout: x0/u/x is NULL");
    if (z && mdl->nz == 0) return fail(SRCB200_E_NULL, "Need to set output or meas. model");
    TpwlDev M = to_dev(*mdl);
    cudaStream_t st = (cudaStream_t)stream;
    const int n = M.n, m = M.m;
    const bool disc = (dt >= 0.0 && M.discr != SRCB200_DISCR_NONE);
    if (M.method == SRCB200_TPWL_NN) {
        const int grid = (int)(batch < 148 * 16 ? batch : 148 * 16);
        if (disc && M.discr == SRCB200_DISCR_ZOH)
            return fail(SRCB200_E_METHOD, "nn rollout with zoh needs a pre-discretised bank: discretise the bank once "
                                          "(srcb200_zoh_batch) and pass it with discr_method NONE");
        bool screened = false;
        if (!disc) if (int e = tpwl_rollout_nn_resident_launch(M, batch, N, x0, u, x, idx, st, &screened)) return e;
        if (!disc && !screened) if (int e = tpwl_rollout_nn_screen_launch(M, batch, N, x0, u, x, idx, st, &screened)) return e;
        if (screened) {
            // resident-entry kernel (tpwl_resident.cu) or the exact two-stage search kernel (tpwl_screen.cu) took it
        } else if (!disc && M.r <= 128) {
            const long long groups = (batch + kMT - 1) / kMT;
            const int g2 = (int)(groups < 148 * 8 ? groups : 148 * 8);
            const size_t smem = sizeof(double) * kMT * (2 * n + m);
            if (M.r == 36) tpwl_rollout_nn_multi_kernel<36><<<g2, kMThreads, smem, st>>>(M, batch, N, x0, u, x, idx);
            else           tpwl_rollout_nn_multi_kernel<0><<<g2, kMThreads, smem, st>>>(M, batch, N, x0, u, x, idx);
        } else if (!disc) {
            const size_t smem = sizeof(double) * (2 * n + m);
            tpwl_rollout_nn_kernel<true><<<grid, kSel, smem, st>>>(M, batch, N, x0, u, dt, x, idx);
        } else {
            const size_t smem = sizeof(double) * (2 * n + m + n * n + n * m + n + discretize_scratch_doubles(n, m));
            if (smem > 227 * 1024) return fail(SRCB200_E_DIM, "rollout: n=%d needs %zu B of shared memory", n, smem);
            SRCB_CUDA(cudaFuncSetAttribute(tpwl_rollout_nn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tpwl_rollout_nn_kernel<false><<<grid, kSel, smem, st>>>(M, batch, N, x0, u, dt, x, idx);
        }
        SRCB_LAUNCH_CHECK("tpwl_rollout_nn_kernel");
    } else {
        // weighting: the whole batch advances in lock step; per time step
        //   weights (batch x P)  ->  DMMA blend W * bank  ->  discretise  ->  affine step
        if (!workspace || workspace_bytes < srcb200_tpwl_rollout_workspace(mdl, batch))
            return fail(SRCB200_E_WORKSPACE, "rollout: workspace too small");
        char* wp = (char*)workspace;
        double* W = (double*)wp;   wp += align_up(sizeof(double) * (size_t)batch * M.P, 256);
        double* Ab = (double*)wp;  wp += align_up(sizeof(double) * (size_t)batch * n * n, 256);
        double* Bb = (double*)wp;  wp += align_up(sizeof(double) * (size_t)batch * n * m, 256);
        double* db = (double*)wp;  wp += align_up(sizeof(double) * (size_t)batch * n, 256);
        void* zws = wp;
        const size_t zws_bytes = (M.discr == SRCB200_DISCR_ZOH) ? align_up(srcb200_zoh_workspace(n, m, batch), 256) : 0;
        void* sws = (char*)workspace + workspace_bytes - align_up(blend_stream_workspace(n * n), 256);
        const long long xs = (long long)(N + 1) * n, us = (long long)N * m;
        SRCB_CUDA(cudaMemcpy2DAsync(x, sizeof(double) * xs, x0, sizeof(double) * n, sizeof(double) * n, batch,
                                    cudaMemcpyDeviceToDevice, st));
        const int grid = (int)(batch < 148 * 16 ? batch : 148 * 16);
        const size_t wsmem = sizeof(double) * (n + M.P);
        SRCB_CUDA(cudaFuncSetAttribute(tpwl_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
        const long long wd = weighting_width(M);
        double* cat = (double*)((char*)zws + zws_bytes);
        double* blended = (double*)((char*)cat + align_up(sizeof(double) * (size_t)M.P * wd, 256));
        const bool fused = !(disc && M.discr == SRCB200_DISCR_ZOH) && batch > kBlendStreamMaxBatch;
        // fe (and an already discrete bank) discretises elementwise: no LU scratch, two CTAs fit on an SM
        const int dscr = (disc && M.discr != SRCB200_DISCR_FE) ? discretize_scratch_doubles(n, m) : 0;
        const size_t fsmem = sizeof(double) * (2 * (size_t)wd + 2 * n + ((m + 1) & ~1) + (size_t)max(dscr, M.P) + 2);
        if (fused && fsmem <= 227 * 1024) {
            // two launches per time step: blend GEMM over the concatenated bank, fused discretise + step + next weights
            SRCB_CUDA(cudaFuncSetAttribute(tpwl_weighting_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            tpwl_concat_bank_kernel<<<(int)(M.P < 148 * 8 ? M.P : 148 * 8), 256, 0, st>>>(M, wd, cat);
            SRCB_LAUNCH_CHECK("tpwl_concat_bank_kernel");
            tpwl_weights_kernel<<<grid, kSel, wsmem, st>>>(M, batch, x, xs, W);
            SRCB_LAUNCH_CHECK("tpwl_weights_kernel");
            const int fgrid = (int)(batch < 148 * 2 ? batch : 148 * 2);
            for (int t = 0; t < N; ++t) {
                if (int e = dgemm_device(0, batch, wd, M.P, 1.0, W, M.P, cat, wd, blended, wd, st)) return e;
                tpwl_weighting_step_kernel<<<fgrid, kWS, fsmem, st>>>(M, batch, disc ? 1 : 0, dt, blended, wd, x + (long long)t * n, xs,
                                                                      u + (long long)t * m, us, x + (long long)(t + 1) * n, xs,
                                                                      (t + 1 < N) ? W : nullptr);
                SRCB_LAUNCH_CHECK("tpwl_weighting_step_kernel");
            }
        } else {
            for (int t = 0; t < N; ++t) {
                // the states of step t live strided inside x: x[b, t, :]
                tpwl_weights_kernel<<<grid, kSel, wsmem, st>>>(M, batch, x + (long long)t * n, xs, W);
                SRCB_LAUNCH_CHECK("tpwl_weights_kernel");
                if (int e = blend(M.A, M.P, (long long)n * n, W, batch, Ab, sws, st)) return e;
                if (int e = blend(M.B, M.P, (long long)n * m, W, batch, Bb, sws, st)) return e;
                if (int e = blend(M.d, M.P, n, W, batch, db, sws, st)) return e;
                if (disc) if (int e = discretize_any(n, m, M.discr, batch, dt, Ab, Bb, db, zws, zws_bytes, st)) return e;
                tpwl_step_kernel<<<grid, 128, 0, st>>>(n, m, batch, Ab, Bb, db, x + (long long)t * n, xs, u + (long long)t * m, us,
                                                       x + (long long)(t + 1) * n, xs);
                SRCB_LAUNCH_CHECK("tpwl_step_kernel");
            }
        }
    }
    if (z) {
        const long long cnt = batch * (long long)(N + 1);
        tpwl_output_kernel<<<(int)((cnt * M.nz + 255) / 256 < 148 * 8 ? (cnt * M.nz + 255) / 256 : 148 * 8), 256, 0, st>>>(M, cnt, x, z);
        SRCB_LAUNCH_CHECK("tpwl_output_kernel");
    }
    return 0;
}
