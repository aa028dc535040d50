// ilqr_fast.cu -- Trunk / Diamond SSM iLQR (n = n_z = 6, cubic basis, m = 4 or 8, Gauss-Newton cost) specialised
// for B200: ONE WARP PER PROBLEM, eight problems per CTA sharing one coefficient table in shared memory.
//
//   * small dense algebra on the FP64 tensor pipe: every 6x6 / 8x6 / 8x8 product of the Riccati sweep
//     (ilqr.py:258-295) and of the discretisation (ssm.py:279-301) is one 8x8x8 DMMA tile (2 x mma.sync.m8n8k4.f64)
//     with the accumulator kept in registers across chained products; vectors ride along in the padding column
//     (P | p), (H | e), (Q_ux | Q_u), (K | k), so Q_x, Q_u, k and p cost no extra instructions.
//   * factorisations with warp shuffles, matrices held one column per lane in registers: the two 6x6 inverses of
//     the implicit-Euler / bilinear discretisation run simultaneously in the two half-warps (Gauss-Jordan with
//     partial pivoting); inv(Q_uu~) is a Gauss-Jordan sweep whose pivots double as the positive-definiteness test
//     (for a symmetric matrix the k-th pivot IS the k-th Cholesky pivot a_kk - sum l_kj^2 of np.linalg.cholesky).
//   * the polynomial model (ssm.py:158-235) is evaluated sparsely: d phi / d x_j of a degree-<=3 monomial is a
//     multiple of a degree-<=2 monomial, so A_c = r_coeff dphi/dx and H = w_coeff dphi/dx are 72 dot products of
//     length 28 against psi = (1, x, x (x) x), accumulated per polynomial degree; the VALUES f, z then cost no
//     coefficient reads at all: for a homogeneous polynomial h of degree p, x . grad h = p h (Euler), hence
//     f_i = sum_j x_j (g1_ij + g2_ij / 2 + g3_ij / 3) with g_p the degree-p part of row i of the Jacobian.  The
//     coefficient table (the dominant shared-memory traffic of a step) shrinks from 108 to 72 rows of 28, spread
//     over the 32 lanes in three rounds and read with conflict-free LDS.128.
//
// Control flow (line search, rho schedule, interrupted sweep on a non-PD Q_uu~, convergence) is identical to the generic kernel in ilqr.cu
// and to the reference (ilqr.py:27-107); results agree with it to rounding (tests/test_ilqr_gpu.py).
#include <cstdlib>
#include "ilqr.cuh"

namespace srcb {
namespace fast {

constexpr int LD = 12;            // tile row stride (doubles): A-/B-fragment loads are bank-conflict free
constexpr int TILE = 8 * LD;
constexpr int WARPS = 8;          // problems in flight per CTA.  Two launch shapes of the solve kernel (template CR):
                                  //   CR = true : 1 CTA per SM,  255 registers, the Jacobian-table rows of a lane in registers
                                  //   CR = false: 2 CTAs per SM, 128 registers, the table read from shared memory every step
constexpr int NJ = 30;            // slots per Jacobian row: [deg1 c, 0 | deg2: 6 | deg3: 21, 0] -- degrees on even boundaries
constexpr int NPD = 73;           // table rows: 72 Jacobian rows (A_c 0..35, H 36..71) + one zero row for idle lanes
constexpr int TS = 30;            // table row stride (doubles): LDS.128 of 8 consecutive rows hit 8 distinct 16 B banks
constexpr int NFEAT = 83;
#ifndef SRCB_GAIN_SHFL
#define SRCB_GAIN_SHFL 0
#endif
constexpr unsigned FULL = 0xffffffffu;

// per-CTA shared block (doubles)
constexpr int SH_T = 0;                         // coefficient table NPD x TS
constexpr int SH_Q = SH_T + NPD * TS;           // Q tile
constexpr int SH_R = SH_Q + TILE;               // R tile
constexpr int SH_QF = SH_R + TILE;              // Qf tile
constexpr int SH_BR = SH_QF + TILE;             // B_r tile (6 x m)
constexpr int SH_ZREF = SH_BR + TILE;           // 8
constexpr int SH_FIDX = SH_ZREF + 8;            // 84 ints = 42 doubles
constexpr int SH_END = SH_FIDX + 44;
// per-warp block (doubles)
constexpr int W_PHI = 0;                        // psi slots 0..29 (see NJ), then PV[72]: x_j * value part of row o at W_PHI + 32
constexpr int W_PV = 32;                        // 72 (+0)
constexpr int W_X = 104;                        // x_t (6), X[6] = 0, X[7] = 1
constexpr int W_DC = 128;                       // d_c (rollout kernel) / c_u (backward pass)
constexpr int W_DD = 136;                       // d_d (rollout kernel) / du (backward pass)
constexpr int W_TILES = 144;
constexpr int NTILES = 11;
constexpr int W_SIZE = W_TILES + NTILES * TILE; // 1200 doubles = 9.4 KB per warp
constexpr size_t SMEM_BYTES = sizeof(double) * (SH_END + WARPS * W_SIZE);

struct Frag { double c0, c1; };

// Optional phase clocks (build with -DSRCB_PHASE_TIMING, read with srcb200_debug_phase): lane 0 of every warp adds
// the clock64 distance between consecutive marks to a global table.  0..7 forward step, 8..15 backward step.
#ifdef SRCB_PHASE_TIMING
__device__ unsigned long long g_phase[32];
// per-warp accumulation in registers, one atomic per phase and PASS (not per step: 2368 warps hammering 14 words would
// distort the loaded measurement)
#define PH_DECL long long ph_t = clock64(); long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PH(i) do { const long long ph_n = clock64(); ph_acc[(i) & 7] += ph_n - ph_t; ph_t = ph_n; } while (0)
#define PH_FLUSH(base) do { if (lane == 0) { _Pragma("unroll") for (int ph_i = 0; ph_i < 8; ++ph_i) if (ph_acc[ph_i]) atomicAdd(&g_phase[(base) + ph_i], (unsigned long long)ph_acc[ph_i]); } } while (0)
#else
#define PH_DECL
#define PH(i)
#define PH_FLUSH(base)
#endif

// Reciprocal without the slow-path branch of __drcp_rn (which splits the basic block and keeps the scheduler from
// overlapping it with the pivot search): MUFU.RCP64H seed (rcp.approx.ftz.f64, ~2^-20) + two Newton steps; within
// 1 ulp for normal inputs, inf / nan for 0 / inf / nan like the division it replaces.
__device__ __forceinline__ double rcp_fast(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

__device__ __forceinline__ void dmma(Frag& c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c.c0), "+d"(c.c1) : "d"(a), "d"(b));
}
// C += op(A) op(B) for 8x8 tiles in shared memory; g = lane >> 2, q = lane & 3
template <bool TA, bool TB>
__device__ __forceinline__ void mma88(Frag& c, const double* __restrict__ A, const double* __restrict__ B, int g, int q) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int k = 4 * s + q;
        dmma(c, TA ? A[k * LD + g] : A[g * LD + k], TB ? B[g * LD + k] : B[k * LD + g]);
    }
}
// C-fragment coordinates (row g, columns 2q, 2q+1) <-> tile.  With the row stride of 12 doubles that makes every DMMA
// fragment LOAD conflict-free, a 16-byte access of this pattern is a 2-way bank conflict (rows g and g + 1 of a quarter
// warp overlap in 8 of the 32 banks: 8 wavefronts instead of 4 -- it was a third of all excess shared-memory wavefronts
// of the kernel, and the shared-memory pipe is the loaded unit).  Two 8-byte accesses in which odd rows take their two
// columns in the opposite order touch 16 distinct 8-byte banks per half warp: 2 + 2 wavefronts.
__device__ __forceinline__ void sts_pair(double* __restrict__ T, int g, int q, double v0, double v1) {
    const int odd = g & 1;
    double* p = T + g * LD + 2 * q;
    p[odd] = odd ? v1 : v0;
    p[odd ^ 1] = odd ? v0 : v1;
}
__device__ __forceinline__ double2 lds_pair(const double* __restrict__ T, int g, int q) {
    const int odd = g & 1;
    const double* p = T + g * LD + 2 * q;
    const double a = p[odd], b = p[odd ^ 1];
    return odd ? make_double2(b, a) : make_double2(a, b);
}
__device__ __forceinline__ void store_frag(double* __restrict__ T, const Frag& c, int g, int q) {
    sts_pair(T, g, q, c.c0, c.c1);
}
__device__ __forceinline__ Frag load_frag(const double* __restrict__ T, int g, int q) {
    const double2 v = lds_pair(T, g, q);
    return Frag{v.x, v.y};
}
__device__ __forceinline__ void zero_tile(double* __restrict__ T, int lane) {
#pragma unroll
    for (int e = lane; e < TILE; e += 32) T[e] = 0.0;
}

extern __shared__ __align__(16) double g_sm[];   // [ CTA-shared block | WARPS per-warp blocks ]

struct Ctx {
    int lane, g, q;
    int off0, off1;  // tile offsets of elements `lane` and `32 + lane` (lane < 4) of a row-major 6 x 6 matrix
    int ws_off;      // this warp's block inside g_sm (offsets, not pointers: accesses stay LDS/STS across calls)
};
#define CTX_SH (g_sm)
#define CTX_WS(c) (g_sm + (c).ws_off)

// ---------------------------------------------------------------------------------------------------------------
// Coefficient table, built once per CTA from the model arrays.  Row pd (A_c[i][j]: pd = 6 i + j with r_coeff,
// H[i][j]: pd = 36 + 6 i + j with w_coeff) holds d/dx_j of the polynomial of output i, sorted by degree:
//   slot 0      : coefficient of x_j itself (degree-1 part: a constant), slot 1: 0
//   slots 2..7  : mult * coeff of the quadratic monomials containing x_j, operand psi = x_0..x_5
//   slots 8..28 : mult * coeff of the cubic monomials containing x_j, operand psi = the 21 products x_a x_b, slot 29: 0
// psi lives in the per-warp block with the same slot numbering (psi_0 = 1, psi_1 = psi_29 = 0).  Row 72 is zero.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int slot_to_q(int slot) {      // q indexes (1, x_0..x_5, 21 quadratic monomials); -1: padding
    if (slot == 0) return 0;
    if (slot >= 2 && slot <= 7) return slot - 1;
    if (slot >= 8 && slot <= 28) return slot - 1;
    return -1;
}
__device__ int find_monomial(const SsmDev& M, int a, int b, int c) {
    // sort ascending with 0xFF (absent) last
    if (a > b) { int t = a; a = b; b = t; }
    if (b > c) { int t = b; b = c; c = t; }
    if (a > b) { int t = a; a = b; b = t; }
    for (int k = 0; k < M.nfeat; ++k) {
        const uint8_t* r = M.mono + k * SRCB200_SSM_MAX_ORDER;
        if (r[0] == a && r[1] == b && r[2] == c) return k;
    }
    return -1;
}

__device__ void build_tables(const SsmDev& M, const double* Qg, const double* Rg, const double* Qfg, double* sh, int m) {
    double* T = sh + SH_T;
    for (int e = threadIdx.x; e < NPD * TS; e += blockDim.x) T[e] = 0.0;
    for (int e = threadIdx.x; e < 4 * TILE + 8; e += blockDim.x) sh[SH_Q + e] = 0.0;
    __syncthreads();
    // Jacobian rows
    for (int e = threadIdx.x; e < 72 * NJ; e += blockDim.x) {
        const int pd = e / NJ, slot = e - pd * NJ;
        const int q = slot_to_q(slot);
        if (q < 0) continue;
        const double* src = pd < 36 ? M.r : M.w;
        const int o = pd < 36 ? pd : pd - 36, i = o / 6, j = o - 6 * i;
        int s0 = 0xFF, s1 = 0xFF;
        if (q >= 1) { s0 = M.mono[(q - 1) * SRCB200_SSM_MAX_ORDER]; s1 = M.mono[(q - 1) * SRCB200_SSM_MAX_ORDER + 1]; }
        const int k = find_monomial(M, s0, s1, j);
        const int mult = 1 + (s0 == j) + (s1 == j);
        T[pd * TS + slot] = (k >= 0) ? (double)mult * src[i * M.nfeat + k] : 0.0;
    }
    for (int e = threadIdx.x; e < 36; e += blockDim.x) {
        const int i = e / 6, j = e - 6 * i;
        if (Qg) sh[SH_Q + i * LD + j] = Qg[e];
        if (Qfg) sh[SH_QF + i * LD + j] = Qfg[e];
    }
    if (Rg) for (int e = threadIdx.x; e < m * m; e += blockDim.x) sh[SH_R + (e / m) * LD + (e % m)] = Rg[e];
    for (int e = threadIdx.x; e < 6 * m; e += blockDim.x) sh[SH_BR + (e / m) * LD + (e % m)] = M.B[e];
    for (int e = threadIdx.x; e < 6; e += blockDim.x) sh[SH_ZREF + e] = M.zref[e];
    int* fidx = reinterpret_cast<int*>(sh + SH_FIDX);
    for (int k = threadIdx.x; k < NFEAT; k += blockDim.x) {
        const uint8_t* r = M.mono + k * SRCB200_SSM_MAX_ORDER;
        const int i0 = r[0], i1 = r[1] == 0xFF ? 7 : r[1], i2 = r[2] == 0xFF ? 7 : r[2];
        fidx[k] = i0 | (i1 << 3) | (i2 << 6);
    }
    __syncthreads();
}

struct Scatter { int o0, o1, o2; };   // per-lane tile offsets of the Jacobian outputs of rounds 0, 1, 2

__device__ __forceinline__ Scatter make_scatter(int lane) {
    Scatter s;
    s.o0 = (lane / 6) * LD + lane % 6;                                   // A_c, o = lane
    s.o1 = ((32 + lane) / 6) * LD + (32 + lane) % 6;                     // A_c, o = 32 + lane (lanes < 4)
    s.o2 = 0;
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// Two 6x6 inverses at once (half-warp h handles matrix h): Gauss-Jordan on [M | I] with partial pivoting, one
// column per lane (j = lane & 15 < 12), rows in registers, implicit row exchange.  The inverse lands in `dst`
// (tile, LD) for each half.  Zero pivots produce inf/nan exactly like the singular case would.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gj6_pair(double (&col)[6], int lane, double* __restrict__ dst) {
    // rows are exchanged physically (static register indices): at step c only rows c..5 are pivot candidates, the
    // pivot row / pivot value / eliminations need no select chains, and the right half ends up as the inverse.
    const int j = lane & 15;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const int src = (lane & 16) | c;
        double cc[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) cc[r] = __shfl_sync(FULL, col[r], src);
        // partial pivoting: first maximal |a_rc|, r >= c
        double best = fabs(cc[c]);
        int p = c;
#pragma unroll
        for (int r = c + 1; r < 6; ++r) {
            const double av = fabs(cc[r]);
            const bool take = av > best;
            best = take ? av : best;
            p = take ? r : p;
        }
#pragma unroll
        for (int r = c + 1; r < 6; ++r) {
            if (p == r) {
                const double t0 = col[c]; col[c] = col[r]; col[r] = t0;
                const double t1 = cc[c];  cc[c] = cc[r];   cc[r] = t1;
            }
        }
        const double pc = __dmul_rn(col[c], __drcp_rn(cc[c]));
        col[c] = pc;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            if (r != c) col[r] = fma(-cc[r], pc, col[r]);
        }
    }
    if (j >= 6 && j < 12) {
#pragma unroll
        for (int r = 0; r < 6; ++r) dst[r * LD + (j - 6)] = col[r];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Lean forward pass (ilqr.py:117-162): the same arithmetic as fwd_fast with a fraction of its instructions.
//   * every matrix-vector product of a step is a DMMA pair against a "vector tile": VT holds x_t, u_t, e_t, du_t in
//     columns 0..3 and VT2 holds d_c, B u, dx in columns 0..2, so A_c x, B u, Q^T e, R^T du (and later A_d x,
//     sep d_c, sep B u) fall out of the accumulator columns of the lanes that need them: (g, 0) owns row g of the
//     state update, (g, 1) owns input g, output g and their cost terms (kept in registers, summed once per pass).
//   * the two 6x6 inverses of the implicit discretisations are in-place Gauss-Jordan sweeps with one matrix ROW per
//     lane (matrix h = lane >> 4): partial pivoting picks a pivot LANE (a REDUX on the high words of |a_rc|), the
//     pivot row is broadcast with shuffles from that lane -- no row exchange, no select chains; the row / column
//     permutation is undone by the addresses of the final stores.
// ---------------------------------------------------------------------------------------------------------------
constexpr int T_AC = 0, T_AD = 1, T_IA = 2, T_SP = 3, T_IMH = 4, T_W0 = 5, T_VT = 6, T_VT2 = 7;

// rows of `a`: lane (h = lane >> 4, r = lane & 15 < 6) holds row r of matrix h.  On return dst_h = inv(matrix h).
// Latency matters more than instruction count here (six strictly sequential pivot steps; measured on B200,
// profiles/lat_microbench_r2.txt: reciprocal 72 cycles, REDUX 23, ballot + ffs 58, 64-bit shuffle 27), so
//   * every row computes the reciprocal of its OWN candidate while the pivot search runs; the pivot row's one is
//     broadcast with the row;
//   * the search key is the high word of |a_rc| with the row index in its low four bits: one REDUX.MAX per matrix
//     yields the winner AND its lane, no ballot / find-first-set.  Candidates that agree in the leading 16 mantissa
//     bits count as tied and the lower row wins -- either is as good a pivot as the exact maximum.
__device__ __forceinline__ void gj6_rows(double (&a)[6], int lane, double* __restrict__ dst) {
    const int r = lane & 15, hs = lane & 16;
    const bool lo = (lane < 16);
    const unsigned tag = 15u - (unsigned)r;
    bool used = (r >= 6);
    int mycol = 0;
    int pl[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const double rown = rcp_fast(a[c]);
        const unsigned key = used ? 0u : ((((unsigned)__double2hiint(a[c]) & 0x7ffffff0u)) | tag);
        const unsigned m0 = __reduce_max_sync(FULL, lo ? key : 0u);
        const unsigned m1 = __reduce_max_sync(FULL, lo ? 0u : key);
        const unsigned mx = lo ? m0 : m1;
        const int prow = 15 - (int)(mx & 15u);
        const int p = hs + prow;
        pl[c] = prow;
        double pr[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) pr[j] = (j == c) ? 0.0 : __shfl_sync(FULL, a[j], p);
        const double rp = __shfl_sync(FULL, rown, p);
        const bool isp = (lane == p);
        const double w = isp ? rp : -__dmul_rn(a[c], rp);      // pivot row: scale ; other rows: -multiplier
        const double z = isp ? 0.0 : 1.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (j != c) a[j] = fma(w, pr[j], __dmul_rn(a[j], z));
        }
        a[c] = w;                                               // the entering identity column e_p
        if (isp) { used = true; mycol = c; }
    }
    // lane r was the pivot of column mycol; register c belongs to the pivot row pl[c] of column c:
    // inv[mycol][pl[c]] = a[c]
    if (r < 6) {
#pragma unroll
        for (int c = 0; c < 6; ++c) dst[mycol * LD + pl[c]] = a[c];
    }
}

// Model evaluation for the lean pass.  A_c -> tile AC, (dlt - s A_c) -> tile AUX (I - h A_c for be / bil,
// I + dt A_c for fe, A_c itself for a discrete model), H -> record.  Returns the polynomial part of f_g in lanes
// (g < 6, q = 0) and the raw output z_g in lanes (g < 6, q = 1).
// The coefficient rows `lane` and `32 + lane` of the Jacobian table live in REGISTERS for the whole pass (ca, cb:
// slot 0 = the degree-1 constant, 1..6 = degree 2, 7..27 = degree 3; 112 registers, the kernel runs 8 warps per SM):
// the table reads were 2/3 of the shared-memory wavefronts of a forward step and the shared-memory pipe was the
// loaded unit.  Only rows 64..71 (lanes < 8) and the broadcast psi operands still come from shared memory.
constexpr int NCR = 28;
__device__ __forceinline__ void load_coeff_rows(int lane, double (&ca)[NCR], double (&cb)[NCR]) {
    const double* T = CTX_SH + SH_T;
    const double* t0 = T + lane * TS;
    const double* t1 = T + (32 + lane) * TS;
    ca[0] = t0[0]; cb[0] = t1[0];
#pragma unroll
    for (int s = 1; s < NCR; ++s) { ca[s] = t0[s + 1]; cb[s] = t1[s + 1]; }
}

template <int CR>
__device__ __forceinline__ double ssm_eval_lean(const Ctx c, const Scatter sc, const double (&ca)[NCR],
                                                const double (&cb)[NCR], double* __restrict__ AC,
                                                double* __restrict__ AUX, double s_aux, double d0, double d1,
                                                double* __restrict__ Hg) {
    const int lane = c.lane;
    double* PSI = CTX_WS(c) + W_PHI;
    double* PV = CTX_WS(c) + W_PV;
    const double* X = CTX_WS(c) + W_X;
    const int* fidx = reinterpret_cast<const int*>(CTX_SH + SH_FIDX);
    if (lane < 21) {
        const int pk = fidx[6 + lane];
        PSI[8 + lane] = __dmul_rn(X[pk & 7], X[(pk >> 3) & 7]);
    } else if (lane < 27) {
        PSI[2 + lane - 21] = X[lane - 21];
    }
    const double xj0 = X[lane % 6], xj1 = X[(32 + lane) % 6], xj2 = X[(64 + lane) % 6];
    __syncwarp();
    const double* t2 = CTX_SH + SH_T + (lane < 8 ? 64 + lane : 72) * TS;   // lanes >= 8: the zero row (one broadcast read)
    const double* t0 = CTX_SH + SH_T + lane * TS;
    const double* t1 = CTX_SH + SH_T + (32 + lane) * TS;
    // The 15 slot pairs are software-pipelined by hand: the 16-byte loads of pair i + 1 (psi and the table rows
    // lane / 32+lane / 64+lane) are issued, then a __syncwarp(), then the six FMAs of pair i.  The barrier is what pins the
    // schedule: left to itself (also with ordered volatile loads) ptxas sinks every table load next to its two FMAs and
    // re-uses one register quad, which serialises a shared-memory latency per FMA pair -- 2.2 k of the 4.9 k cycles of
    // a lone warp's forward step (tools/phase_clocks.py).
    auto ld2 = [](const double* p) { return *reinterpret_cast<const double2*>(p); };
    double2 ps = ld2(PSI + 2), c2 = ld2(t2 + 2), c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
    if (CR < 1) c0 = ld2(t0 + 2);
    if (CR < 2) c1 = ld2(t1 + 2);
    const double g10 = CR >= 1 ? ca[0] : t0[0], g11 = CR >= 2 ? cb[0] : t1[0], g12 = t2[0];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
    double g20 = 0.0, g21 = 0.0, g22 = 0.0;
#pragma unroll
    for (int q = 2; q < NJ; q += 2) {
        double2 psn = ps, c0n = c0, c1n = c1, c2n = c2;
        if (q + 2 < NJ) {
            psn = ld2(PSI + q + 2);
            if (CR < 1) c0n = ld2(t0 + q + 2);
            if (CR < 2) c1n = ld2(t1 + q + 2);
            c2n = ld2(t2 + q + 2);
            __syncwarp();
        }
        const double k0x = CR >= 1 ? ca[q - 1] : c0.x, k1x = CR >= 2 ? cb[q - 1] : c1.x;
        a0 = fma(k0x, ps.x, a0);
        a1 = fma(k1x, ps.x, a1);
        a2 = fma(c2.x, ps.x, a2);
        if (q + 1 < NJ - 1) {                       // slot 29 is padding
            const double k0y = CR >= 1 ? ca[q] : c0.y, k1y = CR >= 2 ? cb[q] : c1.y;
            b0 = fma(k0y, ps.y, b0);
            b1 = fma(k1y, ps.y, b1);
        }
        b2 = fma(c2.y, ps.y, b2);
        if (q == 6) {                               // end of the degree-2 slots (2..7)
            g20 = __dadd_rn(a0, b0); g21 = __dadd_rn(a1, b1); g22 = __dadd_rn(a2, b2);
            a0 = a1 = a2 = b0 = b1 = b2 = 0.0;
        }
        ps = psn; c0 = c0n; c1 = c1n; c2 = c2n;
    }
    const double g30 = __dadd_rn(a0, b0), g31 = __dadd_rn(a1, b1), g32 = __dadd_rn(a2, b2);
    const double j0 = __dadd_rn(__dadd_rn(g10, g20), g30);
    const double j1 = __dadd_rn(__dadd_rn(g11, g21), g31);
    const double j2 = __dadd_rn(__dadd_rn(g12, g22), g32);
    AC[sc.o0] = j0;
    AUX[sc.o0] = __dsub_rn(d0, __dmul_rn(s_aux, j0));
    if (lane < 4) {
        AC[sc.o1] = j1;
        AUX[sc.o1] = __dsub_rn(d1, __dmul_rn(s_aux, j1));
    } else {
        Hg[lane - 4] = j1;
    }
    if (lane < 8) Hg[28 + lane] = j2;
    constexpr double third = 1.0 / 3.0;
    PV[lane] = __dmul_rn(xj0, __dadd_rn(__dadd_rn(g10, __dmul_rn(0.5, g20)), __dmul_rn(third, g30)));
    PV[32 + lane] = __dmul_rn(xj1, __dadd_rn(__dadd_rn(g11, __dmul_rn(0.5, g21)), __dmul_rn(third, g31)));
    if (lane < 8) PV[64 + lane] = __dmul_rn(xj2, __dadd_rn(__dadd_rn(g12, __dmul_rn(0.5, g22)), __dmul_rn(third, g32)));
    __syncwarp();
    double val = 0.0;
    if (c.q < 2 && c.g < 6) {
        const double* pv = PV + 36 * c.q + 6 * c.g;
        const double2 p01 = *reinterpret_cast<const double2*>(pv);
        const double2 p23 = *reinterpret_cast<const double2*>(pv + 2);
        const double2 p45 = *reinterpret_cast<const double2*>(pv + 4);
        val = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(p01.x, p01.y), p23.x), p23.y), p45.x), p45.y);
    }
    return val;
}

// REC = false: open-loop rollout (ssm.py:134-156) -- no cost, no gains, no trajectory record: tr.x receives the states,
// tr.e the outputs z_t (instead of the output errors), nothing else is written.
template <int M, int DISCR, int CR, bool REC = true>
__device__ __noinline__ double fwd_lean(const Ctx c, const IlqrArgs& a, const double* nx, const double* nu,
                                        double alpha, const double* K, const double* k, const Rec tr,
                                        const double* __restrict__ ztar, const double* __restrict__ ulast) {
    constexpr bool IMPL = (DISCR == SRCB200_DISCR_BE || DISCR == SRCB200_DISCR_BIL);
    const int lane = c.lane, g = c.g, q = c.q, N = a.N;
    double* ws = CTX_WS(c);
    double* X = ws + W_X;
    double* AC = ws + W_TILES + T_AC * TILE;    // A_c
    double* AD = ws + W_TILES + T_AD * TILE;    // A_d
    double* IA = ws + W_TILES + T_IA * TILE;    // inv(A_c)
    double* SP = ws + W_TILES + T_SP * TILE;    // sep = inv(A_c) (A_d - I)
    double* IMH = ws + W_TILES + T_IMH * TILE;  // I - h A_c
    double* W0 = ws + W_TILES + T_W0 * TILE;    // inv(I - h A_c) for bil
    double* VT = ws + W_TILES + T_VT * TILE;    // columns: x_t | u_t | e_t | du_t
    double* VT2 = ws + W_TILES + T_VT2 * TILE;  // columns: d_c | B u | dx_t
    const double* Qt = CTX_SH + SH_Q;  const double* Rt = CTX_SH + SH_R;  const double* Qft = CTX_SH + SH_QF;
    const double* Brt = CTX_SH + SH_BR; const double* zref = CTX_SH + SH_ZREF;
    const Scatter sc = make_scatter(lane);
    const double dt = a.dt;
    const bool inc = a.cfg.include_input_var_constraint != 0;
    const bool lf = (q == 0 && g < 6);          // row g of the state update
    const bool lu = (q == 1);                   // input g / output g / their cost terms
    // aux tile of the evaluation: dlt - s A_c
    const double s_aux = (DISCR == SRCB200_DISCR_BE) ? dt : (DISCR == SRCB200_DISCR_BIL) ? 0.5 * dt
                       : (DISCR == SRCB200_DISCR_FE) ? -dt : -1.0;
    const bool addI = (DISCR != SRCB200_DISCR_NONE);
    const double d0 = (addI && lane / 6 == lane % 6) ? 1.0 : 0.0, d1 = (addI && lane == 3) ? 1.0 : 0.0;
    double* AUX = IMPL ? IMH : AD;
    const double zr = (lu && g < 6) ? zref[g] : 0.0;
    const bool dg0 = (q == g), dg1 = (4 + q == g) && (g < 6);   // diagonal flags of the B fragment (k = q / 4 + q, n = g)
    double cacc = 0.0;
    double ca[NCR], cb[NCR];
    if (CR >= 1) load_coeff_rows(lane, ca, cb);

    for (int t = 0; t < 8; ++t) zero_tile(ws + W_TILES + t * TILE, lane);
    if (lane < 3) ws[W_PHI + (lane == 0 ? 0 : (lane == 1 ? 1 : 29))] = (lane == 0) ? 1.0 : 0.0;   // psi_0 = 1, padding slots 0
    __syncwarp();
    if (lane < 8) X[lane] = lane < 6 ? nx[lane] : (lane == 7 ? 1.0 : 0.0);
    if (lf) { const double v = nx[g]; VT[g * LD] = v; tr.x[g] = v; }
    double up = (lu && g < M && ulast) ? ulast[g] : 0.0;
    // prefetch registers of step 0: K[0] as an A fragment, nominal u / k in the input lanes, target in the output lanes
    double pK0 = 0.0, pK1 = 0.0, p_nu = 0.0, p_k = 0.0, p_nx = 0.0, p_zt = 0.0;
    if (g < M && N > 0) {
        if (K) { pK0 = K[g * 6 + q]; if (q < 2) pK1 = K[g * 6 + 4 + q]; }
        if (lu) { p_nu = nu[g]; if (k) p_k = k[g]; }
    }
    if (REC && lu && g < 6) p_zt = ztar[g];
    __syncwarp();

    PH_DECL;
    for (int t = 0; t <= N; ++t) {
        const bool last = (t == N);
        double u = 0.0, du = 0.0;
        if (!last) {
            // u_t = (u_prev[t] + alpha k[t]) + K[t] (x[t] - x_prev[t])          (ilqr.py:140): column 2 of K VT2
            Frag uf{0.0, 0.0};
            if (K) {
                dmma(uf, pK0, VT2[q * LD + g]);
                dmma(uf, pK1, VT2[(4 + q) * LD + g]);
            }
            if (lu) {
                double v = p_nu;
                if (k) v = __dadd_rn(v, __dmul_rn(alpha, p_k));
                if (K) v = __dadd_rn(v, uf.c0);
                u = v;
                du = inc ? __dsub_rn(u, up) : u;
                VT[g * LD + 1] = u;
                VT[g * LD + 3] = du;
                if (REC && g < M) tr.u[t * M + g] = u;
            }
        }
        PH(0);
        const double zt_now = p_zt;
        if (t + 1 <= N) {
            if (t + 1 < N && g < M) {
                if (K) {
                    const double* row = K + ((long long)(t + 1) * M + g) * 6;
                    pK0 = row[q];
                    if (q < 2) pK1 = row[4 + q];
                }
                if (lu) { p_nu = nu[(t + 1) * M + g]; if (k) p_k = k[(t + 1) * M + g]; }
            }
            if (K && lf) p_nx = nx[(t + 1) * 6 + g];
            if (REC && lu && g < 6) p_zt = ztar[(t + 1) * 6 + g];
        }
        // model at x_t
        PH(1);
        // (rollout: the observer Jacobian goes to a scratch tile instead of a record)
        const double val = ssm_eval_lean<CR>(c, sc, ca, cb, AC, AUX, s_aux, d0, d1,
                                             REC ? tr.H + (long long)t * 36 : ws + W_TILES + 8 * TILE);
        PH(2);
        double e = 0.0;
        if (lu && g < 6) {
            e = REC ? __dsub_rn(__dadd_rn(val, zr), zt_now) : __dadd_rn(val, zr);
            if (REC) VT[g * LD + 2] = e;
            if (REC || tr.e) tr.e[t * 6 + g] = e;
        }
        __syncwarp();   // A_c, aux tile, u, e, du visible
        const double vb0 = VT[q * LD + g], vb1 = VT[(4 + q) * LD + g];   // B fragment of the vector tile
        if (last) {
            if (!REC) break;
            // terminal cost .5 e^T Qf e (ilqr.py:164-166)
            Frag fq{0.0, 0.0};
            dmma(fq, Qft[q * LD + g], vb0);
            dmma(fq, Qft[(4 + q) * LD + g], vb1);
            if (lu) cacc = fma(fq.c0, e, cacc);
            break;
        }
        // A_c x, B u, Q^T e, R^T du                                             (ssm.py:168, 203; ilqr.py:168-175)
        Frag fax{0.0, 0.0}, fbu{0.0, 0.0}, fq{0.0, 0.0}, fr{0.0, 0.0};
        dmma(fax, AC[g * LD + q], vb0);        dmma(fax, AC[g * LD + 4 + q], vb1);
        dmma(fbu, Brt[g * LD + q], vb0);       dmma(fbu, Brt[g * LD + 4 + q], vb1);
        if (REC) {
            dmma(fq, Qt[q * LD + g], vb0);         dmma(fq, Qt[(4 + q) * LD + g], vb1);
            dmma(fr, Rt[q * LD + g], vb0);         dmma(fr, Rt[(4 + q) * LD + g], vb1);
        }
        double dc = 0.0;
        if (lf) {
            // f = r phi + B u,  d_c = (f - A_c x) - B u
            dc = __dsub_rn(__dsub_rn(__dadd_rn(val, fbu.c1), fax.c0), fbu.c1);
            *reinterpret_cast<double2*>(VT2 + g * LD) = make_double2(dc, fbu.c1);
        }
        if (REC && lu) cacc = fma(fr.c1, du, fma(fq.c0, e, cacc));
        double xn = 0.0;
        PH(3);
        if (IMPL) {
            // discretisation (ssm.py:279-301): inv(I - h A_c) and inv(A_c), one row per lane
            {
                const int r = lane & 15;
                const double* src = ((lane & 16) ? AC : IMH) + (r < 6 ? r : 6) * LD;
                double row[6];
                const double2 r01 = *reinterpret_cast<const double2*>(src);
                const double2 r23 = *reinterpret_cast<const double2*>(src + 2);
                const double2 r45 = *reinterpret_cast<const double2*>(src + 4);
                row[0] = r01.x; row[1] = r01.y; row[2] = r23.x; row[3] = r23.y; row[4] = r45.x; row[5] = r45.y;
                gj6_rows(row, lane, (lane & 16) ? IA : (DISCR == SRCB200_DISCR_BE ? AD : W0));
            }
            __syncwarp();
            PH(4);
            if (DISCR == SRCB200_DISCR_BIL) {
                // A_d = (I + h A_c) inv(I - h A_c)
                Frag f{0.0, 0.0};
                const double h = 0.5 * dt;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int kk = 4 * s + q;
                    const double av = (g < 6 && kk < 6) ? __dadd_rn(g == kk ? 1.0 : 0.0, __dmul_rn(h, AC[g * LD + kk])) : 0.0;
                    dmma(f, av, W0[kk * LD + g]);
                }
                store_frag(AD, f, g, q);
                __syncwarp();
            }
            // sep = inv(A_c) (A_d - I)
            Frag s{0.0, 0.0};
            {
                double bv0 = AD[q * LD + g], bv1 = AD[(4 + q) * LD + g];
                if (dg0) bv0 = __dsub_rn(bv0, 1.0);
                if (dg1) bv1 = __dsub_rn(bv1, 1.0);
                dmma(s, IA[g * LD + q], bv0);
                dmma(s, IA[g * LD + 4 + q], bv1);
            }
            store_frag(SP, s, g, q);
            if (REC && g < 6 && q < 3)
                *reinterpret_cast<double2*>(tr.A + (long long)t * 36 + g * 6 + 2 * q) = lds_pair(AD, g, q);
            __syncwarp();
            PH(5);
            // B_d = sep B_r ; d_d = sep d_c ; x_{t+1} = (A_d x + B_d u) + d_d with B_d u = sep (B_r u)   (ssm.py:330-333)
            const double sa0 = SP[g * LD + q], sa1 = SP[g * LD + 4 + q];
            Frag bd{0.0, 0.0}, sv{0.0, 0.0}, ax{0.0, 0.0};
            dmma(bd, sa0, Brt[q * LD + g]);       dmma(bd, sa1, Brt[(4 + q) * LD + g]);
            dmma(sv, sa0, VT2[q * LD + g]);       dmma(sv, sa1, VT2[(4 + q) * LD + g]);
            dmma(ax, AD[g * LD + q], vb0);        dmma(ax, AD[g * LD + 4 + q], vb1);
            if (REC && g < 6 && 2 * q < M)
                *reinterpret_cast<double2*>(tr.B + (long long)t * 6 * M + g * M + 2 * q) = make_double2(bd.c0, bd.c1);
            if (lf) xn = __dadd_rn(__dadd_rn(ax.c0, sv.c1), sv.c0);
        } else {
            // fe: A_d = I + dt A_c (aux tile), B_d = dt B_r, d_d = dt d_c ; discrete model: the maps themselves
            const double sb = (DISCR == SRCB200_DISCR_FE) ? dt : 1.0;
            Frag ax{0.0, 0.0}, bu{0.0, 0.0};
            dmma(ax, AD[g * LD + q], vb0);        dmma(ax, AD[g * LD + 4 + q], vb1);
            dmma(bu, __dmul_rn(sb, Brt[g * LD + q]), vb0);
            dmma(bu, __dmul_rn(sb, Brt[g * LD + 4 + q]), vb1);
            if (REC && g < 6 && q < 3)
                *reinterpret_cast<double2*>(tr.A + (long long)t * 36 + g * 6 + 2 * q) = lds_pair(AD, g, q);
            if (REC && g < 6 && 2 * q < M) {
                const double2 b2 = lds_pair(Brt, g, q);
                *reinterpret_cast<double2*>(tr.B + (long long)t * 6 * M + g * M + 2 * q) = make_double2(__dmul_rn(sb, b2.x), __dmul_rn(sb, b2.y));
            }
            if (lf) xn = __dadd_rn(__dadd_rn(ax.c0, bu.c1), __dmul_rn(sb, dc));
        }
        __syncwarp();   // every lane has read this step's fragments of VT / VT2
        if (lf) {
            X[g] = xn;
            VT[g * LD] = xn;
            VT2[g * LD + 2] = K ? __dsub_rn(xn, p_nx) : 0.0;
            tr.x[(long long)(t + 1) * 6 + g] = xn;
        }
        up = u;
        __syncwarp();
        PH(6);
    }
    PH_FLUSH(0);
    // cost = sum over the input / output lanes of their quadratic terms, halved (every term of ilqr.py:164-175 carries 1/2)
    if (!lu) cacc = 0.0;
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) cacc = __dadd_rn(cacc, __shfl_xor_sync(FULL, cacc, off));
    cacc = __shfl_sync(FULL, cacc, 1);
    return __dmul_rn(0.5, cacc);
}

template <int M, int CR>
__device__ __forceinline__ double fwd_dispatch(const Ctx c, const IlqrArgs& a, int discr, const double* nx, const double* nu,
                                               double alpha, const double* K, const double* k, const Rec tr,
                                               const double* __restrict__ ztar, const double* __restrict__ ulast) {
    switch (discr) {
        case SRCB200_DISCR_BE:  return fwd_lean<M, SRCB200_DISCR_BE, CR>(c, a, nx, nu, alpha, K, k, tr, ztar, ulast);
        case SRCB200_DISCR_BIL: return fwd_lean<M, SRCB200_DISCR_BIL, CR>(c, a, nx, nu, alpha, K, k, tr, ztar, ulast);
        case SRCB200_DISCR_FE:  return fwd_lean<M, SRCB200_DISCR_FE, CR>(c, a, nx, nu, alpha, K, k, tr, ztar, ulast);
        default:                return fwd_lean<M, SRCB200_DISCR_NONE, CR>(c, a, nx, nu, alpha, K, k, tr, ztar, ulast);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward pass (ilqr.py:219-300).  The record of step t-1 is fetched into registers while step t computes.
// ---------------------------------------------------------------------------------------------------------------
struct BwdResult { double rho, drho; int pd_fail; };   // pd_fail: horizon index of the failed PD test, -1: none

#define LOAD_STEP(tt)                                                                                        \
    do {                                                                                                     \
        const long long t36 = (long long)(tt) * 36, tB = (long long)(tt) * 6 * M;                            \
        pa = *reinterpret_cast<const double2*>(rc.A + t36 + oA);                                             \
        ph = *reinterpret_cast<const double2*>(rc.H + t36 + oA);                                             \
        pb = *reinterpret_cast<const double2*>(rc.B + tB + oB);                                              \
        pe = rc.e[(tt) * 6 + (lane < 6 ? lane : 0)];                                                         \
        pu = rc.u[(tt) * M + (lane & (M - 1))];                                                              \
        pup = ((tt) == 0) ? (ulast ? ulast[lane & (M - 1)] : 0.0) : rc.u[((tt) - 1) * M + (lane & (M - 1))]; \
    } while (0)

// Solve Q_uu~ X = (Q_ux~ | Q_u) by an un-pivoted Gauss-Jordan sweep on [Q_uu~ | rhs] (one column per lane, rows in
// registers) and return the PD verdict from the pivots.  X replaces the explicit inv(Q_uu~) @ rhs of ilqr.py:289-292
// (same elimination, the product with the identity block is skipped).  Writes -(X) = (K | k) into `out`.
template <int M>
__device__ __forceinline__ bool gj_solve_spd(const double* __restrict__ quut, const double* __restrict__ rhs,
                                             double* __restrict__ out, double* bc, int lane) {
    double col[M];
    const int j = lane;
    // lanes < M: columns of Q_uu~ ; lanes M .. M+6: columns of the right-hand side ; the rest: its zero column 7
    const double* src = (j < M) ? quut + j : rhs + (j - M < 7 ? j - M : 7);
#pragma unroll
    for (int r = 0; r < M; ++r) col[r] = src[r * LD];
    bool pd = true;
#pragma unroll
    for (int c = 0; c < M; ++c) {
#if SRCB_GAIN_SHFL
        double cc[M];
#pragma unroll
        for (int r = 0; r < M; ++r) cc[r] = __shfl_sync(FULL, col[r], c);
#else
        // column c (the multipliers of this step) goes through shared memory: 4 stores + 4 broadcast loads instead
        // of 16 shuffles; two alternating buffers, so one __syncwarp per step orders everything
        double* buf = bc + 8 * (c & 1);
        if (j == c) {
#pragma unroll
            for (int r = 0; r < M; r += 2) *reinterpret_cast<double2*>(buf + r) = make_double2(col[r], col[r + 1]);
        }
        __syncwarp();
        double cc[M];
#pragma unroll
        for (int r = 0; r < M; r += 2) {
            const double2 v = *reinterpret_cast<const double2*>(buf + r);
            cc[r] = v.x; cc[r + 1] = v.y;
        }
#endif
        const double piv = cc[c];
        pd = pd && (piv > 0.0) && !isinf(piv);
        const double pc = __dmul_rn(col[c], rcp_fast(piv));
#pragma unroll
        for (int r = 0; r < M; ++r) col[r] = (r == c) ? pc : fma(-cc[r], pc, col[r]);
    }
    if (j >= M && j < M + 7) {
#pragma unroll
        for (int r = 0; r < M; ++r) out[r * LD + (j - M)] = -col[r];
#pragma unroll
        for (int r = M; r < 8; ++r) out[r * LD + (j - M)] = 0.0;      // K rows beyond m stay zero in the 8 x 8 tile
    }
    return pd;
}

template <int M>
__device__ __noinline__ BwdResult bwd_fast(const Ctx c, const IlqrArgs& a, const Rec rc, const double* __restrict__ ulast,
                                           double* __restrict__ Kout, double* __restrict__ kout, double* __restrict__ ab,
                                           double rho, double drho) {
    const int lane = c.lane, g = c.g, q = c.q, N = a.N;
    double* ws = CTX_WS(c);
    // tiles.  One backward step is five dependency levels separated by one __syncwarp each; within a level all
    // products are independent, so their fragment loads and DMMAs overlap.
    double* P = ws + W_TILES + 0 * TILE;     // (P | p); between levels 2 and 5 it holds the right-hand side (Q_ux~ | Q_u)
    double* A = ws + W_TILES + 1 * TILE;     // A_t with A[6][6] = 1
    double* B = ws + W_TILES + 2 * TILE;     // B_t
    double* H = ws + W_TILES + 3 * TILE;     // (H_t | e_t)
    double* W = ws + W_TILES + 4 * TILE;     // level 1: (H|e)^T Q ; level 3..5: (K | k)
    double* ATP = ws + W_TILES + 5 * TILE;   // level 1: A'^T (P|p) ; level 4..5: K^T Q_uu
    double* BTP = ws + W_TILES + 6 * TILE;   // B^T (P|p)
    double* BTPR = ws + W_TILES + 7 * TILE;  // B^T (P + rho I)
    double* QUU = ws + W_TILES + 8 * TILE;
    double* QUX = ws + W_TILES + 9 * TILE;   // (Q_ux | Q_u)
    double* QT = ws + W_TILES + 10 * TILE;   // Q_uu~
    double* RHS = P;
    double* KT = W;
    double* KQ = ATP;
    double* CU = ws + W_DC;
    double* DU = ws + W_DD;
    const double* Qt = CTX_SH + SH_Q;  const double* Rt = CTX_SH + SH_R;  const double* Qft = CTX_SH + SH_QF;
    const srcb200_ilqr_config& cf = a.cfg;
    const bool sreg = cf.regularize && cf.state_regularization;
    const bool inc = cf.include_input_var_constraint != 0;
    // records move as 16-byte pieces in C-fragment coordinates: lane (g, q) owns columns 2q, 2q+1 of row g
    const bool vA = (g < 6 && q < 3), vB = (g < 6 && 2 * q < M), vK = (g < M && q < 3);
    const int oA = vA ? g * 6 + 2 * q : 0, oB = vB ? g * M + 2 * q : 0;
    double* BC = ws + W_PV;                            // 2 x 2 x 8 doubles: pivot-column broadcast of the gain solve
    int pd_fail = -1;
    double2 pa, ph, pb;                                // record of the next step to process, in registers
    double pe, pu, pup;

    {
        for (int t = 0; t < NTILES; ++t) zero_tile(ws + W_TILES + t * TILE, lane);
        __syncwarp();
        // terminal: (P | p) = (H^T Qf) (H | e)                                   (ilqr.py:177-182)
        H[c.off0] = rc.H[(long long)N * 36 + lane];
        if (lane < 4) H[c.off1] = rc.H[(long long)N * 36 + 32 + lane];
        if (lane < 6) H[lane * LD + 6] = rc.e[N * 6 + lane];
        LOAD_STEP(N - 1);
        __syncwarp();
        Frag pf{0.0, 0.0};
        {
            Frag f{0.0, 0.0};
            mma88<true, false>(f, H, Qft, g, q);
            store_frag(W, f, g, q);
            __syncwarp();
            mma88<false, false>(pf, W, H, g, q);
            if (g >= 6) { pf.c0 = 0.0; pf.c1 = 0.0; }
            if (q == 3) pf.c1 = 0.0;
            store_frag(P, pf, g, q);
        }
        if (lane == 0) A[6 * LD + 6] = 1.0;
        __syncwarp();

        PH_DECL;
        const Frag rfrag = load_frag(Rt, g, q);      // R in C-fragment coordinates: constant over the sweep
        for (int t = N - 1; t >= 0; --t) {
            // ---- level 0: stage A_t, B_t, (H_t | e_t), du from the prefetched registers; fetch step t-1
            if (vA) {
                sts_pair(A, g, q, pa.x, pa.y);
                sts_pair(H, g, q, ph.x, ph.y);
            }
            if (vB) sts_pair(B, g, q, pb.x, pb.y);
            if (lane < 6) H[lane * LD + 6] = pe;   // rows 6,7 / column 7 of this tile never reach rows<6, cols<7
            if (lane < M) DU[lane] = inc ? __dsub_rn(pu, pup) : pu;
            if (t > 0) LOAD_STEP(t - 1);
            __syncwarp();
            PH(8);
            // ---- level 1: W = (H|e)^T Q, A'^T (P|p), B^T (P|p), B^T (P + rho I), c_u = R du
            {
                Frag w{0.0, 0.0}, atp{0.0, 0.0}, btp{0.0, 0.0}, btpr{0.0, 0.0};
                mma88<true, false>(w, H, Qt, g, q);
                mma88<true, false>(atp, A, P, g, q);
                mma88<true, false>(btp, B, P, g, q);
                if (sreg) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const int kk = 4 * s + q;
                        double pv = P[kk * LD + g];
                        if (kk == g && kk < 6) pv = __dadd_rn(pv, rho);     // P + rho I (ilqr.py:266-267)
                        dmma(btpr, B[kk * LD + g], pv);
                    }
                }
                if (lane >= 8 && lane < 8 + M) {
                    const int i = lane - 8;
                    double acc = 0.0;
#pragma unroll
                    for (int jj = 0; jj < M; ++jj) acc = fma(Rt[i * LD + jj], DU[jj], acc);
                    CU[i] = acc;                                                 // c_u = R du
                }
                store_frag(W, w, g, q);
                store_frag(ATP, atp, g, q);
                store_frag(BTP, btp, g, q);
                if (sreg) store_frag(BTPR, btpr, g, q);
            }
            __syncwarp();
            PH(9);
            // ---- level 2: (Q_xx|Q_x), Q_uu, (Q_ux|Q_u), Q_uu~, Q_ux~                 (ilqr.py:258-274)
            Frag qxx{0.0, 0.0};
            {
                mma88<false, false>(qxx, W, H, g, q);                  // (c_xx | c_x) = W (H | e)
                mma88<false, false>(qxx, ATP, A, g, q);                // + (A^T P | A^T p) A'
                Frag quu = rfrag;
                mma88<false, false>(quu, BTP, B, g, q);                // R + (B^T P) B
                Frag qux{(q == 3 && g < M) ? CU[g] : 0.0, 0.0};
                mma88<false, false>(qux, BTP, A, g, q);                // (0 | c_u) + (B^T P | B^T p) A'
                Frag quut, quxt;
                if (sreg) {
                    quut = rfrag;
                    mma88<false, false>(quut, BTPR, B, g, q);
                    quxt = Frag{0.0, 0.0};
                    mma88<false, false>(quxt, BTPR, A, g, q);
                } else {
                    quut = quu;
                    quxt = qux;
                    if (cf.regularize) {
                        if (2 * q == g) quut.c0 = __dadd_rn(quut.c0, rho);
                        if (2 * q + 1 == g) quut.c1 = __dadd_rn(quut.c1, rho);
                    }
                }
                if (g >= M) {   // keep the padding block of Q_uu~ an identity so the sweep is well defined
                    quut.c0 = (2 * q == g) ? 1.0 : 0.0;
                    quut.c1 = (2 * q + 1 == g) ? 1.0 : 0.0;
                }
                if (q == 3) { quxt.c0 = qux.c0; quxt.c1 = 0.0; }       // right-hand side (Q_ux~ | Q_u)
                store_frag(QUU, quu, g, q);
                store_frag(QUX, qux, g, q);
                store_frag(QT, quut, g, q);
                store_frag(RHS, quxt, g, q);                           // P is dead until level 5
            }
            __syncwarp();
            PH(10);
            // ---- level 3: PD test + gains (K | k) = -Q_uu~^-1 (Q_ux~ | Q_u)         (ilqr.py:276-292)
            const bool pd = gj_solve_spd<M>(QT, RHS, KT, BC, lane);
            if (!pd && pd_fail < 0) pd_fail = t;
            if (!pd && cf.regularize) {
                // ilqr.py:282-287: raise rho and LEAVE the sweep -- the reference then falls through to the decrease
                // of ilqr.py:298 and returns; there is no restart.  K_s = k_s = 0 for every s <= t (their zero
                // initialisation), hence zero line-search scalars (evaluated literally at t: 0 * inf = nan like numpy).
                rho_update(cf, true, rho, drho);
                double za = 0.0, zb = 0.0;
                if (lane < M) {
                    za = __dmul_rn(0.0, QUX[lane * LD + 6]);
                    double v = 0.0;
#pragma unroll
                    for (int i = 0; i < M; ++i) v = __dadd_rn(v, __dmul_rn(0.0, QUU[i * LD + lane]));
                    zb = __dmul_rn(v, 0.0);
                }
#pragma unroll
                for (int off = 1; off < 8; off <<= 1) {
                    za = __dadd_rn(za, __shfl_xor_sync(FULL, za, off));
                    zb = __dadd_rn(zb, __shfl_xor_sync(FULL, zb, off));
                }
                for (int e = lane; e < (t + 1) * M * 6; e += 32) Kout[e] = 0.0;
                for (int e = lane; e < (t + 1) * M; e += 32) kout[e] = 0.0;
                for (int e = lane; e < 2 * t; e += 32) ab[e] = 0.0;
                if (lane == 0) { ab[2 * t] = za; ab[2 * t + 1] = zb; }
                break;
            }
            __syncwarp();
            PH(11);
            // ---- level 4: K^T Q_uu ; gains to global
            {
                Frag kq{0.0, 0.0};
                mma88<true, false>(kq, KT, QUU, g, q);
                if (vK) *reinterpret_cast<double2*>(Kout + (long long)t * M * 6 + g * 6 + 2 * q) = lds_pair(KT, g, q);
                if (lane < M) kout[t * M + lane] = KT[lane * LD + 6];
                store_frag(KQ, kq, g, q);
            }
            __syncwarp();
            PH(12);
            // ---- level 5: (P | p) = (((Q_xx|Q_x) + KQ (K|k)) + K^T (Q_ux|Q_u)) + Q_ux^T (K|k)   (ilqr.py:294-295)
            mma88<false, false>(qxx, KQ, KT, g, q);
            mma88<true, false>(qxx, KT, QUX, g, q);
            mma88<true, false>(qxx, QUX, KT, g, q);
            // line-search scalars: a_t = k . Q_u, b_t = (k^T Q_uu) . k              (ilqr.py:69-71)
            double pa = 0.0, pb = 0.0;
            if (lane < M) {
                const double kv = KT[lane * LD + 6];
                pa = __dmul_rn(kv, QUX[lane * LD + 6]);
                pb = __dmul_rn(KQ[6 * LD + lane], kv);
            }
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) {
                pa = __dadd_rn(pa, __shfl_xor_sync(FULL, pa, off));
                pb = __dadd_rn(pb, __shfl_xor_sync(FULL, pb, off));
            }
            if (lane == 0) { ab[2 * t] = pa; ab[2 * t + 1] = pb; }
            if (g >= 6) { qxx.c0 = 0.0; qxx.c1 = 0.0; }
            if (q == 3) qxx.c1 = 0.0;
            __syncwarp();                       // every lane has read the right-hand side held in P's tile
            store_frag(P, qxx, g, q);
            PH(13);
            // the (K | k) tile is W next step: its column 7 must be zero again, rows >= M too (they are: K rows >= M = 0)
        }
        PH_FLUSH(8);
        rho_update(cf, false, rho, drho);         // ilqr.py:298 -- after a complete AND after an interrupted sweep
    }
    return BwdResult{rho, drho, pd_fail};
}

// ---------------------------------------------------------------------------------------------------------------
// Solve kernel (ilqr.py:27-107): one warp per problem AND per iteration.  A task is "the next iteration of problem
// b"; after it the warp saves the solver state (rho, drho, cost, counters: 8 doubles in the problem's scratch) and
// puts b back at the end of a task queue unless the solve finished, then takes the oldest waiting task of the
// highest non-empty priority class.  The
// iteration counts of a batch are very uneven (5 .. 50); handing out whole problems leaves most of the GPU idle
// while the last long solves finish, handing out iterations round-robin keeps every warp busy until the total work
// is done.  Everything a task needs lives in global memory (records, gains), so any warp on any SM can resume it.
//
// A problem is in a queue at most once, so at most `batch` ring slots are occupied; cap = 2 * batch + kIlqrQueueWaiters.
// ---------------------------------------------------------------------------------------------------------------
// Two priority classes, one ticket ring each.  Iteration counts correlate with the cost of the initial rollout, and a
// batch finishes earliest when its long solves never wait ("longest first"): after its first task a problem whose
// initial cost is above the running mean of the batch goes to the HIGH ring, the others to the LOW ring, and a free
// warp serves HIGH first.  (Replaying the measured traces of the bench workload: FIFO 134 pass-times, this rule 119,
// an oracle that knows every length 118.)  Scheduling only: results do not depend on it.
//
// The queue primitives (counters, rings, pop / push of one id by one thread) are shared with the generic kernel:
// ilqr.cuh, namespace ilqrq.
using namespace ilqrq;

__device__ __forceinline__ int queue_pop(int* q, int cap, int lane, int stop_at) {
    int id = -1;
    if (lane == 0) id = pop_one(q, cap, stop_at);
    id = __shfl_sync(FULL, id, 0);
    __threadfence();            // acquire: what the previous owner of this problem wrote is visible (L1 invalidated)
    return id;
}

__device__ __forceinline__ void queue_push(int* q, int cap, int lane, int id, int cls) {
    __threadfence();            // release: this warp's records / gains / state before the id becomes visible
    __syncwarp();
    if (lane == 0) push_one(q, cap, id, cls);
}

// launch shape S: 0 = 2 CTAs x 8 warps, table in shared memory; 1 = 12 warps, one row in registers; 2 = 8 warps, two rows
constexpr int shape_warps(int s) { return s == 1 ? 12 : 8; }
template <int M, int S>
__global__ void __launch_bounds__(shape_warps(S) * 32, S == 0 ? 2 : 1)
ilqr_ssm_fast_kernel(const __grid_constant__ SsmDev Mdl, const __grid_constant__ IlqrArgs a) {
    build_tables(Mdl, a.Q, a.R, a.Qf, g_sm, M);
    Ctx c;
    c.lane = threadIdx.x & 31;
    c.g = c.lane >> 2;
    c.q = c.lane & 3;
    c.off0 = (c.lane / 6) * LD + c.lane % 6;
    c.off1 = ((32 + c.lane) / 6) * LD + (32 + c.lane) % 6;
    const int warp = threadIdx.x >> 5;
    c.ws_off = SH_END + warp * W_SIZE;
    const int lane = c.lane, N = a.N;
    const srcb200_ilqr_config& cf = a.cfg;
    const int discr = Mdl.discr;
    // a warp may resume a problem with the backward pass before it ever ran a forward pass: start from a clean work area
    for (int e = lane; e < W_SIZE; e += 32) CTX_WS(c)[e] = 0.0;
    __syncwarp();

    while (true) {
        const long long b = queue_pop(a.work_counter, a.queue_cap, lane, a.stop_at);
        if (b < 0) break;
        double* wsb = a.ws + b * a.L.total;
        double* kbuf = wsb + a.L.k;
        double* ab = wsb + a.L.ab;
        double* Kbuf = a.oK + b * (long long)N * M * 6;
        const double* ztar = a.z_target + (a.shared_target ? 0 : b * (long long)(N + 1) * 6);
        const double* ulast = a.u_last ? a.u_last + b * M : nullptr;
        double* trace = a.otrace ? a.otrace + b * (long long)(cf.max_iter + 1) * 4 : nullptr;
        double* sv = wsb + a.L.state;           // [rho, drho, cost, (fails, cur), (status, trials), (it, started)]
        int* svi = reinterpret_cast<int*>(sv + 3);

        double rho, drho, cost;
        int fails, cur, status, trials, it, cls;
        if (svi[5] != 0x5ca1ab1e) {
            // first task of this problem: nominal rollout (ilqr.py:38-40)
            rho = cf.rho0; drho = cf.drho0;
            fails = 0; cur = 0; status = 0; trials = 0; it = 0;
            const Rec nom = rec_at(wsb + a.L.rec, a.L), tr0 = rec_at(wsb, a.L);
            if (lane < 6) nom.x[lane] = a.x0[b * 6 + lane];
            for (int e = lane; e < N * M; e += 32) nom.u[e] = a.u_init ? a.u_init[b * (long long)N * M + e] : 0.0;
            __syncwarp();
            __threadfence_block();
            cost = fwd_dispatch<M, S>(c, a, discr, nom.x, nom.u, 1.0, nullptr, nullptr, tr0, ztar, ulast);
            if (a.ocost0 && lane == 0) a.ocost0[b] = cost;
            // priority class: initial cost above the running mean of the batch -> expected to need many iterations
            cls = 1;
            if (lane == 0) cls = classify(a.work_counter, cost, a.prio_frac);
            cls = __shfl_sync(FULL, cls, 0);
        } else {
            rho = sv[0]; drho = sv[1]; cost = sv[2];
            fails = svi[0]; cur = svi[1]; status = svi[2]; trials = svi[3]; it = svi[4]; cls = svi[6];
        }

        bool conv = false, stop = false;
        if (it <= cf.max_iter) {                // one pass of the `while not converged and nbr_iter <= max_iter` loop
            const Rec rcur = rec_at(wsb + (cur ? a.L.rec : 0), a.L), rtrial = rec_at(wsb + (cur ? 0 : a.L.rec), a.L);
            const BwdResult br = bwd_fast<M>(c, a, rcur, ulast, Kbuf, kbuf, ab, rho, drho);
            rho = br.rho; drho = br.drho;
            const int pd_fail = br.pd_fail;
            const double rho_bwd = rho;
            if (pd_fail >= 0) status |= SRCB200_ILQR_ST_NONPD;
            {
                __syncwarp();
                const double prev_cost = cost;
                double alpha = cf.alpha0;
                bool improved = false, failed = false;
                double cost_t = cost, alpha_acc = 0.0;
                while (!improved && !failed) {
                    improved = true;
                    cost_t = fwd_dispatch<M, S>(c, a, discr, rcur.x, rcur.u, alpha, Kbuf, kbuf, rtrial, ztar, ulast);
                    ++trials;
                    double dc = 0.0;
                    const double a2 = __dmul_rn(__dmul_rn(alpha, alpha), 0.5);
                    for (int t = 0; t < N; ++t)
                        dc = __dadd_rn(dc, __dadd_rn(__dmul_rn(alpha, ab[2 * t]), __dmul_rn(a2, ab[2 * t + 1])));
                    alpha_acc = alpha;
                    if (cf.do_linesearch) {
                        const double ratio = __ddiv_rn(__dsub_rn(cost_t, prev_cost), dc);
                        if (ratio <= cf.improv_lb || ratio > cf.improv_ub) {
                            alpha = __dmul_rn(cf.alpha_scaling, alpha);
                            improved = false;
                            if (alpha < cf.alpha_min) {
                                rho_update(cf, true, rho, drho);
                                rho = __dadd_rn(rho, cf.rho_increase_fp);
                                failed = true;
                            }
                        }
                    }
                }
                if (!failed) {
                    cur ^= 1;
                    cost = cost_t;
                    const double dJ = __dsub_rn(prev_cost, cost);
                    conv = (dJ < cf.epsilon) && (dJ >= 0.0);
                    if (conv) status |= SRCB200_ILQR_ST_CONVERGED;
                    fails = 0;
                } else {
                    ++fails;
                    if (fails >= cf.counter_limit) { conv = true; status |= SRCB200_ILQR_ST_ABANDONED; }
                }
                if (trace && lane == 0) {
                    trace[it * 4 + 0] = cost;
                    trace[it * 4 + 1] = failed ? 0.0 : alpha_acc;
                    trace[it * 4 + 2] = rho_bwd;
#ifdef SRCB_PHASE_TIMING
                    { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); trace[it * 4 + 3] = (double)gt; }   // timeline study
#else
                    trace[it * 4 + 3] = (double)pd_fail;
#endif
                }
                ++it;
                if (!isfinite(cost)) { status |= SRCB200_ILQR_ST_NONFINITE; stop = true; }
            }
        }
        if (!conv && !stop && it <= cf.max_iter) {
            // not finished: save the state and hand the problem to whichever warp is free next
            __syncwarp();
            if (lane == 0) {
                sv[0] = rho; sv[1] = drho; sv[2] = cost;
                svi[0] = fails; svi[1] = cur; svi[2] = status; svi[3] = trials; svi[4] = it; svi[5] = 0x5ca1ab1e; svi[6] = cls;
            }
            queue_push(a.work_counter, a.queue_cap, lane, (int)b, cls);
            continue;
        }
        if (!conv && !stop && it > cf.max_iter) status |= SRCB200_ILQR_ST_MAXITER;

        __syncwarp();
        const Rec fin = rec_at(wsb + (cur ? a.L.rec : 0), a.L);
        for (int e = lane; e < (N + 1) * 6; e += 32) a.ox[b * (long long)(N + 1) * 6 + e] = fin.x[e];
        for (int e = lane; e < N * M; e += 32) a.ou[b * (long long)N * M + e] = fin.u[e];
        if (lane == 0) {
            a.ocost[b] = cost;
            if (a.orho) a.orho[b] = rho;
            a.oiter[b] = it;
            a.ostatus[b] = status;
            if (a.otrials) a.otrials[b] = trials;
            __threadfence();
            atomicSub(a.work_counter + Q_REMAINING, 1);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Open-loop rollout (ssm.py:134-156) for the same model shape: one warp per trajectory, re-linearised every step.
// ---------------------------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(WARPS * 32, 2)
ssm_rollout_fast_kernel(const __grid_constant__ SsmDev Mdl, const __grid_constant__ IlqrArgs a, long long batch,
                        const double* __restrict__ x0, const double* __restrict__ u, double* __restrict__ xo,
                        double* __restrict__ zo) {
    build_tables(Mdl, nullptr, nullptr, nullptr, g_sm, M);
    Ctx c;
    c.lane = threadIdx.x & 31;
    c.g = c.lane >> 2;
    c.q = c.lane & 3;
    c.off0 = (c.lane / 6) * LD + c.lane % 6;
    c.off1 = ((32 + c.lane) / 6) * LD + (32 + c.lane) % 6;
    c.ws_off = SH_END + (threadIdx.x >> 5) * W_SIZE;
    const int N = a.N, discr = Mdl.discr;
    // every trajectory costs the same: static striding is perfectly balanced, no work counter needed.  The step is the
    // iLQR kernel's lean forward step without cost, gains and record (fwd_lean<.., REC = false>).
    for (long long b = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); b < batch; b += (long long)gridDim.x * WARPS) {
        Rec tr;
        tr.x = xo + b * (long long)(N + 1) * 6;
        tr.e = zo ? zo + b * (long long)(N + 1) * 6 : nullptr;
        tr.u = nullptr; tr.H = nullptr; tr.A = nullptr; tr.B = nullptr; tr.idx = nullptr;
        const double* xb = x0 + b * 6;
        const double* ub = u + b * (long long)N * M;
        switch (discr) {
            case SRCB200_DISCR_BE:  fwd_lean<M, SRCB200_DISCR_BE, 0, false>(c, a, xb, ub, 1.0, nullptr, nullptr, tr, nullptr, nullptr); break;
            case SRCB200_DISCR_BIL: fwd_lean<M, SRCB200_DISCR_BIL, 0, false>(c, a, xb, ub, 1.0, nullptr, nullptr, tr, nullptr, nullptr); break;
            case SRCB200_DISCR_FE:  fwd_lean<M, SRCB200_DISCR_FE, 0, false>(c, a, xb, ub, 1.0, nullptr, nullptr, tr, nullptr, nullptr); break;
            default:                fwd_lean<M, SRCB200_DISCR_NONE, 0, false>(c, a, xb, ub, 1.0, nullptr, nullptr, tr, nullptr, nullptr); break;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Batched evaluation + linearisation on the FP64 tensor pipe (kernel "b" of BASELINE.json): per state ONE dense
// contraction   [r_coeff; w_coeff] (16 x 84, zero padded)  x  [phi | dphi/dx_1..6 | 0] (84 x 8)   = 42 DMMA tiles,
// coefficient fragments resident in registers for the whole launch, the monomial / derivative fragments generated
// on the fly from x.  Output tile: rows 0..5 (f | A_c), rows 8..13 (z | H).  A warp walks states grid-stride.
// ---------------------------------------------------------------------------------------------------------------
constexpr int EV_WARPS = 4;
constexpr int EV_KS = 21;   // k4 steps (84 / 4)

template <int M>
__global__ void __launch_bounds__(EV_WARPS * 32, 4)
ssm_eval_dmma_kernel(const __grid_constant__ SsmDev Mdl, long long count, const double* __restrict__ xg,
                     const double* __restrict__ ug, double dt, double* __restrict__ Ao, double* __restrict__ Bo,
                     double* __restrict__ dout, double* __restrict__ Ho, double* __restrict__ co, double* __restrict__ zo) {
    __shared__ double s_xe[EV_WARPS][8];
    __shared__ double s_u[EV_WARPS][8];
    __shared__ double s_br[6 * 8];
    __shared__ __align__(16) double s_tiles[EV_WARPS][6 * TILE];   // A_c, A_d, inv(A_c), sep, B_d, W0 (be / bil only)
    __shared__ double s_dc[EV_WARPS][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int nf = Mdl.nfeat, discr = Mdl.discr;
    for (int e = threadIdx.x; e < 6 * M; e += blockDim.x) s_br[e] = Mdl.B[e];
    // ---- coefficient fragments a[mt][s][lane] = C_mt[g][4 s + q], shared by the warps of the CTA (conflict-free LDS)
    __shared__ double s_af[2][EV_KS][32];
    for (int e = threadIdx.x; e < EV_KS * 32; e += blockDim.x) {
        const int s = e >> 5, l = e & 31, gg = l >> 2, k = 4 * s + (l & 3);
        const bool ok = (gg < 6) && (k < nf);
        s_af[0][s][l] = ok ? Mdl.r[gg * nf + k] : 0.0;
        s_af[1][s][l] = ok ? Mdl.w[gg * nf + k] : 0.0;
    }
    // ---- recipes of the B fragment: entry(k = 4 s + q, col = g) = mult * xe[a] * xe[b] * xe[c]   (xe[7] = 1)
    unsigned rec[EV_KS];
#pragma unroll
    for (int s = 0; s < EV_KS; ++s) {
        const int k = 4 * s + q;
        unsigned r = 0;   // mult = 0
        if (k < nf && g < 7) {
            int idx[3];
            for (int t = 0; t < 3; ++t) { const int v = Mdl.mono[k * SRCB200_SSM_MAX_ORDER + t]; idx[t] = (v == 0xFF) ? 7 : v; }
            if (g == 0) {
                r = (unsigned)idx[0] | ((unsigned)idx[1] << 3) | ((unsigned)idx[2] << 6) | (1u << 9);
            } else {
                const int var = g - 1;
                int mult = 0, rest[3] = {7, 7, 7}, nr = 0;
                bool removed = false;
                for (int t = 0; t < 3; ++t) {
                    if (idx[t] == var) ++mult;
                    if (idx[t] == var && !removed) { removed = true; continue; }
                    rest[nr++] = idx[t];
                }
                if (mult > 0) r = (unsigned)rest[0] | ((unsigned)rest[1] << 3) | ((unsigned)rest[2] << 6) | ((unsigned)mult << 9);
            }
        }
        rec[s] = r;
    }
    __syncthreads();
    double* xe = s_xe[warp];
    double* su = s_u[warp];
    double* AC = s_tiles[warp];          double* AD = AC + TILE;       double* IA = AC + 2 * TILE;
    double* SP = AC + 3 * TILE;          double* W0 = AC + 5 * TILE;
    double* DC = s_dc[warp];
    const bool want_dyn = (Ao || Bo || dout);
    const bool need_inv = want_dyn && dt >= 0.0 && (discr == SRCB200_DISCR_BE || discr == SRCB200_DISCR_BIL);
    for (int t = 0; t < 6; ++t) zero_tile(AC + t * TILE, lane);
    const long long nwarps = (long long)gridDim.x * EV_WARPS;
    const long long st0 = (long long)blockIdx.x * EV_WARPS + warp;
    double px = (st0 < count && lane < 6) ? xg[st0 * 6 + lane] : 0.0;            // state of the next iteration, prefetched
    double pu = (st0 < count && ug && lane < M) ? ug[st0 * M + lane] : 0.0;
    for (long long st = st0; st < count; st += nwarps) {
        asm volatile("" ::: "memory");   // keep the coefficient fragments in shared memory (no hoisting into registers)
        if (lane < 8) {
            xe[lane] = lane < 6 ? px : (lane == 7 ? 1.0 : 0.0);
            su[lane] = pu;
        }
        {
            const long long nx = st + nwarps;
            px = (nx < count && lane < 6) ? xg[nx * 6 + lane] : 0.0;
            pu = (nx < count && ug && lane < M) ? ug[nx * M + lane] : 0.0;
        }
        __syncwarp();
        // three independent accumulation chains per output tile (the DMMA latency would otherwise serialise the
        // 21 k-steps), summed at the end
        Frag p0[3] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}}, p1[3] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int s = 0; s < EV_KS; ++s) {
            const unsigned r = rec[s];
            double b = xe[r & 7];
            b = __dmul_rn(b, xe[(r >> 3) & 7]);
            b = __dmul_rn(b, xe[(r >> 6) & 7]);
            b = __dmul_rn(b, (double)(r >> 9));
            dmma(p0[s % 3], s_af[0][s][lane], b);
            dmma(p1[s % 3], s_af[1][s][lane], b);
        }
        Frag c0{__dadd_rn(__dadd_rn(p0[0].c0, p0[1].c0), p0[2].c0), __dadd_rn(__dadd_rn(p0[0].c1, p0[1].c1), p0[2].c1)};
        Frag c1{__dadd_rn(__dadd_rn(p1[0].c0, p1[1].c0), p1[2].c0), __dadd_rn(__dadd_rn(p1[0].c1, p1[1].c1), p1[2].c1)};
        // ---- epilogue.  lane (g, q) holds columns 2q, 2q+1 of row g: col 0 = value, col 1 + j = d/dx_j
        // observation outputs
        if (g < 6) {
            if (Ho) {
                if (q == 0) Ho[st * 36 + g * 6 + 0] = c1.c1;
                else {
                    Ho[st * 36 + g * 6 + 2 * q - 1] = c1.c0;
                    if (q < 3) Ho[st * 36 + g * 6 + 2 * q] = c1.c1;
                }
            }
        }
        if (zo || co) {
            // H x : partial sums inside the quad
            double hx = 0.0;
            if (q == 0) hx = __dmul_rn(c1.c1, xe[0]);
            else { hx = __dmul_rn(c1.c0, xe[2 * q - 1]); if (q < 3) hx = fma(c1.c1, xe[2 * q], hx); }
            hx += __shfl_xor_sync(FULL, hx, 1);
            hx += __shfl_xor_sync(FULL, hx, 2);
            if (g < 6 && q == 0) {
                if (zo) zo[st * 6 + g] = __dadd_rn(c1.c0, Mdl.zref[g]);
                if (co) co[st * 6 + g] = __dsub_rn(c1.c0, hx);
            }
        }
        if (want_dyn) {
            // f = r phi + B u ; d_c = (f - A_c x) - B u : quad-level partial sums
            double ax = 0.0, bu = 0.0;
            if (q == 0) ax = __dmul_rn(c0.c1, xe[0]);
            else { ax = __dmul_rn(c0.c0, xe[2 * q - 1]); if (q < 3) ax = fma(c0.c1, xe[2 * q], ax); }
            if (g < 6) {
#pragma unroll
                for (int jj = 0; jj < M / 4; ++jj) bu = fma(s_br[g * M + q * (M / 4) + jj], su[q * (M / 4) + jj], bu);
            }
            ax += __shfl_xor_sync(FULL, ax, 1);  ax += __shfl_xor_sync(FULL, ax, 2);
            bu += __shfl_xor_sync(FULL, bu, 1);  bu += __shfl_xor_sync(FULL, bu, 2);
            const double fpoly = __shfl_sync(FULL, c0.c0, lane & ~3);     // column 0 lives in lane q = 0 of the quad
            const double dc = __dsub_rn(__dsub_rn(__dadd_rn(fpoly, bu), ax), bu);
            const bool cont = (dt < 0.0) || discr == SRCB200_DISCR_NONE;
            if (cont || discr == SRCB200_DISCR_FE) {
                const double sc = cont ? 1.0 : dt;
                if (g < 6) {
                    // A (continuous, or I + dt A_c), straight from the accumulators
                    auto put = [&](int col, double v) {
                        double o = cont ? v : __dmul_rn(dt, v);
                        if (!cont && col == g) o = __dadd_rn(1.0, o);
                        if (Ao) Ao[st * 36 + g * 6 + col] = o;
                    };
                    if (q == 0) put(0, c0.c1);
                    else { put(2 * q - 1, c0.c0); if (q < 3) put(2 * q, c0.c1); }
                    if (Bo) {
#pragma unroll
                        for (int jj = 0; jj < M / 4; ++jj) {
                            const int col = q * (M / 4) + jj;
                            Bo[st * 6 * M + g * M + col] = cont ? s_br[g * M + col] : __dmul_rn(sc, s_br[g * M + col]);
                        }
                    }
                    if (dout && q == 0) dout[st * 6 + g] = cont ? dc : __dmul_rn(dt, dc);
                }
            } else if (need_inv) {
                // be / bil: A_c to its tile, the two 6x6 inverses by the shuffle sweep, products on the tensor pipe
                if (g < 6) {
                    if (q == 0) AC[g * LD + 0] = c0.c1;
                    else { AC[g * LD + 2 * q - 1] = c0.c0; if (q < 3) AC[g * LD + 2 * q] = c0.c1; }
                    if (q == 0) DC[g] = dc;
                }
                __syncwarp();
                const double h = (discr == SRCB200_DISCR_BE) ? dt : 0.5 * dt;
                const int hm = lane >> 4, j = lane & 15;
                double col[6];
#pragma unroll
                for (int r2 = 0; r2 < 6; ++r2) {
                    double v = 0.0;
                    if (j < 6) {
                        const double av = AC[r2 * LD + j];
                        v = hm ? av : __dsub_rn(r2 == j ? 1.0 : 0.0, __dmul_rn(h, av));
                    } else if (j < 12) {
                        v = (r2 == j - 6) ? 1.0 : 0.0;
                    }
                    col[r2] = v;
                }
                gj6_pair(col, lane, hm ? IA : (discr == SRCB200_DISCR_BE ? AD : W0));
                __syncwarp();
                if (discr == SRCB200_DISCR_BIL) {
                    Frag f{0.0, 0.0};
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2) {
                        const int kk = 4 * s2 + q;
                        const double av = (g < 6 && kk < 6) ? __dadd_rn(g == kk ? 1.0 : 0.0, __dmul_rn(h, AC[g * LD + kk])) : 0.0;
                        dmma(f, av, W0[kk * LD + g]);
                    }
                    store_frag(AD, f, g, q);
                    __syncwarp();
                }
                Frag sp{0.0, 0.0};
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int kk = 4 * s2 + q;
                    const double bv = (kk < 6 && g < 6) ? __dsub_rn(AD[kk * LD + g], kk == g ? 1.0 : 0.0) : 0.0;
                    dmma(sp, IA[g * LD + kk], bv);
                }
                store_frag(SP, sp, g, q);
                __syncwarp();
                if (Ao) {
                    for (int e2 = lane; e2 < 36; e2 += 32) Ao[st * 36 + e2] = AD[(e2 / 6) * LD + e2 % 6];
                }
                if (Bo) {
                    for (int e2 = lane; e2 < 6 * M; e2 += 32) {
                        const int i = e2 / M, jj = e2 % M;
                        double acc = 0.0;
#pragma unroll
                        for (int kk = 0; kk < 6; ++kk) acc = fma(SP[i * LD + kk], s_br[kk * M + jj], acc);
                        Bo[st * 6 * M + e2] = acc;
                    }
                }
                if (dout && lane < 6) {
                    double acc = 0.0;
#pragma unroll
                    for (int kk = 0; kk < 6; ++kk) acc = fma(SP[lane * LD + kk], DC[kk], acc);
                    dout[st * 6 + lane] = acc;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Batched evaluation + linearisation, second generation (continuous / fe / discrete outputs): EIGHT STATES PER WARP
// as the M dimension of the DMMA tiles, and the SPARSE structure of the contraction.  d phi / d x_j of a degree-<=3
// monomial is a multiple of an entry of psi = (x, x (x) x) (8 + 24 padded slots), the same psi for every j, so
//     J_j (8 states x 16 outputs)  =  Psi (8 x 32)  x  T_j (32 x 16),     j = 0..5,
// with T_j the multiplicity-scaled coefficients of the monomials containing x_j (rows of r_coeff for outputs 0..5,
// of w_coeff for 6..11): 6 x 8 k-steps x 2 n-tiles = 96 DMMAs per EIGHT states (12 per state; the dense kernel above
// issues 42 per state, 35 % of them padding).  The Psi fragment is built once per group and reused by all 96; the
// T_j fragments are read conflict-free from a 24 KB table.  Degree-2 and degree-3 parts accumulate separately, which
// gives the VALUES by Euler's theorem (x . grad h = p h for a homogeneous h of degree p) without any extra
// contraction.  A lane ends up owning whole rows (over j) of A_c / H of its state, so f, z, d_c = (f - A x) - B u and
// c = C(x) - H x are in-lane dot products and the rows leave as 16-byte stores.
// ---------------------------------------------------------------------------------------------------------------
constexpr int E2_WARPS = 4;
constexpr int E2_TAB = 6 * 2 * 8 * 32;            // T_j fragments: [j][nt][ks][lane]
constexpr int E2_G1 = 6 * 2 * 32 * 2;             // degree-1 constants in C-fragment layout: [j][nt][lane] x 2
constexpr size_t E2_SMEM = sizeof(double) * (SH_END + E2_TAB + E2_G1 + E2_WARPS * (8 * 8 + 8 * 8));

template <int M>
__global__ void __launch_bounds__(E2_WARPS * 32, 3)
ssm_eval_sparse_dmma_kernel(const __grid_constant__ SsmDev Mdl, long long count, const double* __restrict__ xg,
                            const double* __restrict__ ug, double dt, int fe, double* __restrict__ Ao,
                            double* __restrict__ Bo, double* __restrict__ dout, double* __restrict__ Ho,
                            double* __restrict__ co, double* __restrict__ zo) {
    build_tables(Mdl, nullptr, nullptr, nullptr, g_sm, M);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    double* TAB = g_sm + SH_END;
    double* G1 = TAB + E2_TAB;
    double* XE = G1 + E2_G1 + warp * 128;          // 8 states x 8: x_0..x_5, 0, 0
    double* UE = XE + 64;                          // 8 states x 8 inputs
    const double* T = g_sm + SH_T;
    const double* Brt = g_sm + SH_BR;
    const double* zref = g_sm + SH_ZREF;
    const int* fidx = reinterpret_cast<const int*>(g_sm + SH_FIDX);
    // fragment tables from the Jacobian table of build_tables (row 6 i + j: d f_i / d x_j, 36 + 6 i + j: d z_i / d x_j;
    // slot 0: degree-1 constant, 2..7: degree 2 against x_0..x_5, 8..28: degree 3 against the 21 products)
    for (int e = threadIdx.x; e < E2_TAB; e += blockDim.x) {
        const int l = e & 31, ks = (e >> 5) & 7, nt = (e >> 8) & 1, j = e >> 9;
        const int k = 4 * ks + (l & 3), out = 8 * nt + (l >> 2);
        double v = 0.0;
        if (out < 12) {
            const int pd = (out < 6) ? 6 * out + j : 36 + 6 * (out - 6) + j;
            const int slot = (k < 6) ? 2 + k : ((k >= 8 && k < 29) ? k : -1);
            if (slot >= 0) v = T[pd * TS + slot];
        }
        TAB[e] = v;
    }
    for (int e = threadIdx.x; e < E2_G1; e += blockDim.x) {
        const int h = e & 1, l = (e >> 1) & 31, nt = (e >> 6) & 1, j = e >> 7;
        const int out = 8 * nt + 2 * (l & 3) + h;
        double v = 0.0;
        if (out < 12) v = T[((out < 6) ? 6 * out + j : 36 + 6 * (out - 6) + j) * TS];
        G1[e] = v;
    }
    // psi recipe of this lane: slot 4 ks + q of the 32-slot operand (0..5: x, 8..28: products)
    int ra[8], rb[8];                              // indices into the padded state row (6, 7 read zeros; 8: the constant 1)
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int k = 4 * ks + q;
        if (k < 6) { ra[ks] = k; rb[ks] = -1; }
        else if (k >= 8 && k < 29) { const int pk = fidx[6 + k - 8]; ra[ks] = pk & 7; rb[ks] = (pk >> 3) & 7; }
        else { ra[ks] = 6; rb[ks] = -1; }
    }
    __syncthreads();
    const double sdt = fe ? dt : 1.0;
    const long long groups = (count + 7) / 8;
    const long long gstride = (long long)gridDim.x * E2_WARPS;
    for (long long grp = (long long)blockIdx.x * E2_WARPS + warp; grp < groups; grp += gstride) {
        const long long st0 = grp * 8;
        // stage the 8 states / inputs of the group
        {
            const int sidx = lane >> 2, c2 = (lane & 3) * 2;          // two consecutive entries per lane
            const long long st = st0 + sidx;
            double v0 = 0.0, v1 = 0.0, u0 = 0.0, u1 = 0.0;
            if (st < count) {
                if (c2 < 6) { v0 = xg[st * 6 + c2]; v1 = xg[st * 6 + c2 + 1]; }
                if (ug && c2 < M) { u0 = ug[st * M + c2]; u1 = ug[st * M + c2 + 1]; }
            }
            *reinterpret_cast<double2*>(XE + sidx * 8 + c2) = make_double2(v0, v1);
            *reinterpret_cast<double2*>(UE + sidx * 8 + c2) = make_double2(u0, u1);
        }
        __syncwarp();
        // Psi fragment: A operand (row = state g, k = 4 ks + q)
        double psi[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const double a = XE[g * 8 + ra[ks]];
            psi[ks] = (rb[ks] >= 0) ? __dmul_rn(a, XE[g * 8 + rb[ks]]) : a;
        }
        const double2 x01 = *reinterpret_cast<const double2*>(XE + g * 8);
        const double2 x23 = *reinterpret_cast<const double2*>(XE + g * 8 + 2);
        const double2 x45 = *reinterpret_cast<const double2*>(XE + g * 8 + 4);
        const double xs[6] = {x01.x, x01.y, x23.x, x23.y, x45.x, x45.y};
        // rows owned by this lane: tile 0 -> outputs 2q, 2q+1 ; tile 1 -> outputs 8 + 2q, 9 + 2q  (valid for q < 2)
        double J[2][2][6];                         // [tile][half][j]
        double val[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        constexpr double third = 1.0 / 3.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const double* tb = TAB + ((j * 2 + nt) * 8) * 32 + lane;
                Frag a2{0.0, 0.0}, a3{0.0, 0.0}, b3{0.0, 0.0};
                dmma(a2, psi[0], tb[0 * 32]);
                dmma(a3, psi[2], tb[2 * 32]);
                dmma(b3, psi[3], tb[3 * 32]);
                dmma(a2, psi[1], tb[1 * 32]);
                dmma(a3, psi[4], tb[4 * 32]);
                dmma(b3, psi[5], tb[5 * 32]);
                dmma(a3, psi[6], tb[6 * 32]);
                dmma(b3, psi[7], tb[7 * 32]);
                const double2 g1 = *reinterpret_cast<const double2*>(G1 + ((j * 2 + nt) * 32 + lane) * 2);
                const double g30 = __dadd_rn(a3.c0, b3.c0), g31 = __dadd_rn(a3.c1, b3.c1);
                J[nt][0][j] = __dadd_rn(__dadd_rn(g1.x, a2.c0), g30);
                J[nt][1][j] = __dadd_rn(__dadd_rn(g1.y, a2.c1), g31);
                // Euler: value += x_j (g1 + g2 / 2 + g3 / 3)
                val[nt][0] = fma(xs[j], __dadd_rn(__dadd_rn(g1.x, __dmul_rn(0.5, a2.c0)), __dmul_rn(third, g30)), val[nt][0]);
                val[nt][1] = fma(xs[j], __dadd_rn(__dadd_rn(g1.y, __dmul_rn(0.5, a2.c1)), __dmul_rn(third, g31)), val[nt][1]);
            }
        }
        const long long st = st0 + g;
        if (st < count) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int out = 8 * nt + 2 * q + h;
                    if (out >= 12) continue;
                    const double* Jr = J[nt][h];
                    double jx = 0.0;                                   // (row of the Jacobian) . x
#pragma unroll
                    for (int j = 0; j < 6; ++j) jx = fma(Jr[j], xs[j], jx);
                    if (out < 6) {
                        // dynamics row `out`: f = r phi + B u ; d = (f - A x) - B u               (ssm.py:168, 203)
                        double bu = 0.0;
#pragma unroll
                        for (int k2 = 0; k2 < M; ++k2) bu = fma(Brt[out * LD + k2], UE[g * 8 + k2], bu);
                        const double dc = __dsub_rn(__dsub_rn(__dadd_rn(val[nt][h], bu), jx), bu);
                        if (Ao) {
                            double r6[6];
#pragma unroll
                            for (int j = 0; j < 6; ++j) {
                                const double v = fe ? __dmul_rn(dt, Jr[j]) : Jr[j];
                                r6[j] = (fe && j == out) ? __dadd_rn(1.0, v) : v;
                            }
                            double2* dst = reinterpret_cast<double2*>(Ao + st * 36 + out * 6);
                            dst[0] = make_double2(r6[0], r6[1]); dst[1] = make_double2(r6[2], r6[3]); dst[2] = make_double2(r6[4], r6[5]);
                        }
                        if (Bo) {
#pragma unroll
                            for (int k2 = 0; k2 < M; k2 += 2)
                                *reinterpret_cast<double2*>(Bo + st * 6 * M + out * M + k2) =
                                    make_double2(__dmul_rn(sdt, Brt[out * LD + k2]), __dmul_rn(sdt, Brt[out * LD + k2 + 1]));
                        }
                        if (dout) dout[st * 6 + out] = fe ? __dmul_rn(dt, dc) : dc;
                    } else {
                        const int i = out - 6;
                        if (Ho) {
                            double2* dst = reinterpret_cast<double2*>(Ho + st * 36 + i * 6);
                            dst[0] = make_double2(Jr[0], Jr[1]); dst[1] = make_double2(Jr[2], Jr[3]); dst[2] = make_double2(Jr[4], Jr[5]);
                        }
                        if (zo) zo[st * 6 + i] = __dadd_rn(val[nt][h], zref[i]);
                        if (co) co[st * 6 + i] = __dsub_rn(val[nt][h], jx);       // c = C(x) - H x  (ssm.py:228-235)
                    }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace fast

int ssm_eval_dmma_launch(const SsmDev& M, long long count, const double* x, const double* u, double dt, double* A,
                         double* B, double* d, double* H, double* c, double* z, cudaStream_t st, bool* handled) {
    *handled = false;
    const char* env = getenv("SRCB200_ILQR_GENERIC");
    if (env && env[0] == '1') return 0;
    if (!(M.n == 6 && M.nz == 6 && M.order == 3 && M.nfeat == fast::NFEAT && (M.m == 4 || M.m == 8))) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool want_dyn = (A || B || d);
    const bool cont = (dt < 0.0) || M.discr == SRCB200_DISCR_NONE;
    const char* dense = getenv("SRCB200_SSM_EVAL_DENSE");
    if ((!want_dyn || cont || M.discr == SRCB200_DISCR_FE) && !(dense && dense[0] == '1')) {
        // sparse contraction, eight states per warp (no 6 x 6 inverses needed for these outputs)
        const long long groups = (count + 7) / 8;
        const long long ctas2 = (groups + fast::E2_WARPS - 1) / fast::E2_WARPS;
        const int grid2 = (int)(ctas2 < 3LL * sms ? ctas2 : 3LL * sms);
        const int fe = (!cont && M.discr == SRCB200_DISCR_FE) ? 1 : 0;
        if (M.m == 8) {
            SRCB_CUDA(cudaFuncSetAttribute(fast::ssm_eval_sparse_dmma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast::E2_SMEM));
            fast::ssm_eval_sparse_dmma_kernel<8><<<grid2, fast::E2_WARPS * 32, fast::E2_SMEM, st>>>(M, count, x, u, dt, fe, A, B, d, H, c, z);
        } else {
            SRCB_CUDA(cudaFuncSetAttribute(fast::ssm_eval_sparse_dmma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast::E2_SMEM));
            fast::ssm_eval_sparse_dmma_kernel<4><<<grid2, fast::E2_WARPS * 32, fast::E2_SMEM, st>>>(M, count, x, u, dt, fe, A, B, d, H, c, z);
        }
        SRCB_LAUNCH_CHECK("ssm_eval_sparse_dmma_kernel");
        *handled = true;
        return 0;
    }
    const long long ctas = (count + fast::EV_WARPS - 1) / fast::EV_WARPS;
    const int grid = (int)(ctas < 5LL * sms ? ctas : 5LL * sms);
    if (M.m == 8) fast::ssm_eval_dmma_kernel<8><<<grid, fast::EV_WARPS * 32, 0, st>>>(M, count, x, u, dt, A, B, d, H, c, z);
    else          fast::ssm_eval_dmma_kernel<4><<<grid, fast::EV_WARPS * 32, 0, st>>>(M, count, x, u, dt, A, B, d, H, c, z);
    SRCB_LAUNCH_CHECK("ssm_eval_dmma_kernel");
    *handled = true;
    return 0;
}

// Rollout dispatch (called from ssm.cu): returns handled = true when the specialised kernel ran.
int ssm_rollout_fast_launch(const SsmDev& M, long long batch, int N, const double* x0, const double* u, double dt,
                            double* x, double* z, cudaStream_t st, bool* handled) {
    *handled = false;
    const char* env = getenv("SRCB200_ILQR_GENERIC");
    if (env && env[0] == '1') return 0;
    if (!(M.n == 6 && M.nz == 6 && M.order == 3 && M.nfeat == fast::NFEAT && (M.m == 4 || M.m == 8))) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long ctas = (batch + fast::WARPS - 1) / fast::WARPS;
    const int grid = (int)(ctas < 2LL * sms ? ctas : 2LL * sms);
    IlqrArgs a;
    memset(&a, 0, sizeof(a));
    a.n = 6; a.m = M.m; a.nz = 6; a.N = N; a.batch = batch; a.dt = dt;
    if (M.m == 8) {
        SRCB_CUDA(cudaFuncSetAttribute(fast::ssm_rollout_fast_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast::SMEM_BYTES));
        fast::ssm_rollout_fast_kernel<8><<<grid, fast::WARPS * 32, fast::SMEM_BYTES, st>>>(M, a, batch, x0, u, x, z);
    } else {
        SRCB_CUDA(cudaFuncSetAttribute(fast::ssm_rollout_fast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast::SMEM_BYTES));
        fast::ssm_rollout_fast_kernel<4><<<grid, fast::WARPS * 32, fast::SMEM_BYTES, st>>>(M, a, batch, x0, u, x, z);
    }
    SRCB_LAUNCH_CHECK("ssm_rollout_fast_kernel");
    *handled = true;
    return 0;
}


#ifdef SRCB_PHASE_TIMING
extern "C" int srcb200_debug_phase(unsigned long long* host32, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host32, fast::g_phase, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(fast::g_phase, z, sizeof(z)); }
    return 0;
}
#endif

// Dispatch: Gauss-Newton SSM problems with the Trunk/Diamond shape go to the specialised kernel.
int ilqr_ssm_fast_launch(const SsmDev& M, const IlqrArgs& a, cudaStream_t st, bool* handled) {
    *handled = false;
    const char* env = getenv("SRCB200_ILQR_GENERIC");
    if (env && env[0] == '1') return 0;
    if (!(M.n == 6 && M.nz == 6 && M.order == 3 && M.nfeat == fast::NFEAT && (M.m == 4 || M.m == 8) && a.gn && !a.index_lin) || a.batch > (1LL << 29))
        return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // Persistent launch, three shapes (SRCB200_ILQR_SHAPE forces one): 0 two CTAs of 8 warps per SM, 128 registers, Jacobian
    // table read from shared memory; 1: one CTA of 12 warps, 168 registers, table row `lane` in registers; 2: one CTA of
    // 8 warps, 255 registers, rows `lane` and `32 + lane` in registers (each warp ~1.8 x faster, half as many run).  Small
    // batches are spread over the SMs: the task queue feeds any number of warps.
    // Without the switch the shape follows the batch (tools/shape_sweep.py, N = 100, one B200, ms per batch, shapes 0 / 1 / 2:
    // 512: 27.3 / 23.3 / 21.7, 1024: 28.1 / 25.2 / 22.5, 2048: 34.5 / 30.6 / 28.3, 2368: 37.5 / 34.0 / 32.3, 3072: 39.6 / 35.3 /
    // 39.0, 3584: 41.2 / 39.5 / 43.2, 4096: 45.1 / 45.0 / 48.5, 6144: 61.6 / 60.7 / 69.2, 8192: 77.2 / 78.3 / -): few problems
    // are bound by the speed of a single warp (registers instead of shared-memory reads), many by the SM's throughput.
    const char* shp = getenv("SRCB200_ILQR_SHAPE");
    const int cr = (shp && shp[0] >= '0' && shp[0] <= '2') ? shp[0] - '0'
                   : (a.batch <= 2560 ? 2 : (a.batch <= 3840 ? 1 : 0));
    ilqrq::queue_init_kernel<<<(a.queue_cap + 255) / 256, 256, 0, st>>>(a.work_counter, a.queue_cap, (int)a.batch, a.ws,
                                                                                 a.L.total, a.L.state);
    SRCB_LAUNCH_CHECK("ilqr_queue_init_kernel");
    // Large batches: the throughput shape runs until `handover` problems are left, then its warps stop taking tasks (every
    // problem is suspended in global memory between iterations anyway) and a second launch in shape 2 -- half as many
    // warps, each ~1.8 x faster -- finishes the tail, which is bound by the speed of single warps (SRCB200_ILQR_HANDOVER=0
    // disables, =n sets the threshold).
    long long handover = (cr == 2 || shp) ? 0 : 1776;     // tools/handover_sweep.py: 4096 problems 44.6 -> 42.6 ms (1776 / 2368), 8192: 77.0 -> 76.1
    if (const char* ho = getenv("SRCB200_ILQR_HANDOVER")) handover = (cr == 2) ? 0 : atoll(ho);
    if (handover >= a.batch) handover = 0;
    auto launch = [&](int shape, int stop_at) -> int {
        IlqrArgs aa = a;
        aa.stop_at = stop_at;
        const int nw = fast::shape_warps(shape);
        const size_t smem = sizeof(double) * (fast::SH_END + nw * fast::W_SIZE);
        const long long slots = (long long)sms * (shape == 0 ? 2 : 1);
        int grid = (int)(a.batch < slots ? a.batch : slots);
        if (grid * nw > kIlqrQueueWaiters) grid = kIlqrQueueWaiters / nw;
#define SRCB_LAUNCH_SHAPE(MM, CRR)                                                                                         \
    do {                                                                                                                   \
        SRCB_CUDA(cudaFuncSetAttribute(fast::ilqr_ssm_fast_kernel<MM, CRR>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                       (int)smem));                                                                        \
        fast::ilqr_ssm_fast_kernel<MM, CRR><<<grid, nw * 32, smem, st>>>(M, aa);                                           \
    } while (0)
        if (M.m == 8) { if (shape == 0) SRCB_LAUNCH_SHAPE(8, 0); else if (shape == 1) SRCB_LAUNCH_SHAPE(8, 1); else SRCB_LAUNCH_SHAPE(8, 2); }
        else          { if (shape == 0) SRCB_LAUNCH_SHAPE(4, 0); else if (shape == 1) SRCB_LAUNCH_SHAPE(4, 1); else SRCB_LAUNCH_SHAPE(4, 2); }
#undef SRCB_LAUNCH_SHAPE
        SRCB_LAUNCH_CHECK("ilqr_ssm_fast_kernel");
        return 0;
    };
    if (const char* chain = getenv("SRCB200_ILQR_CHAIN")) {
        // experiments: "shape:stop_at,shape:stop_at,...", the last entry with stop_at 0
        const char* pch = chain;
        while (*pch) {
            const int shape = *pch - '0';
            const long long stop = atoll(pch + 2);
            if (shape < 0 || shape > 2) break;
            if (stop < a.batch) if (int e = launch(shape, (int)stop)) return e;
            while (*pch && *pch != ',') ++pch;
            if (*pch == ',') ++pch;
            if (stop == 0) break;
        }
    } else {
        if (int e = launch(cr, (int)handover)) return e;
        if (handover > 0) if (int e = launch(2, 0)) return e;
    }
    SRCB_LAUNCH_CHECK("ilqr_ssm_fast_kernel");
    *handled = true;
    return 0;
}

}  // namespace srcb
