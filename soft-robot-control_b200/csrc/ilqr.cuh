// ilqr.cuh -- structures shared by the generic iLQR kernels (ilqr.cu) and the register/tensor-core specialised
// Trunk/Diamond SSM kernel (ilqr_fast.cu).
#pragma once
#include "ssm.cuh"
#include "tpwl.cuh"

namespace srcb {

// ---------------------------------------------------------------------------------------------------------------
// Per-problem global scratch
// ---------------------------------------------------------------------------------------------------------------
struct Layout {                 // offsets in doubles inside one trajectory record / the per-problem workspace
    long long x, u, e, H, A, B, idx, rec;       // record fields, record size
    long long k, ab, cxx, state, total;         // backward outputs (K goes straight to the result buffer); constant c_xx;
                                                // solver state of a suspended problem (fast kernel, 8 doubles)
};

__host__ __device__ inline Layout make_layout(int n, int m, int nz, int N, bool gn, bool index_lin) {
    Layout L;
    long long o = 0;
    L.x = o;  o += (long long)(N + 1) * n;
    L.u = o;  o += (long long)N * m;
    L.e = o;  o += (long long)(N + 1) * nz;
    L.H = o;  o += gn ? (long long)(N + 1) * nz * n : 0;
    L.A = o;  o += index_lin ? 0 : (long long)N * n * n;
    L.B = o;  o += index_lin ? 0 : (long long)N * n * m;
    L.idx = o; o += index_lin ? (N + 1) / 2 + 1 : 0;     // N int32 packed into doubles
    L.rec = (o + 1) & ~1LL;
    o = 2 * L.rec;
    L.k = o;  o += (long long)N * m;
    L.ab = o; o += 2LL * N;
    L.cxx = o; o += gn ? 0 : (long long)n * n;  // constant-H mode: H^T Q H lives here (global, L2) instead of shared memory
    o = (o + 1) & ~1LL;
    L.state = o; o += 8;
    L.total = (o + 1) & ~1LL;
    return L;
}

struct Rec {                    // one trajectory record resolved to pointers
    double* x; double* u; double* e; double* H; double* A; double* B; int* idx;
};
__device__ inline Rec rec_at(double* base, const Layout& L) {
    Rec r;
    r.x = base + L.x; r.u = base + L.u; r.e = base + L.e; r.H = base + L.H; r.A = base + L.A; r.B = base + L.B;
    r.idx = reinterpret_cast<int*>(base + L.idx);
    return r;
}

// Shared-memory plan of the TPWL nearest-neighbour forward pass (ilqr_fwd_tpwl.cuh), in doubles from its base
struct FwdNNPlan {
    int pre0, pre1, PRE, Acur, LDA, xf, dh, cand, red, misc, qf, vf, end;
    int screen;                 // 1: FP32 screening banks are resident
};
constexpr int kFwdNNCandCap = 512;

__host__ __device__ inline FwdNNPlan make_fwdnn(int n, int m, int nz, int P, int r, bool useq, bool usev, int nt) {
    FwdNNPlan F;
    int o = 0;
    auto take = [&o](int cnt) { const int at = o; o += (cnt + 1) & ~1; return at; };
    F.PRE = (n + 2 * m + m * n + nz + 1) & ~1;
    F.pre0 = take(F.PRE);
    F.pre1 = take(F.PRE);
    F.LDA = n; while ((F.LDA & 15) != 4 && (F.LDA & 15) != 12) ++F.LDA;      // conflict-free row stride of A
    F.Acur = take(n * F.LDA + n * m + n);
    F.xf = take((n + 1) / 2);               // FP32 copy of the state
    F.dh = take((P + 1) / 2);               // P floats
    F.cand = take(kFwdNNCandCap / 2 + 2);   // ints: [count, pad, candidates...]
    F.red = take(2 * (nt / 32) + 2);
    F.misc = take(16);
    const int nb = (useq ? 1 : 0) + (usev ? 1 : 0);
    const long long bank = ((long long)P * r + 1) / 2;          // doubles per FP32 bank
    F.screen = (nb > 0 && nb * bank * 8 <= 150 * 1024) ? 1 : 0;
    F.qf = o; if (F.screen && useq) o += (int)((bank + 1) & ~1LL);
    F.vf = o; if (F.screen && usev) o += (int)((bank + 1) & ~1LL);
    F.end = o;
    return F;
}

// Task queues of the fast kernel behind the per-problem scratch: 64 ints of counters + two slot rings (priority classes).  The ring must be
// longer than (problems that can be queued) + (warps that can wait on a ticket at the same time), see ilqr_fast.cu.
constexpr int kIlqrQueueWaiters = 8192;
inline long long ilqr_queue_cap(long long batch) { return 2 * batch + kIlqrQueueWaiters; }
inline size_t ilqr_queue_bytes(long long batch) { return 256 + 2 * sizeof(int) * (size_t)ilqr_queue_cap(batch); }   // two rings

struct IlqrArgs {
    int n, m, nz, N, gn, index_lin, shared_target;
    long long batch;
    double dt;
    srcb200_ilqr_config cfg;
    const double *x0, *u_init, *z_target, *u_last, *Q, *R, *Qf, *Hc;
    double *ox, *ou, *oK, *ocost, *ocost0, *orho, *otrace;
    int *oiter, *ostatus, *otrials;
    double* ws;
    Layout L;
    int model_scratch;          // doubles of model scratch in shared memory
    double prio_frac;           // HIGH class: initial cost > prio_frac * running mean
    int* work_counter;          // fast kernel's task queue: ints [head, tail, remaining, pad..64) then `queue_cap` slots
    int queue_cap;
    int stop_at;                // fast kernel: warps stop taking tasks once this many problems (or fewer) are unfinished
                                // (0: run to the end) -- the tail of a large batch is handed to a launch shape with faster warps
};

// ---------------------------------------------------------------------------------------------------------------
// Task queues of the persistent solve kernels (ilqr_fast.cu: a task = one iteration of one problem on one warp;
// ilqr_impl.cuh: the same on one CTA).  Two priority classes, one ticket ring each.
// Counters (ints at q): per ring c in {HIGH = 0, LOW = 1}: head q[4c], tail q[4c+1], avail q[4c+2]; q[8] = problems
// not finished; q[10..11] = sum of initial costs (double), q[12] = their count.  Rings at q + 64 + c * cap.
// push: p = tail++, slot[p % cap] = id, fence, avail++.   pop: acquire one unit of `avail` (so a committed entry
// exists for every ticket), t = head++, wait for slot[t % cap] (its push has at least reserved it), take it.
// ---------------------------------------------------------------------------------------------------------------
namespace ilqrq {
constexpr int Q_REMAINING = 8, Q_CSUM = 10, Q_CCNT = 12;
constexpr int kStarted = 0x5ca1ab1e;        // marker in a problem's saved state: its first task has run

static __global__ void queue_init_kernel(int* q, int cap, int batch, double* ws, long long total, long long state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        q[64 + i] = -1;                         // HIGH ring: empty
        q[64 + cap + i] = i < batch ? i : -1;   // LOW ring: every problem's first task
    }
    if (i < batch) reinterpret_cast<int*>(ws + i * total + state + 3)[5] = 0;      // "not started"
    if (i == 0) {
        for (int k = 0; k < 16; ++k) q[k] = 0;
        q[4 + 1] = batch;                       // LOW tail
        q[4 + 2] = batch;                       // LOW avail
        q[Q_REMAINING] = batch;
    }
}

__device__ __forceinline__ bool try_acquire(int* avail) {
    if (*(volatile int*)avail <= 0) return false;
    if (atomicSub(avail, 1) >= 1) return true;
    atomicAdd(avail, 1);
    return false;
}

// one thread: next task id (HIGH ring first), or -1 when every problem is finished
__device__ __forceinline__ int pop_one(int* q, int cap, int stop_at = 0) {
    unsigned ns = 256;                          // back off: an idle warp / CTA must not compete with working ones
    while (true) {
        // hand-over: nothing is taken any more, the tasks stay in the rings for the next launch
        if (stop_at > 0 && *(volatile int*)(q + Q_REMAINING) <= stop_at) return -1;
        int cls = -1;
        if (try_acquire(q + 2)) cls = 0;
        else if (try_acquire(q + 4 + 2)) cls = 1;
        if (cls >= 0) {
            const int t = atomicAdd(q + 4 * cls, 1);
            volatile int* slot = q + 64 + cls * cap + (t % cap);
            int v;
            while ((v = *slot) < 0) __nanosleep(64);      // the push that owns this ticket is between tail++ and the store
            *slot = -1;
            return v;
        }
        if (*(volatile int*)(q + Q_REMAINING) <= 0) return -1;
        __nanosleep(ns);
        if (ns < 8192) ns <<= 1;
    }
}

// one thread, after the caller's release fence
__device__ __forceinline__ void push_one(int* q, int cap, int id, int cls) {
    const int p = atomicAdd(q + 4 * cls + 1, 1);
    atomicExch(q + 64 + cls * cap + (p % cap), id);
    __threadfence();
    atomicAdd(q + 4 * cls + 2, 1);
}

// one thread: priority class from the initial cost -- above the running mean of the batch: HIGH (0), else LOW (1)
__device__ __forceinline__ int classify(int* q, double cost, double frac = 1.0) {
    if (!isfinite(cost)) return 1;
    double* csum = reinterpret_cast<double*>(q + Q_CSUM);
    const double sprev = atomicAdd(csum, cost);
    const int cprev = atomicAdd(q + Q_CCNT, 1);
    return (cost * (double)(cprev + 1) > frac * (sprev + cost)) ? 0 : 1;
}
}  // namespace ilqrq

// rho schedule (ilqr.py:198-217), including the `dhro` typo: drho is never lowered.
__device__ __forceinline__ void rho_update(const srcb200_ilqr_config& c, bool increase, double& rho, double& drho) {
    if (increase) {
        drho = fmax(__dmul_rn(drho, c.rho_scaling), c.rho_scaling);
        rho = fmax(__dmul_rn(rho, drho), c.rho_min);
        if (rho > c.rho_max) rho = c.rho_max;
    } else {
        const double dhro = fmin(__ddiv_rn(drho, c.rho_scaling), __ddiv_rn(1.0, c.rho_scaling));
        rho = __dmul_rn(rho, dhro);
        if (rho <= c.rho_min) rho = c.rho_min;
    }
}


// implemented in ilqr_fast.cu: returns 1 if the problem was dispatched to the specialised SSM kernel
int ilqr_ssm_fast_launch(const SsmDev& M, const IlqrArgs& a, cudaStream_t st, bool* handled);

}  // namespace srcb
