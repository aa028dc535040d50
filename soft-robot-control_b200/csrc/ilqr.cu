// ilqr.cu -- C ABI of the batched iLQR (sofacontrol/lqr/ilqr.py:27-300); the kernels live in ilqr_impl.cuh
// (generic, one CTA per problem; instantiated in ilqr_ssm.cu / ilqr_tpwl.cu) and ilqr_fast.cu (Trunk/Diamond SSM).
#include "ilqr.cuh"

namespace srcb {
int ilqr_solve_ssm(const SsmDev&, const srcb200_ilqr_config*, const srcb200_ilqr_problem*, const srcb200_ilqr_result*, void*, size_t, cudaStream_t);
int ilqr_solve_tpwl(const TpwlDev&, const srcb200_ilqr_config*, const srcb200_ilqr_problem*, const srcb200_ilqr_result*, void*, size_t, cudaStream_t);
int ilqr_forward_ssm(const SsmDev&, const srcb200_ilqr_config*, const srcb200_ilqr_problem*, const double*, const double*, double,
                     const double*, const double*, double*, double*, double*, double*, double*, double*, void*, size_t, cudaStream_t);
int ilqr_forward_tpwl(const TpwlDev&, const srcb200_ilqr_config*, const srcb200_ilqr_problem*, const double*, const double*, double,
                      const double*, const double*, double*, double*, double*, double*, double*, double*, void*, size_t, cudaStream_t);
int ilqr_backward_ssm(const SsmDev&, const srcb200_ilqr_config*, const srcb200_ilqr_problem*, const double*, const double*,
                      const double*, const double*, double*, double*, double*, double*, double*, double*, int32_t*, void*, size_t, cudaStream_t);
int ilqr_backward_tpwl(const TpwlDev&, const srcb200_ilqr_config*, const srcb200_ilqr_problem*, const double*, const double*,
                       const double*, const double*, double*, double*, double*, double*, double*, double*, int32_t*, void*, size_t, cudaStream_t);

static int check_problem(const srcb200_ilqr_problem* pr, bool need_x0) {
    if (!pr) return fail(SRCB200_E_NULL, "ilqr problem is NULL");
    if (pr->batch < 0 || pr->N < 1) return fail(SRCB200_E_DIM, "ilqr: batch=%lld N=%d", (long long)pr->batch, pr->N);
    if ((need_x0 && !pr->x0) || !pr->z_target || !pr->Q || !pr->R || !pr->Qf)
        return fail(SRCB200_E_NULL, "ilqr: x0/z_target/Q/R/Qf is NULL");
    return 0;
}
}  // namespace srcb

using namespace srcb;

#define DISPATCH_MODEL(kind, model, CALL_SSM, CALL_TPWL)                                            \
    if ((kind) == SRCB200_ILQR_MODEL_SSM) {                                                         \
        const srcb200_ssm_model* mm_ = (const srcb200_ssm_model*)(model);                           \
        if (int e = check_ssm_model(mm_)) return e;                                                 \
        SsmDev M = to_dev(*mm_);                                                                    \
        return CALL_SSM;                                                                            \
    } else if ((kind) == SRCB200_ILQR_MODEL_TPWL) {                                                 \
        const srcb200_tpwl_model* mm_ = (const srcb200_tpwl_model*)(model);                         \
        if (int e = check_tpwl_model(mm_)) return e;                                                \
        TpwlDev M = to_dev(*mm_);                                                                   \
        if (M.discr == SRCB200_DISCR_ZOH && pr->dt >= 0.0)                                          \
            return fail(SRCB200_E_METHOD, "iLQR on a TPWL model with zoh needs a pre-discretised bank "        \
                                          "(pre_discretize(dt)): per-step expm is not available inside the solver"); \
        return CALL_TPWL;                                                                           \
    }                                                                                               \
    return fail(SRCB200_E_DIM, "unknown model_kind %d", (int)(kind));

extern "C" size_t srcb200_ilqr_workspace_bytes(int32_t model_kind, const void* model, const srcb200_ilqr_problem* pr) {
    if (!model || !pr || pr->batch <= 0) return 0;
    int n, m, nz;
    bool il;
    if (model_kind == SRCB200_ILQR_MODEL_SSM) {
        SsmDev M = to_dev(*(const srcb200_ssm_model*)model);
        n = M.n; m = M.m; nz = M.nz; il = false;
    } else {
        TpwlDev M = to_dev(*(const srcb200_tpwl_model*)model);
        n = M.n; m = M.m; nz = M.nz; il = (M.method == SRCB200_TPWL_NN) && (M.discr == SRCB200_DISCR_NONE || pr->dt < 0.0);   // == TpwlPolicy::index_lin
    }
    const Layout L = make_layout(n, m, nz, pr->N, pr->gauss_newton != 0, il);
    return sizeof(double) * (size_t)L.total * (size_t)pr->batch + ilqr_queue_bytes(pr->batch);   // + the task queue of the fast kernel
}

extern "C" int srcb200_ilqr_solve_batch(int32_t model_kind, const void* model, const srcb200_ilqr_config* cfg,
                                        const srcb200_ilqr_problem* pr, const srcb200_ilqr_result* res, void* ws,
                                        size_t ws_bytes, void* stream) {
    if (!model || !cfg) return fail(SRCB200_E_NULL, "ilqr: model/cfg is NULL");
    if (int e = check_problem(pr, true)) return e;
    if (pr->batch == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_MODEL(model_kind, model, ilqr_solve_ssm(M, cfg, pr, res, ws, ws_bytes, st),
                   ilqr_solve_tpwl(M, cfg, pr, res, ws, ws_bytes, st));
}

extern "C" int srcb200_ilqr_forward_pass(int32_t model_kind, const void* model, const srcb200_ilqr_config* cfg,
                                         const srcb200_ilqr_problem* pr, const double* x_prev, const double* u_prev,
                                         double alpha, const double* K, const double* k, double* x, double* u,
                                         double* cost, double* A, double* B, double* d, void* ws, size_t ws_bytes,
                                         void* stream) {
    if (!model || !cfg) return fail(SRCB200_E_NULL, "ilqr: model/cfg is NULL");
    if (int e = check_problem(pr, false)) return e;
    if (pr->batch == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_MODEL(model_kind, model,
                   ilqr_forward_ssm(M, cfg, pr, x_prev, u_prev, alpha, K, k, x, u, cost, A, B, d, ws, ws_bytes, st),
                   ilqr_forward_tpwl(M, cfg, pr, x_prev, u_prev, alpha, K, k, x, u, cost, A, B, d, ws, ws_bytes, st));
}

extern "C" int srcb200_ilqr_backward_pass(int32_t model_kind, const void* model, const srcb200_ilqr_config* cfg,
                                          const srcb200_ilqr_problem* pr, const double* x, const double* u,
                                          const double* A, const double* B, double* K, double* k, double* Q_u,
                                          double* Q_uu, double* rho, double* drho, int32_t* restarts, void* ws,
                                          size_t ws_bytes, void* stream) {
    if (!model || !cfg) return fail(SRCB200_E_NULL, "ilqr: model/cfg is NULL");
    if (int e = check_problem(pr, false)) return e;
    if (pr->batch == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_MODEL(model_kind, model,
                   ilqr_backward_ssm(M, cfg, pr, x, u, A, B, K, k, Q_u, Q_uu, rho, drho, restarts, ws, ws_bytes, st),
                   ilqr_backward_tpwl(M, cfg, pr, x, u, A, B, K, k, Q_u, Q_uu, rho, drho, restarts, ws, ws_bytes, st));
}
