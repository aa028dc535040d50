// ilqr_ssm.cu -- generic iLQR kernels instantiated for the SSM polynomial model policy (see ilqr_impl.cuh).
#include "ilqr_impl.cuh"

namespace srcb {
int ilqr_solve_ssm(const SsmDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const srcb200_ilqr_result* res,
                   void* ws, size_t ws_bytes, cudaStream_t st) { return solve_impl<SsmPolicy>(M, cfg, pr, res, ws, ws_bytes, st); }
int ilqr_forward_ssm(const SsmDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const double* xp,
                     const double* up, double alpha, const double* K, const double* k, double* x, double* u, double* cost,
                     double* A, double* B, double* d, void* ws, size_t ws_bytes, cudaStream_t st) {
    return forward_impl<SsmPolicy>(M, cfg, pr, xp, up, alpha, K, k, x, u, cost, A, B, d, ws, ws_bytes, st);
}
int ilqr_backward_ssm(const SsmDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const double* x,
                      const double* u, const double* A, const double* B, double* K, double* k, double* Qu, double* Quu,
                      double* rho, double* drho, int32_t* restarts, void* ws, size_t ws_bytes, cudaStream_t st) {
    return backward_impl<SsmPolicy>(M, cfg, pr, x, u, A, B, K, k, Qu, Quu, rho, drho, restarts, ws, ws_bytes, st);
}
}  // namespace srcb
