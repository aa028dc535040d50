// ilqr_tpwl_diamond.cu -- the iLQR kernels instantiated for the TPWL bank policy with the Diamond robot's dimensions
// fixed at compile time (n = 72, m = 4, n_z = 6; see TpwlPolicyT in ilqr_impl.cuh).  Its own translation unit so that
// it compiles in parallel with the run-time-dimension instantiation (ilqr_tpwl.cu).
#include "ilqr_impl.cuh"

namespace srcb {
int ilqr_solve_tpwl_diamond(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr,
                            const srcb200_ilqr_result* res, void* ws, size_t ws_bytes, cudaStream_t st) {
    return solve_impl<TpwlPolicyDiamond>(M, cfg, pr, res, ws, ws_bytes, st);
}
int ilqr_forward_tpwl_diamond(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const double* xp,
                              const double* up, double alpha, const double* K, const double* k, double* x, double* u,
                              double* cost, double* A, double* B, double* d, void* ws, size_t ws_bytes, cudaStream_t st) {
    return forward_impl<TpwlPolicyDiamond>(M, cfg, pr, xp, up, alpha, K, k, x, u, cost, A, B, d, ws, ws_bytes, st);
}
int ilqr_backward_tpwl_diamond(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const double* x,
                               const double* u, const double* A, const double* B, double* K, double* k, double* Qu, double* Quu,
                               double* rho, double* drho, int32_t* restarts, void* ws, size_t ws_bytes, cudaStream_t st) {
    return backward_impl<TpwlPolicyDiamond>(M, cfg, pr, x, u, A, B, K, k, Qu, Quu, rho, drho, restarts, ws, ws_bytes, st);
}
}  // namespace srcb
