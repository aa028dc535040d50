// ilqr_tpwl.cu -- generic iLQR kernels instantiated for the TPWL bank policy (see ilqr_impl.cuh); the Diamond shape
// is forwarded to the compile-time-dimension instantiation in ilqr_tpwl_diamond.cu.
#include <cstdlib>
#include "ilqr_impl.cuh"

namespace srcb {
// the Diamond shape (n = 72, m = 4, n_z = 6) runs the instantiation with compile-time dimensions
int ilqr_solve_tpwl_diamond(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr,
                            const srcb200_ilqr_result* res, void* ws, size_t ws_bytes, cudaStream_t st);
int ilqr_forward_tpwl_diamond(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const double* xp,
                              const double* up, double alpha, const double* K, const double* k, double* x, double* u,
                              double* cost, double* A, double* B, double* d, void* ws, size_t ws_bytes, cudaStream_t st);
int ilqr_backward_tpwl_diamond(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const double* x,
                               const double* u, const double* A, const double* B, double* K, double* k, double* Qu, double* Quu,
                               double* rho, double* drho, int32_t* restarts, void* ws, size_t ws_bytes, cudaStream_t st);
static bool diamond_shape(const TpwlDev& M) {
    const char* env = getenv("SRCB200_ILQR_GENERIC");
    return M.n == 72 && M.m == 4 && M.nz == 6 && !(env && env[0] == '1');
}
int ilqr_solve_tpwl(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const srcb200_ilqr_result* res,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
    if (diamond_shape(M))
        return ilqr_solve_tpwl_diamond(M, cfg, pr, res, ws, ws_bytes, st);
    return solve_impl<TpwlPolicy>(M, cfg, pr, res, ws, ws_bytes, st);
}
int ilqr_forward_tpwl(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const double* xp,
                      const double* up, double alpha, const double* K, const double* k, double* x, double* u, double* cost,
                      double* A, double* B, double* d, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (diamond_shape(M)) return ilqr_forward_tpwl_diamond(M, cfg, pr, xp, up, alpha, K, k, x, u, cost, A, B, d, ws, ws_bytes, st);
    return forward_impl<TpwlPolicy>(M, cfg, pr, xp, up, alpha, K, k, x, u, cost, A, B, d, ws, ws_bytes, st);
}
int ilqr_backward_tpwl(const TpwlDev& M, const srcb200_ilqr_config* cfg, const srcb200_ilqr_problem* pr, const double* x,
                       const double* u, const double* A, const double* B, double* K, double* k, double* Qu, double* Quu,
                       double* rho, double* drho, int32_t* restarts, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (diamond_shape(M)) return ilqr_backward_tpwl_diamond(M, cfg, pr, x, u, A, B, K, k, Qu, Quu, rho, drho, restarts, ws, ws_bytes, st);
    return backward_impl<TpwlPolicy>(M, cfg, pr, x, u, A, B, K, k, Qu, Quu, rho, drho, restarts, ws, ws_bytes, st);
}
}  // namespace srcb
