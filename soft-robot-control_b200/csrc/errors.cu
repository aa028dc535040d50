// errors.cu -- error string plumbing + version / device check for libsrcb200.
#include <stdarg.h>
#include "common.cuh"

namespace srcb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();   // clear the sticky launch error so the next call reports its own
    return (int)e;
}
}  // namespace srcb

__global__ void srcb_probe_kernel(int* out) { if (out) *out = SRCB200_ABI_VERSION; }

extern "C" int srcb200_abi_version(void) { return SRCB200_ABI_VERSION; }
extern "C" const char* srcb200_last_error_string(void) { return srcb::g_err; }
extern "C" int srcb200_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { srcb::cuda_fail(e, "cudaGetDevice"); return SRCB200_E_NOGPU; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { srcb::cuda_fail(e, "cudaGetDeviceProperties"); return SRCB200_E_NOGPU; }
    if (p.major != 10) return srcb::fail(SRCB200_E_NOGPU, "device %s is sm_%d%d; libsrcb200 is built for sm_100a only",
                                         p.name, p.major, p.minor);
    srcb_probe_kernel<<<1, 1>>>(nullptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) { srcb::cuda_fail(e, "probe kernel launch"); return SRCB200_E_NOGPU; }
    return 0;
}
