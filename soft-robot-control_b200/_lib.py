"""ctypes binding of libsrcb200.so (C ABI: include/srcb200.h) + device-buffer helpers.

torch is used here only for device memory, streams and (elsewhere) torch.distributed.  There is NO CPU fallback:
if the library is missing, or no sm_100 GPU is visible, every compute entry point raises.
"""
import ctypes as C
import os

import numpy as np

from . import _build

_LIB = None

# ---- constants mirrored from include/srcb200.h
DISCR = {'fe': 0, 'be': 1, 'bil': 2, 'zoh': 3, 'none': 4}
TPWL_METHOD = {'nn': 0, 'weighting': 1}
E_NULL, E_DIM, E_METHOD, E_WORKSPACE, E_NOGPU = -1, -2, -3, -4, -5
ILQR_MODEL_SSM, ILQR_MODEL_TPWL = 0, 1
ABI_VERSION = 3
ST_CONVERGED, ST_MAXITER, ST_ABANDONED, ST_NONPD, ST_NONFINITE = 1, 2, 4, 8, 16
SSM_MAX_ORDER = 4

c_dp = C.c_void_p  # device pointers travel as plain addresses


class SsmModel(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("nz", C.c_int32), ("order", C.c_int32),
                ("nfeat", C.c_int32), ("discr_method", C.c_int32),
                ("r_coeff", c_dp), ("w_coeff", c_dp), ("v_coeff", c_dp), ("B_r", c_dp), ("z_ref", c_dp),
                ("mono", c_dp)]


class TpwlModel(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("nz", C.c_int32), ("P", C.c_int32),
                ("method", C.c_int32), ("discr_method", C.c_int32),
                ("wq", C.c_double), ("wv", C.c_double), ("beta", C.c_double),
                ("qT", c_dp), ("vT", c_dp), ("A", c_dp), ("B", c_dp), ("d", c_dp), ("H", c_dp), ("z_ref", c_dp)]


class IlqrConfig(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("include_input_var_constraint", C.c_int32),
                ("do_linesearch", C.c_int32), ("regularize", C.c_int32), ("state_regularization", C.c_int32),
                ("counter_limit", C.c_int32), ("epsilon", C.c_double),
                ("alpha0", C.c_double), ("alpha_scaling", C.c_double), ("improv_lb", C.c_double),
                ("improv_ub", C.c_double), ("alpha_min", C.c_double),
                ("rho0", C.c_double), ("drho0", C.c_double), ("rho_scaling", C.c_double),
                ("rho_increase_fp", C.c_double), ("rho_max", C.c_double), ("rho_min", C.c_double)]


class IlqrProblem(C.Structure):
    _fields_ = [("batch", C.c_int64), ("N", C.c_int32), ("gauss_newton", C.c_int32), ("dt", C.c_double),
                ("x0", c_dp), ("u_init", c_dp), ("z_target", c_dp), ("u_last", c_dp),
                ("Q", c_dp), ("R", c_dp), ("Qf", c_dp), ("H_const", c_dp),
                ("shared_target", C.c_int32), ("_pad", C.c_int32)]


class IlqrResult(C.Structure):
    _fields_ = [("x", c_dp), ("u", c_dp), ("K", c_dp), ("cost", c_dp), ("cost0", c_dp), ("rho", c_dp),
                ("iterations", c_dp), ("status", c_dp), ("trials", c_dp), ("trace", c_dp)]


_SIGS = {
    "srcb200_abi_version": (C.c_int, []),
    "srcb200_last_error_string": (C.c_char_p, []),
    "srcb200_device_check": (C.c_int, []),
    "srcb200_ssm_eval_linearize_batch": (C.c_int, [C.POINTER(SsmModel), C.c_int64, c_dp, c_dp, C.c_double,
                                                   c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "srcb200_ssm_map_batch": (C.c_int, [C.POINTER(SsmModel), C.c_int32, C.c_int32, C.c_int64, c_dp, c_dp, c_dp,
                                        c_dp]),
    "srcb200_ssm_rollout_batch": (C.c_int, [C.POINTER(SsmModel), C.c_int64, C.c_int32, c_dp, c_dp, C.c_double,
                                            c_dp, c_dp, c_dp]),
    "srcb200_tpwl_nearest_batch": (C.c_int, [C.POINTER(TpwlModel), C.c_int64, c_dp, c_dp, c_dp, c_dp]),
    "srcb200_tpwl_weights_batch": (C.c_int, [C.POINTER(TpwlModel), C.c_int64, c_dp, c_dp, c_dp]),
    "srcb200_tpwl_linearize_workspace": (C.c_size_t, [C.POINTER(TpwlModel), C.c_int64]),
    "srcb200_tpwl_linearize_batch": (C.c_int, [C.POINTER(TpwlModel), C.c_int64, c_dp, C.c_double, c_dp, c_dp,
                                               c_dp, c_dp, c_dp, C.c_size_t, c_dp]),
    "srcb200_tpwl_rollout_workspace": (C.c_size_t, [C.POINTER(TpwlModel), C.c_int64]),
    "srcb200_tpwl_rollout_batch": (C.c_int, [C.POINTER(TpwlModel), C.c_int64, C.c_int32, c_dp, c_dp, C.c_double,
                                             c_dp, c_dp, c_dp, c_dp, C.c_size_t, c_dp]),
    "srcb200_tpwl_output_batch": (C.c_int, [C.POINTER(TpwlModel), C.c_int64, c_dp, c_dp, c_dp]),
    "srcb200_discretize_batch": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_double, c_dp, c_dp,
                                           c_dp, c_dp, c_dp, c_dp, c_dp]),
    "srcb200_zoh_workspace": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "srcb200_zoh_batch": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                    c_dp, C.c_size_t, c_dp]),
    "srcb200_ilqr_workspace_bytes": (C.c_size_t, [C.c_int32, c_dp, C.POINTER(IlqrProblem)]),
    "srcb200_ilqr_solve_batch": (C.c_int, [C.c_int32, c_dp, C.POINTER(IlqrConfig), C.POINTER(IlqrProblem),
                                           C.POINTER(IlqrResult), c_dp, C.c_size_t, c_dp]),
    "srcb200_ilqr_forward_pass": (C.c_int, [C.c_int32, c_dp, C.POINTER(IlqrConfig), C.POINTER(IlqrProblem),
                                            c_dp, c_dp, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                            c_dp, c_dp, C.c_size_t, c_dp]),
    "srcb200_ilqr_backward_pass": (C.c_int, [C.c_int32, c_dp, C.POINTER(IlqrConfig), C.POINTER(IlqrProblem),
                                             c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                             c_dp, C.c_size_t, c_dp]),
    "srcb200_pod_gram": (C.c_int, [C.c_int64, C.c_int64, c_dp, C.c_int64, c_dp, C.c_int64, C.c_int32, c_dp]),
    "srcb200_sym_eig_psd": (C.c_int, [C.c_int32, c_dp, C.c_int64, c_dp, c_dp, C.c_int64, c_dp, c_dp]),
    "srcb200_dgemm": (C.c_int, [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_double, c_dp, C.c_int64, c_dp,
                                C.c_int64, c_dp, C.c_int64, c_dp]),
    "srcb200_ekf_predict_batch": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                            c_dp]),
    "srcb200_ekf_update_batch": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "srcb200_dlqr_riccati_batch": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, c_dp, c_dp, c_dp, c_dp, C.c_int32,
                                             C.c_double, C.c_int32, C.c_int32, c_dp, c_dp, c_dp, c_dp]),
    "srcb200_tvlqr_batch": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                      c_dp]),
    "srcb200_tpwl_bank_point_batch": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                                c_dp, c_dp, c_dp, c_dp]),
    "srcb200_gusto_accuracy_batch": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_double, c_dp, c_dp, c_dp,
                                               c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "srcb200_mpc_shift_batch": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_dp, c_dp,
                                          c_dp, c_dp, c_dp, c_dp, c_dp]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


class Srcb200Error(RuntimeError):
    """A libsrcb200 call failed (negative code = argument error, positive = cudaError_t)."""

    def __init__(self, code, msg):
        super().__init__("libsrcb200 error %d: %s" % (code, msg))
        self.code = code


def library_path():
    return _build.LIBPATH


def lib():
    """Loads (never builds implicitly on a box without nvcc) libsrcb200.so and declares the signatures."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("libsrcb200.so is not built (%s missing). Run `python -c \"import __graft_entry__ as g; "
                           "g.build()\"` -- there is no CPU fallback for the hot path." % path)
    L = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)   # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if L.srcb200_abi_version() != ABI_VERSION:
        raise RuntimeError("libsrcb200 ABI mismatch")
    _LIB = L
    return L


def check(code):
    if code != 0:
        msg = lib().srcb200_last_error_string()
        msg = msg.decode() if msg else ""
        if code == E_METHOD:
            raise RuntimeError(msg)          # the reference raises RuntimeError for bad discr/tpwl methods
        raise Srcb200Error(code, msg)


_DEVICE_OK = False


def require_gpu():
    """Fails loudly when there is no B200 / the kernels cannot load -- the product path has no fallback."""
    global _DEVICE_OK
    if _DEVICE_OK:
        return
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("sofacontrol_b200 needs a CUDA device (sm_100a); no CPU fallback exists for the hot path")
    torch.cuda.current_device()
    check(lib().srcb200_device_check())
    _DEVICE_OK = True


# ---- device-buffer helpers ----------------------------------------------------------------------------------
def torch_mod():
    import torch
    return torch


def to_dev(a, dtype=None):
    """numpy / list / torch tensor -> contiguous CUDA tensor (float64 unless dtype given)."""
    require_gpu()
    torch = torch_mod()
    dtype = dtype or torch.float64
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous()
    np_dtype = {torch.float64: np.float64, torch.int32: np.int32, torch.uint8: np.uint8}[dtype]
    arr = np.ascontiguousarray(np.asarray(a), dtype=np_dtype)
    if not arr.flags.writeable:
        arr = arr.copy()
    return torch.from_numpy(arr).cuda()


def empty(shape, dtype=None):
    require_gpu()
    torch = torch_mod()
    return torch.empty(shape, device="cuda", dtype=dtype or torch.float64)


def zeros(shape, dtype=None):
    torch = torch_mod()
    return torch.zeros(shape, device="cuda", dtype=dtype or torch.float64)


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr():
    return torch_mod().cuda.current_stream().cuda_stream


def to_host(t):
    return t.cpu().numpy()
