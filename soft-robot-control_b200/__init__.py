"""sofacontrol_b200 -- B200-native (sm_100a) implementation of the soft-robot-control hot path.

Mirrors the reference's module layout for that path only:
    sofacontrol_b200.SSM.ssm      <- sofacontrol/SSM/ssm.py        (SSM, SSMDynamics)
    sofacontrol_b200.tpwl.tpwl    <- sofacontrol/tpwl/tpwl.py      (TPWL, TPWLATV)
    sofacontrol_b200.lqr.ilqr     <- sofacontrol/lqr/ilqr.py       (iLQR)      + lqr.config (iLQRConfig)
    sofacontrol_b200.mor.pod      <- sofacontrol/mor/pod.py        (POD, compute_POD, ...)
    sofacontrol_b200.utils        <- sofacontrol/utils.py          (QuadraticCost, qv2x, x2qv, ...)
Every numerical method runs hand-written CUDA kernels from libsrcb200.so (C ABI in include/srcb200.h) through
ctypes; torch tensors are used only as device buffers.  There is no CPU fallback: without the built library or a
B200 the compute calls raise.
"""
__version__ = "0.1.0"
