"""Import the UNMODIFIED reference modules from /root/reference (only possible in the build container).

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py and by the `-m "not gpu"` tests that pin the
numpy restatements in oracle/*.py against the real reference.  Never imported by the product package.

`sofacontrol/utils.py:5` does `import osqp` at module top; osqp is absent here and only
`Polyhedron(with_reproject=True)` touches it, so an empty stub module is injected.
`sofacontrol/SSM/ssm.py` needs jax (absent here): `load_ssm()` registers the stand-in of oracle/jax_shim.py (numpy
float64 for `jax.numpy`, identity `jit`, exact forward-mode dual-number `jacobian`) and imports the file UNMODIFIED.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("SRC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sofacontrol"))


def load():
    """Returns a namespace with the reference modules: tpwl, pod, ilqr, config, lqr, traj_tracking_lqr, observer, utils,
    measurement_models."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "osqp" not in sys.modules:
        sys.modules["osqp"] = types.ModuleType("osqp")
    if "control" not in sys.modules:            # lqr/lqr.py:1 imports python-control for CLQR only (absent here)
        try:
            import control  # noqa: F401
        except ImportError:
            sys.modules["control"] = types.ModuleType("control")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import numpy as _np
    if not hasattr(_np, "infty"):               # lqr/lqr.py:15 uses the NumPy-1.x alias `np.infty` (removed in 2.0)
        _np.infty = _np.inf
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import sofacontrol.utils as utils
        import sofacontrol.mor.pod as pod
        import sofacontrol.tpwl.tpwl as tpwl
        import sofacontrol.lqr.ilqr as ilqr
        import sofacontrol.lqr.config as config
        import sofacontrol.measurement_models as measurement_models
        import sofacontrol.lqr.lqr as lqr
        import sofacontrol.lqr.traj_tracking_lqr as traj_tracking_lqr
        import sofacontrol.tpwl.observer as observer
    return types.SimpleNamespace(utils=utils, pod=pod, tpwl=tpwl, ilqr=ilqr, config=config, lqr=lqr,
                                 traj_tracking_lqr=traj_tracking_lqr, observer=observer,
                                 measurement_models=measurement_models, root=REFERENCE_ROOT)


def load_ssm():
    """The reference's sofacontrol.SSM.ssm module, imported unmodified on top of oracle/jax_shim.py."""
    load()
    from oracle import jax_shim
    jax_shim.install()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import sofacontrol.SSM.ssm as ssm
    return ssm
