"""CPU-only: the C-ABI library loads and exports every symbol include/srcb200.h declares; host-side logic of the
Python boundary (struct layouts, monomial tables, synthetic generators, loud failure without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(REPO, "include", "srcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srcb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sofacontrol_b200 import _lib
    names = _header_functions()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.library_path())
    for n in names:
        assert hasattr(lib, n), "missing symbol %s" % n
    assert sorted(_lib.EXPORTED_SYMBOLS) == names          # the ctypes table binds exactly the header's API
    assert _lib.lib().srcb200_abi_version() == _lib.ABI_VERSION == 3


def test_struct_layouts_match_header_sizes():
    from sofacontrol_b200 import _lib
    assert ctypes.sizeof(_lib.SsmModel) == 6 * 4 + 6 * 8
    assert ctypes.sizeof(_lib.TpwlModel) == 6 * 4 + 3 * 8 + 7 * 8
    assert ctypes.sizeof(_lib.IlqrConfig) == 6 * 4 + 12 * 8
    assert ctypes.sizeof(_lib.IlqrProblem) == 8 + 4 + 4 + 8 + 8 * 8 + 8
    assert ctypes.sizeof(_lib.IlqrResult) == 10 * 8


def test_argument_errors_without_gpu():
    """Argument validation happens before any launch, so the error contract is testable without a device."""
    from sofacontrol_b200 import _lib
    L = _lib.lib()
    assert L.srcb200_ssm_rollout_batch(None, 1, 1, None, None, 0.1, None, None, None) == _lib.E_NULL
    bad = _lib.SsmModel(n=9, m=4, nz=9, order=3, nfeat=83, discr_method=0)
    assert L.srcb200_ssm_eval_linearize_batch(bad, 1, None, None, 0.1, None, None, None, None, None, None, None) == _lib.E_DIM
    assert b"dims out of range" in L.srcb200_last_error_string()
    zoh = _lib.SsmModel(n=6, m=4, nz=6, order=3, nfeat=83, discr_method=3, r_coeff=8, w_coeff=8, v_coeff=8, B_r=8, z_ref=8, mono=8)
    assert L.srcb200_ssm_eval_linearize_batch(zoh, 1, None, None, 0.1, None, None, None, None, None, None, None) == _lib.E_METHOD
    with pytest.raises(RuntimeError):
        _lib.check(_lib.E_METHOD)
    t = _lib.TpwlModel(n=7, m=2, nz=0, P=4)
    assert L.srcb200_tpwl_nearest_batch(t, 1, None, None, None, None) == _lib.E_DIM
    assert L.srcb200_dgemm(0, 4, 4, 4, 1.0, None, 4, None, 4, None, 4, None) == _lib.E_NULL
    assert L.srcb200_pod_gram(10, 4, 8, 2, 8, 4, 0, None) == _lib.E_DIM     # ldx < ns


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    s = synth.trunk_ssm(4)
    m = SSMDynamics(s['z_ref'], model=s['model'], params=s['params'])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.get_jacobians(np.zeros(6), np.zeros(4), 0.01)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.rollout(np.zeros(6), np.zeros((3, 4)), 0.01)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "soft-robot-control_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_monomial_table_and_counts():
    from sofacontrol_b200.SSM.ssm import monomial_table
    from sofacontrol_b200 import synth
    from oracle.ssm_np import monomial_index_table
    for dim, order in ((6, 3), (3, 2), (4, 4), (8, 3)):
        t = monomial_table(dim, order)
        o = monomial_index_table(dim, order)
        assert t.shape == (synth.num_monomials(dim, order), 4)
        assert np.array_equal(np.where(t[:, :order] == 0xFF, -1, t[:, :order].astype(np.int64)), o)
    assert synth.num_monomials(6, 3) == 83


def test_ssm_class_surface_and_struct_parsing():
    from sofacontrol_b200 import synth
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    s = synth.trunk_ssm(8)
    m = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    assert (m.state_dim, m.input_dim, m.output_dim, m.SSM_order, m.ROM_order) == (6, 8, 6, 3, 3)
    assert m.r_coeff.shape == (6, 83) and m.B_r.shape == (6, 8) and m.Ts == 0.01
    assert m.maps['f_nl'] == m.reduced_dynamics and 'f_nl_d' not in m.maps
    assert m.C_map == m.reduced_to_observed and m.W_map == m.observed_to_reduced
    assert np.array_equal(m.get_ref_point(), s['z_ref'])
    for name in ("update_state", "get_jacobians", "get_continuous_jacobians", "get_discrete_jacobians",
                 "get_observer_jacobians", "update_observer_state", "discretize_dynamics", "update_dynamics", "rollout",
                 "x_to_zfyf", "x_to_zy", "zfyf_to_zy", "zy_to_zfyf", "compute_RO_state", "get_state_dim"):
        assert callable(getattr(m, name))
    A, B, d = np.eye(6), np.ones((6, 8)), np.arange(6.0)
    assert np.array_equal(SSMDynamics.update_dynamics(np.ones(6), np.ones(8), A, B, d), np.ones(6) + 8.0 + d)


def test_tpwl_class_surface_host_side():
    from sofacontrol_b200 import synth
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    data, Hf = synth.tpwl_bank(seed=1, r=4, m=2, P=9, num_nodes=6, tip_node=1)
    m = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, Cf=Hf, discr_method='be')
    assert m.state_dim == 8 and m.input_dim == 2 and m.num_points == 9 and m.output_dim == 6 and m.meas_dim == 6
    assert m.H.shape == (6, 8) and m.C.shape == (6, 8) and m.nonlinear_observer is False
    assert m.pre_discretized_dt is None and m.ref_point is None
    assert m.get_sim_params() == {'beta_weighting': None, 'discr_method': 'be', 'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}
    assert np.array_equal(m.zy_to_zfyf(z=np.zeros(6)), m.z_ref)
    bad = dict(data); bad['rom_info'] = dict(data['rom_info'], type='other')
    with pytest.raises(NotImplementedError):
        TPWLATV(bad)


def test_ilqr_config_defaults_match_reference_fields():
    from sofacontrol_b200.lqr.config import iLQRConfig
    from oracle.ilqr_np import Config
    a, b = vars(iLQRConfig()), vars(Config())
    for k, v in b.items():
        assert a[k] == v
    assert set(a) == set(b)
