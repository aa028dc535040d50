"""TPWL bank construction on B200 -- the numerical part of sofacontrol/tpwl/tpwl_utils.py (TPWLSnapshotData.add_point,
add_continuous_TPWL: 84-117, 263-276) and its target containers.  The SOFA-driven snapshot collection
(save_snapshot / evaluate_point, driven by the simulator) is out of scope; `TPWLSnapshotData.add_point` takes the
full-order matrices of a snapshot, projects them onto the POD basis (compute_RO_matrix: DMMA GEMMs) and appends the
continuous-time linearisation (A_c, B_c, d_c) built by csrc/control.cu: bank_point_kernel; `add_points` does a
whole list of snapshots in one launch.
"""
import numpy as np

from .. import _lib as L
from .. import utils as scutils


class Target:
    """tpwl_utils.py:5-25."""

    def __init__(self):
        self.t = None
        self.u = None
        self.z = None
        self.x = None
        self.Hf = None

    def load_target_file(self, file):
        data = scutils.load_data(file)
        self.t = data.get('t')
        self.u = data.get('u')
        self.z = data.get('z')
        self.Hf = data.get('Hf')


class DynamicsTarget(Target):
    """tpwl_utils.py:28-38."""

    def __init__(self):
        super(DynamicsTarget, self).__init__()
        self.A = None
        self.B = None
        self.x = None


def continuous_tpwl_points(K, D, M, H, f, q):
    """(A_c, B_c, d_c) of stored points from reduced matrices K, D, M (count, r, r), H (count, r, m), f, q (count, r):
    extract_AB (utils.py:251-286) + the affine term of add_continuous_TPWL (tpwl_utils.py:263-276)."""
    L.require_gpu()
    K = np.asarray(K, dtype=np.float64)
    single = (K.ndim == 2)
    r, m = K.shape[-1], np.asarray(H).shape[-1]
    dev = [L.to_dev(np.asarray(a, dtype=np.float64).reshape((-1,) + s)) for a, s in
           ((K, (r, r)), (D, (r, r)), (M, (r, r)), (H, (r, m)), (f, (r,)), (q, (r,)))]
    cnt = dev[0].shape[0]
    A, B, d = L.empty((cnt, 2 * r, 2 * r)), L.empty((cnt, 2 * r, m)), L.empty((cnt, 2 * r))
    L.check(L.lib().srcb200_tpwl_bank_point_batch(r, m, cnt, *[L.ptr(a) for a in dev], L.ptr(A), L.ptr(B), L.ptr(d),
                                                  L.stream_ptr()))
    res = (L.to_host(A), L.to_host(B), L.to_host(d))
    return tuple(a[0] for a in res) if single else res


class SnapshotPoint:
    """The fields of utils.Point that add_point reads (utils.py:21-50): full-order state, input and matrices."""

    def __init__(self, t=0.0, q=None, v=None, u=None, K=None, D=None, M=None, H=None, b=None, f=None, S=None,
                 q_next=None, v_next=None, dt=-1):
        self.t, self.q, self.v, self.u = t, q, v, u
        self.K, self.D, self.M, self.H, self.b, self.f, self.S = K, D, M, H, b, f, S
        self.q_next, self.v_next, self.dt = q_next, v_next, dt


class TPWLSnapshotData:
    """tpwl_utils.py:41-117, 263-276: collects the stored points of a TPWL model in the reference's dict schema
    (utils.py:53-68 + tpwl_utils.py:53-61)."""

    def __init__(self, rom, config=None, info=None, Hf=None):
        self.dict = {k: [] for k in ('t', 'q', 'v', 'u', 'q+', 'v+', 'K', 'D', 'M', 'S', 'H', 'b', 'f', 'A_c', 'B_c',
                                     'd_c', 'A_d', 'B_d', 'd_d', 'z', 'z_est')}
        self.dict['dt'] = -1
        self.rom = rom
        self.dict['rom_info'] = self.rom.get_info()
        self.config = config
        self.info = dict() if info is None else info
        self.saved_tpwl_steps = []
        self.Hf = Hf

    def add_point(self, point):
        """tpwl_utils.py:84-117 (save_continuous_TPWL path)."""
        self.add_points([point])

    def add_points(self, points):
        if self.dict['dt'] == -1 and points:
            self.dict['dt'] = points[0].dt
        rom = self.rom
        for p in points:
            self.saved_tpwl_steps.append(p.t)
            self.dict['q'].append(rom.compute_RO_state(qf=p.q))
            self.dict['v'].append(rom.compute_RO_state(vf=p.v))
            self.dict['u'].append(p.u)
            self.dict['K'].append(rom.compute_RO_matrix(p.K))
            self.dict['D'].append(rom.compute_RO_matrix(p.D))
            self.dict['M'].append(rom.compute_RO_matrix(p.M))
            self.dict['b'].append(rom.compute_RO_matrix(p.b, left=True) if p.b is not None else None)
            self.dict['f'].append(rom.compute_RO_matrix(p.f, left=True))
            self.dict['H'].append(rom.compute_RO_matrix(p.H, left=True))
            self.dict['S'].append(rom.compute_RO_matrix(p.S) if p.S is not None else None)
            if p.q_next is not None:
                self.dict['q+'].append(rom.compute_RO_state(qf=p.q_next))
                self.dict['v+'].append(rom.compute_RO_state(vf=p.v_next))
        k = len(points)
        if k:
            sl = slice(len(self.dict['K']) - k, None)
            A, B, d = continuous_tpwl_points(np.array(self.dict['K'][sl]), np.array(self.dict['D'][sl]),
                                             np.array(self.dict['M'][sl]), np.array(self.dict['H'][sl]),
                                             np.array(self.dict['f'][sl]), np.array(self.dict['q'][sl]))
            self.dict['A_c'] += list(A)
            self.dict['B_c'] += list(B)
            self.dict['d_c'] += list(d)

    def as_arrays(self):
        """utils.dict_lists_to_array (utils.py:338-344) on the fields a TPWL model reads."""
        out = dict(self.dict)
        for k in ('q', 'v', 'u', 'A_c', 'B_c', 'd_c'):
            out[k] = np.asarray(out[k])
        return out
