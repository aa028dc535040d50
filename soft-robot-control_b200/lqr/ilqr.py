"""Batched iLQR on B200 -- drop-in for sofacontrol/lqr/ilqr.py (class iLQR).

`iLQR(dt, model, cost_params, planning_horizon, **kwargs)`, `set_target`, `set_u_last`,
`ilqr_computation(x0, u_warmstart=None) -> (x, u, K)`, `forward_pass`, `dlqr_recursion`, the cost helpers and
`update_regularization` keep the reference's names, argument meaning and return shapes (ilqr.py:6-300).  All of them
run csrc/ilqr.cu (one CTA per problem) -- no numpy arithmetic on the host.

Extension (BASELINE configs 3/4): a leading batch axis.  `x0 (Bt, n)`, `z_target (Bt, N+1, n_z)` or a shared
`(N+1, n_z)`, `u_warmstart (Bt, N, m)`, `u_last (Bt, m)` solve Bt independent problems in one launch and return
`x (Bt, N+1, n)`, `u (Bt, N, m)`, `K (Bt, N, m, n)`; per-problem `cost`, `iterations`, `status`, `rho`, `trials`
land in `self.info`.

kwargs (new, all optional):
    gauss_newton : None (default) -> True for models with `nonlinear_observer` (SSM: H_t = dC/dx at x_t, the
                   H-property adapter of SURVEY.md App. C.2 -- with the literal constant zero `model.H` of the SSM
                   class the reference solve is degenerate), False for TPWL (constant model.H, exactly ilqr.py:177-196).
    trace        : record the per-iteration branch trace in self.info['trace'].
"""
import numpy as np

from .. import _lib as L
from .config import iLQRConfig


def _model_kind(model):
    from ..SSM.ssm import SSM
    from ..tpwl.tpwl import TPWL
    if isinstance(model, SSM):
        return L.ILQR_MODEL_SSM
    if isinstance(model, TPWL):
        return L.ILQR_MODEL_TPWL
    raise TypeError("sofacontrol_b200.lqr.iLQR needs a sofacontrol_b200 SSM or TPWL model (the kernels evaluate "
                    "the model on the device); got %r" % type(model))


class iLQR:
    def __init__(self, dt, model, cost_params, planning_horizon, **kwargs):
        self.params = iLQRConfig()
        self.dt = dt
        self.model = model
        self.planning_horizon = planning_horizon
        self.cost_params = cost_params

        self.state_dim = model.get_state_dim()
        self.input_dim = model.get_input_dim()

        self.z_target = None
        self.u_last = np.zeros(self.input_dim)  # For receding horizon

        self._kind = _model_kind(model)
        gn = kwargs.get('gauss_newton', None)
        self.gauss_newton = bool(getattr(model, 'nonlinear_observer', False)) if gn is None else bool(gn)
        self.want_trace = bool(kwargs.get('trace', False))
        self.rho = self.params.rho0
        self.drho = self.params.drho0
        self.info = {}
        self._ws = None

    def set_target(self, z_target):
        self.z_target = z_target.copy()

    def set_u_last(self, u_last):
        self.u_last = u_last.copy()

    # ---- plumbing -------------------------------------------------------------------------------------------
    def _cfg(self):
        p = self.params
        return L.IlqrConfig(max_iter=int(p.max_iter), include_input_var_constraint=int(bool(p.include_input_var_constraint)),
                            do_linesearch=int(bool(p.do_linesearch)), regularize=int(bool(p.regularize)),
                            state_regularization=int(bool(p.state_regularization)), counter_limit=int(p.counter_limit),
                            epsilon=float(p.epsilon), alpha0=float(p.alpha0), alpha_scaling=float(p.alpha_scaling),
                            improv_lb=float(p.improv_lb), improv_ub=float(p.improv_ub), alpha_min=float(p.alpha_min),
                            rho0=float(p.rho0), drho0=float(p.drho0), rho_scaling=float(p.rho_scaling),
                            rho_increase_fp=float(p.rho_increase_fp), rho_max=float(p.rho_max),
                            rho_min=float(p.rho_min))

    def _model_handle(self):
        if self._kind == L.ILQR_MODEL_SSM:
            return self.model.device_model()
        m = self.model
        if (m.tpwl_method == 'nn' and m.discr_method == 'zoh' and
                not (m.pre_discretized_dt is not None and self.dt == m.pre_discretized_dt)):
            return m._zoh_bank_model(self.dt)      # same matrices the reference recomputes with expm at every step
        return m.device_model(self.dt)

    def _nz(self):
        return int(self.model.get_output_dim())

    def _problem(self, batch, x0=None, u_init=None, z_target=None, u_last=None):
        """Builds the srcb200_ilqr_problem; arguments are CUDA tensors (or None)."""
        nz, n = self._nz(), self.state_dim
        c = self.cost_params
        host = [np.asarray(c.Q, dtype=np.float64), np.asarray(c.R, dtype=np.float64),
                np.asarray(c.Qf if c.Qf is not None else np.zeros((nz, nz)), dtype=np.float64)]
        if not self.gauss_newton:
            H = getattr(self.model, 'H', None)
            host.append(np.asarray(H if H is not None else np.zeros((nz, n)), dtype=np.float64))
        cache = getattr(self, '_cost_cache', None)
        if cache is None or len(cache[0]) != len(host) or not all(np.array_equal(a, b) for a, b in zip(cache[0], host)):
            cache = ([h.copy() for h in host], [L.to_dev(h) for h in host])     # re-upload only when the values change
            self._cost_cache = cache
        keep = dict(Q=cache[1][0], R=cache[1][1], Qf=cache[1][2])
        Hc = cache[1][3] if not self.gauss_newton else None
        shared = int(z_target.dim() == 2)
        pr = L.IlqrProblem(batch=batch, N=int(self.planning_horizon), gauss_newton=int(self.gauss_newton),
                           dt=float(self.dt), x0=L.ptr(x0), u_init=L.ptr(u_init), z_target=L.ptr(z_target),
                           u_last=L.ptr(u_last), Q=L.ptr(keep['Q']), R=L.ptr(keep['R']), Qf=L.ptr(keep['Qf']),
                           H_const=L.ptr(Hc), shared_target=shared)
        pr._keep = (keep, Hc, x0, u_init, z_target, u_last)
        return pr

    def _workspace(self, handle, pr):
        need = int(L.lib().srcb200_ilqr_workspace_bytes(self._kind, C_addr(handle), pr))
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = L.empty((need // 8 + 1,))
        return self._ws

    def _targets_dev(self, batch):
        if self.z_target is None:
            raise RuntimeError('set_target must be called before solving')
        zt = np.asarray(self.z_target, dtype=np.float64)
        N, nz = self.planning_horizon, self._nz()
        if zt.shape[-2:] != (N + 1, nz):
            raise ValueError('z_target must be (N+1, n_z) or (Bt, N+1, n_z)')
        if zt.ndim == 3 and zt.shape[0] != batch:
            raise ValueError('z_target batch mismatch')
        return L.to_dev(zt)

    def _u_last_dev(self, batch):
        ul = np.asarray(self.u_last, dtype=np.float64)
        if ul.ndim == 1:
            ul = np.broadcast_to(ul, (batch, self.input_dim))
        return L.to_dev(ul)

    # ---- the solver -----------------------------------------------------------------------------------------
    def solve_device(self, x0, z_target, u_init=None, u_last=None):
        """Device-resident entry: CUDA tensors x0 (Bt, n), z_target (Bt, N+1, n_z) or (N+1, n_z), optional
        u_init (Bt, N, m), u_last (Bt, m).  Returns a dict of CUDA tensors."""
        L.require_gpu()
        torch = L.torch_mod()
        Bt, N, n, m = x0.shape[0], int(self.planning_horizon), self.state_dim, self.input_dim
        handle = self._model_handle()
        pr = self._problem(Bt, x0, u_init, z_target, u_last)
        ws = self._workspace(handle, pr)
        out = dict(x=L.empty((Bt, N + 1, n)), u=L.empty((Bt, N, m)), K=L.empty((Bt, N, m, n)), cost=L.empty((Bt,)),
                   cost0=L.empty((Bt,)), rho=L.empty((Bt,)), iterations=L.empty((Bt,), torch.int32),
                   status=L.empty((Bt,), torch.int32), trials=L.empty((Bt,), torch.int32))
        if self.want_trace:
            out['trace'] = L.zeros((Bt, int(self.params.max_iter) + 1, 4))
        res = L.IlqrResult(x=L.ptr(out['x']), u=L.ptr(out['u']), K=L.ptr(out['K']), cost=L.ptr(out['cost']),
                           cost0=L.ptr(out['cost0']), rho=L.ptr(out['rho']), iterations=L.ptr(out['iterations']),
                           status=L.ptr(out['status']), trials=L.ptr(out['trials']), trace=L.ptr(out.get('trace')))
        cfg = self._cfg()
        L.check(L.lib().srcb200_ilqr_solve_batch(self._kind, C_addr(handle), cfg, pr, res, L.ptr(ws), ws.numel() * 8,
                                                 L.stream_ptr()))
        return out

    def solve_pinned(self, x0_h, z_target_h, out_h=None, u_init_h=None, u_last_h=None):
        """End-to-end batched solve on PINNED host torch tensors: async H2D of the inputs, one kernel launch, async
        D2H of (x, u, K, cost, iterations, status) into pinned outputs (allocated on first use, reusable), one sync."""
        torch = L.torch_mod()
        up = lambda t: None if t is None else t.cuda(non_blocking=True)
        out = self.solve_device(up(x0_h), up(z_target_h), up(u_init_h), up(u_last_h))
        if out_h is None:
            out_h = {k: torch.empty(out[k].shape, dtype=out[k].dtype, pin_memory=True)
                     for k in ('x', 'u', 'K', 'cost', 'iterations', 'status')}
        for k, v in out_h.items():
            v.copy_(out[k], non_blocking=True)
        torch.cuda.synchronize()
        return out_h

    def solve_pinned_stream(self, batches, depth=2):
        """End-to-end solves of a SEQUENCE of batches on pinned host tensors, double-buffered: the device->host copy of
        batch i (x, u, K, cost, iterations, status: 203 MB at 4096 x 100) runs on a side stream while batch i + 1 is
        already being solved.  `batches`: iterable of (x0_h, z_target_h) pinned tensors.  Yields the pinned output dict
        of every batch in order (each dict is reused `depth` batches later: consume it before asking for that one)."""
        torch = L.torch_mod()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        # pinned output buffers are allocated once per solver and depth (cudaHostAlloc of 2 x 203 MB is far slower than a solve)
        cache = getattr(self, '_pin_rings', None)
        if cache is None:
            cache = self._pin_rings = {}
        ring = cache.setdefault(depth, [None] * depth)      # (pinned outputs, completion event, device outputs kept alive)
        pending = []
        for i, (x0_h, zt_h) in enumerate(batches):
            out = self.solve_device(x0_h.cuda(non_blocking=True), zt_h.cuda(non_blocking=True))
            slot = i % depth
            if ring[slot] is None or any(ring[slot][0][k].shape != out[k].shape for k in ring[slot][0]):
                ring[slot] = [{k: torch.empty(out[k].shape, dtype=out[k].dtype, pin_memory=True)
                               for k in ('x', 'u', 'K', 'cost', 'iterations', 'status')}, torch.cuda.Event(), None]
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                for k, v in ring[slot][0].items():
                    out[k].record_stream(side)
                    v.copy_(out[k], non_blocking=True)
                ring[slot][1].record(side)
            ring[slot][2] = out
            pending.append(slot)
            if len(pending) == depth:                      # hand out the oldest batch once its copy has landed
                s0 = pending.pop(0)
                ring[s0][1].synchronize()
                yield ring[s0][0]
        for s0 in pending:
            ring[s0][1].synchronize()
            yield ring[s0][0]

    def ilqr_computation(self, x0, u_warmstart=None):
        """ilqr.py:27-107.  Returns (x, u, K): the optimal sequence and the stabilising gains of the last backward
        pass.  x0 (n,) solves one problem with the reference's return shapes; x0 (Bt, n) solves a batch."""
        x0 = np.asarray(x0, dtype=np.float64)
        single = (x0.ndim == 1)
        x0d = L.to_dev(x0.reshape(-1, self.state_dim))
        Bt = x0d.shape[0]
        ud = None
        if u_warmstart is not None:
            uw = np.asarray(u_warmstart, dtype=np.float64)
            ud = L.to_dev(np.broadcast_to(uw, (Bt,) + uw.shape[-2:]) if uw.ndim == 2 else uw)
        out = self.solve_device(x0d, self._targets_dev(Bt), ud, self._u_last_dev(Bt))
        host = {k: L.to_host(v) for k, v in out.items()}
        self.info = {k: (v[0] if single else v) for k, v in host.items() if k not in ('x', 'u', 'K')}
        self.rho = self.info['rho']
        x, u, K = host['x'], host['u'], host['K']
        return (x[0], u[0], K[0]) if single else (x, u, K)

    def is_converged_calculation(self, prev_cost, cost):
        """ilqr.py:109-115 (scalar predicate; the kernel applies the same test per problem)."""
        return bool(((prev_cost - cost) < self.params.epsilon) and ((prev_cost - cost) >= 0))

    def forward_pass(self, x_prev, u_prev, alpha=1., K=None, k=None):
        """ilqr.py:117-162 -> (x, u, cost, A, B, d); a leading batch axis on x_prev/u_prev is accepted."""
        L.require_gpu()
        x_prev = np.asarray(x_prev, dtype=np.float64)
        single = (x_prev.ndim == 2)
        N, n, m = int(self.planning_horizon), self.state_dim, self.input_dim
        xp = L.to_dev(x_prev.reshape(-1, N + 1, n))
        Bt = xp.shape[0]
        up = L.to_dev(np.asarray(u_prev, dtype=np.float64).reshape(-1, N, m))
        Kd = None if K is None else L.to_dev(np.asarray(K, dtype=np.float64).reshape(-1, N, m, n))
        kd = None if k is None else L.to_dev(np.asarray(k, dtype=np.float64).reshape(-1, N, m))
        handle = self._model_handle()
        pr = self._problem(Bt, None, None, self._targets_dev(Bt), self._u_last_dev(Bt))
        ws = self._workspace(handle, pr)
        x, u, cost = L.empty((Bt, N + 1, n)), L.empty((Bt, N, m)), L.empty((Bt,))
        A, B, d = L.empty((Bt, N, n, n)), L.empty((Bt, N, n, m)), L.empty((Bt, N, n))
        L.check(L.lib().srcb200_ilqr_forward_pass(self._kind, C_addr(handle), self._cfg(), pr, L.ptr(xp), L.ptr(up),
                                                  float(alpha), L.ptr(Kd), L.ptr(kd), L.ptr(x), L.ptr(u), L.ptr(cost),
                                                  L.ptr(A), L.ptr(B), L.ptr(d), L.ptr(ws), ws.numel() * 8, L.stream_ptr()))
        res = [L.to_host(t) for t in (x, u, cost, A, B, d)]
        return tuple(r[0] for r in res) if single else tuple(res)

    def dlqr_recursion(self, x, u, A, B, d):
        """ilqr.py:219-300 -> (K, k, Q_u, Q_uu); updates self.rho / self.drho like the reference."""
        L.require_gpu()
        torch = L.torch_mod()
        x = np.asarray(x, dtype=np.float64)
        single = (x.ndim == 2)
        N, n, m = int(self.planning_horizon), self.state_dim, self.input_dim
        xd = L.to_dev(x.reshape(-1, N + 1, n))
        Bt = xd.shape[0]
        ud = L.to_dev(np.asarray(u, dtype=np.float64).reshape(-1, N, m))
        Ad = L.to_dev(np.asarray(A, dtype=np.float64).reshape(-1, N, n, n))
        Bd = L.to_dev(np.asarray(B, dtype=np.float64).reshape(-1, N, n, m))
        handle = self._model_handle()
        pr = self._problem(Bt, None, None, self._targets_dev(Bt), self._u_last_dev(Bt))
        ws = self._workspace(handle, pr)
        K, k = L.empty((Bt, N, m, n)), L.empty((Bt, N, m))
        Qu, Quu = L.empty((Bt, N, m)), L.empty((Bt, N, m, m))
        rho = L.to_dev(np.broadcast_to(np.asarray(self.rho, dtype=np.float64), (Bt,)))
        drho = L.to_dev(np.broadcast_to(np.asarray(self.drho, dtype=np.float64), (Bt,)))
        pd_fail = L.empty((Bt,), torch.int32)
        L.check(L.lib().srcb200_ilqr_backward_pass(self._kind, C_addr(handle), self._cfg(), pr, L.ptr(xd), L.ptr(ud),
                                                   L.ptr(Ad), L.ptr(Bd), L.ptr(K), L.ptr(k), L.ptr(Qu), L.ptr(Quu),
                                                   L.ptr(rho), L.ptr(drho), L.ptr(pd_fail), L.ptr(ws), ws.numel() * 8,
                                                   L.stream_ptr()))
        rho_h, drho_h = L.to_host(rho), L.to_host(drho)
        self.rho, self.drho = (rho_h[0], drho_h[0]) if single else (rho_h, drho_h)
        self.info['pd_fail_step'] = L.to_host(pd_fail)     # -1: every Q_uu~ was PD
        res = [L.to_host(t) for t in (K, k, Qu, Quu)]
        return tuple(r[0] for r in res) if single else tuple(res)

    # ---- cost helpers (ilqr.py:164-196): evaluated through a zero-gain forward pass / the model's device maps
    def terminal_cost(self, x):
        z = self.model.x_to_zfyf(x, zf=True)
        e = z - np.asarray(self.z_target)[-1, :]
        Qf = np.asarray(self.cost_params.Qf)
        return .5 * e.T @ Qf @ e

    def step_cost(self, x, u, step, u_prev_step=None):
        z = self.model.x_to_zfyf(x, zf=True)
        e = z - np.asarray(self.z_target)[step, :]
        du = u if u_prev_step is None else (u - u_prev_step)
        return .5 * e.T @ self.cost_params.Q @ e + .5 * du.T @ self.cost_params.R @ du

    def _H_at(self, x):
        """The output Jacobian the cost derivatives use: constant model.H (ilqr.py:178,187 as written), or in
        Gauss-Newton mode H(x) = dC/dx from the model's device evaluation (SURVEY.md App. C.2 adapter)."""
        if self.gauss_newton:
            return np.asarray(self.model.get_observer_jacobians(np.asarray(x, dtype=np.float64))[0])
        return np.asarray(self.model.H)

    def terminal_cost_vectors(self, x):
        """ilqr.py:177-182 -> (c, c_x, c_xx) with Qf.  Single-state helper like the reference's; the solver kernels
        form the same quantities per step on the device."""
        z = self.model.x_to_zfyf(x, zf=True)
        H = self._H_at(x)
        e = z - np.asarray(self.z_target)[-1, :]
        Qf = np.asarray(self.cost_params.Qf)
        return (.5 * e.T @ Qf @ e, H.T @ Qf @ e, H.T @ Qf @ H)

    def step_cost_vectors(self, x, u, step, u_prev_step=None):
        """ilqr.py:184-196 -> (c, c_x, c_xx, c_u, c_uu)."""
        z = self.model.x_to_zfyf(x, zf=True)
        H = self._H_at(x)
        e = z - np.asarray(self.z_target)[step, :]
        Q, R = np.asarray(self.cost_params.Q), np.asarray(self.cost_params.R)
        du = u if u_prev_step is None else (u - u_prev_step)
        return (.5 * e.T @ Q @ e + .5 * du.T @ R @ du, H.T @ Q @ e, H.T @ Q @ H, R @ du, R)

    def update_regularization(self, increase=True):
        """ilqr.py:198-217 on the host-side scalars (the kernel carries its own per-problem copy); keeps the
        reference's `dhro` typo: drho never shrinks."""
        p = self.params
        if increase:
            self.drho = np.max((self.drho * p.rho_scaling, p.rho_scaling))
            self.rho = np.max((self.rho * self.drho, p.rho_min))
            if self.rho > p.rho_max:
                self.rho = p.rho_max
        else:
            self.dhro = np.min((self.drho / p.rho_scaling, 1.0 / p.rho_scaling))
            self.rho = self.rho * self.dhro
            if self.rho <= p.rho_min:
                self.rho = p.rho_min


def C_addr(struct):
    """Address of a ctypes structure as a void* argument."""
    import ctypes
    return ctypes.addressof(struct)
