"""Batched receding-horizon iLQR (closed-loop Monte Carlo), BASELINE.json configs[3].

The reference's `ilqr` controller is single shot (policy recomputed only at t_step == 0,
sofacontrol/tpwl/controllers.py:59-66, 185-206); the receding-horizon variant is only hinted at by the commented
`rh_ilqr` (examples/hardware/diamond.py:564-566) and by the warm-start hooks `u_warmstart` / `set_u_last`
(sofacontrol/lqr/ilqr.py:24-27, 46-47, 145-149).  This driver assembles exactly those hooks, for a whole batch, in the
order of TemplateController.evaluate (tpwl/controllers.py:85-117): observer update with the previous input and the
new measurement, then the policy from the belief state, then the input.

    every control step k:   belief  = observer(u_{k-1}, y_k)            (full state, EKF, or SSM output map)
                            target  = z_ref[k : k + N + 1]
                            u_init  = previous plan shifted by one step (last input repeated)
                            u_last  = input applied at step k-1
                            solve Bt iLQR problems (one launch of the iLQR kernel), apply u[:, 0] to the plant
                            (a reduced-order model stepped once on the device) + optional process noise.

Per step: one solver launch, one plant-step launch, the observer's launches and ONE glue launch
(srcb200_mpc_shift_batch: shifted warm start, applied input + log, next target window); all buffers are allocated
once.  torch is used for the noise draw only.
"""
import numpy as np

from . import _lib as L


class FullStateBelief:
    """FullStateObserver (tpwl/observer.py:3-30): the controller sees the plant state."""

    def reset(self, x0):
        self.x = x0

    def update(self, u_prev, x_plant, z_plant, dt):
        self.x = x_plant
        return self.x


class EKFBelief:
    """DiscreteEKFObserver in the loop: y = C x_plant + y_ref (+ measurement noise) -> predict with u_{k-1}, update
    with y_k (tpwl/observer.py:83-126).  `ekf` is a sofacontrol_b200.tpwl.observer.DiscreteEKFObserver."""

    def __init__(self, ekf, meas_noise_std=0.0, seed=5):
        self.ekf = ekf
        self.noise = float(meas_noise_std)
        self.seed = seed

    def reset(self, x0):
        torch = L.torch_mod()
        self.ekf.initialize_reduced(L.to_host(x0))
        self._Ct = L.to_dev(np.ascontiguousarray(np.asarray(self.ekf.C, dtype=np.float64).T))
        self._yref = L.to_dev(np.asarray(self.ekf.dyn_sys.y_ref, dtype=np.float64))
        self._gen = torch.Generator(device="cuda").manual_seed(self.seed)

    def update(self, u_prev, x_plant, z_plant, dt):
        from .mor.pod import dgemm_device
        torch = L.torch_mod()
        y = dgemm_device(x_plant, self._Ct) + self._yref
        if self.noise > 0.0:
            y = y + self.noise * torch.randn(y.shape, device="cuda", dtype=torch.float64, generator=self._gen)
        self.ekf.predict_device(u_prev, dt)
        self.ekf.update_device(y.contiguous())
        return self.ekf.x_dev


class SSMOutputBelief:
    """SSMObserver + compute_RO_state (SSM/controllers.py:186-187, 302-309): the plant's tip output arrives as [v; q],
    the SSM convention is [q; v]; belief x = W_map(z - z_ref) (batched map kernel)."""

    def __init__(self, ssm):
        self.ssm = ssm

    def reset(self, x0):
        torch = L.torch_mod()
        h = self.ssm.get_output_dim() // 2
        self._perm = torch.tensor(list(range(h, 2 * h)) + list(range(h)), device="cuda")
        self.x = x0

    def update(self, u_prev, x_plant, z_plant, dt):
        z = z_plant.index_select(1, self._perm).contiguous()
        self.x = self.ssm._map_device(1, True, z)
        return self.x


class RecedingHorizonILQR:
    def __init__(self, solver, plant=None, observer=None, process_noise_std=0.0, seed=4):
        """solver: sofacontrol_b200.lqr.ilqr.iLQR (planning_horizon = N); plant: model stepped in closed loop
        (defaults to the solver's model); observer: FullStateBelief (default) / EKFBelief / SSMOutputBelief."""
        self.solver = solver
        self.plant = plant if plant is not None else solver.model
        self.observer = observer if observer is not None else FullStateBelief()
        self.noise = float(process_noise_std)
        self.seed = seed

    def run(self, x0, z_ref, steps, x0_plant=None):
        """x0 (Bt, n) initial belief (= plant state unless x0_plant is given); z_ref (steps + N + 1, n_z) shared or
        (Bt, steps + N + 1, n_z); returns a dict of host arrays: x (Bt, steps+1, n_plant), u (Bt, steps, m),
        iterations / cost / status (Bt, steps)."""
        dev = lambda a: None if a is None else L.to_dev(np.asarray(a, dtype=np.float64))
        out = self.run_device(dev(x0), dev(z_ref), steps, dev(x0_plant))
        return {k: L.to_host(v) for k, v in out.items()}

    def run_device(self, x0, z_ref, steps, x0_plant=None):
        torch = L.torch_mod()
        s = self.solver
        N, m = int(s.planning_horizon), s.input_dim
        nz = int(s.model.get_output_dim())
        Bt = x0.shape[0]
        need = steps + N + 1
        if z_ref.shape[-2] < need:
            raise ValueError("z_ref needs at least steps + N + 1 = %d rows" % need)
        if z_ref.dim() == 2:
            z_ref = z_ref[None].expand(Bt, -1, -1)
        z_ref = z_ref[:, :need].contiguous()
        xp = (x0 if x0_plant is None else x0_plant).contiguous()
        npl = xp.shape[1]
        xs = L.empty((Bt, steps + 1, npl))
        us = L.zeros((Bt, steps, m))
        its = L.empty((Bt, steps), torch.int32)
        costs = L.empty((Bt, steps))
        stat = L.empty((Bt, steps), torch.int32)
        xs[:, 0] = xp
        u_warm = [L.zeros((Bt, N, m)), L.zeros((Bt, N, m))]
        u_applied = L.zeros((Bt, m))
        z_win = z_ref[:, :N + 1].contiguous()
        self.observer.reset(x0.contiguous())
        belief = x0.contiguous()
        gen = torch.Generator(device="cuda").manual_seed(self.seed)
        lib = L.lib()
        for k in range(steps):
            sol = s.solve_device(belief, z_win, u_warm[k & 1] if k > 0 else None, u_applied)
            # one glue launch: u_applied = plan[:, 0] (logged), warm start = shifted plan, next target window
            L.check(lib.srcb200_mpc_shift_batch(Bt, N, m, nz, steps, k, L.ptr(sol['u']), L.ptr(z_ref),
                                                L.ptr(u_warm[(k + 1) & 1]), L.ptr(u_applied), L.ptr(z_win), L.ptr(us),
                                                L.stream_ptr()))
            xn, zn = self.plant.rollout_device(xp, u_applied.view(Bt, 1, m), s.dt)
            xp = xn[:, 1]
            if self.noise > 0.0:
                xp = xp + self.noise * torch.randn(xp.shape, device="cuda", dtype=torch.float64, generator=gen)
            xp = xp.contiguous()
            belief = self.observer.update(u_applied, xp, None if zn is None else zn[:, 1].contiguous(), s.dt)
            xs[:, k + 1] = xp
            its[:, k] = sol['iterations']
            costs[:, k] = sol['cost']
            stat[:, k] = sol['status']
        return dict(x=xs, u=us, iterations=its, cost=costs, status=stat)
