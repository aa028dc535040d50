"""Batched receding-horizon iLQR (closed-loop Monte Carlo), BASELINE.json configs[3].

The reference's `ilqr` controller is single shot (policy recomputed only at t_step == 0,
sofacontrol/tpwl/controllers.py:59-66, 185-206); the receding-horizon variant is only hinted at by the commented
`rh_ilqr` (examples/hardware/diamond.py:564-566) and by the warm-start hooks `u_warmstart` / `set_u_last`
(sofacontrol/lqr/ilqr.py:24-27, 46-47, 145-149).  This driver assembles exactly those hooks, for a whole batch:

    every control step k:   x0      = current states (Bt, n)
                            target  = z_ref[k : k + N + 1]
                            u_init  = previous plan shifted by one step (last input repeated)
                            u_last  = input applied at step k-1
                            solve Bt iLQR problems (one launch of the iLQR kernel), apply u[:, 0] to the plant
                            (the same reduced-order model stepped once on the device) + optional process noise.

Everything stays on the device; torch is used for slicing / shifting buffers and for the noise draw.
"""
import numpy as np

from . import _lib as L


class RecedingHorizonILQR:
    def __init__(self, solver, plant=None, process_noise_std=0.0, seed=4):
        """solver: sofacontrol_b200.lqr.ilqr.iLQR (planning_horizon = N); plant: model stepped in closed loop
        (defaults to the solver's model)."""
        self.solver = solver
        self.plant = plant if plant is not None else solver.model
        self.noise = float(process_noise_std)
        self.seed = seed

    def run(self, x0, z_ref, steps):
        """x0 (Bt, n); z_ref (steps + N + 1, n_z) shared or (Bt, steps + N + 1, n_z); returns a dict of host arrays:
        x (Bt, steps+1, n), u (Bt, steps, m), iterations (Bt, steps), cost (Bt, steps), status (Bt, steps)."""
        out = self.run_device(L.to_dev(np.asarray(x0, dtype=np.float64)), L.to_dev(np.asarray(z_ref, dtype=np.float64)), steps)
        return {k: L.to_host(v) for k, v in out.items()}

    def run_device(self, x0, z_ref, steps):
        torch = L.torch_mod()
        s = self.solver
        N, n, m = int(s.planning_horizon), s.state_dim, s.input_dim
        Bt = x0.shape[0]
        shared = (z_ref.dim() == 2)
        need = steps + N + 1
        if z_ref.shape[-2] < need:
            raise ValueError("z_ref needs at least steps + N + 1 = %d rows" % need)
        xs = L.empty((Bt, steps + 1, n))
        us = L.empty((Bt, steps, m))
        its = L.empty((Bt, steps), torch.int32)
        costs = L.empty((Bt, steps))
        stat = L.empty((Bt, steps), torch.int32)
        xs[:, 0] = x0
        x = x0.contiguous()
        u_plan = None
        u_last = L.zeros((Bt, m))
        gen = torch.Generator(device="cuda").manual_seed(self.seed)
        for k in range(steps):
            zt = (z_ref[k:k + N + 1] if shared else z_ref[:, k:k + N + 1]).contiguous()
            u_init = None
            if u_plan is not None:
                u_init = torch.cat((u_plan[:, 1:], u_plan[:, -1:]), dim=1).contiguous()   # shifted warm start
            sol = s.solve_device(x, zt, u_init, u_last)
            u_plan = sol['u']
            u0 = u_plan[:, 0].contiguous()
            xn, _ = self.plant.rollout_device(x, u0[:, None, :].contiguous(), s.dt, want_z=False)
            x = xn[:, 1].contiguous()
            if self.noise > 0.0:
                x = x + self.noise * torch.randn(x.shape, device="cuda", dtype=torch.float64, generator=gen)
            u_last = u0
            xs[:, k + 1] = x
            us[:, k] = u0
            its[:, k] = sol['iterations']
            costs[:, k] = sol['cost']
            stat[:, k] = sol['status']
        return dict(x=xs, u=us, iterations=its, cost=costs, status=stat)
