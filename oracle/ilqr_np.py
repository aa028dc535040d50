"""numpy restatement of sofacontrol/lqr/ilqr.py (+ lqr/config.py) with a branch trace.
TEST INFRASTRUCTURE ONLY (oracle/__init__.py).  PINNED bit-for-bit against the imported reference class
(tests/test_oracle_vs_reference.py).  Every arithmetic expression keeps the reference's operand order
(Python's left-to-right `*` / `@` chain), because the solver is branchy and the CUDA kernel is compared
against this trace decision by decision.

Reference quirks reproduced on purpose: loop bound `nbr_iter <= max_iter`; the `dhro` typo in
the decrease branch (drho never shrinks); rho gets the scaled increase AND +rho_increase_fp on a line-search
failure; the input-variation penalty ignores cross-time Hessian terms; P is never symmetrised; explicit inverse
of Q_uu_tilde; p/P updates use the un-regularised Q_uu/Q_ux.
"""
import numpy as np


class Config:
    """lqr/config.py:1-31 -- same field names and defaults."""

    def __init__(self):
        self.max_iter = 50
        self.epsilon = 0.1
        self.include_input_var_constraint = True
        self.do_linesearch = True
        self.regularize = True
        self.alpha0 = 1.
        self.alpha_scaling = 0.5
        self.improv_lb = 1e-4
        self.improv_ub = 100
        self.alpha_min = 5e-2
        self.counter_limit = 5
        self.rho0 = 0.
        self.drho0 = 0.
        self.rho_scaling = 1.5
        self.rho_increase_fp = 10.
        self.rho_max = 1e5
        self.rho_min = 1e-3
        self.state_regularization = True


class ILQRNP:
    def __init__(self, dt, model, cost_params, planning_horizon):
        self.params = Config()
        self.dt, self.model, self.cost = dt, model, cost_params
        self.N = planning_horizon
        self.n, self.m = model.get_state_dim(), model.get_input_dim()
        self.z_target = None
        self.u_last = np.zeros(self.m)
        self.trace = []

    def set_target(self, z_target):
        self.z_target = z_target.copy()

    def set_u_last(self, u_last):
        self.u_last = u_last.copy()

    # ---- costs: ilqr.py:164-196
    def _zerr(self, x, step):
        return self.model.x_to_zfyf(x, zf=True) - self.z_target[step, :]

    def terminal_cost(self, x):
        e = self._zerr(x, -1)
        return .5 * e.T @ self.cost.Qf @ e

    def step_cost(self, x, u, step, u_prev_step=None):
        e = self._zerr(x, step)
        du = u if u_prev_step is None else (u - u_prev_step)
        return .5 * e.T @ self.cost.Q @ e + .5 * du.T @ self.cost.R @ du

    def _u_prev(self, u, t):
        if not self.params.include_input_var_constraint:
            return None
        return self.u_last if t == 0 else u[t - 1]

    # ---- forward pass: ilqr.py:117-162
    def forward_pass(self, x_prev, u_prev, alpha=1., K=None, k=None):
        N, n, m = self.N, self.n, self.m
        cost = 0
        x = np.zeros((N + 1, n)); u = np.zeros((N, m))
        A = np.zeros((N, n, n)); B = np.zeros((N, n, m)); d = np.zeros((N, n))
        x[0] = x_prev[0]
        if K is None:
            K = np.zeros((N, m, n))
        if k is None:
            k = np.zeros((N, m))
        for t in range(N):
            u[t] = u_prev[t] + alpha * k[t] + K[t] @ (x[t] - x_prev[t])
            cost += self.step_cost(x[t], u[t], t, u_prev_step=self._u_prev(u, t))
            A[t], B[t], d[t] = self.model.get_jacobians(x[t], u=u[t], dt=self.dt)
            x[t + 1] = self.model.update_dynamics(x[t], u[t], A[t], B[t], d[t])
        cost += self.terminal_cost(x[-1])
        return x, u, cost, A, B, d

    # ---- regularisation schedule: ilqr.py:198-217
    def update_regularization(self, increase=True):
        P = self.params
        if increase:
            self.drho = np.max((self.drho * P.rho_scaling, P.rho_scaling))
            self.rho = np.max((self.rho * self.drho, P.rho_min))
            if self.rho > P.rho_max:
                self.rho = P.rho_max
        else:
            self.dhro = np.min((self.drho / P.rho_scaling, 1.0 / P.rho_scaling))  # sic: `dhro`
            self.rho = self.rho * self.dhro
            if self.rho <= P.rho_min:
                self.rho = P.rho_min

    # ---- backward pass: ilqr.py:219-300
    def dlqr_recursion(self, x, u, A, B, d):
        """Literal control flow of ilqr.py:236-300.  On a failed Cholesky with `regularize` the reference bumps rho
        (ilqr.py:286), `break`s out of the `for` (287) and then FALLS THROUGH to the rho decrease (298) and the
        `break` of the `while` (299): there is no restart (the comment at 287 says otherwise; the code wins).  It
        returns with K[t] = k[t] = 0 for t <= t_fail, Q_u[t_fail] / Q_uu[t_fail] written, Q_u / Q_uu zero below.
        With `regularize=False` a non-PD Q_uu is inverted as it is (no break)."""
        N, n, m = self.N, self.n, self.m
        Pm = self.params
        self._last_pd_fail = -1                     # horizon index of the failed PD test (-1: none); trace only
        Q_u = np.zeros((N, m)); Q_uu = np.zeros((N, m, m))
        K = np.zeros((N, m, n)); k = np.zeros((N, m))
        # terminal_cost_vectors (ilqr.py:177-182): z first, then model.H is read
        e = self._zerr(x[-1], -1)
        H = self.model.H
        P = H.T @ self.cost.Qf @ H
        p = H.T @ self.cost.Qf @ e
        for t in reversed(range(N)):
            # step_cost_vectors (ilqr.py:186-196)
            e = self._zerr(x[t], t)
            H = self.model.H
            c_xx = H.T @ self.cost.Q @ H
            c_x = H.T @ self.cost.Q @ e
            up = self._u_prev(u, t)
            c_u = self.cost.R @ u[t] if up is None else self.cost.R @ (u[t] - up)
            c_uu = self.cost.R
            Q_x = c_x + A[t].T @ p
            Q_u[t] = c_u + B[t].T @ p
            Q_xx = c_xx + A[t].T @ P @ A[t]
            Q_uu[t] = c_uu + B[t].T @ P @ B[t]
            Q_ux = B[t].T @ P @ A[t]
            if Pm.regularize:
                if Pm.state_regularization:
                    Preg = P + self.rho * np.eye(n)
                    Q_uu_t = c_uu + B[t].T @ Preg @ B[t]
                    Q_ux_t = B[t].T @ Preg @ A[t]
                else:
                    Q_uu_t = Q_uu[t] + self.rho * np.eye(m)
                    Q_ux_t = Q_ux
            else:
                Q_uu_t, Q_ux_t = Q_uu[t], Q_ux
            try:
                np.linalg.cholesky(Q_uu_t)
                pos_def = True
            except np.linalg.LinAlgError:
                pos_def = False
            if not pos_def:
                if self._last_pd_fail < 0:
                    self._last_pd_fail = t
                if Pm.regularize:
                    self.update_regularization(increase=True)
                    break                            # ilqr.py:287 -- leaves the for loop only
            inv = np.linalg.inv(Q_uu_t)
            K[t] = - inv @ Q_ux_t
            k[t] = - inv @ Q_u[t]
            p = Q_x + K[t].T @ Q_uu[t] @ k[t] + K[t].T @ Q_u[t] + Q_ux.T @ k[t]
            P = Q_xx + K[t].T @ Q_uu[t] @ K[t] + K[t].T @ Q_ux + Q_ux.T @ K[t]
        self.update_regularization(increase=False)   # ilqr.py:298, reached on BOTH paths
        return K, k, Q_u, Q_uu

    # ---- outer loop: ilqr.py:27-115
    def ilqr_computation(self, x0, u_warmstart=None):
        Pm = self.params
        self.rho, self.drho = Pm.rho0, Pm.drho0
        self.trace = []
        fails = 0
        x_prev = np.zeros((self.N + 1, self.n))
        x_prev[0] = x0
        if u_warmstart is None:
            u_warmstart = np.zeros((self.N, self.m))
        x, u, cost, A, B, d = self.forward_pass(x_prev, u_warmstart)
        self.initial_cost = cost
        conv = False
        it = 0
        K = None
        while not conv and it <= Pm.max_iter:
            K, k, Q_u, Q_uu = self.dlqr_recursion(x, u, A, B, d)
            ev = {'it': it, 'pd_fail_t': self._last_pd_fail, 'rho_after_bwd': float(self.rho), 'trials': []}
            prev_cost = cost
            alpha = Pm.alpha0
            improved = failed = False
            while not improved and not failed:
                improved = True
                xt, ut, ct, At, Bt, dt_ = self.forward_pass(x, u, alpha=alpha, K=K, k=k)
                delta_cost = 0
                for t in range(self.N):
                    delta_cost += alpha * k[t].T @ Q_u[t] + alpha ** 2 * .5 * k[t].T @ Q_uu[t] @ k[t]
                ratio = None
                if Pm.do_linesearch:
                    ratio = (ct - prev_cost) / delta_cost
                    if ratio <= Pm.improv_lb or ratio > Pm.improv_ub:
                        alpha = Pm.alpha_scaling * alpha
                        improved = False
                        if alpha < Pm.alpha_min:
                            self.update_regularization(increase=True)
                            self.rho += Pm.rho_increase_fp
                            failed = True
                ev['trials'].append((float(ct), float(delta_cost), None if ratio is None else float(ratio)))
            if not failed:
                x, u, cost, A, B, d = xt, ut, ct, At, Bt, dt_
                conv = ((prev_cost - cost) < Pm.epsilon) and ((prev_cost - cost) >= 0)
                fails = 0
            else:
                fails += 1
                if fails >= Pm.counter_limit:
                    conv = True
            ev.update(accepted=not failed, cost=float(cost), rho=float(self.rho), conv=bool(conv))
            self.trace.append(ev)
            it += 1
        self.iterations = it
        self.final_cost = cost
        return x, u, K


def riccati_noise_floor(model, cost, z_target, x, u, A, B, u_last=None, rho=0.0):
    """How far is the reference's own FP64 backward pass from exact arithmetic?  Runs the value recursion of
    ilqr.py:258-295 (fixed rho) once in float64 and once in numpy longdouble (x87 80-bit here) on identical inputs
    and returns the relative differences of K and k.  The un-symmetrised recursion amplifies rounding noise along
    the horizon (measured ~1e6 over N = 100 on the Trunk figure-8), which bounds how closely ANY independent
    implementation can reproduce the reference's gains -- see DESIGN.md "Parity tolerance"."""
    N, m = u.shape
    n = x.shape[1]
    Hs, Es = [], []
    for t in range(N + 1):
        z = model.x_to_zfyf(x[t], zf=True)
        Hs.append(np.array(model.H, dtype=np.float64))
        Es.append(z - z_target[t])
    u_last = np.zeros(m) if u_last is None else u_last

    def run(dt):
        c = lambda a: np.asarray(a, dtype=dt)
        Q, R, Qf = c(cost.Q), c(cost.R), c(cost.Qf)
        P = c(Hs[N]).T @ Qf @ c(Hs[N])
        p = c(Hs[N]).T @ Qf @ c(Es[N])
        Ks, ks = [], []
        for t in reversed(range(N)):
            H, e, At, Bt = c(Hs[t]), c(Es[t]), c(A[t]), c(B[t])
            c_xx, c_x = H.T @ Q @ H, H.T @ Q @ e
            c_u = R @ c(u[t] - (u_last if t == 0 else u[t - 1]))
            Q_x, Q_u = c_x + At.T @ p, c_u + Bt.T @ p
            Q_xx, Q_uu, Q_ux = c_xx + At.T @ P @ At, R + Bt.T @ P @ Bt, Bt.T @ P @ At
            Preg = P + c(rho) * np.eye(n, dtype=dt)           # state regularisation (ilqr.py:266-267)
            Q_uu_t, Q_ux_t = R + Bt.T @ Preg @ Bt, Bt.T @ Preg @ At
            inv = c(np.linalg.inv(np.asarray(Q_uu_t, dtype=np.float64)))
            if dt is not np.float64:
                for _ in range(3):                     # Newton refinement of the inverse in extended precision
                    inv = inv + inv @ (np.eye(m, dtype=dt) - Q_uu_t @ inv)
            K, k = -inv @ Q_ux_t, -inv @ Q_u
            p = Q_x + K.T @ Q_uu @ k + K.T @ Q_u + Q_ux.T @ k
            P = Q_xx + K.T @ Q_uu @ K + K.T @ Q_ux + Q_ux.T @ K
            Ks.append(K); ks.append(k)
        return np.array(Ks[::-1]), np.array(ks[::-1])

    K64, k64 = run(np.float64)
    K80, k80 = run(np.longdouble)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    return rel(K64, K80), rel(k64, k80)
