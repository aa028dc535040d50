"""GuSTO model adapter for the SSM class -- drop-in for sofacontrol/scp/models/ssm.py (SSMGuSTO, 35-93).
Every method forwards to the batched device entry points of sofacontrol_b200.SSM.ssm and therefore also accepts a
stack of points ((count, n_x), (count, n_u)): the whole trajectory of an SCP iteration is one launch."""
import numpy as np

from .template import TemplateModel


class SSMGuSTO(TemplateModel):
    def __init__(self, dyn_sys):
        super(SSMGuSTO, self).__init__()
        self.dyn_sys = dyn_sys
        if self.dyn_sys.H is not None:
            self.H = self.dyn_sys.H
        else:
            raise RuntimeError('dyn_sys must have output model specified')
        self.n_x = self.dyn_sys.get_state_dim()
        self.n_u = self.dyn_sys.get_input_dim()
        self.n_z = self.H.shape[0]
        self.nonlinear_observer = self.dyn_sys.nonlinear_observer

    def get_continuous_dynamics(self, x, u):
        """scp/models/ssm.py:35-55 -> (f, A, B) with f = A x + B u + d."""
        A, B, d = self.dyn_sys.get_continuous_jacobians(x, u=u)
        if np.asarray(x).ndim == 1:
            f = A @ x + B @ u + d
        else:
            f = np.einsum('bij,bj->bi', A, x) + np.einsum('bij,bj->bi', B, u) + d
        return f, A, B

    def get_discrete_dynamics(self, x, u, dt):
        return self.dyn_sys.get_jacobians(x, dt=dt, u=u)

    def get_observer_jacobians(self, x, u, dt):
        return self.dyn_sys.get_observer_jacobians(x)

    def get_characteristic_vals(self):
        return np.ones(self.n_x), np.ones(self.n_x)

    def rollout(self, x0, u, dt):
        return self.dyn_sys.rollout(x0, u, dt)
