"""GuSTO model adapter for the TPWL class -- drop-in for sofacontrol/scp/models/tpwl.py (TPWLGuSTO, 32-96); batched
like scp/models/ssm.py here."""
import numpy as np

from ... import utils as scutils
from .template import TemplateModel


class TPWLGuSTO(TemplateModel):
    def __init__(self, dyn_sys):
        super(TPWLGuSTO, self).__init__()
        self.dyn_sys = dyn_sys
        if self.dyn_sys.H is not None:
            self.H = self.dyn_sys.H
        else:
            raise RuntimeError('dyn_sys must have output model specified')
        self.n_x = self.dyn_sys.get_state_dim()
        self.n_u = self.dyn_sys.get_input_dim()
        self.n_z = self.H.shape[0]
        self.nonlinear_observer = False

    def get_continuous_dynamics(self, x, u):
        """scp/models/tpwl.py:32-50 -> (f, A, B), f = A x + B u + d of the selected / blended bank entry."""
        A, B, d = self.dyn_sys.get_jacobians(x)
        if np.asarray(x).ndim == 1:
            f = A @ x + B @ u + d
        else:
            f = np.einsum('bij,bj->bi', A, x) + np.einsum('bij,bj->bi', B, u) + d
        return f, A, B

    def get_discrete_dynamics(self, x, u, dt):
        return self.dyn_sys.get_jacobians(x, dt=dt)

    def pre_discretize(self, dt):
        self.dyn_sys.pre_discretize(dt)

    def get_characteristic_vals(self):
        """scp/models/tpwl.py:67-84: max |x| and max |f| over the stored points -- evaluated as one batch."""
        D = self.dyn_sys.tpwl_dict
        x = scutils.qv2x(np.asarray(D['q']), np.asarray(D['v']))
        f, _, _ = self.get_continuous_dynamics(x, np.asarray(D['u'], dtype=np.float64))
        return np.abs(x).max(axis=0), np.abs(f).max(axis=0)

    def rollout(self, x0, u, dt):
        return self.dyn_sys.rollout(x0, u, dt)
