"""SSMObserver -- drop-in for sofacontrol/SSM/controllers.py:302-309: the measurement arrives as [v; q] and the SSM
output convention is [q; v]; the belief is that reordered measurement (the reduced state follows from
SSMDynamics.compute_RO_state(z), the batched W_map kernel)."""
from ..utils import vq2qv


class SSMObserver:
    def __init__(self, dyn_sys):
        self.z = None
        self.x = None
        self.dyn_sys = dyn_sys

    def update(self, u, y, dt, x=None):
        self.z = vq2qv(y)

    def reduced_state(self):
        """Extension: x = W_map(z - z_ref) of the current belief (SSM/controllers.py:186-187 does this in the
        controller); accepts a batch of measurements."""
        return self.dyn_sys.compute_RO_state(self.z if self.z.ndim == 1 else self.z.T).T \
            if self.z.ndim == 2 else self.dyn_sys.compute_RO_state(self.z)
