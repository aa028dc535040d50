import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sofacontrol_b200.synth as synth
from sofacontrol_b200.SSM.ssm import SSMDynamics
from sofacontrol_b200.lqr.ilqr import iLQR
from sofacontrol_b200.utils import QuadraticCost
from oracle.ssm_np import SSMDynamicsNP, GaussNewtonSSM
from oracle.ilqr_np import ILQRNP
from oracle.utils_np import QuadraticCost as QCo
gi = np.load('tests/golden/ssm_ilqr.npz')
m = 8
s = synth.trunk_ssm(m)
model = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
Q, R, Qf = synth.trunk_ilqr_costs(6, m)
sol = iLQR(0.02, model, QuadraticCost(Q, R, Qf), 100, trace=True)
sol.set_target(gi['trunk_zt'])
x, u, K = sol.ilqr_computation(np.zeros(6))
rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
print('rel x,u,K', rel(x, gi['trunk_x']), rel(u, gi['trunk_u']), rel(K, gi['trunk_K']))
print('info', {k: v for k, v in sol.info.items() if k != 'trace'})
o = ILQRNP(0.02, GaussNewtonSSM(SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])), QCo(Q, R, Qf), 100)
o.set_target(gi['trunk_zt'])
xo, uo, Ko = o.ilqr_computation(np.zeros(6))
print('oracle iters', o.iterations, 'cost', o.final_cost)
tr = sol.info['trace']
for ev in o.trace:
    i = ev['it']
    print(i, 'oracle: acc', ev['accepted'], 'ntr', len(ev['trials']), 'cost %.12g' % ev['cost'], 'rho %.6g' % ev['rho_after_bwd'], 'pd_fail_t', ev['pd_fail_t'],
          '| gpu: alpha', tr[i, 1], 'cost %.12g' % tr[i, 0], 'rho %.6g' % tr[i, 2], 'restarts', tr[i, 3], ' ratios', [('%.3g' % t[2]) for t in ev['trials']])
