"""CPU oracle for the soft-robot-control hot path.  TEST INFRASTRUCTURE ONLY.

Plain numpy FP64 restatements of the reference algorithms (each function cites the reference file:line it
follows).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package, and only as the checker / reported CPU baseline -- never as the thing shipped.  The product
package (`sofacontrol_b200`) must not import it.

Pinning status (see DESIGN.md "Oracle") -- every module is PINNED to the unmodified reference code:
  * tpwl_np, ilqr_np, pod_np, utils_np : bit-for-bit against the reference modules imported from /root/reference
    (tests/test_oracle_vs_reference.py, runs in the build container), including the non-PD branch of
    dlqr_recursion, and against the committed golden vectors under tests/golden/ that the imported reference
    generated (oracle/make_golden.py).
  * ssm_np : against sofacontrol/SSM/ssm.py imported UNMODIFIED on top of oracle/jax_shim.py (a stand-in for the
    un-vendored jax: numpy float64 as jax.numpy, identity jit, exact forward-mode dual-number jacobian).  Maps,
    Jacobians, discretisations, rollouts on the reference fixtures SSM_model.mat / u_big.csv and iLQR solves agree to
    1e-14 .. 1e-12 (bitwise on most inputs; the lambdified basis evaluates x**3 with libm pow, the restatement as
    (x*x)*x like XLA's integer_pow).  Golden vectors ssm_units.npz / ssm_module_test.npz / ssm_ilqr.npz are outputs
    of the reference class itself.
  * observer_np, lqr_np (closed-loop observers, infinite-horizon gains) : see their headers.
"""
