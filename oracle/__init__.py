"""CPU oracle for the soft-robot-control hot path.  TEST INFRASTRUCTURE ONLY.

Plain numpy FP64 restatements of the reference algorithms (each function cites the reference file:line it
follows).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package, and only as the checker / reported CPU baseline -- never as the thing shipped.  The product
package (`sofacontrol_b200`) must not import it.

Pinning status (see DESIGN.md "Oracle"):
  * tpwl_np, ilqr_np, pod_np, utils_np : PINNED -- checked bit-for-bit against the unmodified reference modules
    imported from /root/reference (tests/test_oracle_vs_reference.py, runs in the build container) and against the
    committed golden vectors under tests/golden/ that the imported reference generated (oracle/make_golden.py).
  * ssm_np : "parity unpinned" by any reference-owned test -- sofacontrol/SSM/ssm.py needs jax (un-vendored,
    version unpinned) and cannot run here.  The restatement follows ssm.py line by line with analytic Jacobians,
    is cross-checked against sympy differentiation of the reference's own basis construction, and is anchored on
    the reference fixtures SSM_model.mat / u_big.csv / z_big.csv / rest_qv.pkl (golden rollout vectors).
"""
