"""N > 1 host logic on CPU: world-size-2 gloo process groups (127.0.0.1).  Covers the batch sharding used by
bench.py / callers, the max-over-ranks timing reduction, the all-gather of sharded results and the row-sharded POD
flow (Gram all-reduce -> replicated eigen-solve -> sharded back-projection) with torch CPU stand-ins for the two
DMMA kernels (no GPU in this container)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sofacontrol_b200 import parallel
    from sofacontrol_b200.mor import pod
    import sofacontrol_b200.synth as synth
    out = {}
    # 1) batch sharding: disjoint cover
    total = 4097
    sl = parallel.shard_slice(total, rank, world)
    out["slice"] = (sl.start, sl.stop)
    # 2) timing reduction
    out["tmax"] = parallel.max_over_ranks(1.0 + rank)
    # 3) gather of sharded per-problem results
    local = torch.arange(sl.start, sl.stop, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
    full = parallel.gather_sharded(local, total)
    out["gather_ok"] = bool(torch.equal(full[:, 0], torch.arange(total, dtype=torch.float64)))
    # 4) row-sharded POD
    X, _, _ = synth.pod_snapshots(400, 60, seed=5)
    rows = parallel.shard_rows(400, rank, world, multiple=8)
    Xl = torch.from_numpy(X[rows].copy())
    U, nb, S = pod.compute_POD_sharded(Xl, 5e-5, gram=lambda a: a.t() @ a, gemm=lambda a, b: a @ b)
    out["rows"] = (rows.start, rows.stop)
    out["U"] = U.numpy()
    out["nb"] = nb
    out["S"] = S.numpy()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # sharding covers the batch exactly once
    assert res[0]["slice"] == (0, 2049) and res[1]["slice"] == (2049, 4097)
    assert res[0]["tmax"] == 2.0 and res[1]["tmax"] == 2.0
    assert res[0]["gather_ok"] and res[1]["gather_ok"]
    # POD: same answer as the single-process oracle
    from oracle import pod_np
    import sofacontrol_b200.synth as synth
    X, _, _ = synth.pod_snapshots(400, 60, seed=5)
    _, Uo, nbo, So = pod_np.compute_POD(X, 5e-5)
    assert res[0]["nb"] == res[1]["nb"] == nbo
    U = np.vstack([res[0]["U"], res[1]["U"]])
    assert res[0]["rows"][1] == res[1]["rows"][0] and res[1]["rows"][1] == 400
    assert np.abs(res[0]["S"][:nbo] - So[:nbo]).max() / So[0] < 1e-9
    assert pod_np.subspace_angle(Uo, U)[0] < 1e-8


def test_shard_helpers_edge_cases():
    from sofacontrol_b200 import parallel
    for total, world in ((0, 4), (3, 8), (4096, 8), (10, 3)):
        sl = [parallel.shard_slice(total, r, world) for r in range(world)]
        assert sl[0].start == 0 and sl[-1].stop == total
        assert all(a.stop == b.start for a, b in zip(sl, sl[1:]))
        sizes = [s.stop - s.start for s in sl]
        assert max(sizes) - min(sizes) <= 1
    r = [parallel.shard_rows(1001, k, 4, multiple=128) for k in range(4)]
    assert r[0].start == 0 and r[-1].stop == 1001 and all(a.stop == b.start for a, b in zip(r, r[1:]))
