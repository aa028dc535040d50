"""Run under torchrun on >= 2 GPUs: row-sharded POD (DMMA Gram per rank + ONE NCCL all-reduce + replicated eigen-solve +
sharded back-projection) vs the single-process numpy SVD of the same matrix.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from sofacontrol_b200 import parallel
    from sofacontrol_b200.mor import pod
    import sofacontrol_b200.synth as synth
    from oracle import pod_np
    X, _, _ = synth.pod_snapshots(4884, 612, seed=5)
    rows = parallel.shard_rows(4884, rank, world, multiple=4)
    Xl = torch.from_numpy(X[rows].copy()).cuda()
    U, nb, S = pod.compute_POD_sharded(Xl, 5e-5)
    Ufull = parallel.gather_sharded(U, 4884) if (rows.stop - rows.start) * world == 4884 else None
    gathered = [torch.zeros((4884 // world + 8, nb), device="cuda", dtype=torch.float64) for _ in range(world)]
    pad = torch.zeros((4884 // world + 8, nb), device="cuda", dtype=torch.float64)
    pad[:U.shape[0]] = U
    dist.all_gather(gathered, pad)
    sizes = [parallel.shard_rows(4884, r, world, multiple=4) for r in range(world)]
    Uall = torch.cat([g[:s.stop - s.start] for g, s in zip(gathered, sizes)]).cpu().numpy()
    if rank == 0:
        _, Uo, nbo, So = pod_np.compute_POD(X, 5e-5)
        ang = pod_np.subspace_angle(Uo, Uall)[0]
        out = {"world": world, "modes": int(nb), "modes_svd": int(nbo), "subspace_angle": ang,
               "sigma_relerr": float(np.abs(S.cpu().numpy()[:nbo] - So[:nbo]).max() / So[0]),
               "ok": bool(nb == nbo and ang < 1e-8)}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
