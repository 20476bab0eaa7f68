"""Multi-GPU check of the sample-sharded ALS (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

Every rank owns a contiguous shard of one global sample set (Philox indexed by the global sample id); the per-core
moments are all-reduced over NCCL inside every micro-step.  Checks: all ranks end with bit-identical cores, and the
sharded fit equals the single-GPU fit of the full sample set to 1e-10 on fitted values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

from bsde_solver._backend import BASIS_POLY, DeviceInputs, get_backend
from bsde_solver._shard import shard_range


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = get_backend()
    d, p, r, N, sweeps = 12, 4, 4, 1 << 17, 6
    ranks = np.array([1] + [r] * (d - 1) + [1], dtype=np.int32)
    rng = np.random.RandomState(3)
    init = np.concatenate([(rng.rand(ranks[i] * p * ranks[i + 1]) - 0.5) * 2 for i in range(d)])
    sig = np.array([0.6])

    def data(lo, n):
        x, _ = be.paths_simulate(1, sig, np.zeros(d), n, 1, 1.0, seed=77, sample_offset=lo, keep_xi=False)
        xT = x[1].contiguous()
        return xT, be.terminal(1, sig, xT)

    # single-GPU reference on the full sample set, before the communicator exists (world = 1 inside the library)
    xT, y = data(0, N)
    full = init.copy()
    be.fit_als(DeviceInputs(BASIS_POLY, p, xT), y, sweeps, ranks, 0, full)
    fit_full = be.tt_eval(DeviceInputs(BASIS_POLY, p, xT), ranks, full)

    be.init_distributed()
    lo, hi = shard_range(N, rank, world)
    xs, ys = data(lo, hi - lo)
    assert torch.equal(xs, xT[:, lo:hi]), "shard does not regenerate its slice of the global stream"
    for use_graph in (1, 0):
        be.set_option("use_graph", use_graph)
        cores = init.copy()
        be.fit_als(DeviceInputs(BASIS_POLY, p, xs), ys, sweeps, ranks, 0, cores)
        t = torch.from_numpy(cores).cuda()
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        identical = all(torch.equal(g, gathered[0]) for g in gathered)
        fit = be.tt_eval(DeviceInputs(BASIS_POLY, p, xT), ranks, cores)
        err = float(torch.linalg.vector_norm(fit - fit_full) / torch.linalg.vector_norm(fit_full))
        if rank == 0:
            print(f"world {world} use_graph {use_graph} fused peer all-reduce {getattr(be, 'p2p', False)}: replicas bit-identical {identical}, "
                  f"fitted values vs single GPU {err:.2e}")
        assert identical and err < 1e-10
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
