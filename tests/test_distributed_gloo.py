"""CPU, world_size 2 over gloo: the host-side logic of the sample-sharded path.  Each rank owns a contiguous shard
of the samples; the only exchange per micro-step is an all-reduce (sum) of the packed normal equations [A | b | y.y],
after which every rank solves redundantly.  The test runs that protocol with the numpy oracle as the local compute and
checks it against the single-process fit, and checks the shard bookkeeping used by the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle.tt_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_als(rank, world, port, x, y, init, p, n_iter, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bsde_solver._shard import shard_range

        lo, hi = shard_range(x.shape[0], rank, world)
        xs, ys = x[lo:hi], y[lo:hi]
        d = x.shape[1]
        phis = [O.poly_eval(xs[:, i], p) for i in range(d)]
        cores = [c.copy() for c in init]
        O.orthonormalize(cores, "right", start=1)
        Rs = O.right_interfaces(cores, phis)
        Ls = [np.ones((xs.shape[0], 1))] + [None] * (d - 1)

        def micro(j):
            V = O.design_matrix(Ls[j], phis[j], Rs[j])
            n = V.shape[1]
            packed = torch.from_numpy(np.concatenate([(V.T @ V).ravel(), V.T @ ys, [ys @ ys]]))
            dist.all_reduce(packed, op=dist.ReduceOp.SUM)          # the one collective of the path
            packed = packed.numpy()
            return O.solver(packed[: n * n].reshape(n, n), packed[n * n: n * n + n], "lstsq").reshape(cores[j].shape)

        for _ in range(n_iter):
            for j in range(d - 1):
                Q, S = np.linalg.qr(O.left_unfold(micro(j)))
                cores[j + 1] = (S @ O.right_unfold(cores[j + 1])).reshape(cores[j + 1].shape)
                cores[j] = Q.reshape(cores[j].shape)
                Ls[j + 1] = np.einsum("sa,sab->sb", Ls[j], O.core_times_phi(cores[j], phis[j]))
            for j in range(d - 1, 0, -1):
                Q, S = np.linalg.qr(O.right_unfold(micro(j)).T)
                cores[j - 1] = (O.left_unfold(cores[j - 1]) @ S.T).reshape(cores[j - 1].shape)
                cores[j] = Q.T.reshape(cores[j].shape)
                Rs[j - 1] = np.einsum("sab,sb->sa", O.core_times_phi(cores[j], phis[j]), Rs[j])
        # replicas must agree bit for bit: every rank solved the same reduced system
        flat = torch.from_numpy(np.concatenate([c.ravel() for c in cores]))
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        if rank == 0:
            out["identical"] = all(bool((g == gathered[0]).all()) for g in gathered)
            out["cores"] = [c.copy() for c in cores]
    finally:
        dist.destroy_process_group()


def test_sharded_normal_equation_allreduce_matches_single_process():
    rng = np.random.RandomState(11)
    N, d, p, r, n_iter = 3001, 5, 3, 2, 4
    x = rng.rand(N, d) * 2 - 1
    y = np.sin(x.sum(axis=1))
    ranks = [1] + [r] * (d - 1) + [1]
    init = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)]
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    ref = O.als(phis, y, n_iter=n_iter, cores=[c.copy() for c in init], randomize=False)
    ref_fit = O.tt_eval(ref, phis)

    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_sharded_als, args=(2, port, x, y, init, p, n_iter, out), nprocs=2, join=True)
    assert out["identical"]
    fit = O.tt_eval(out["cores"], phis)
    assert np.linalg.norm(fit - ref_fit) / np.linalg.norm(ref_fit) < 1e-9


def test_shard_ranges_partition_the_samples():
    from bsde_solver._shard import shard_range

    for n in (1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
