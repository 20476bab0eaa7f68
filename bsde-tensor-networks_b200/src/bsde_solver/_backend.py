"""Process-wide handle on the CUDA backend: one context on this process's GPU, device memory through torch.

torch is plumbing here (allocation, streams, torch.distributed for the NCCL bootstrap); every numerical operation
goes through the C ABI of libbsde_b200.so.  Nothing in this module computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import BASIS_LEGENDRE, BASIS_PHI, BASIS_POLY, BackendError, check

_backend = None


def _torch():
    import torch

    return torch


class DeviceInputs:
    """Per-asset regression inputs resident on the GPU, sample-contiguous.

    kind BASIS_POLY / BASIS_LEGENDRE: `data` is X [d][N] and the basis is evaluated inside the kernels;
    kind BASIS_PHI: `data` is PHI [d][p][N] (and `ddata` the derivative values), the reference's explicit lists.
    """

    def __init__(self, kind, p, data, ddata=None, coef=None, d2data=None):
        self.kind, self.p, self.data, self.ddata, self.d2data = int(kind), int(p), data, ddata, d2data
        self.coef = None if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
        self.d = int(data.shape[0])
        self.N = int(data.shape[-1])

    def __len__(self):
        return self.d


def _ptr(t):
    """Device pointer of a torch tensor / host pointer of a numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        if not t.flags["C_CONTIGUOUS"]:
            raise ValueError("the C ABI takes contiguous buffers (got a strided numpy view)")
        return t.ctypes.data_as(C.c_void_p)
    if not t.is_contiguous():
        raise ValueError("the C ABI takes contiguous buffers (got a strided tensor view; call .contiguous())")
    return C.c_void_p(t.data_ptr())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Backend:
    def __init__(self, device=None):
        torch = _torch()
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise BackendError("no CUDA device visible: the B200 backend has no CPU path")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        self.tdev = torch.device("cuda", self.device)
        ctx = C.c_void_p()
        check(self.lib.bsde_ctx_create(self.device, C.byref(ctx)))
        self.ctx = ctx
        self.rank, self.world = 0, 1
        # The library works on its own stream (CUDA graphs cannot be captured on the legacy default stream that
        # torch uses by default).  Every call is fenced against torch's current stream on both sides, so tensors
        # produced by torch before the call and tensors read by torch after it are ordered correctly.
        self.stream = torch.cuda.Stream(device=self.tdev)
        check(self.lib.bsde_ctx_set_stream(self.ctx, C.c_void_p(self.stream.cuda_stream)))

    # ---- plumbing ----------------------------------------------------------------------------------------
    def _enter(self):
        self.stream.wait_stream(_torch().cuda.current_stream(self.tdev))

    def _exit(self):
        _torch().cuda.current_stream(self.tdev).wait_stream(self.stream)

    def _call(self, fn, *args):
        """One C-ABI call on the library's stream, fenced against torch's current stream on both sides."""
        self._enter()
        try:
            check(fn(self.ctx, *args))
        finally:
            self._exit()

    def empty(self, *shape):
        torch = _torch()
        return torch.empty(*shape, dtype=torch.float64, device=self.tdev)

    def to_device(self, a):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            return a.to(device=self.tdev, dtype=torch.float64).contiguous()
        return torch.from_numpy(_f64(a)).to(self.tdev)

    def sync(self):
        check(self.lib.bsde_ctx_sync(self.ctx))

    def launch_count(self):
        out = C.c_int64()
        check(self.lib.bsde_ctx_launch_count(self.ctx, C.byref(out)))
        return out.value

    def set_option(self, name, value):
        check(self.lib.bsde_ctx_set_option(self.ctx, name.encode(), int(value)))

    def get_stat(self, name):
        out = C.c_double()
        check(self.lib.bsde_ctx_get_stat(self.ctx, name.encode(), C.byref(out)))
        return out.value

    def init_distributed(self):
        """Sample-sharded data parallelism: bootstrap the library's NCCL communicator over torch.distributed."""
        torch = _torch()
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            check(self.lib.bsde_comm_unique_id(uid.ctypes.data_as(C.c_void_p)))
        t = torch.from_numpy(uid)
        if dist.get_backend() == "nccl":
            t = t.to(self.tdev)
        dist.broadcast(t, src=0)
        uid = t.cpu().numpy().copy()
        check(self.lib.bsde_comm_init(self.ctx, uid.ctypes.data_as(C.c_void_p), rank, world))
        self.rank, self.world = rank, world
        # fused all-reduce over NVLink peer memory (one node, <= 8 ranks); every rank must take the same decision
        self.p2p = False
        want = os.environ.get("BSDE_P2P", "1") != "0" and world <= 8 and dist.get_backend() == "nccl"
        mine = np.zeros(128, dtype=np.uint8)
        ok = 0
        if want:
            ok = 1 if self.lib.bsde_comm_p2p_export(self.ctx, mine.ctypes.data_as(C.c_void_p)) == 0 else 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.tdev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            handles = [torch.empty(128, dtype=torch.uint8, device=self.tdev) for _ in range(world)]
            dist.all_gather(handles, torch.from_numpy(mine).to(self.tdev))
            allh = np.concatenate([h.cpu().numpy() for h in handles])
            ok = 1 if self.lib.bsde_comm_p2p_open(self.ctx, allh.ctypes.data_as(C.c_void_p), 1) == 0 else 0
            flag = torch.tensor([ok], dtype=torch.int32, device=self.tdev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                self.p2p = True
            else:
                check(self.lib.bsde_comm_p2p_open(self.ctx, None, 0))
        dist.barrier()

    # ---- the hot path ------------------------------------------------------------------------------------
    def fit_als(self, inp: DeviceInputs, y, n_iter, ranks, solver_id, cores_flat):
        ranks = _i32(ranks)
        info = np.zeros(4, dtype=np.int32)
        self._call(self.lib.bsde_fit_als, inp.kind, inp.d, inp.p, inp.N, _ptr(inp.data), _ptr(inp.coef), _ptr(y),
                                    int(n_iter), _ptr(ranks), int(solver_id), _ptr(cores_flat), _ptr(info))
        return info

    def fit_als_host(self, kind, d, p, N, inputs_host, coef, y_host, n_iter, ranks, solver_id, cores_flat):
        ranks = _i32(ranks)
        info = np.zeros(4, dtype=np.int32)
        coef = None if coef is None else _f64(coef)
        self._call(self.lib.bsde_fit_als_host, int(kind), int(d), int(p), int(N), _ptr(inputs_host), _ptr(coef),
                                         _ptr(y_host), int(n_iter), _ptr(ranks), int(solver_id), _ptr(cores_flat),
                                         _ptr(info))
        return info

    def fit_als_host_assets(self, kind, p, blocks, coef, y_host, n_iter, ranks, solver_id, cores_flat):
        """One contiguous host block per asset (N values of x_j, or its (p, N) basis rows), consumed where it lies."""
        ranks = _i32(ranks)
        info = np.zeros(4, dtype=np.int32)
        coef = None if coef is None else _f64(coef)
        d = len(blocks)
        N = int(blocks[0].shape[-1])
        ptrs = (C.c_void_p * d)(*[b.ctypes.data for b in blocks])
        self._call(self.lib.bsde_fit_als_host_assets, int(kind), d, int(p), N, C.cast(ptrs, C.c_void_p), _ptr(coef),
                                                _ptr(y_host), int(n_iter), _ptr(ranks), int(solver_id), _ptr(cores_flat), _ptr(info))
        return info

    def fit_salsa(self, inp: DeviceInputs, y, n_iter, ranks, init_ranks, omega, freq_omega, min_sv, max_rank, solver_id,
                  do_reg, cores_buf):
        state = np.zeros(8, dtype=np.float64)
        self._call(self.lib.bsde_fit_salsa, inp.kind, inp.d, inp.p, inp.N, _ptr(inp.data), _ptr(inp.coef), _ptr(y),
                                      int(n_iter), _ptr(ranks), _ptr(_i32(init_ranks)), float(omega), float(freq_omega),
                                      float(min_sv), int(max_rank), int(solver_id), int(bool(do_reg)), _ptr(cores_buf),
                                      cores_buf.size, _ptr(state))
        return state

    def tt_eval(self, inp: DeviceInputs, ranks, cores_flat, out=None):
        out = self.empty(inp.N) if out is None else out
        self._call(self.lib.bsde_tt_eval, inp.kind, inp.d, inp.p, inp.N, _ptr(inp.data), _ptr(inp.coef),
                                    _ptr(_i32(ranks)), _ptr(_f64(cores_flat)), _ptr(out))
        return out

    def tt_grad(self, inp: DeviceInputs, ranks, cores_flat, out=None):
        out = self.empty(inp.d, inp.N) if out is None else out
        self._call(self.lib.bsde_tt_grad, inp.kind, inp.d, inp.p, inp.N, _ptr(inp.data), _ptr(inp.ddata),
                                    _ptr(inp.coef), _ptr(_i32(ranks)), _ptr(_f64(cores_flat)), _ptr(out))
        return out

    def tt_hessian(self, inp: DeviceInputs, ranks, cores_flat, out=None):
        out = self.empty(inp.d, inp.d, inp.N) if out is None else out
        self._call(self.lib.bsde_tt_hessian, inp.kind, inp.d, inp.p, inp.N, _ptr(inp.data), _ptr(inp.ddata), _ptr(inp.d2data),
                                       _ptr(inp.coef), _ptr(_i32(ranks)), _ptr(_f64(cores_flat)), _ptr(out))
        return out

    def tt_hessian(self, inp: DeviceInputs, ranks, cores_flat, out=None):
        out = self.empty(inp.d, inp.d, inp.N) if out is None else out
        self._call(self.lib.bsde_tt_hessian, inp.kind, inp.d, inp.p, inp.N, _ptr(inp.data), _ptr(inp.ddata), _ptr(inp.d2data),
                                       _ptr(inp.coef), _ptr(_i32(ranks)), _ptr(_f64(cores_flat)), _ptr(out))
        return out

    # ---- paths, targets ----------------------------------------------------------------------------------
    def paths_simulate(self, model_id, params, x0, N, steps, T, seed=0, sample_offset=0, xi=None, keep_xi=True):
        """x [steps+1][d][N] (and xi in the same layout).  `xi` given => replay of host-supplied increments.
        x0: (d,) one starting point for all samples, or (N, d) one per sample (path.py:6)."""
        x0 = _f64(x0)
        per_sample = x0.ndim == 2
        if per_sample and x0.shape[0] != N:
            raise ValueError(f"X0 has {x0.shape[0]} rows for {N} samples")
        d = x0.shape[-1]
        x = self.empty(steps + 1, d, N)
        if per_sample:
            x[0].copy_(self.to_device(np.ascontiguousarray(x0.T)))
            x0 = None
        replay = xi is not None
        if replay:
            xi = self.to_device(xi)
        elif keep_xi:
            xi = self.empty(steps + 1, d, N)
        self._call(self.lib.bsde_paths_simulate, int(model_id), _ptr(_f64(params)), d, int(N), int(steps), float(T),
                                           _ptr(x0), int(seed), int(sample_offset), int(replay), _ptr(xi), _ptr(x))
        return x, xi

    def terminal(self, model_id, params, x):
        d, N = x.shape
        out = self.empty(N)
        self._call(self.lib.bsde_terminal, int(model_id), _ptr(_f64(params)), d, N, _ptr(x), _ptr(out))
        return out

    def target(self, model_id, params, x, y, y_next, grad, xi, dt, apply_sigma, out=None):
        d, N = x.shape
        out = self.empty(N) if out is None else out
        self._call(self.lib.bsde_target, int(model_id), _ptr(_f64(params)), d, N, _ptr(x), _ptr(y), _ptr(y_next),
                                   _ptr(grad), _ptr(xi), float(dt), int(bool(apply_sigma)), _ptr(out))
        return out

    def mean(self, v):
        out = C.c_double()
        self._call(self.lib.bsde_mean, _ptr(v), int(v.numel()), C.byref(out))
        return out.value

    def transpose(self, src, rows, cols):
        dst = self.empty(cols, rows)
        self._call(self.lib.bsde_transpose, _ptr(src), int(rows), int(cols), _ptr(dst))
        return dst

    # ---- verification hooks ------------------------------------------------------------------------------
    def normal_eq(self, inp: DeviceInputs, y, ranks, cores_flat, j):
        ranks = _i32(ranks)
        n = int(ranks[j]) * inp.p * int(ranks[j + 1])
        A = np.zeros((n, n))
        b = np.zeros(n)
        self._call(self.lib.bsde_normal_eq, inp.kind, inp.d, inp.p, inp.N, _ptr(inp.data), _ptr(inp.coef), _ptr(y),
                                      _ptr(ranks), _ptr(_f64(cores_flat)), int(j), _ptr(A), _ptr(b))
        return A, b

    def solve(self, A, b, solver_id):
        A, b = _f64(A), _f64(b)
        n = b.shape[0]
        x = np.zeros(n)
        info = np.zeros(4, dtype=np.int32)
        self._call(self.lib.bsde_solve, n, _ptr(A), _ptr(b), int(solver_id), _ptr(x), _ptr(info))
        return x, info

    def orthonormalize(self, p, ranks, cores_flat, mode, start, stop):
        """p: one basis size for every core, or a sequence of per-core sizes (a ragged TensorTrain shape)."""
        ranks = _i32(ranks)
        if np.ndim(p) == 0:
            self._call(self.lib.bsde_orthonormalize, len(ranks) - 1, int(p), _ptr(ranks), _ptr(cores_flat), int(mode),
                                               int(start), int(stop))
        else:
            shape = _i32(p)
            if shape.shape[0] != len(ranks) - 1:
                raise ValueError("one basis size per core expected")
            self._call(self.lib.bsde_orthonormalize_shape, len(ranks) - 1, _ptr(shape), _ptr(ranks), _ptr(cores_flat), int(mode),
                                                     int(start), int(stop))
        return cores_flat

    def salsa_move(self, side, core_a, core_b, min_sv, expand=False, do_reg=True):
        """One SVD move of SALSA on host cores (verification hook): returns (core_a, core_b, gamma_or_theta, adapted)."""
        ra, p, k = core_a.shape
        rb = core_b.shape[2]
        a = np.zeros(ra * p * (k + 1))
        b = np.zeros((k + 1) * p * rb)
        a[: core_a.size] = _f64(core_a).ravel()
        b[: core_b.size] = _f64(core_b).ravel()
        reg = np.zeros(k + 2)
        rank, adapted = C.c_int(), C.c_int()
        self._call(self.lib.bsde_salsa_move, int(side), ra, p, k, rb, float(min_sv), int(bool(expand)), int(bool(do_reg)),
                                       _ptr(a), _ptr(b), _ptr(reg), C.byref(rank), C.byref(adapted))
        kn = rank.value
        return (a[: ra * p * kn].reshape(ra, p, kn), b[: kn * p * rb].reshape(kn, p, rb), reg[:kn] if do_reg else None, bool(adapted.value))

    def basis_eval(self, kind, p, x, coef, deriv):
        N = int(x.numel())
        out = self.empty(p, N)
        coef = None if coef is None else _f64(coef)
        self._call(self.lib.bsde_basis_eval, int(kind), int(p), N, _ptr(x), _ptr(coef), int(deriv), _ptr(out))
        return out


def get_backend() -> Backend:
    """The process-wide backend (created on first use, on cuda:LOCAL_RANK)."""
    global _backend
    if _backend is None:
        _backend = Backend()
    return _backend


__all__ = ["Backend", "DeviceInputs", "get_backend", "BackendError", "BASIS_PHI", "BASIS_POLY", "BASIS_LEGENDRE"]
