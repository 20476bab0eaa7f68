#!/usr/bin/env python
"""bench.py — samples·sweeps/s of the tensor-train regression (ALS) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "C3"): HJB law, d = 100 assets, 2^20 Monte-Carlo samples per GPU, TT rank 4,
4 monomials per asset, ALS with 10 sweeps per fit (one sweep = 2(d-1) = 198 micro-steps).  Samples are synthetic:
X_T of the HJB forward SDE (sigma = sqrt(2), X_0 = 0, T = 1) drawn by the library's own counter-based path
kernel, y = g(X_T) = log(1/2 + 1/2 |x|^2).  One *step* = one ALS fit of the batch.  With N GPUs every rank owns its
own 2^20 samples (weak scaling) and the per-core normal equations are all-reduced over NCCL every micro-step.

Printed JSON line (rank 0): `value` = whole-job samples·sweeps/s with inputs resident in HBM; `e2e` = the same
through the host-buffer entry point (bsde_fit_als_host: pinned host X and y copied in, cores copied out, inside
the timed region); `roofline` = the assembly kernel against the FP64 tensor-pipe peak; `cpu_baseline` = the numpy
oracle port on this host's cores on a bounded sample.  Further keys of the same line, so that BENCH / SCALE alone say
what a fit that actually fits costs: `converging` = the same kernel shape on a target the train can
represent (Black-Scholes law, y = |x|^2: the fit converges and most solves are rank-deficient, i.e. take the truncated
path); `full_solve` = seconds and Y0 against the analytic price of the largest explicit backward solve that converges
(Black-Scholes d = 20, N = 2^20, 50 steps, rank 4, 4 monomials; strong scaling over the GPUs); `e2e_phis` = the reference's
own calling convention ALS(phis, y) with host lists of (N, p) arrays; `multi_gpu_check` (N > 1) = replicas bit-identical
and the sharded fit against the single-GPU fit of the union.
`--impl reference` times the UNMODIFIED reference (pip-installed into baseline/_ref, run on oracle/shim's stand-in for the
absent opt_einsum) on a bounded sample whose true size is in `config`, with the oracle port beside it.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), os.path.join(ROOT, "bsde-tensor-networks_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

D, P, R, N_PER_GPU, N_SWEEPS = 100, 4, 4, 1 << 20, 10
# FP64 peak of this pool's B200, measured by bench/micro/fp64_peak.cu (mma.sync f64 back to back, all SMs) in round 1:
# 37.0 TFLOP/s (profiles/fp64_peak_r01.txt); cuBLAS DGEMM 8192^3 reaches 35.5.  MEASURED_PEAKS.json has no FP64 entry.
FP64_PEAK_TFLOPS = 37.0
CPU_SAMPLE_N = 1 << 12


def algorithmic_flops_per_sample(n):
    """symmetric half of A^T A (one multiply + one add each) plus A^T y."""
    return n * (n + 1) + 2 * n


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def initial_cores(seed=7):
    rng = np.random.RandomState(seed)
    ranks = np.array([1] + [R] * (D - 1) + [1], dtype=np.int32)
    cores = np.concatenate([(rng.rand(ranks[i] * P * ranks[i + 1]) - 0.5) * 2 for i in range(D)])
    return ranks, cores


def cpu_port(n_samples, n_sweeps, seed=0):
    """The oracle (numpy restatement of the reference algorithm with cached interfaces) on the same workload shape."""
    import oracle.tt_oracle as O

    rng = np.random.RandomState(seed)
    x = rng.randn(n_samples, D) * np.sqrt(2.0)
    y = np.log(0.5 + 0.5 * np.sum(x**2, axis=1))
    phis = [O.poly_eval(x[:, i], P) for i in range(D)]
    ranks = [1] + [R] * (D - 1) + [1]
    init = [(rng.rand(ranks[i], P, ranks[i + 1]) - 0.5) * 2 for i in range(D)]
    t0 = time.perf_counter()
    O.als(phis, y, n_iter=n_sweeps, cores=init, randomize=False)
    return time.perf_counter() - t0


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


REF_SAMPLE_N, REF_SWEEPS = 1 << 10, 1       # unmodified reference: 2^10 samples, one sweep per step (~5 s per step on 8 cores)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this host's cores, one bounded fit per step.
    The unmodified reference (baseline/_ref + oracle/shim) when it is installed, else the numpy oracle port."""
    if rank != 0:
        return
    import oracle.run_reference as RR

    port = {}
    for n in (1 << 12, 1 << 14):
        t = cpu_port(n, 2)
        port[f"N=2^{int(np.log2(n))}"] = n * 2 / t
    common = {"impl": "reference", "metric": "samples*sweeps/s (TT regression, ALS)", "unit": "samples*sweeps/s", "n_gpus": args.gpus,
              "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
              "data": "synthetic", "gpu_launches": 0}
    if RR.available():
        times = RR.time_als(D, REF_SAMPLE_N, R, P, REF_SWEEPS, args.steps, args.warmup)
        t = float(sum(times))
        value = REF_SAMPLE_N * REF_SWEEPS * args.steps / t
        sample = (f"UNMODIFIED reference ALS (baseline/_ref, bsde_solver.core.optimisation.als.ALS on oracle/shim's opt_einsum stand-in), "
                  f"d={D} r={R} p={P}, N=2^{int(np.log2(REF_SAMPLE_N))} samples (NOT 2^20: the reference recomputes every interface per "
                  f"micro-step and cannot allocate the full batch), {REF_SWEEPS} sweep per step")
        kind, n_ref, sw_ref = "reference", REF_SAMPLE_N, REF_SWEEPS
    else:
        t = 0.0
        for _ in range(args.warmup):
            cpu_port(CPU_SAMPLE_N, N_SWEEPS)
        for _ in range(args.steps):
            t += cpu_port(CPU_SAMPLE_N, N_SWEEPS)
        value = CPU_SAMPLE_N * N_SWEEPS * args.steps / t
        sample = f"numpy oracle port (cached interfaces; baseline/_ref is not installed), d={D} r={R} p={P}, N=2^{int(np.log2(CPU_SAMPLE_N))} samples, {N_SWEEPS} sweeps per step"
        kind, n_ref, sw_ref = "port", CPU_SAMPLE_N, N_SWEEPS
    cfg = dict(workload_config(world), samples_per_gpu=n_ref, global_samples=n_ref, sweeps_per_step=sw_ref,
               workload=f"C3 ALS fit shape (HJB law, d={D}, rank {R}, {P} monomials/asset) on a BOUNDED sample of N=2^{int(np.log2(n_ref))} samples, {sw_ref} sweep(s) per step, host CPU",
               l2="n/a (CPU)", parallelism="host threads (BLAS)")
    print(json.dumps(dict(common, value=value, ms_per_step=1e3 * t / args.steps, config=cfg,
                          cpu_baseline={"value": value, "unit": "samples*sweeps/s", "cores": host_threads(), "kind": kind, "sample": sample,
                                        "oracle_port_samples_sweeps_per_s": port,
                                        "note": "the port caches the interfaces (O(d) per sweep); the reference recontracts them per micro-step (O(d^2))"},
                          e2e={"value": value, "unit": "samples*sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})))


def workload_config(world):
    return {"workload": f"C3 ALS fit: HJB d={D}, N=2^20 samples per GPU, rank {R}, {P} monomials/asset, {N_SWEEPS} sweeps/fit",
            "d": D, "rank": R, "basis": P, "samples_per_gpu": N_PER_GPU, "global_samples": N_PER_GPU * world,
            "sweeps_per_step": N_SWEEPS, "micro_steps_per_sweep": 2 * (D - 1), "parallelism": f"dp{world} (samples sharded)",
            "l2": "inputs (839 MB/GPU) and interfaces (3.3 GB/GPU) exceed the 126 MB L2; no flush needed"}


def load_full_solve():
    """bench/full_solve.py as a module (bench.py itself shadows the name `bench`)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bsde_full_solve", os.path.join(ROOT, "bench", "full_solve.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CONV_D = 20       # largest asset count at which the rank-4 / 4-monomial ALS fit of Black-Scholes data converges (d >= 50 does not,
                  # neither here nor in the numpy restatement of the reference: DESIGN.md section 9)


def converging_leg(be, torch, dist, rank, world, N, barrier):
    """The C3 kernel shape (r = p = 4, n = 64) on a problem the train DOES fit: the terminal regression of the full solve
    below — X_T of the Black-Scholes forward SDE (X0 = (1, 0.5, ...), sigma = 0.4, 50 Euler steps), y = g(X_T) = |x|^2.
    Once the fit converges most normal equations are numerically rank-deficient and their solves truncate like
    numpy.linalg.lstsq (vh_lstsq.cuh) instead of taking the LDL^T fast path."""
    from bsde_solver._backend import BASIS_POLY, DeviceInputs

    d = CONV_D
    x0 = np.tile([1.0, 0.5], d // 2)
    x, _ = be.paths_simulate(0, np.array([0.05, 0.4]), x0, N, 50, 1.0, seed=4242, sample_offset=rank * N, keep_xi=False)
    xT = x[50].contiguous()
    del x
    y = be.terminal(0, np.array([0.05, 0.4]), xT)
    inp = DeviceInputs(BASIS_POLY, P, xT)
    rng = np.random.RandomState(11)
    ranks = np.array([1] + [R] * (d - 1) + [1], dtype=np.int32)
    init = np.concatenate([(rng.rand(ranks[i] * P * ranks[i + 1]) - 0.5) * 2 for i in range(d)])
    info = None

    def fit():
        nonlocal info
        cores = init.copy()
        info = be.fit_als(inp, y, N_SWEEPS, ranks, 0, cores)
        return cores

    for _ in range(2):
        fit()
    barrier()
    K = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        cores = fit()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    res = float(torch.linalg.vector_norm(be.tt_eval(inp, ranks, cores) - y) / torch.linalg.vector_norm(y))
    micro = 2 * (d - 1) * N_SWEEPS
    flops = N * N_SWEEPS * ((2 * (d - 2)) * algorithmic_flops_per_sample(R * P * R) + 2 * algorithmic_flops_per_sample(P * R))
    return {"workload": f"Black-Scholes law d={d}, N=2^{int(np.log2(N))} per GPU, rank {R}, {P} monomials, y = |x|^2, {N_SWEEPS} sweeps/fit "
                        f"({micro} micro-steps/fit)",
            "value": world * N * N_SWEEPS * K / (ms * 1e-3), "unit": "samples*sweeps/s", "ms_per_fit": ms / K, "fits_timed": K,
            "us_per_micro_step": 1e3 * ms / K / micro,
            "fraction_of_fp64_peak_whole_fit": flops / (ms / K * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
            "fit_relative_residual": res,
            "solves_per_fit": {"cholesky": int(info[0]), "truncated": int(info[1]), "lu": int(info[2]), "min_numerical_rank": int(info[3])}}


def host_list_leg(be, torch, xT, y, ranks, init):
    """The reference's calling convention at C3 size: ALS(phis, y) with host lists of (N, p) arrays, twice — the arrays this
    package's own PolynomialBasis produces (recognised, shipped as points) and anonymous copies of them (explicit values)."""
    from bsde_solver.core.calculus.basis import PolynomialBasis
    from bsde_solver.core.optimisation import als as A
    from bsde_solver.core.tensor.tensor_core import TensorCore
    from bsde_solver.core.tensor.tensor_train import TensorTrain

    xh, yh = xT.cpu().numpy(), y.cpu().numpy()
    basis = PolynomialBasis(P)
    tagged = [TensorCore(basis.eval(xh[i]), name=f"phi_{i+1}", indices=("batch", f"m_{i+1}")) for i in range(D)]
    tt_ranks = tuple(int(r) for r in ranks)

    def replay(self):                       # same initial cores as the other legs (the reference draws them from numpy's stream)
        off = 0
        for k, name in enumerate(self.cores):
            n = self.cores[name].size
            self.cores[name][:] = init[off:off + n].reshape(self.cores[name].shape)
            off += n
        return self

    out = {}
    orig = TensorTrain.randomize
    TensorTrain.randomize = replay
    try:
        for label, phis in (("tagged", tagged), ("plain", None)):
            if phis is None:
                phis = [np.asfortranarray(np.array(t)) for t in tagged]          # anonymous F-ordered (N, p) copies, as the reference builds them
                del tagged
            A.ALS(phis, yh, n_iter=N_SWEEPS, ranks=tt_ranks)                      # warm-up (workspaces, plans)
            torch.cuda.synchronize()
            reps = 2
            t0 = time.perf_counter()
            for _ in range(reps):
                A.ALS(phis, yh, n_iter=N_SWEEPS, ranks=tt_ranks)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            h2d = sum(int(np.asarray(p_).nbytes) for p_ in phis) if label == "plain" else int(xh.nbytes)
            out[label] = {"value": xh.shape[1] * N_SWEEPS / dt, "unit": "samples*sweeps/s", "ms_per_step": 1e3 * dt,
                          "h2d_bytes_per_step": h2d + int(yh.nbytes), "inputs": A.last_fit_info.get("inputs"),
                          "timing": "host wall clock around ALS(phis, y) incl. the H2D copies from pageable memory and the cores coming back"}
    finally:
        TensorTrain.randomize = orig
    return out


def multi_gpu_check(be, torch, dist, rank, world, local, N_check=1 << 17, sweeps=6):
    """What tests/multigpu_check.py asserts, inside the driver-visible run: every rank ends a sharded fit with bit-identical
    cores, and that fit equals the single-GPU fit of the union of the shards (fresh single-GPU context on rank 0).  Run on
    a well-posed fit (HJB law with sigma = 0.6, d = 12, C3 core shape): on the C3 throughput data the fit explains nothing
    (residual 0.9995) and on the converging workload most systems are rank-deficient — in both the cores are not a function
    of the data to better than the fit residual, whatever the summation order."""
    from bsde_solver._backend import BASIS_POLY, Backend, DeviceInputs

    d = 12
    sig = np.array([0.6])
    rng = np.random.RandomState(3)
    ranks = np.array([1] + [R] * (d - 1) + [1], dtype=np.int32)
    init = np.concatenate([(rng.rand(ranks[i] * P * ranks[i + 1]) - 0.5) * 2 for i in range(d)])

    def data(b, lo, n):
        x, _ = b.paths_simulate(1, sig, np.zeros(d), n, 1, 1.0, seed=77, sample_offset=lo, keep_xi=False)
        xT = x[1].contiguous()
        return xT, b.terminal(1, sig, xT)

    xs, ys = data(be, rank * N_check, N_check)
    cores = init.copy()
    be.fit_als(DeviceInputs(BASIS_POLY, P, xs), ys, sweeps, ranks, 0, cores)
    t = torch.from_numpy(cores).cuda()
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    identical = bool(all(torch.equal(g, gathered[0]) for g in gathered))
    err = res = None
    if rank == 0:
        single = Backend(device=local)                    # world = 1 inside the library: no exchange
        xu, yu = data(single, 0, N_check * world)
        full = init.copy()
        single.fit_als(DeviceInputs(BASIS_POLY, P, xu), yu, sweeps, ranks, 0, full)
        inp = DeviceInputs(BASIS_POLY, P, xu)
        f_single = single.tt_eval(inp, ranks, full)
        f_shard = single.tt_eval(inp, ranks, cores)
        err = float(torch.linalg.vector_norm(f_shard - f_single) / torch.linalg.vector_norm(f_single))
        res = float(torch.linalg.vector_norm(f_shard - yu) / torch.linalg.vector_norm(yu))
    dist.barrier()
    return {"replicas_identical": identical, "vs_single_gpu": err, "fit_relative_residual": res,
            "what": f"{sweeps}-sweep fit (HJB law sigma=0.6, d={d}, rank {R}, {P} monomials) on 2^{int(np.log2(N_check))} samples per rank: cores "
                    "all-gathered and compared bit for bit; fitted values of the sharded fit vs the single-GPU fit of the union of the shards (relative L2)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=N_PER_GPU, help="samples per GPU (default: the C3 workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the converging-fit, full-solve, host-list and multi-GPU-check legs")
    ap.add_argument("--full-solve-assets", type=int, default=20)
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE",
                    help="library option for A/B measurements (bsde_ctx_set_option), e.g. fuse_interface=0")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from bsde_solver._backend import BASIS_POLY, DeviceInputs, get_backend

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = get_backend()
    if world > 1:
        be.init_distributed()
    for kv in args.option:
        name, _, val = kv.partition("=")
        be.set_option(name, int(val))
    N = args.samples

    # ---- synthetic workload, generated on the device by the library's own kernels ----
    x, _ = be.paths_simulate(1, np.array([np.sqrt(2.0)]), np.zeros(D), N, 1, 1.0, seed=2024, sample_offset=rank * N, keep_xi=False)
    xT = x[1].contiguous()                      # [d][N]
    y = be.terminal(1, np.array([np.sqrt(2.0)]), xT)
    del x
    inp = DeviceInputs(BASIS_POLY, P, xT)
    ranks, init = initial_cores()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solve_paths = {}

    def fit_device():
        cores = init.copy()
        info = be.fit_als(inp, y, N_SWEEPS, ranks, 0, cores)
        solve_paths.update(cholesky=int(info[0]), truncated=int(info[1]), lu=int(info[2]), min_numerical_rank=int(info[3]))
        return cores

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        fit_device()
    barrier()
    sampler.start()
    l0 = be.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cores = fit_device()
    e1.record()
    barrier()
    sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = be.launch_count() - l0
    sweep_graph = {"edges": int(be.get_stat("graph_edges")), "programmatic_edges": int(be.get_stat("graph_pdl_edges")),
                   "note": "kernel -> kernel edges of the captured ALS sweep; programmatic = dependent launch (DESIGN.md section 4a)"}
    value = world * N * N_SWEEPS * args.steps / (ms * 1e-3)
    fit_quality = float(torch.linalg.vector_norm(be.tt_eval(inp, ranks, cores) - y) / torch.linalg.vector_norm(y))

    # ---- end to end through the host-buffer entry point ----
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty(xT.shape, dtype=torch.float64, pin_memory=True)
        y_host = torch.empty(y.shape, dtype=torch.float64, pin_memory=True)
        x_host.copy_(xT)
        y_host.copy_(y)
        torch.cuda.synchronize()
        xh, yh = x_host.numpy(), y_host.numpy()

        def fit_host():
            cores = init.copy()
            be.fit_als_host(BASIS_POLY, D, P, N, xh, None, yh, N_SWEEPS, ranks, 0, cores)
            return cores

        for _ in range(max(1, args.warmup - 1)):
            fit_host()
        barrier()
        sampler.start()
        e0.record()
        for _ in range(args.steps):
            fit_host()
        e1.record()
        barrier()
        sampler.stop()
        ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        ms2 = float(ms2.item())
        e2e = {"value": world * N * N_SWEEPS * args.steps / (ms2 * 1e-3), "unit": "samples*sweeps/s",
               "h2d_bytes_per_step": int(xh.nbytes + yh.nbytes + init.nbytes), "d2h_bytes_per_step": int(init.nbytes + 16),
               "ms_per_step": ms2 / args.steps}

    # ---- per-kernel device time (eager launches bracketed by CUDA events on the library's stream) ----
    be.set_option("profile", 1)
    fit_device()
    stats = {k: (be.get_stat(f"{k}_ms"), int(be.get_stat(f"{k}_n"))) for k in ("assembly", "reduce", "allreduce", "solve", "interface")}
    solve_phase_cycles = [be.get_stat(f"solve_phase{k}") for k in range(7)]
    if os.environ.get("BSDE_SOLVE_TQ"):
        print("solve tq", [be.get_stat(f"solve_phase{k}") / max(1, stats["solve"][1]) for k in range(8, 16)], file=sys.stderr)
    be.set_option("profile", 0)
    asm_ms, asm_n = stats["assembly"]
    n_int, n_edge = R * P * R, P * R
    per_sweep_flops = (2 * (D - 2)) * algorithmic_flops_per_sample(n_int) + 2 * algorithmic_flops_per_sample(n_edge)
    asm_flops_per_launch = N * per_sweep_flops / (2 * (D - 1))                 # average over the sweep's launches
    achieved = asm_flops_per_launch / (asm_ms / asm_n * 1e-3) / 1e12
    total_prof = sum(v[0] for v in stats.values())
    roofline = {"kernel": "k_assemble_r4 (normal-equation moments on the FP64 tensor pipe, mma.sync m8n8k4 f64 = DMMA.8x8x4; interface update of the previous micro-step fused in)",
                "bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved / FP64_PEAK_TFLOPS, "traffic": None,
                "peak_source": "FP64 DMMA peak measured on this pool by bench/micro/fp64_peak.cu (profiles/fp64_peak_r01.txt); MEASURED_PEAKS.json has no FP64 figure",
                "avg_launch_us": 1e3 * asm_ms / asm_n, "launches_timed": asm_n,
                "algorithmic_flops_per_launch": asm_flops_per_launch,
                "algorithmic_bytes_per_launch": N * (88 + 32),
                "note": "achieved counts the reference algorithm's n(n+1)+2n flops per sample; the kernel executes ~0.36x of them (Kronecker moment factorisation), so frac can exceed 1",
                "share_of_step": {k: v[0] / total_prof for k, v in stats.items()},
                "class_us_per_launch": {k: (1e3 * v[0] / v[1] if v[1] else None) for k, v in stats.items()},
                "solve_kernel_phase_cycles_per_launch": dict(zip(("expand_moments", "solve", "qr", "push_factor", "solve:load", "solve:factor", "solve:back_substitution"),
                                                                 [c / max(1, stats["solve"][1]) for c in solve_phase_cycles]))}
    # what the kernel EXECUTES on the tensor pipe (interior launches: 14 DMMA.8x8x4 tiles of 512 flops per 4 samples), next to
    # the algorithmic count above; the two edge launches of a sweep do less, so this overstates the sweep average by < 1 %
    exec_tflops = N * 1792.0 / (asm_ms / asm_n * 1e-3) / 1e12
    roofline["executed"] = {"tensor_flops_per_sample": 1792, "tflops": exec_tflops, "frac": exec_tflops / FP64_PEAK_TFLOPS,
                            "note": "DMMA flops issued per second / FP64 peak; the same pipe also carries 11 DMUL per 4 samples and ~128 "
                                    "FP64 operations per sample of staging (ncu: DMMA 61 % + FP64 15 % of the pipe busy, profiles/kernels_r02_s3.txt)"}
    try:
        with open(os.path.join(ROOT, "profiles", "assembly_traffic.json")) as f:
            roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")
        roofline["traffic_source"] = "static: ncu --set full capture of this kernel kept in profiles/assembly_traffic.json (not measured in this run)"
    except Exception:
        pass

    # ---- what a fit that actually fits costs; a whole backward solve; the reference's calling convention; replica check ----
    converging = full_solve = e2e_phis = mg_check = None
    if not args.no_extras and N == N_PER_GPU:
        converging = converging_leg(be, torch, dist, rank, world, N, barrier)
        if world == 1 and not args.no_e2e:
            e2e_phis = host_list_leg(be, torch, xT, y, ranks, init)
        del xT, y, inp
        torch.cuda.empty_cache()
        if world > 1:
            mg_check = multi_gpu_check(be, torch, dist, rank, world, local)
        fs = load_full_solve().run(samples=1 << 20, assets=args.full_solve_assets, steps=50, rank=R, degree=P, sweeps=N_SWEEPS, seed=0,
                                   model_name="bs")
        if rank == 0:
            full_solve = {k: fs[k] for k in ("workload", "n_gpus", "scaling", "paths_s", "solve_s", "total_s", "samples_sweeps_per_s", "Y0",
                                             "Y0_reference", "Y0_reference_kind", "Y0_rel_err", "step_time_s_min_median_max")}
            full_solve["note"] = ("largest explicit solve at rank 4 / 4 monomials that converges; BASELINE configs[2] itself (HJB d=100) diverges in the "
                                  "reference algorithm too (DESIGN.md §9), as does Black-Scholes at d=100")

    # ---- CPU baseline on this host (rank 0, single-GPU runs only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t = cpu_port(1 << 14, 2)
        cpu = {"value": (1 << 14) * 2 / t, "unit": "samples*sweeps/s", "cores": host_threads(), "kind": "port",
               "sample": f"same fit shape (d={D}, r={R}, p={P}) on N=2^14 samples, 2 sweeps, numpy oracle port with cached interfaces; {t:.1f} s"}

    if rank == 0:
        print(json.dumps({
            "metric": "samples*sweeps/s (TT regression, ALS)", "value": value, "unit": "samples*sweeps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world) if N == N_PER_GPU else dict(workload_config(world), samples_per_gpu=N, workload="reduced-N debugging run (not the benchmark)"),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "fit_relative_residual": fit_quality, "solves_per_fit": solve_paths, "sweep_graph": sweep_graph,
            "converging": converging, "full_solve": full_solve, "e2e_phis": e2e_phis, "multi_gpu_check": mg_check,
            "allreduce": (None if world == 1 else ("fused into the reduction kernel over NVLink peer memory" if getattr(be, "p2p", False) else "ncclAllReduce")),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
