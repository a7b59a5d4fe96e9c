#!/usr/bin/env python
"""bench.py -- baseline-Gibbs-iterations/sec on the HERA-like shape (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one Gibbs iteration (GCR solve for all times + power-spectrum draw) of every
baseline resident on the GPU (1024 baselines / 8 GPUs = 128 per GPU; weak scaling: every rank
holds its own 128).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "baseline-Gibbs-iterations/sec at Nfreq=384, Ntimes=1024, Nfg=32 on 1/8 B200"
UNIT = "baseline-iterations/s"
FP64_PEAK_NOMINAL_TFLOPS = 37.2  # 148 SMs x 64 FP64 FMA/clk x 2 x 1.965 GHz


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--baselines-per-gpu", type=int, default=128)
    ap.add_argument("--nfreq", type=int, default=384)
    ap.add_argument("--ntimes", type=int, default=1024)
    ap.add_argument("--nfg", type=int, default=32)
    ap.add_argument("--e2e-baselines", type=int, default=32)
    ap.add_argument("--e2e-iters", type=int, default=16)
    ap.add_argument("--substreams", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def make_baseline(seed, nt, nf, nm):
    """Synthetic HERA-like baseline: flat-ish EoR delay spectrum, 32 smooth foreground modes with a
    steep amplitude spectrum, white noise, ~5 % of the channels flagged at all times."""
    rng = np.random.default_rng(seed)
    x = np.linspace(-1.0, 1.0, nf)
    F = np.linalg.qr(np.polynomial.legendre.legvander(x, nm - 1))[0].astype(np.complex128)
    crandn = lambda *s: (rng.standard_normal(s) + 1j * rng.standard_normal(s)) / np.sqrt(2)  # noqa: E731
    eor = crandn(nt, nf)
    amps = crandn(nt, nm) * np.logspace(3, 0, nm)
    sigma = 0.5
    vis = eor + amps @ F.T + sigma * crandn(nt, nf)
    flags = np.ones(nf, dtype=bool)
    flags[rng.choice(nf, max(1, nf // 20), replace=False)] = False
    ninv_diag = np.full(nf, 1.0 / sigma ** 2)
    lam0sq = np.ones(nf)  # S_initial = identity (run-hydra-pspec.py:425)
    return vis, flags, F, ninv_diag, lam0sq


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=3)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower() == "active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "power_w_max": max(float(s[2]) for s in self.samples)}


# ------------------------------------------------------------------------------------------------
def _oracle_worker(args):
    seed, nt, nf, nm, niter = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import hydra_oracle as ho
    vis, flags, F, ninv_diag, lam0sq = make_baseline(seed, nt, nf, nm)
    t0 = time.perf_counter()
    for it in range(niter):
        # each Gibbs iteration restarts from S_initial: the per-iteration cost of the reference algorithm
        # (2 x sqrtm, pinv, Ntimes preconditioned CG solves) without the risk of its CG stagnating on a
        # later, coloured S sample (1e5 iterations per time)
        ho.gibbs_sample_with_fg(vis, flags, np.eye(nf), F, np.diag(ninv_diag), np.zeros((2, nf)), Niter=1,
                                seed=seed + it, solver="cg")
    return time.perf_counter() - t0


def cpu_reference_throughput(nt, nf, nm, steps, warmup):
    """The reference's CPU algorithm (oracle port: sqrtm + pinv + per-time preconditioned CG,
    hydra_pspec/pspec.py:325-374, 151-235) with one single-threaded process per host core -- the
    `mpirun -n <cores>` layout of run-hydra-pspec.py.  Each step = `cores` baselines x 1 iteration."""
    import multiprocessing as mp
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        if warmup:
            pool.map(_oracle_worker, [(1000 + i, 64, nf // 4, max(nm // 4, 1), 1) for i in range(cores)])
        t0 = time.perf_counter()
        pool.map(_oracle_worker, [(i, nt, nf, nm, steps) for i in range(cores)])
        dt = time.perf_counter() - t0
    return cores * steps / dt, cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    t_wall = time.perf_counter()
    val, cores, dt = cpu_reference_throughput(args.ntimes, args.nfreq, args.nfg, steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": f"HERA-like baselines Nfreq={args.nfreq} Ntimes={args.ntimes} Nfg={args.nfg}",
                   "sample": f"{cores} baselines x {steps} Gibbs iterations, one single-threaded process per core"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} baselines x {steps} iterations of the oracle port (scipy sqrtm/pinv/CG), "
                                   f"wall {time.perf_counter() - t_wall:.1f} s"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from hydra_pspec_b200 import _lib, pspec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (hydra_pspec_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    # one process per GPU: page-locked e2e buffers next to this rank's GPU (single-GPU runs keep all host cores:
    # the CPU baseline leg uses them)
    numa_node = _lib.bind_to_device_numa(local_rank, verbose=True) if world > 1 else None
    # stdout carries exactly one JSON line: anything libraries print at the fd level while the job runs
    # (NCCL's "NCCL version ..." banner on a box with NCCL_DEBUG set) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nt, nf, nm, B = args.ntimes, args.nfreq, args.nfg, args.baselines_per_gpu
    K, W = args.steps, max(args.warmup, 3)
    KP = 3  # extra, untimed steps on a single stream for the per-kernel CUDA-event timings
    N = nf + nm
    # a non-default torch stream: the engine launches on it, and torch.cuda.Event times it
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- resident chains (baselines are independent: rank r holds global baselines r*B .. r*B+B-1)
    eng = pspec.GibbsEngine(B, nt, nf, nm, max_iters=K + W + KP, rng="philox", cg_compat=False, refresh_omega=True,
                            keep=(), seed=1234 + rank, device=local_rank, stream=stream, substreams=args.substreams)
    t_load = time.perf_counter()
    for c in range(B):
        vis, flags, F, ninv_diag, lam0sq = make_baseline(rank * B + c, nt, nf, nm)
        eng.load_chain(c, vis, flags, F, ninv_diag, lam0sq)
    t_load = time.perf_counter() - t_load

    eng.run(W)
    eng.sync()
    torch.cuda.synchronize()
    barrier()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record()
        eng.run(K)
        ev1.record()
        torch.cuda.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    # per-kernel durations: same engine, sub-batches off so that the kernels do not overlap
    eng.set_substreams(1)
    eng.set_profile(True)
    eng.kernel_ms(reset=True)
    eng.run(KP)
    kms = eng.kernel_ms(reset=True)
    eng.set_profile(False)
    bad = int(np.count_nonzero(eng.info()))
    ps_last = eng.signal_ps(0, W + K + KP - 1, 1)
    finite = bool(np.all(np.isfinite(ps_last)))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * K / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (k_solve): 8 N^2 T real flops per baseline-iteration
    solve_ms, solve_launches = kms["solve"]
    flops_per_launch = 8.0 * N * N * nt * B
    achieved = flops_per_launch / (solve_ms / max(solve_launches, 1) * 1e-3) * 1e-12
    peak = _lib.lib().hp_fp64_peak_tflops(local_rank, 0.3)
    roofline = {
        "bound": "tensor", "kernel": "k_solve", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak if peak > 0 else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of one k_solve launch of this exact workload, from the
        # `ncu --set full` capture summarised in profiles/r1_v9_summary.md (1.090 GB + 0.857 GB); the algorithmic
        # bytes of a launch are Rfix in + X out + W once = 16 B * (2 T N + N^2 / 2) * B = 1.94 GB
        "traffic": 1.947e9 if (nt, nf, nm, B) == (1024, 384, 32, 128) else None,
        "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
        "peak_source": "FP64 DMMA.8x8x4 issue loop measured in this run (hp_fp64_peak_tflops); "
                       "MEASURED_PEAKS.json has no FP64 entry; nominal 37.2",
        "flops_per_launch": flops_per_launch,
        "step_share": {k: v[0] / max(sum(x[0] for x in kms.values()), 1e-9) for k, v in kms.items()},
        "kernel_ms_per_step": {k: v[0] / KP for k, v in kms.items()},
        "kernel_timing": f"CUDA events around each launch, {KP} extra steps on one stream after the timed region",
    }
    # algorithmic flops of a step: complex Cholesky N^3/6 complex MACs (= 4 N^3 / 3 real flops) + two triangular
    # solves for T right-hand sides (N^2 T complex MACs = 8 N^2 T); the explicit W = L^-1 is an implementation choice
    step_flops = (4.0 * N ** 3 / 3 + 8.0 * N * N * nt) * B
    roofline["step_tflops"] = step_flops * K / (ms_max * 1e-3) * 1e-12
    roofline["step_frac_of_peak"] = roofline["step_tflops"] / peak if peak > 0 else None
    # the other kernels of the step against their own rooflines (CUDA-event times of the same untimed steps)
    post_ms = kms["post"][0] / KP
    chol_ms = kms["chol"][0] / KP
    post_bytes = 16.0 * nt * (N + 2 * nf) * B          # X in, flags*vis in, frequency-space signal out
    roofline["other_kernels"] = {
        "k_post_fft": {"bound": "hbm", "achieved": post_bytes / (post_ms * 1e-3) * 1e-9 if post_ms > 0 else None, "unit": "GB/s",
                       "algorithmic_bytes_per_launch": post_bytes},
        "k_chol_col + k_trinv": {"bound": "tensor", "unit": "TFLOP/s",
                                 "achieved": (8.0 * N ** 3 / 3) * B / (chol_ms * 1e-3) * 1e-12 if chol_ms > 0 else None,
                                 "flops_per_step": (8.0 * N ** 3 / 3) * B,
                                 "peak": peak, "note": "4 N^3/3 (Cholesky) + 4 N^3/3 (explicit inverse of the factor)"},
    }
    _ck = roofline["other_kernels"]["k_chol_col + k_trinv"]
    _ck["frac"] = _ck["achieved"] / peak if (_ck["achieved"] and peak > 0) else None
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        if roofline["other_kernels"]["k_post_fft"]["achieved"]:
            roofline["other_kernels"]["k_post_fft"]["peak"] = peaks["hbm_gbs"]
            roofline["other_kernels"]["k_post_fft"]["frac"] = roofline["other_kernels"]["k_post_fft"]["achieved"] / peaks["hbm_gbs"]
        hbm_bytes = 16.0 * (3 * nt * N + 3 * nt * nf) * B  # Rfix, eta, X + Ssc, Sf, z per baseline-iteration
        roofline["hbm"] = {"algorithmic_gbs": hbm_bytes * K / (ms_max * 1e-3) * 1e-9, "peak_gbs": peaks["hbm_gbs"]}
    except Exception:
        pass
    eng.close()

    # ---- end to end through the public API with host buffers (full reference outputs)
    e2e = None
    if not args.no_e2e:
        Be, Ke = min(args.e2e_baselines, B), args.e2e_iters
        host = [make_baseline(10_000 + rank * Be + c, nt, nf, nm) for c in range(Be)]
        pin = []
        for vis, flags, F, nd, l0sq in host:
            pv = _lib.pinned_empty(vis.shape, np.complex128)
            pv[...] = vis
            pin.append((pv, flags, F, nd, l0sq))
        bufs = None

        def one_call():
            nonlocal bufs
            e = pspec.GibbsEngine(Be, nt, nf, nm, max_iters=Ke, rng="philox", keep=("cr", "fg", "chisq"),
                                  seed=99 + rank, device=local_rank, stream=stream)
            if bufs is None:
                bufs = e.host_buffers(Ke)   # page-locked destination arrays, allocated once by the caller
            for c, (pv, flags, F, nd, l0sq) in enumerate(pin):
                e.load_chain(c, pv, flags, F, nd, l0sq)
            e.run_to_host(Ke, bufs)         # compute overlapped with the device-to-host copies
            e.close()

        one_call()  # warm-up (allocation paths, first-touch)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            one_call()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        # the same call keeping only what a power-spectrum analysis consumes (signal_ps + ln_post per iteration):
        # shows how much of the end-to-end time is the reference's full return set crossing PCIe
        bufs_ps = None

        def one_call_ps():
            nonlocal bufs_ps
            e = pspec.GibbsEngine(Be, nt, nf, nm, max_iters=Ke, rng="philox", keep=(), seed=99 + rank, device=local_rank,
                                  stream=stream)
            if bufs_ps is None:
                bufs_ps = e.host_buffers(Ke)
            for c, (pv, flags, F, nd, l0sq) in enumerate(pin):
                e.load_chain(c, pv, flags, F, nd, l0sq)
            e.run_to_host(Ke, bufs_ps)
            e.close()

        one_call_ps()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            one_call_ps()
        torch.cuda.synchronize()
        dt_ps = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt_ps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_ps = float(tt.item())
        h2d = sum(pv.nbytes + F.nbytes + nd.nbytes + l0sq.nbytes + flags.size for pv, flags, F, nd, l0sq in pin) / Ke
        d2h = sum(v.nbytes for v in bufs.values()) / Ke
        e2e = {"value": world * Be * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "call": f"GibbsEngine(create) + load_chain x{Be} from pinned host arrays + run_to_host({Ke} iterations): "
                       "signal_cr/fg_amps/chisq/signal_ps/ln_post of every iteration (the reference's full return set, "
                       "10.3 MB per baseline-iteration) land in pinned host arrays; PCIe-bound",
               "pcie_gbs": d2h * Ke / dt * 1e-9, "numa_node": numa_node,
               "power_spectrum_only": {"value": world * Be * Ke / dt_ps, "unit": UNIT,
                                       "d2h_bytes_per_step": int(sum(v.nbytes for v in bufs_ps.values()) / Ke),
                                       "call": "same call with keep=(): only signal_ps + ln_post are read back"}}

    # ---- CPU baseline on this box (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, dtc = cpu_reference_throughput(nt, nf, nm, 1, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cores} baselines x 1 Gibbs iteration of the oracle port, one single-threaded process per "
                         f"core ({dtc:.1f} s)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": f"HERA-like (BASELINE.json configs[3]): {B} baselines per GPU "
                                   f"(1024 / 8), Nfreq={nf} Ntimes={nt} Nfg={nm}, time-invariant flags (5 %), "
                                   "diagonal noise, device Philox draws, exact solves",
                       "baselines_per_gpu": B, "substreams": args.substreams, "parallelism": f"baseline-sharded x{world}, no hot-path collective",
                       "l2": f"per-step working set ~{B * 16 * (4 * nt * N + 4 * nt * nf) / 2**30:.1f} GiB per GPU >> 126 MB L2",
                       "outputs_kept": "signal_ps + ln_post per iteration (value); full reference return set (e2e)",
                       "load_s": round(t_load, 2)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk.summary(), "chol_failures": bad, "finite": finite,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
