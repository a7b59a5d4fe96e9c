"""Throughput of the BASELINE.json configurations that are not the bench.py headline.

    python profiles/scripts/bench_configs.py 2 [baselines] [hold]   configs[2]: Nfreq=256 Ntimes=512 Nfg=16, per-time RFI flags
                                                              (hold > 1: every random mask is kept for `hold` consecutive times)
    python profiles/scripts/bench_configs.py 4 [baselines]   configs[4]: Nfreq=1024 Nfg=64 Ntimes=1024, dense noise covariance

Device Philox draws, exact solves, signal_ps + ln_post kept (like bench.py's `value`).  Prints one JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from hydra_pspec_b200 import pspec  # noqa: E402
from bench import make_baseline  # noqa: E402

cfg = int(sys.argv[1])
K, W = 5, 2
if cfg == 2:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    nt, nf, nm = 512, 256, 16
    eng = pspec.GibbsEngine(B, nt, nf, nm, max_iters=K + W, rng="philox", keep=(), seed=7, time_flags=True)
elif cfg == 4:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    nt, nf, nm = 1024, 1024, 64
    eng = pspec.GibbsEngine(B, nt, nf, nm, max_iters=K + W, rng="philox", keep=(), seed=7, dense_noise=True, substreams=2)
else:
    raise SystemExit("config must be 2 or 4")
t0 = time.perf_counter()
rng = np.random.default_rng(99)
if cfg == 4:
    Xn = (rng.standard_normal((nf, 2 * nf)) + 1j * rng.standard_normal((nf, 2 * nf))) / np.sqrt(2)
    Ncov = 0.25 * (Xn @ Xn.conj().T / (2 * nf) + 0.2 * np.eye(nf))
    Ninv = np.linalg.inv(Ncov)
for c in range(B):
    vis, flags, F, ninv_diag, lam0sq = make_baseline(c, nt, nf, nm)
    if cfg == 2:
        fl = np.broadcast_to(flags, (nt, nf)).copy()
        hold = int(sys.argv[3]) if len(sys.argv) > 3 else 1
        rnd = np.random.default_rng(1000 + c).random(((nt + hold - 1) // hold, nf)) > 0.05
        fl &= np.repeat(rnd, hold, axis=0)[:nt]
        eng.load_chain(c, vis, fl, F, ninv_diag, lam0sq)
    else:
        eng.load_chain(c, vis, flags, F, np.real(np.diagonal(Ninv)).copy(), lam0sq, ninv_dense=Ninv)
load_s = time.perf_counter() - t0
eng.run(W)
eng.sync()
eng.set_profile(True)
t0 = time.perf_counter()
eng.run(K)
eng.sync()
dt = (time.perf_counter() - t0) / K
N = nf + nm
chol_flop = B * (nt if cfg == 2 else 1) * (4.0 / 3.0) * N ** 3      # complex Cholesky: N^3/6 complex MACs x 8
solve_flop = B * 8.0 * N * N * nt
out = {"config": cfg, "baselines": B, "Nfreq": nf, "Ntimes": nt, "Nfg": nm, "ms_per_step": dt * 1e3,
       "baseline_iterations_per_s": B / dt, "load_s": load_s,
       "algorithmic_tflops": (chol_flop + (0 if cfg == 2 else solve_flop)) / dt / 1e12,
       "kernel_ms_per_step": {k: v[0] / K for k, v in eng.kernel_ms().items()},
       "chol_failures": int(np.count_nonzero(eng.info())), "finite": bool(np.all(np.isfinite(eng.signal_ps(0))))}
print(json.dumps(out))
eng.close()
