"""Throughput of the per-time-flags path at BASELINE.json configs[2]:
128 baselines, Nfreq=256, Ntimes=512, Nfg=16, ~5 % random RFI flags per time, device Philox draws."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from hydra_pspec_b200 import pspec  # noqa: E402
from bench import make_baseline  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nt, nf, nm = 512, 256, 16
K, W = 5, 2
eng = pspec.GibbsEngine(B, nt, nf, nm, max_iters=K + W, rng="philox", keep=(), seed=7, time_flags=True)
t0 = time.perf_counter()
for c in range(B):
    vis, flags, F, ninv_diag, lam0sq = make_baseline(c, nt, nf, nm)
    rng = np.random.default_rng(1000 + c)
    fl = np.broadcast_to(flags, (nt, nf)).copy()
    fl &= rng.random((nt, nf)) > 0.05
    eng.load_chain(c, vis, fl, F, ninv_diag, lam0sq)
print("load s", time.perf_counter() - t0)
eng.run(W)
eng.sync()
eng.set_profile(True)
t0 = time.perf_counter()
eng.run(K)
eng.sync()
dt = (time.perf_counter() - t0) / K
flop = B * nt * (8.0 / 3.0) * (nf + nm) ** 3
print(f"step {dt * 1e3:.2f} ms  {B / dt:.1f} baseline-it/s  chol {flop / dt / 1e12:.2f} TFLOP/s (8N^3/3 per time)")
print(eng.kernel_ms())
print("bad", int(np.count_nonzero(eng.info())), "finite", bool(np.all(np.isfinite(eng.signal_ps(0)))))
eng.close()
