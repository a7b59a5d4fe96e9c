"""Where the end-to-end time goes: raw pinned D2H bandwidth, and the phases of bench.py's e2e call
(create / load / run_to_host / close) for several iteration counts."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402
from hydra_pspec_b200 import _lib, pspec  # noqa: E402

nt, nf, nm, Be = 1024, 384, 32, 32
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4):
        h.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"raw D2H pinned: {4 * (1 << 30) / dt / 1e9:.1f} GB/s")
del x, h
host = [bench.make_baseline(100 + c, nt, nf, nm) for c in range(Be)]
pin = []
for vis, flags, F, nd, l0 in host:
    pv = _lib.pinned_empty(vis.shape, np.complex128); pv[...] = vis; pin.append((pv, flags, F, nd, l0))
for Ke in (16, 32):
    bufs = None
    for rep in range(3):
        t0 = time.perf_counter()
        e = pspec.GibbsEngine(Be, nt, nf, nm, max_iters=Ke, rng="philox", keep=("cr", "fg", "chisq"), seed=1)
        t1 = time.perf_counter()
        if bufs is None:
            bufs = e.host_buffers(Ke)
        for c, (pv, flags, F, nd, l0) in enumerate(pin):
            e.load_chain(c, pv, flags, F, nd, l0)
        e.sync(); t2 = time.perf_counter()
        e.run_to_host(Ke, bufs); t3 = time.perf_counter()
        e.close(); t4 = time.perf_counter()
        nb = sum(v.nbytes for v in bufs.values())
        print(f"Ke={Ke} rep{rep}: create {1e3 * (t1 - t0):.1f} load {1e3 * (t2 - t1):.1f} run_to_host {1e3 * (t3 - t2):.1f} "
              f"({nb / 1e9 / (t3 - t2):.1f} GB/s) close {1e3 * (t4 - t3):.1f} ms -> {Be * Ke / (t4 - t0):.0f} it/s")
    del bufs
