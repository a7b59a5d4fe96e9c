"""Phase timers of k_solve3 at the headline shape (HP_S3_TIMERS=1 selects the instrumented instantiation):

    python profiles/scripts/solve3_timers.py [substreams]
"""
import ctypes as C
import os
import sys
from pathlib import Path

os.environ["HP_S3_TIMERS"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from hydra_pspec_b200 import _lib, pspec  # noqa: E402
from bench import make_baseline  # noqa: E402

B, nt, nf, nm, K = 128, 1024, 384, 32, 4
eng = pspec.GibbsEngine(B, nt, nf, nm, max_iters=K + 2, rng="philox", keep=(), seed=7,
                        substreams=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for c in range(B):
    eng.load_chain(c, *make_baseline(c, nt, nf, nm))
eng.run(2)
eng.sync()
L = C.CDLL(str(_lib.LIB_PATH))
out = (C.c_ulonglong * 8)()
L.hp_solve3_timers(out, 1)
eng.run(K)
eng.sync()
L.hp_solve3_timers(out, 0)
nw = 8 * B * (nt // 16) * K   # warp-tiles
names = ["pass 1 strips", "philox + y store", "barrier (y complete)", "pass 2 strips", "x store", "wait y buffer",
         "wait right-hand sides", "whole loop"]
for n_, v in zip(names, out):
    print(f"{n_:24s} {v / nw:10.0f} cycles per warp-tile")
print("k-steps per warp-tile: 351 (pass 1 + pass 2 = 2808 / 8); DMMA-pipe time if alone: 351 x 12 x 16 = 67 k cycles")
eng.close()
