"""Phase timers of k_solve2 at the headline shape (HP_S2_TIMERS=1 selects the instrumented instantiation):

    HP_S2_TIMERS=1 python profiles/scripts/solve2_timers.py [substreams]
"""
import ctypes as C
import os
import sys
from pathlib import Path

os.environ["HP_S2_TIMERS"] = "1"
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
out = (C.c_ulonglong * 16)()
L.hp_solve2_timers(out, 1)
eng.run(K)
eng.sync()
L.hp_solve2_timers(out, 0)
ntile = B * (nt // 16) * K
nw = 8 * ntile   # consumer-warp tiles
names = ["wait W block", "wait rhs rows", "wait exchange barrier", "philox draws", "hand-over", "collect + finish", "pass 1",
         "pass 2", "consumer barrier", "blocks acquired", "early probes ok", "whole loop"]
for n_, v in zip(names, out):
    print(f"{n_:24s} {v / nw:10.0f} per warp-tile")
print(f"blocks per tile expected {2 * 91}, probe success rate {out[10] / max(out[9], 1):.3f}")
eng.close()
