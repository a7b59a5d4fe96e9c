"""Phase timers of k_solve at the headline shape.  Needs the library built with the timers:

    cd hydra_pspec_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \\
        -DHP_SOLVE_TIMERS -c hp_solve.cu -o hp_solve.o && nvcc -shared -o libhydra_pspec_b200.so hp_*.o -lcudart

(rebuild with build.sh afterwards: touch hp_solve.cu first)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from hydra_pspec_b200 import _lib, pspec  # noqa: E402
from bench import make_baseline  # noqa: E402

B, nt, nf, nm, K = 128, 1024, 384, 32, 4
eng = pspec.GibbsEngine(B, nt, nf, nm, max_iters=K + 2, rng="philox", keep=(), seed=7)
for c in range(B):
    eng.load_chain(c, *make_baseline(c, nt, nf, nm))
eng.run(2)
eng.sync()
L = C.CDLL(str(_lib.LIB_PATH))
out = (C.c_ulonglong * 8)()
L.hp_solve_timers(out, 1)
eng.run(K)
eng.sync()
L.hp_solve_timers(out, 0)
nct = B * (nt // 16) * K
names = ["staging", "pass 1", "pass 2", "epilogue", "(ring waits)", "(row exchange + barrier)"]
tot = sum(out[:4])
for n_, v in zip(names, out):
    print(f"{n_:26s} {100 * v / tot:5.1f} %   {v / nct:9.0f} cycles per CTA")
print(f"total {tot / nct:.0f} cycles per CTA")
eng.close()
