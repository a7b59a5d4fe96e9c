"""Phase timers of k_post_fft at the headline shape.  Needs the library built with the timers:

    cd hydra_pspec_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \\
        -DHP_POST_TIMERS -c hp_fft.cu -o hp_fft.o && nvcc -shared -o libhydra_pspec_b200.so hp_*.o -lcudart

(rebuild with build.sh afterwards: touch hp_fft.cu first)."""
import ctypes as C
import sys
from pathlib import Path

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
L.hp_post_timers(out, 1)
eng.run(K)
eng.sync()
L.hp_post_timers(out, 0)
nct = B * (nt // 8) * K
names = ["tables + f + X load", "FFT 1 (5 passes)", "post-twiddle + Sf store", "F f (DMMA)", "residual (data load)", "FFT 2 + E sums"]
tot = sum(out[:6])
for n_, v in zip(names, out):
    print(f"{n_:26s} {100 * v / tot:5.1f} %   {v / nct:9.0f} cycles per CTA")
print(f"total {tot / nct:.0f} cycles per CTA")
eng.close()
