"""Phase timers of k_pt_cholsolve at configs[2].  Needs the library built with the timers:

    cd hydra_pspec_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \\
        -DHP_PT_TIMERS -c hp_pertime.cu -o hp_pertime.o && nvcc -shared -o libhydra_pspec_b200.so hp_*.o -lcudart

(rebuild with build.sh afterwards: touch hp_pertime.cu first).  Prints the share of every phase and the time per
(baseline, time) system."""
import sys, ctypes as C, subprocess, json
sys.path.insert(0, "/root/repo")
sys.argv = ["x", "2"]
exec(open("/root/repo/profiles/scripts/bench_configs.py").read().replace("eng.close()", ""))
from hydra_pspec_b200 import _lib
L = C.CDLL(str(_lib.LIB_PATH))
out = (C.c_ulonglong * 8)()
L.hp_pt_timers(out, 0)
tot = sum(out)
names = ["prologue", "j-loop", "generation", "diag", "y_k+store", "trsm", "noise+backward", "X store"]
for n_, v in zip(names, out):
    print(f"{n_:16s} {100 * v / tot:5.1f} %   {v / (B * nt * (K + W)) / 1.965e3:8.2f} us per item")
