"""What limits a real DMMA k-step?  Runs profiles/microbench/dmma_mix.cu (built as libdmma_mix.so) and prints cycles per
DMMA per scheduler for the instruction mixes of the solve kernels."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
L = C.CDLL(str(ROOT / "profiles" / "microbench" / "libdmma_mix.so"))
L.dmma_mix.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
modes = [(0, "same operands"), (1, "distinct register operands"), (3, "distinct + 4 DADD interleaved"),
         (19, "distinct + 4 DADD up front"), (5, "distinct + 8 LDS.64, no DADD"), (7, "distinct + 4 DADD + 8 LDS.64"),
         (15, "distinct + 4 DADD + 4 LDS.64 + 2 LDG.128 (L2 stream)"), (33, "distinct, 16 DMMA (4M), no DADD"),
         (37, "distinct, 16 DMMA (4M) + 8 LDS.64")]
for warps in (4, 8, 16):
    for m, name in modes:
        c, t = C.c_double(), C.c_double()
        rc = L.dmma_mix(m, warps, 4000, C.byref(c), C.byref(t))
        print(f"{name:56s} warps/SM={warps:2d}  rc={rc}  {c.value:6.2f} cycles/DMMA/scheduler  {t.value:6.2f} TFLOP/s", flush=True)
