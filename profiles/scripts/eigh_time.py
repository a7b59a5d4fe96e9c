"""Time hp_eigh_batch (csrc/hp_eigh.cu) against numpy.linalg.eigh: 384 x 384 Hermitian matrices, batches of 1 / 16 / 128."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from hydra_pspec_b200 import pspec

rng = np.random.default_rng(0)
n = 384
X = rng.standard_normal((128, n, n)) + 1j * rng.standard_normal((128, n, n))
S = X @ np.conj(np.swapaxes(X, 1, 2))
pspec.device_eigh(S[:2])
for b in (1, 16, 128):
    t = time.perf_counter()
    w, V = pspec.device_eigh(S[:b])
    print("device_eigh batch", b, "%.3f s" % (time.perf_counter() - t))
t = time.perf_counter()
np.linalg.eigh(S[:4])
print("numpy eigh x4 %.3f s" % (time.perf_counter() - t))
