import time, numpy as np, sys
sys.path.insert(0, '/root/repo')
import bench
from hydra_pspec_b200 import _lib, pspec
nt, nf, nm = 1024, 384, 32
for Be, Ke in [(8, 8), (32, 8)]:
    host = [bench.make_baseline(100 + c, nt, nf, nm) for c in range(Be)]
    pin = []
    for vis, flags, F, nd, l0 in host:
        pv = _lib.pinned_empty(vis.shape, np.complex128); pv[...] = vis; pin.append((pv, flags, F, nd, l0))
    outs = {k: _lib.pinned_empty((Be, Ke) + shp, dt) for k, shp, dt in [("cr", (nt, nf), np.complex128), ("fg", (nt, nm), np.complex128), ("chisq", (nt, nf), np.float64), ("ps", (nf,), np.float64), ("lnp", (), np.float64)]}
    for rep in range(2):
        t0 = time.perf_counter()
        e = pspec.GibbsEngine(Be, nt, nf, nm, max_iters=Ke, rng="philox", keep=("cr", "fg", "chisq"), seed=1)
        t1 = time.perf_counter()
        for c, (pv, flags, F, nd, l0) in enumerate(pin): e.load_chain(c, pv, flags, F, nd, l0)
        e.sync(); t2 = time.perf_counter()
        e.run(Ke); e.sync(); t3 = time.perf_counter()
        L = _lib.lib()
        for c in range(Be):
            for key, buf in (("ps", 0), ("lnp", 1), ("cr", 2), ("fg", 3), ("chisq", 4)):
                dst = outs[key][c]; _lib.check(L.hp_engine_read(e._h, c, buf, 0, Ke, _lib.ptr(dst), dst.nbytes))
        t4 = time.perf_counter()
        e.close(); t5 = time.perf_counter()
        print(f"Be={Be} Ke={Ke} rep{rep}: create {1e3*(t1-t0):.1f} load {1e3*(t2-t1):.1f} run {1e3*(t3-t2):.1f} read {1e3*(t4-t3):.1f} ({sum(v.nbytes for v in outs.values())/1e9/(t4-t3):.1f} GB/s) close {1e3*(t5-t4):.1f} ms -> {Be*Ke/(t5-t0):.0f} it/s")
