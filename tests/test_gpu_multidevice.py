"""Sharding baselines over GPUs must not change any sample: the Philox draws of a chain depend on
(seed, global baseline index, iteration) only.  Runs under torchrun-free conditions: one process drives two engines on
two devices (or, on a one-GPU box, two engines on the same device holding the two shards)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _baselines(nbl, nt, nf, nm):
    import bench
    out = []
    for i in range(nbl):
        vis, flags, F, nd, l0 = bench.make_baseline(500 + i, nt, nf, nm)
        out.append(dict(vis=vis, flags=flags, fgmodes=F, ninv_diag=nd, lam0sq=l0))
    return out


def _run(bls, ids, device, niter, nt, nf, nm, seed=7):
    from hydra_pspec_b200 import pspec
    eng = pspec.GibbsEngine(len(ids), nt, nf, nm, max_iters=niter, rng="philox", keep=("cr",), seed=seed, device=device)
    eng.set_chain_ids(np.asarray(ids, dtype=np.int32))
    for c, gi in enumerate(ids):
        b = bls[gi]
        eng.load_chain(c, b["vis"] * b["flags"], b["flags"], b["fgmodes"], b["ninv_diag"], b["lam0sq"])
    eng.run(niter)
    ps = np.stack([eng.signal_ps(c) for c in range(len(ids))])
    cr = np.stack([eng.signal_cr(c) for c in range(len(ids))])
    eng.close()
    return ps, cr


def test_samples_do_not_depend_on_the_sharding():
    import torch
    from hydra_pspec_b200 import driver
    nt, nf, nm, nbl, niter = 32, 96, 6, 5, 4
    bls = _baselines(nbl, nt, nf, nm)
    ps1, cr1 = _run(bls, list(range(nbl)), 0, niter, nt, nf, nm)
    ndev = torch.cuda.device_count()
    for world in (2, 3):
        shards = driver.split_data_for_scatter(list(range(nbl)), world)
        parts = [_run(bls, sh, r % ndev, niter, nt, nf, nm) for r, sh in enumerate(shards)]
        ps = np.concatenate([p[0] for p in parts])
        cr = np.concatenate([p[1] for p in parts])
        assert np.array_equal(ps, ps1), world      # bit for bit
        assert np.array_equal(cr, cr1), world


def test_run_baselines_uses_global_chain_ids():
    """driver.run_baselines on one rank == the same baselines run as one engine with ids 0..n-1 and the same seed."""
    from hydra_pspec_b200 import driver
    nt, nf, nm, nbl, niter = 32, 96, 6, 3, 3
    bls = _baselines(nbl, nt, nf, nm)
    for b in bls:
        b["vis"] = b["vis"] * b["flags"]
    ps, lp = driver.run_baselines(bls, Niter=niter, seed=7, device=0)
    ps1, _ = _run(bls, list(range(nbl)), 0, niter, nt, nf, nm, seed=7)
    assert np.array_equal(ps, ps1)
