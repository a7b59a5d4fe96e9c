"""Multi-process host logic of the baseline sharding (world_size 2, gloo, CPU only)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FakeEngine:
    """Stands in for GibbsEngine on a machine without a GPU: 'samples' encode the loaded data."""

    def __init__(self, nchains, ntimes, nfreqs, nmodes, max_iters, **kw):
        self.n, self.nf, self.it, self.data, self.seed = nchains, nfreqs, max_iters, {}, kw.get("seed", 0)

    def load_chain(self, c, vis, flags, fgmodes, ninv_diag, lam0sq, ps_prior=None):
        self.data[c] = float(np.real(vis[0, 0]))

    def run(self, n):
        pass

    def signal_ps(self, c):
        return np.full((self.it, self.nf), self.data[c]) + np.arange(self.it)[:, None]

    def ln_post(self, c):
        return np.full(self.it, -self.data[c])

    def close(self):
        pass


def _worker(rank, world, port, nbl, q):
    import torch.distributed as dist
    from hydra_pspec_b200 import driver
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bls = [dict(vis=np.full((4, 6), 10.0 * i, complex), flags=np.ones(6, bool), fgmodes=np.ones((6, 2), complex),
                    ninv_diag=np.ones(6), lam0sq=np.ones(6)) for i in range(nbl)]
        ps, lp = driver.run_baselines(bls, Niter=3, engine_factory=FakeEngine)
        local = np.arange(driver.shard_counts(nbl, world)[rank] * 2, dtype=float).reshape(-1, 2) + 100 * rank
        g = driver.gather_samples(local, nbl)
        q.put((rank, ps, lp, g))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nbl", [2, 5])
def test_sharded_run_and_gather(nbl):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nbl, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ps, lp, g in res:
        # every rank ends up with all baselines, in global order
        assert ps.shape == (nbl, 3, 6) and lp.shape == (nbl, 3)
        assert np.array_equal(ps[:, 0, 0], 10.0 * np.arange(nbl))
        assert np.array_equal(lp[:, 0], -10.0 * np.arange(nbl))
        c0 = (nbl + 1) // 2
        want = np.concatenate([np.arange(c0 * 2, dtype=float).reshape(-1, 2),
                               np.arange((nbl - c0) * 2, dtype=float).reshape(-1, 2) + 100])
        assert np.array_equal(g, want)


def test_split_matches_reference_rule():
    from hydra_pspec_b200 import driver
    assert [len(x) for x in driver.split_data_for_scatter(list(range(10)), 4)] == [3, 3, 2, 2]
    assert driver.split_data_for_scatter(list("abcde"), 2) == [["a", "b", "c"], ["d", "e"]]
    assert driver.shard_counts(1024, 8) == [128] * 8
    with pytest.raises(ValueError):
        driver.split_data_for_scatter([1, 2], 3)
