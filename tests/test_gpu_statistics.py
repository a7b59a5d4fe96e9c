"""Device-RNG (Philox) chains against an oracle chain: posterior mean and variance of every delay
bin must agree within Monte-Carlo error (north_star, second correctness criterion).  GPU only."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import hydra_oracle as ho  # noqa: E402  (checker only)


def crandn(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)


def batch_mean_stats(x, nbatch=20):
    """mean, and standard error of the mean from batch means (handles autocorrelation)."""
    n = (x.shape[0] // nbatch) * nbatch
    b = x[:n].reshape(nbatch, -1, *x.shape[1:]).mean(axis=1)
    return x[:n].mean(axis=0), b.std(axis=0, ddof=1) / np.sqrt(nbatch)


def make_problem(seed, nt, nf, nm, nflag):
    rng = np.random.default_rng(seed)
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    p_true = 1.0 + 3.0 * rng.random(nf)
    fop = ho.fourier_operator(nf)
    S_true = fop.conj().T @ np.diag(p_true / nf ** 2) @ fop * nf  # arbitrary overall scale
    sig = 0.7
    vis = crandn(rng, nt, nf) @ np.linalg.cholesky(S_true).T + sig * crandn(rng, nt, nf) + (6 * crandn(rng, nt, nm)) @ F.T
    flags = np.ones(nf, dtype=bool)
    if nflag:
        flags[rng.choice(nf, nflag, replace=False)] = False
    prior = np.zeros((2, nf))
    prior[0, nf // 2] = 1e3
    prior[1, nf // 2] = 1e-3
    return vis, flags, F, np.eye(nf) / sig ** 2, prior


def oracle_chain(vis, flags, F, Ninv, prior, niter, seed):
    """Proper Gibbs chain with the reference's conditional draws but NEW fluctuation terms in every
    iteration (the reference re-uses them, pspec.py:195-197; see test_frozen_omega_semantics)."""
    rng = np.random.default_rng(seed)
    nt, nf = vis.shape
    S = np.eye(nf, dtype=complex)
    visf = vis * flags
    ps = np.zeros((niter, nf))
    for i in range(niter):
        oma, omb = crandn(rng, nt, nf), crandn(rng, nt, nf)
        _, S, ps[i], _, _, _ = ho.gibbs_step_fgmodes(visf, flags, S, F, Ninv, prior, oma, omb, rng.uniform(size=nf),
                                                     solver="direct")
    return ps


@pytest.mark.parametrize("nflag", [0, 2])
def test_posterior_moments_agree(nflag):
    from hydra_pspec_b200 import pspec
    nt, nf, nm = 24, 16, 2
    vis, flags, F, Ninv, prior = make_problem(3 + nflag, nt, nf, nm, nflag)
    burn, niter = 100, 3000
    ps_o = oracle_chain(vis, flags, F, Ninv, prior, 1200, seed=17)[burn:]
    out = pspec.gibbs_sample_with_fg(vis, flags, np.eye(nf), F, Ninv, prior, Niter=niter, seed=4242, verbose=False,
                                     rng="philox")
    ps_d = out[2][burn:]
    assert np.all(np.isfinite(ps_d)) and np.all(ps_d > 0)
    lo, ld = np.log(ps_o), np.log(ps_d)  # the conditionals are heavy tailed; compare moments of log ps
    mo, so = batch_mean_stats(lo)
    md, sd = batch_mean_stats(ld)
    z = (md - mo) / np.sqrt(so ** 2 + sd ** 2)
    assert np.max(np.abs(z)) < 5.0, z
    vo, svo = batch_mean_stats((lo - mo) ** 2)
    vd, svd = batch_mean_stats((ld - md) ** 2)
    zv = (vd - vo) / np.sqrt(svo ** 2 + svd ** 2)
    assert np.max(np.abs(zv)) < 5.0, zv
    # and the sampler actually moves: posterior width of a bin is ~ 1/sqrt(Ntimes)
    assert np.all(np.sqrt(vd) > 0.3 / np.sqrt(nt))


def test_posterior_moments_agree_per_time_flags():
    """Same criterion for the per-time-flags path (k_pt_cholsolve): the white draw added between the forward
    and the backward substitution must give every time its own correctly distributed constrained realisation."""
    from hydra_pspec_b200 import pspec
    nt, nf, nm = 24, 16, 2
    vis, _, F, Ninv, prior = make_problem(21, nt, nf, nm, 0)
    rng = np.random.default_rng(77)
    flags = rng.random((nt, nf)) > 0.15
    flags[:, 5] = False
    burn, niter = 100, 3000
    ps_o = oracle_chain(vis, flags, F, Ninv, prior, 1200, seed=19)[burn:]
    out = pspec.gibbs_sample_with_fg(vis, flags, np.eye(nf), F, Ninv, prior, Niter=niter, seed=777, verbose=False,
                                     rng="philox")
    ps_d = out[2][burn:]
    assert np.all(np.isfinite(ps_d)) and np.all(ps_d > 0)
    lo, ld = np.log(ps_o), np.log(ps_d)
    mo, so = batch_mean_stats(lo)
    md, sd = batch_mean_stats(ld)
    z = (md - mo) / np.sqrt(so ** 2 + sd ** 2)
    assert np.max(np.abs(z)) < 5.0, z
    vo, svo = batch_mean_stats((lo - mo) ** 2)
    vd, svd = batch_mean_stats((ld - md) ** 2)
    zv = (vd - vo) / np.sqrt(svo ** 2 + svd ** 2)
    assert np.max(np.abs(zv)) < 5.0, zv


def test_frozen_omega_semantics():
    """refresh_omega=False reproduces the reference's re-use of the fluctuation draws: with the
    power spectrum held fixed by a tight prior, consecutive GCR solutions are identical."""
    from hydra_pspec_b200 import pspec
    nt, nf, nm = 16, 16, 2
    vis, flags, F, Ninv, _ = make_problem(9, nt, nf, nm, 0)
    ninv_diag = np.real(np.diag(Ninv)).copy()
    for refresh, same in ((False, True), (True, False)):
        eng = pspec.GibbsEngine(1, nt, nf, nm, max_iters=1, rng="philox", refresh_omega=refresh, keep=(), seed=5)
        eng.load_chain(0, vis, flags, F, ninv_diag, np.ones(nf))
        eng.gcr()
        a = eng.last_gcr(0)
        eng.gcr()
        b = eng.last_gcr(0)
        eng.close()
        assert np.array_equal(a, b) == same


def test_chains_are_independent_and_seeded():
    from hydra_pspec_b200 import pspec
    nt, nf, nm = 16, 16, 2
    vis, flags, F, Ninv, prior = make_problem(10, nt, nf, nm, 1)
    ninv_diag = np.real(np.diag(Ninv)).copy()

    def run(seed):
        eng = pspec.GibbsEngine(3, nt, nf, nm, max_iters=4, rng="philox", keep=(), seed=seed)
        for c in range(3):
            eng.load_chain(c, vis, flags, F, ninv_diag, np.ones(nf), ps_prior=prior)
        eng.run(4)
        out = [eng.signal_ps(c) for c in range(3)]
        eng.close()
        return out

    a, b, c = run(1), run(1), run(2)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)                   # same seed -> bit-identical
    assert not np.array_equal(a[0], c[0])             # different seed
    assert not np.array_equal(a[0], a[1])             # identical data, different chain id -> different draws


def test_driver_run_baselines_single_rank():
    from hydra_pspec_b200 import driver
    nt, nf, nm = 16, 16, 2
    bls = []
    for i in range(3):
        vis, flags, F, Ninv, prior = make_problem(20 + i, nt, nf, nm, 1)
        bls.append(dict(vis=vis, flags=flags, fgmodes=F, ninv_diag=np.real(np.diag(Ninv)).copy(), lam0sq=np.ones(nf),
                        ps_prior=prior))
    ps, lp = driver.run_baselines(bls, Niter=5, seed=3)
    assert ps.shape == (3, 5, nf) and lp.shape == (3, 5)
    assert np.all(np.isfinite(ps)) and np.all(ps > 0) and np.all(np.isfinite(lp))
