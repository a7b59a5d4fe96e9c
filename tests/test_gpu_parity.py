"""Parity of the CUDA path (through the C ABI / drop-in API) with the reference.  GPU only.

Three layers:
  * committed golden vectors produced by the unmodified reference (tests/golden/make_golden.py):
    numpy draw streams injected, solver="reference-cg"  -> tolerance 1e-10 (north_star);
  * the CPU oracle on fresh seeded inputs, exact solver on both sides -> 1e-10;
  * size-independent properties at larger sizes (residual of the linear system).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import hydra_oracle as ho  # noqa: E402  (checker only)

TOL = 1e-10  # relative, complex128 (north_star)
KEYS = ["signal_cr", "signal_S", "signal_ps", "fg_amps", "chisq", "ln_post"]


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    if b.size == 0:
        return 0.0
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def crandn(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)


@pytest.mark.parametrize("name", ["A_testdata", "B_defaults", "C_nonuniform"])
def test_chain_matches_reference_golden(golden_dir, name):
    from hydra_pspec_b200 import pspec
    g = np.load(golden_dir / f"chain_{name}.npz", allow_pickle=True)
    out = pspec.gibbs_sample_with_fg(g["vis"], g["flags"], g["S_initial"], g["fgmodes"], g["Ninv"], g["ps_prior"],
                                     Niter=int(g["Niter"]), seed=int(g["seed"]), verbose=False)
    for o, k in zip(out[:6], KEYS):
        assert rel(o, g[k]) < (1e-8 if k == "chisq" else TOL) * (50 if name == "A_testdata" else 1), k


def test_map_estimate_matches_reference_golden(golden_dir):
    from hydra_pspec_b200 import pspec
    g = np.load(golden_dir / "chain_E_map.npz", allow_pickle=True)
    np.random.seed(1234)  # the reference run drew sample_S's uniforms from this global state
    out = pspec.gibbs_sample_with_fg(g["vis"], g["flags"], g["S_initial"], g["fgmodes"], g["Ninv"], g["ps_prior"],
                                     Niter=5, map_estimate=True, verbose=False)
    assert out[0].shape[0] == 1  # map_estimate forces Niter = 1 (pspec.py:572-574)
    for o, k in zip(out[:5], KEYS[:5]):
        assert rel(o, g[k]) < TOL, k


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("nt,nf,nm,nflag,seed", [(16, 32, 4, 0, 1), (21, 45, 5, 3, 2), (40, 120, 12, 5, 3),
                                                 (33, 96, 0, 2, 4), (9, 37, 3, 1, 5), (5, 77, 2, 0, 6)])
def test_chain_matches_oracle_exact(nt, nf, nm, nflag, seed, dense, monkeypatch):
    """Fresh inputs, numpy draw streams, exact solver on both sides.  Runs through the fused FFT
    kernels and through the dense-transform path (which Nfreqs = 37 takes in any case)."""
    from hydra_pspec_b200 import pspec
    monkeypatch.setattr(pspec, "_FORCE_DENSE_TRANSFORMS", dense)
    rng = np.random.default_rng(seed)
    F = np.linalg.qr(crandn(rng, nf, max(nm, 1)))[0][:, :nm]
    fop = ho.fourier_operator(nf)
    p0 = 0.5 + rng.random(nf)
    S0 = fop.conj().T @ np.diag(p0 / nf ** 2) @ fop
    sig = 0.3 + rng.random(nf)
    vis = crandn(rng, nt, nf) * sig + crandn(rng, nt, nf) @ np.linalg.cholesky(S0 + 1e-13 * np.eye(nf)).T
    if nm:
        vis = vis + (5 * crandn(rng, nt, nm)) @ F.T
    flags = np.ones(nf, dtype=bool)
    flags[rng.choice(nf, nflag, replace=False)] = False
    prior = np.zeros((2, nf))
    prior[0, nf // 2 - 1:nf // 2 + 2] = 50.0
    prior[1, nf // 2 - 1:nf // 2 + 2] = 0.05
    Ninv = np.diag(1.0 / sig ** 2)
    want = ho.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=3, seed=seed, solver="direct")
    got = pspec.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=3, seed=seed, verbose=False, solver="exact")
    for o, w, k in zip(got[:6], want, KEYS):
        assert rel(o, w) < TOL, k


def test_general_initial_covariance():
    """S_initial that is not delay-diagonal: first iteration runs in its eigenbasis."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(11)
    nt, nf, nm = 12, 24, 3
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    Xs = crandn(rng, nf, 2 * nf)
    S0 = Xs @ Xs.conj().T / (2 * nf)
    vis = crandn(rng, nt, nf) @ np.linalg.cholesky(S0).T + 0.5 * crandn(rng, nt, nf) + (4 * crandn(rng, nt, nm)) @ F.T
    Ninv = np.eye(nf) * 4.0
    flags = np.ones(nf, dtype=bool)
    flags[5] = False
    prior = np.zeros((2, nf))
    want = ho.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=2, seed=3, solver="direct")
    got = pspec.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=2, seed=3, verbose=False, solver="exact")
    for o, w, k in zip(got[:6], want, KEYS):
        assert rel(o, w) < TOL, k


def test_dense_noise_matches_reference_golden(golden_dir):
    """Non-diagonal Ninv and a non delay-diagonal S_initial (golden case D, no flags)."""
    from hydra_pspec_b200 import pspec
    g = np.load(golden_dir / "chain_D_dense.npz", allow_pickle=True)
    out = pspec.gibbs_sample_with_fg(g["vis"], g["flags"], g["S_initial"], g["fgmodes"], g["Ninv"], g["ps_prior"],
                                     Niter=int(g["Niter"]), seed=int(g["seed"]), verbose=False)
    for o, k in zip(out[:6], KEYS):
        assert rel(o, g[k]) < 5e-9, k   # the reference's CG is close to stagnation here


@pytest.mark.parametrize("nflag,general_s", [(0, False), (3, False), (2, True)])
def test_dense_noise_matches_oracle_exact(nflag, general_s):
    """Dense Hermitian Ninv with flagged channels (flags on rows and columns), exact solves."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(21 + nflag)
    nt, nf, nm = 18, 40, 4
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    Xn = crandn(rng, nf, 3 * nf)
    Ncov = Xn @ Xn.conj().T / (3 * nf) * 0.3
    Ninv = np.linalg.inv(Ncov)
    fop = ho.fourier_operator(nf)
    if general_s:
        Xs = crandn(rng, nf, 2 * nf)
        S0 = Xs @ Xs.conj().T / (2 * nf)
    else:
        S0 = fop.conj().T @ np.diag((0.5 + rng.random(nf)) / nf ** 2) @ fop
    vis = crandn(rng, nt, nf) @ np.linalg.cholesky(Ncov).T + (5 * crandn(rng, nt, nm)) @ F.T + crandn(rng, nt, nf)
    flags = np.ones(nf, dtype=bool)
    flags[rng.choice(nf, nflag, replace=False)] = False
    prior = np.zeros((2, nf))
    want = ho.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=2, seed=8, solver="direct", symmetric_flags=True)
    got = pspec.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=2, seed=8, verbose=False, solver="exact")
    for o, w, k in zip(got[:6], want, KEYS):
        assert rel(o, w) < TOL, k
    got = pspec.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=40, verbose=False, rng="philox", seed=1)
    assert np.all(np.isfinite(got[2])) and np.all(got[2] > 0)


def test_gcr_and_sample_S_functions(golden_dir):
    """The two parity units north_star names: one GCR solve, one S draw."""
    from hydra_pspec_b200 import pspec
    g = np.load(golden_dir / "chain_C_nonuniform.npz", allow_pickle=True)
    nf = g["vis"].shape[1]
    nm = g["fgmodes"].shape[1]
    mats = pspec.build_matrices(nf + nm, g["flags"], g["S_initial"], g["Ninv"], g["fgmodes"])
    cr = pspec.gcr_fgmodes(g["vis"] * g["flags"], g["flags"], mats, g["fgmodes"])
    want = np.concatenate([g["signal_cr"][0], g["fg_amps"][0]], axis=1)
    assert rel(cr, want) < TOL
    fn = np.load(golden_dir / "functions.npz")
    np.random.seed(77)
    got = pspec.sample_S(s=fn["sampleS_s"], prior=fn["sampleS_prior"])
    assert rel(got, fn["sampleS_out"]) < TOL
    np.random.seed(78)
    got = pspec.sample_S(s=fn["sampleS_s"])
    assert rel(got, fn["sampleS_out_noprior"]) < TOL


def test_error_behaviour():
    from hydra_pspec_b200 import pspec
    vis = np.zeros((4, 8), complex)
    with pytest.raises(AssertionError):
        pspec.gibbs_sample_with_fg(vis, np.ones(7, bool), np.eye(8), np.ones((8, 2)), np.eye(8), np.zeros((2, 8)),
                                   Niter=1, verbose=False)
    bad_prior = np.zeros((2, 8))
    bad_prior[0, 3] = 1.0  # upper bound set, lower bound 0 -> "prior_min must be greater than zero"
    with pytest.raises(ValueError):
        pspec.gibbs_sample_with_fg(vis + 1, np.ones(8, bool), np.eye(8), np.ones((8, 2)), np.eye(8), bad_prior,
                                   Niter=1, verbose=False)
    with pytest.raises(ValueError):
        pspec.sample_S()


def test_full_size_linear_system_residual():
    """HERA-like shape (Nfreq=384, Nfg=32): the solve satisfies the reference's A x = b."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(5)
    nt, nf, nm = 32, 384, 32
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    fop = ho.fourier_operator(nf)
    p0 = 0.5 + rng.random(nf)
    S0 = fop.conj().T @ np.diag(p0 / nf ** 2) @ fop
    vis = crandn(rng, nt, nf) + (30 * crandn(rng, nt, nm)) @ F.T
    flags = np.ones(nf, dtype=bool)
    flags[[10, 11, 200]] = False
    Ninv = np.eye(nf) * 2.0
    mats_dev = pspec.build_matrices(nf + nm, flags, S0, Ninv, F)
    x = pspec.gcr_fgmodes(vis * flags, flags, mats_dev, F, solver="exact")
    oma, omb = ho.reference_gcr_draws(nt, nf)
    # reference operators without sqrtm/pinv: Sh = F^H diag(sqrt(p0)/n^1.5 ...) -- build from the definition
    Sh = fop.conj().T @ np.diag(np.sqrt(p0 / nf) / nf) @ fop
    Ni = np.diag(flags * 2.0).astype(complex)
    Nih = np.sqrt(Ni)
    A = np.zeros((nf + nm, nf + nm), complex)
    A[:nf, :nf] = np.eye(nf) + S0 @ Ni
    A[:nf, nf:] = S0 @ Ni @ F
    A[nf:, :nf] = F.conj().T @ Ni
    A[nf:, nf:] = F.conj().T @ Ni @ F
    worst = 0.0
    for t in range(nt):
        z = Ni @ (flags * vis[t]) + Nih @ omb[t]
        b = np.concatenate([S0 @ z + Sh @ oma[t], F.conj().T @ z])
        worst = max(worst, np.linalg.norm(A @ x[t] - b) / np.linalg.norm(b))
    assert worst < 1e-11


@pytest.mark.parametrize("nt,nf,nm,seed", [(2, 4, 1, 1), (3, 8, 2, 2), (5, 6, 0, 3), (17, 64, 40, 4), (33, 31, 3, 5)])
def test_edge_shapes_match_oracle(nt, nf, nm, seed):
    """Smallest shapes the API accepts, a system smaller than one 32x32 block, more foreground modes than a block,
    a prime Nfreqs (dense transforms), Ntimes not a multiple of any tile size."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(100 + seed)
    F = np.linalg.qr(crandn(rng, nf, max(nm, 1)))[0][:, :nm]
    fop = ho.fourier_operator(nf)
    S0 = fop.conj().T @ np.diag((0.5 + rng.random(nf)) / nf ** 2) @ fop
    sig = 0.3 + rng.random(nf)
    vis = crandn(rng, nt, nf) * sig + crandn(rng, nt, nf) @ np.linalg.cholesky(S0 + 1e-13 * np.eye(nf)).T
    if nm:
        vis = vis + (5 * crandn(rng, nt, nm)) @ F.T
    flags = np.ones(nf, dtype=bool)
    flags[rng.integers(nf)] = False
    prior = np.zeros((2, nf))
    Ninv = np.diag(1.0 / sig ** 2)
    want = ho.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=3, seed=seed, solver="direct")
    got = pspec.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=3, seed=seed, verbose=False, solver="exact")
    for o, w, k in zip(got[:6], want, KEYS):
        assert rel(o, w) < TOL, k


def test_prior_on_every_bin_and_wide_dynamic_range():
    """Every delay bin prior-bounded (inversion sampler everywhere) and an initial spectrum spanning six decades."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(321)
    nt, nf, nm = 14, 40, 4
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    fop = ho.fourier_operator(nf)
    p0 = 10.0 ** rng.uniform(-3, 3, nf)
    S0 = fop.conj().T @ np.diag(p0 / nf ** 2) @ fop
    vis = crandn(rng, nt, nf) * 0.5 + (5 * crandn(rng, nt, nm)) @ F.T + crandn(rng, nt, nf) @ np.linalg.cholesky(
        S0 + 1e-12 * np.eye(nf)).T
    flags = np.ones(nf, dtype=bool)
    flags[[3, 30]] = False
    prior = np.zeros((2, nf))
    prior[0, :] = 1e4
    prior[1, :] = 1e-4
    Ninv = np.eye(nf) * 4.0
    want = ho.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=3, seed=9, solver="direct")
    got = pspec.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=3, seed=9, verbose=False, solver="exact")
    for o, w, k in zip(got[:6], want, KEYS):
        assert rel(o, w) < 1e-8, k   # cond(A) ~ 1e6 on the oracle's (unwhitened) side
    assert np.all(got[2] >= 1e-4) and np.all(got[2] <= 1e4)


def _headline_baseline(seed):
    import bench  # the synthetic HERA-like baseline of the benchmark (repo root is on sys.path via conftest)
    return bench.make_baseline(seed, 1024, 384, 32)


def test_headline_shape_chain_matches_oracle():
    """BASELINE.json configs[3] shape (Ntimes=1024, Nfreq=384, Nfg=32: 13 block rows, 64 time tiles): two Gibbs
    iterations with the reference's numpy draws injected and exact solves, every output against the oracle."""
    from hydra_pspec_b200 import pspec
    nf = 384
    vis, flags, F, nd, _ = _headline_baseline(3)
    Ninv, prior = np.diag(nd), np.zeros((2, nf))
    want = ho.gibbs_sample_with_fg(vis, flags, np.eye(nf), F, Ninv, prior, Niter=2, seed=5, solver="direct")
    got = pspec.gibbs_sample_with_fg(vis, flags, np.eye(nf), F, Ninv, prior, Niter=2, seed=5, verbose=False, solver="exact")
    for o, w, k in zip(got[:6], want, KEYS):
        assert rel(o, w) < TOL, k


def test_headline_shape_batch_substreams_matches_oracle():
    """The same shape through gibbs_sample_batch: 8 chains on 4 sub-streams (two distinct baselines, each loaded four
    times), numpy draws, exact solves."""
    from hydra_pspec_b200 import pspec
    nf = 384
    data = [_headline_baseline(11), _headline_baseline(12)]
    prior = np.zeros((2, nf))
    want = [ho.gibbs_sample_with_fg(v, fl, np.eye(nf), F, np.diag(nd), prior, Niter=2, seed=8, solver="direct")
            for v, fl, F, nd, _ in data]
    bls = [dict(vis=data[c % 2][0], flags=data[c % 2][1], S_initial=np.eye(nf), fgmodes=data[c % 2][2],
                Ninv=np.diag(data[c % 2][3]), ps_prior=prior) for c in range(8)]
    got = pspec.gibbs_sample_batch(bls, Niter=2, seed=8, rng="numpy", solver="exact", substreams=4)
    for c in range(8):
        for o, w, k in zip(got[c][:6], want[c % 2], KEYS):
            assert rel(o, w) < TOL, (c, k)


def test_long_chain_through_the_output_ring(monkeypatch, tmp_path):
    """Niter >> ring depth (3 device slots) and a staging area of two iterations: every iteration's signal_cr / fg_amps /
    chisq must still arrive (bounded device output ring), in RAM and through memory-mapped .npy files."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(77)
    nt, nf, nm, niter = 20, 48, 4, 11
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    vis = crandn(rng, nt, nf) + (10 * crandn(rng, nt, nm)) @ F.T
    flags = np.ones(nf, dtype=bool)
    flags[[7, 30]] = False
    prior = np.zeros((2, nf))
    Ninv = np.eye(nf) * 1.5
    want = ho.gibbs_sample_with_fg(vis, flags, np.eye(nf), F, Ninv, prior, Niter=niter, seed=1, solver="direct")
    per_iter = nt * nf * 24 + nt * nm * 16 + nf * 8 + 8
    monkeypatch.setattr(pspec, "_STAGING_BYTES", 2 * per_iter)          # two iterations per chunk
    got = pspec.gibbs_sample_with_fg(vis, flags, np.eye(nf), F, Ninv, prior, Niter=niter, seed=1, verbose=False, solver="exact")
    for o, w, k in zip(got[:6], want, KEYS):
        assert rel(o, w) < TOL, k
    # batch of three copies: in RAM ...
    bls = [dict(vis=vis, flags=flags, S_initial=np.eye(nf), fgmodes=F, Ninv=Ninv, ps_prior=prior) for _ in range(3)]
    monkeypatch.setattr(pspec, "_STAGING_BYTES", 2 * 3 * per_iter)
    outs = pspec.gibbs_sample_batch(bls, Niter=niter, seed=1, rng="numpy", solver="exact", substreams=2)
    for o in outs:
        for a, w, k in zip(o[:6], want, KEYS):
            assert rel(a, w) < TOL, k
    # ... and through memory-mapped files when the host budget says the arrays do not fit
    monkeypatch.setenv("HP_HOST_BUDGET_GB", "0")
    for c, b in enumerate(bls):
        b["out_dir"] = tmp_path / f"bl{c}"
        b["out_dir"].mkdir()
    outs = pspec.gibbs_sample_batch(bls, Niter=niter, seed=1, rng="numpy", solver="exact", write_Niter=4)
    for c, o in enumerate(outs):
        assert isinstance(o[0], np.memmap)
        for a, w, k in zip(o[:6], want, KEYS):
            assert rel(np.asarray(a), w) < TOL, k
        d = bls[c]["out_dir"]
        assert rel(np.load(d / "gcr-eor.npy"), want[0]) < TOL and rel(np.load(d / "chisq.npy"), want[4]) < TOL
        assert rel(np.load(d / "dps-eor.npy"), want[2]) < TOL and rel(np.load(d / "ln-post.npy"), want[5]) < TOL
    with pytest.raises(MemoryError):
        pspec.gibbs_sample_batch([dict(b, out_dir=None) for b in bls], Niter=niter, seed=1, rng="numpy", solver="exact")


def test_device_ring_read_window():
    """hp_engine_read of a big output: the last `ring_iters` iterations are readable, older ones are refused."""
    from hydra_pspec_b200 import pspec, _lib
    rng = np.random.default_rng(5)
    nt, nf, nm = 16, 32, 2
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    vis = crandn(rng, nt, nf)
    eng = pspec.GibbsEngine(1, nt, nf, nm, max_iters=8, rng="philox", keep=("cr",), seed=3, ring_iters=3)
    eng.load_chain(0, vis, np.ones(nf, bool), F, np.ones(nf), np.ones(nf))
    eng.run(8)
    last = eng.signal_cr(0, 5, 3)
    assert last.shape == (3, nt, nf) and np.all(np.isfinite(last))
    with pytest.raises(_lib.HydraLibError):
        eng.signal_cr(0, 4, 1)
    eng.close()


@pytest.mark.parametrize("nt,nf,nm,seed", [(20, 128, 8, 1), (12, 256, 16, 2), (9, 384, 32, 3), (17, 128, 0, 4), (8, 256, 6, 5),
                                           (11, 512, 20, 6), (10, 1024, 40, 7), (9, 1024, 64, 8)])
def test_register_fft_shapes_match_oracle(nt, nf, nm, seed):
    """Nfreqs = 128 / 256 / 384 / 512 / 1024 take k_post_fft2 (register-resident FFTs, plans 4.4.4.2 / 8.8.4 / 6.4.4.4 / 8.8.4.2 /
    8.8.4.4): partial time tiles, Nmodes = 0, Nmodes not a multiple of 4 and beyond the 32 pre-fetched foreground k-steps,
    flagged channels (second transform), three iterations."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(500 + seed)
    F = np.linalg.qr(crandn(rng, nf, max(nm, 1)))[0][:, :nm]
    fop = ho.fourier_operator(nf)
    S0 = fop.conj().T @ np.diag((0.5 + rng.random(nf)) / nf ** 2) @ fop
    sig = 0.3 + rng.random(nf)
    vis = crandn(rng, nt, nf) * sig + crandn(rng, nt, nf) @ np.linalg.cholesky(S0 + 1e-13 * np.eye(nf)).T
    if nm:
        vis = vis + (5 * crandn(rng, nt, nm)) @ F.T
    flags = np.ones(nf, dtype=bool)
    flags[rng.choice(nf, 5, replace=False)] = False
    prior = np.zeros((2, nf))
    Ninv = np.diag(1.0 / sig ** 2)
    want = ho.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=3, seed=seed, solver="direct")
    got = pspec.gibbs_sample_with_fg(vis, flags, S0, F, Ninv, prior, Niter=3, seed=seed, verbose=False, solver="exact")
    for o, w, k in zip(got[:6], want, KEYS):
        assert rel(o, w) < TOL, k


def test_verbose_prints_the_reference_table(capsys):
    """pspec.py:602-604: header and one row per iteration (Iter, Time, Info, |Ax - b|, Chisq, ln Post)."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(3)
    nt, nf, nm = 6, 32, 3
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    vis = crandn(rng, nt, nf) + (3 * crandn(rng, nt, nm)) @ F.T
    flags = np.ones(nf, dtype=bool)
    flags[4] = False
    out = pspec.gibbs_sample_with_fg(vis, flags, np.eye(nf), F, np.eye(nf), np.zeros((2, nf)), Niter=3, seed=1, verbose=True)
    text = capsys.readouterr().out.splitlines()
    assert text[0].split() == "Iter Time [s] Info |Ax - b| Chisq ln Post".split()
    rows = [ln.split() for ln in text[2:5]]
    assert [r[0] for r in rows] == ["1", "2", "3"]
    for r, lp, cs in zip(rows, out[5], out[4]):
        assert abs(float(r[-1]) - lp) <= 0.05 + 1e-6 * abs(lp)
        assert abs(float(r[-2]) - cs[:, flags].mean()) <= 1e-3 * max(1.0, cs[:, flags].mean())


def test_read_ahead_gives_identical_samples_and_guards_the_state():
    """hp_host_sink.read_ahead: the next iterations are computed under a chunk's last copies.  Same samples bit for bit
    (Philox draws, 4 chains on 2 sub-streams, 9 iterations through a 3-slot ring in chunks of 2); while read-ahead iterations
    are pending the calls that would see the advanced chain state are refused."""
    from hydra_pspec_b200 import pspec, _lib
    rng = np.random.default_rng(77)
    nt, nf, nm, nch, niter = 20, 64, 4, 4, 9
    F = np.linalg.qr(crandn(rng, nf, nm))[0]
    vis = crandn(rng, nch, nt, nf) + (3 * crandn(rng, nch, nt, nm)) @ F.T
    flags = np.ones(nf, dtype=bool)
    flags[[3, 40]] = False

    def run(read_ahead):
        eng = pspec.GibbsEngine(nch, nt, nf, nm, max_iters=niter, rng="philox", keep=("cr", "fg", "chisq"), ring_iters=3, seed=5,
                                substreams=2)
        for c in range(nch):
            eng.load_chain(c, vis[c] * flags, flags, F, np.ones(nf), np.full(nf, 1.0 / nf))
        stage = eng.host_buffers(2, iter_major=True)
        out = {k: [] for k in ("signal_cr", "fg_amps", "chisq", "signal_ps", "ln_post")}
        done = 0
        while done < niter:
            c = min(2, niter - done)
            ra = read_ahead if done + c < niter else 0
            eng.run_to_host(c, stage, first_iter=done, iter_major=True, read_ahead=ra)
            if ra:
                with pytest.raises(_lib.HydraLibError):
                    eng.signal_S(0)
                with pytest.raises(_lib.HydraLibError):
                    eng.run(1)
            for k in out:
                big = k in ("signal_cr", "fg_amps", "chisq")
                out[k].append(np.array(stage[k][:c] if big else stage[k][:, :c]))
            done += c
        S = eng.signal_S(0)
        eng.close()
        return {k: np.concatenate(v, axis=0 if k in ("signal_cr", "fg_amps", "chisq") else 1) for k, v in out.items()}, S

    a, Sa = run(0)
    for ra in (1, 2, 5):
        b, Sb = run(ra)
        for k in a:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
        np.testing.assert_array_equal(Sa, Sb)
