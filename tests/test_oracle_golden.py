"""Pin the CPU oracle against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import hydra_oracle as ho

CHAINS = ["A_testdata", "B_defaults", "C_nonuniform", "D_dense", "E_map"]


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b))


@pytest.fixture(scope="module")
def fn(golden_dir):
    return np.load(golden_dir / "functions.npz")


def test_fourier_operator(fn):
    for n in (4, 5, 16):
        assert rel(ho.fourier_operator(n), fn[f"fourier_operator_{n}"]) < 1e-15


def test_covariance_from_pspec(fn):
    out = ho.covariance_from_pspec(fn["cov_ps"], ho.fourier_operator(12))
    assert rel(out, fn["cov_out"]) < 1e-14


def test_sprior(fn):
    assert rel(ho.sprior(fn["sprior_in"], 2, 10.0), fn["sprior_out"]) < 1e-14


def test_gcr_draw_stream(fn):
    oma, omb = ho.reference_gcr_draws(6, 6)
    for idx in (0, 5):
        assert np.array_equal(oma[idx], fn[f"gcrdraw_a_{idx}"])
        assert np.array_equal(omb[idx], fn[f"gcrdraw_b_{idx}"])


def test_inversion_sample_invgamma(fn):
    for (alpha, beta, lo, hi), u, want in zip(fn["invsamp_cases"], fn["invsamp_u"], fn["invsamp_out"]):
        got = ho.inversion_sample_invgamma(alpha, beta, lo, hi, u)
        assert abs(got - want) <= 1e-14 * abs(want), (alpha, beta, lo, hi)


def test_inversion_sample_errors():
    for args in [(5.0, 1.0, 0.0, 1.0), (5.0, 1.0, 1.0, -1.0), (5.0, 1.0, 1.0, np.inf), (5.0, 1.0, 2.0, 1.0)]:
        with pytest.raises(ValueError):
            ho.inversion_sample_invgamma(*args, u=0.5)


def test_sample_S(fn):
    got, _ = ho.sample_S(s=fn["sampleS_s"], prior=fn["sampleS_prior"], u=fn["sampleS_u"])
    assert rel(got, fn["sampleS_out"]) < 1e-13
    got, _ = ho.sample_S(s=fn["sampleS_s"], prior=None, u=fn["sampleS_u_noprior"])
    assert rel(got, fn["sampleS_out_noprior"]) < 1e-13


@pytest.mark.parametrize("name", CHAINS)
def test_chain_matches_reference(golden_dir, name):
    g = np.load(golden_dir / f"chain_{name}.npz", allow_pickle=True)
    map_est = bool(g["map_estimate"])
    draws = None
    seed = int(g["seed"])
    if map_est:
        nt, nf = g["vis"].shape
        draws = (np.zeros((nt, nf), complex), np.zeros((nt, nf), complex), g["u_used"])
    cr, S, ps, fg, chisq, lnp = ho.gibbs_sample_with_fg(
        g["vis"], g["flags"], g["S_initial"], g["fgmodes"], g["Ninv"], g["ps_prior"],
        Niter=int(g["Niter"]), seed=seed, map_estimate=map_est, solver="cg", draws=draws)
    # the oracle calls the same scipy CG as the reference: agreement is at round-off level
    assert rel(cr, g["signal_cr"]) < 1e-11
    assert rel(fg, g["fg_amps"]) < 1e-11
    assert rel(ps, g["signal_ps"]) < 1e-11
    assert rel(S, g["signal_S"]) < 1e-11
    assert rel(chisq, g["chisq"]) < 1e-9
    if not map_est:
        # map_estimate: a delay bin with beta ~ 0 makes S_sample singular (cond ~ 1e16), so the
        # reference's ln_post (through np.linalg.inv) is round-off noise there.
        assert rel(lnp, g["ln_post"]) < 1e-10


@pytest.mark.parametrize("name", ["A_testdata", "B_defaults", "C_nonuniform", "D_dense"])
def test_theta_chain_matches_reference(golden_dir, name):
    """Direct solve x scalar CG model (what the CUDA cg_compat mode computes) replays the
    reference chain, including iterations where the reference's CG stagnates (case B, it 3)."""
    g = np.load(golden_dir / f"chain_{name}.npz", allow_pickle=True)
    out = ho.gibbs_sample_with_fg(
        g["vis"], g["flags"], g["S_initial"], g["fgmodes"], g["Ninv"], g["ps_prior"],
        Niter=int(g["Niter"]), seed=int(g["seed"]), solver="theta")
    for o, k in zip(out, ["signal_cr", "signal_S", "signal_ps", "fg_amps", "chisq", "ln_post"]):
        assert rel(o, g[k]) < 5e-9, k


@pytest.mark.parametrize("name", ["A_testdata", "C_nonuniform"])
def test_cg_theta_model(golden_dir, name):
    """theta * (direct solve) reproduces the reference's CG output to 1e-10."""
    g = np.load(golden_dir / f"chain_{name}.npz", allow_pickle=True)
    vis, flags, F = g["vis"], g["flags"], g["fgmodes"]
    nt, nf = vis.shape
    mats = ho.build_matrices(flags, g["S_initial"], g["Ninv"], F)
    oma, omb = ho.reference_gcr_draws(nt, nf)
    want = np.concatenate([g["signal_cr"][0], g["fg_amps"][0]], axis=1)
    worst_plain = 0.0
    for t in range(nt):
        d = (vis * flags)[t]
        b = ho.gcr_rhs(d, flags, mats, F, oma[t], omb[t])
        xd = ho.gcr_solve_1d(d, flags, mats, F, oma[t], omb[t], "direct")
        th = ho.cg_theta(np.vdot(b, xd), np.linalg.norm(b))
        assert rel(th * xd, want[t]) < 1e-10
        worst_plain = max(worst_plain, rel(xd, want[t]))
    # and the plain direct solve differs from the reference by CG's own stopping error
    assert worst_plain < 5e-8


def test_per_time_flags_extension_matches_reference_time_by_time(golden_dir):
    """The oracle's 2-D flag extension against the unmodified reference's gcr_fgmodes_1d called one time
    at a time with that time's operators (tests/golden/make_golden_pertime.py)."""
    g = np.load(golden_dir / "gcr_pertime.npz")
    nt = g["vis"].shape[0]
    oma, omb = ho.reference_gcr_draws(*g["vis"].shape)
    mats = [ho.build_matrices(g["flags"][t], g["S"], g["Ninv"], g["fgmodes"]) for t in range(nt)]
    out = ho.gcr_fgmodes(g["vis"] * g["flags"], [g["flags"][t] for t in range(nt)], mats, g["fgmodes"], oma, omb, solver="cg")
    assert np.max(np.abs(out - g["cr"])) / np.max(np.abs(g["cr"])) < 1e-12
