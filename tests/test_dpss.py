"""DPSS foreground modes (reference: hydra_pspec/dpss.py) against the unmodified reference's output
(golden made in the build container by importing /root/reference/hydra_pspec/dpss.py)."""
import numpy as np


def test_dpss_modes_and_fit_match_reference(golden_dir):
    from hydra_pspec_b200 import dpss
    g = np.load(golden_dir / "dpss_fit.npz")
    nm, alpha = int(g["nmodes"]), float(g["alpha"])
    modes = dpss.dpss_modes(g["freqs"].size, nm, alpha)
    np.testing.assert_allclose(modes, g["modes"], rtol=0, atol=1e-14)
    assert modes.shape == (nm, g["freqs"].size)  # .T is the (Nfreqs, Nmodes) fgmodes of the sampler
    m2, amps = dpss.dpss_fit_modes(g["d"], g["w"], g["freqs"], g["cov"], nmodes=nm, alpha=alpha)
    np.testing.assert_allclose(m2, g["modes"], rtol=0, atol=1e-14)
    # the reference stops L-BFGS-B at its default tolerance: agreement to ~1e-5 of the amplitude scale
    scale = np.max(np.abs(g["amps"]))
    assert np.max(np.abs(amps - g["amps"])) < 2e-4 * scale
    _, amps_t = dpss.dpss_fit_modes(g["d"], g["w"], g["freqs"], g["cov"], nmodes=nm, alpha=alpha, taper=g["taper"])
    assert np.max(np.abs(amps_t - g["amps_taper"])) < 2e-4 * scale


def test_dpss_fit_is_the_exact_minimiser():
    from hydra_pspec_b200 import dpss
    rng = np.random.default_rng(3)
    nf, nm = 40, 5
    freqs = np.linspace(100.0, 110.0, nf)
    w = (rng.random(nf) > 0.1).astype(float)
    X = rng.standard_normal((nf, 2 * nf))
    cov = X @ X.T / (2 * nf) + 0.1 * np.eye(nf)
    d = rng.standard_normal(nf) + 1j * rng.standard_normal(nf)
    modes, amps = dpss.dpss_fit_modes(d, w, freqs, cov, nmodes=nm, alpha=2.0)
    invcov = np.linalg.inv(cov)

    def loglike(p):
        m = np.sum(p[0::2, None] * modes + 1j * p[1::2, None] * modes, axis=0)
        x = w * (d - m)
        return 0.5 * np.real(x.conj() @ invcov @ x)
    f0 = loglike(amps)
    for _ in range(20):
        assert loglike(amps + 1e-3 * rng.standard_normal(amps.size)) >= f0 - 1e-12
