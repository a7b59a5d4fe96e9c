#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the reference checkout does not exist on the GPU
box): ``python tests/golden/make_golden.py``.  Output: ``tests/golden/*.npz``.

pyuvdata / astropy are not installed in this image; ``hydra_pspec/utils.py``
imports them at module scope for its uvh5 helpers, none of which the Gibbs hot
path (``hydra_pspec/pspec.py``) calls, so they are replaced by empty stub modules.
"""
import sys
import types
import warnings
from pathlib import Path

import numpy as np
import scipy.special

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


def load_reference():
    for name in ["pyuvdata", "pyuvdata.utils", "astropy", "astropy.units"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pyuvdata"].UVData = type("UVData", (), {})
    sys.modules["pyuvdata"].utils = sys.modules["pyuvdata.utils"]
    sys.modules["astropy"].units = sys.modules["astropy.units"]
    sys.modules["astropy.units"].Quantity = type("Quantity", (), {})
    sys.path.insert(0, str(REF))
    import hydra_pspec
    return hydra_pspec


def cplx_normal(rng, shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)


def run_chain(hp, name, vis, flags, S0, F, Ninv, ps_prior, Niter, seed, map_estimate=False):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cr, S, ps, fg, chisq, lnp, _ = hp.pspec.gibbs_sample_with_fg(
            vis.copy(), flags.copy(), S0.copy(), F.copy(), Ninv.copy(), ps_prior.copy(),
            Niter=Niter, seed=seed, verbose=False, nproc=1, map_estimate=map_estimate)
    np.savez_compressed(
        HERE / f"chain_{name}.npz", vis=vis, flags=flags, S_initial=S0, fgmodes=F, Ninv=Ninv,
        ps_prior=ps_prior, Niter=Niter, seed=seed, map_estimate=map_estimate,
        signal_cr=cr, signal_S=S, signal_ps=ps, fg_amps=fg, chisq=chisq, ln_post=lnp)
    print(f"chain_{name}: cr {cr.shape} ps[0,:3]={ps[0, :3]} lnpost={lnp}")


def main():
    hp = load_reference()
    td = REF / "test_data" / "0-1"
    S_td = np.load(td / "eor-cov.npy")
    F_td = np.load(td / "fgmodes.npy")
    Ncov_td = np.load(td / "noise-cov.npy")
    noise_td = np.load(td / "noise.npy")

    # ---- case A: the reference's own test_data operators (config.yaml: Nfgmodes=12, prior
    # 0.1..2 on +-3 bins, seed 7123689), first 24 times, three flagged channels.
    rng = np.random.default_rng(20240601)
    nt, nf, nm = 24, 120, 12
    F = F_td[:, :nm]
    Lc = np.linalg.cholesky((S_td + S_td.conj().T) / 2)
    eor = cplx_normal(rng, (nt, nf)) @ Lc.T
    amps = cplx_normal(rng, (nt, nm)) * np.logspace(1.5, -1, nm)
    vis = eor + amps @ F.T + noise_td[:nt]
    flags = np.ones(nf, dtype=bool)
    flags[[17, 18, 77]] = False
    ps_prior = np.zeros((2, nf))
    ps_prior[0, nf // 2 - 3:nf // 2 + 4] = 2.0
    ps_prior[1, nf // 2 - 3:nf // 2 + 4] = 0.1
    run_chain(hp, "A_testdata", vis, flags, S_td, F, np.linalg.inv(Ncov_td), ps_prior, 3, 7123689)

    # ---- case B: run-hydra-pspec.py defaults: S_initial = I, Legendre fgmodes, Ninv = I/100,
    # no flags, no prior.
    rng = np.random.default_rng(7)
    nt, nf, nm = 10, 32, 4
    F = np.array([scipy.special.legendre(i)(np.linspace(-1.0, 1.0, nf)) for i in range(nm)]).T
    vis = 3.0 * cplx_normal(rng, (nt, nf)) + (20 * cplx_normal(rng, (nt, nm))) @ F.T
    run_chain(hp, "B_defaults", vis, np.ones(nf, dtype=bool), np.eye(nf), F, np.eye(nf) / 100.0,
              np.zeros((2, nf)), 4, 11)

    # ---- case C: non-uniform diagonal noise, flags, odd Nfreqs, complex fgmodes, prior on 2 bins.
    rng = np.random.default_rng(8)
    nt, nf, nm = 7, 45, 5
    F = np.linalg.qr(cplx_normal(rng, (nf, nm)))[0]
    fop = hp.utils.fourier_operator(nf)
    p0 = 0.5 + rng.random(nf)
    S0 = fop.conj().T @ np.diag(p0 / nf ** 2) @ fop
    sig = 0.3 + rng.random(nf)
    vis = cplx_normal(rng, (nt, nf)) * sig + (5 * cplx_normal(rng, (nt, nm))) @ F.T \
        + cplx_normal(rng, (nt, nf)) @ np.linalg.cholesky(S0 + 1e-12 * np.eye(nf)).T
    flags = np.ones(nf, dtype=bool)
    flags[[0, 20, 21, 44]] = False
    ps_prior = np.zeros((2, nf))
    ps_prior[0, [22, 23]] = 50.0
    ps_prior[1, [22, 23]] = 0.5
    run_chain(hp, "C_nonuniform", vis, flags, S0, F, np.diag(1.0 / sig ** 2), ps_prior, 3, 99)

    # ---- case D: general (non delay-diagonal) S_initial and dense Hermitian Ninv, no flags.
    # S and N are kept close to white: for coloured S the reference's CG stagnates (its A is
    # not Hermitian; once arg(b^H A^-1 b) exceeds ~0.04 rad it runs to maxiter=1e5).
    rng = np.random.default_rng(9)
    nt, nf, nm = 12, 24, 3
    F = np.linalg.qr(cplx_normal(rng, (nf, nm)))[0]
    Xs = cplx_normal(rng, (nf, nf))
    S0 = 2.0 * np.eye(nf) + 0.004 * (Xs + Xs.conj().T)
    Xn = cplx_normal(rng, (nf, nf))
    Ncov = 0.25 * np.eye(nf) + 0.003 * (Xn + Xn.conj().T)
    vis = cplx_normal(rng, (nt, nf)) @ np.linalg.cholesky(S0).T \
        + cplx_normal(rng, (nt, nf)) @ np.linalg.cholesky(Ncov).T + (4 * cplx_normal(rng, (nt, nm))) @ F.T
    run_chain(hp, "D_dense", vis, np.ones(nf, dtype=bool), S0, F, np.linalg.inv(Ncov),
              np.zeros((2, nf)), 2, 5)

    # ---- case E: map_estimate (no random terms, Niter forced to 1), case-B inputs.
    rng = np.random.default_rng(7)
    nt, nf, nm = 10, 32, 4
    F = np.array([scipy.special.legendre(i)(np.linspace(-1.0, 1.0, nf)) for i in range(nm)]).T
    vis = 3.0 * cplx_normal(rng, (nt, nf)) + (20 * cplx_normal(rng, (nt, nm))) @ F.T
    np.random.seed(1234)  # map_estimate skips np.random.seed(seed) (pspec.py:572-577)
    u_map = np.random.RandomState(1234).uniform(size=nf)
    run_chain(hp, "E_map", vis, np.ones(nf, dtype=bool), np.eye(nf), F, np.eye(nf) / 100.0,
              np.zeros((2, nf)), 1, None, map_estimate=True)
    d = dict(np.load(HERE / "chain_E_map.npz", allow_pickle=True))
    d["u_used"] = u_map[None, :]
    d["seed"] = -1
    np.savez_compressed(HERE / "chain_E_map.npz", **d)

    # ---- function-level vectors
    out = {}
    for n in (4, 5, 16):
        out[f"fourier_operator_{n}"] = hp.utils.fourier_operator(n)
    ps = np.random.default_rng(3).random(12) + 0.1
    out["cov_ps"] = ps
    out["cov_out"] = hp.pspec.covariance_from_pspec(ps, hp.utils.fourier_operator(12))
    sig = cplx_normal(np.random.default_rng(4), (9, 16))
    out["sprior_in"] = sig
    out["sprior_out"] = hp.pspec.sprior(sig, 2, 10.0)
    # inversion_sample_invgamma: reference draws u internally from np.random.uniform()
    cases, us, res = [], [], []
    for k, (alpha, beta, lo, hi) in enumerate([(24.0, 12.0, 0.1, 2.0), (204.0, 150.0, 0.1, 2.0),
                                                (1024.0, 900.0, 0.5, 1.5), (8.0, 3.0, 1e-3, 1e3),
                                                (204.0, 30.0, 0.1, 2.0), (50.0, 400.0, 0.1, 2.0)]):
        for rep in range(4):
            seed = 1000 + 10 * k + rep
            np.random.seed(seed)
            r = hp.pspec.inversion_sample_invgamma(alpha, beta, lo, hi)
            cases.append((alpha, beta, lo, hi))
            us.append(np.random.RandomState(seed).uniform())
            res.append(float(r))
    out["invsamp_cases"] = np.array(cases)
    out["invsamp_u"] = np.array(us)
    out["invsamp_out"] = np.array(res)
    # sample_S with and without prior
    s = cplx_normal(np.random.default_rng(5), (13, 20)) * np.linspace(0.5, 2, 20)
    prior = np.zeros((2, 20))
    prior[0, 8:12] = 60.0
    prior[1, 8:12] = 1.0
    np.random.seed(77)
    out["sampleS_s"] = s
    out["sampleS_prior"] = prior
    out["sampleS_out"] = hp.pspec.sample_S(s=s, prior=prior)
    out["sampleS_u"] = np.random.RandomState(77).uniform(size=20)
    np.random.seed(78)
    out["sampleS_out_noprior"] = hp.pspec.sample_S(s=s)
    out["sampleS_u_noprior"] = np.random.RandomState(78).uniform(size=20)
    # gcr draws for three time indices
    for idx in (0, 5):
        np.random.seed(912983 + idx)
        omi, omj = np.random.randn(6, 1), np.random.randn(6, 1)
        omk, oml = np.random.randn(6, 1), np.random.randn(6, 1)
        out[f"gcrdraw_a_{idx}"] = ((omi + 1j * omj) / 2 ** 0.5)[:, 0]
        out[f"gcrdraw_b_{idx}"] = ((omk + 1j * oml) / 2 ** 0.5)[:, 0]
    np.savez_compressed(HERE / "functions.npz", **out)
    print("functions.npz written")


if __name__ == "__main__":
    main()
