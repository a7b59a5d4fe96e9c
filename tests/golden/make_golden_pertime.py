#!/usr/bin/env python
"""Golden for the per-time-flags extension (BASELINE.json configs[2]): the UNMODIFIED reference's
`build_matrices` + `gcr_fgmodes_1d` (pspec.py:151-235, 325-374), called one time at a time with that
time's flag vector -- the per-time GCR the reference's own FIXMEs ask for (run-hydra-pspec.py:527,
pspec.py:449).  Build container only.  Output: tests/golden/gcr_pertime.npz."""
import warnings
from pathlib import Path

import numpy as np

from make_golden import load_reference, cplx_normal

HERE = Path(__file__).resolve().parent


def main():
    hp = load_reference()
    rng = np.random.default_rng(4242)
    nt, nf, nm = 8, 40, 5
    F = np.linalg.qr(cplx_normal(rng, (nf, nm)))[0]
    fop = hp.utils.fourier_operator(nf)
    p0 = 0.5 + rng.random(nf)
    S = fop.conj().T @ np.diag(p0 / nf ** 2) @ fop
    sig = 0.3 + rng.random(nf)
    Ninv = np.diag(1.0 / sig ** 2)
    vis = cplx_normal(rng, (nt, nf)) * sig + (5 * cplx_normal(rng, (nt, nm))) @ F.T \
        + cplx_normal(rng, (nt, nf)) @ np.linalg.cholesky(S + 1e-12 * np.eye(nf)).T
    flags = np.ones((nt, nf), dtype=bool)
    for t in range(1, nt):  # well separated flagged channels (scipy's sqrtm returns NaN for adjacent zero eigenvalues)
        flags[t, rng.choice(np.arange(0, nf, 4), size=3, replace=False) + (t % 3)] = False
    cr = np.zeros((nt, nf + nm), dtype=complex)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(nt):
            mats = hp.pspec.build_matrices(nf + nm, flags[t], S, Ninv, F)
            assert np.all(np.isfinite(mats[0][3]))
            x, _, info = hp.pspec.gcr_fgmodes_1d(t, (vis * flags)[t], flags[t], mats, F)
            assert info == 0
            cr[t] = x
    np.savez_compressed(HERE / "gcr_pertime.npz", vis=vis, flags=flags, S=S, fgmodes=F, Ninv=Ninv, cr=cr)
    print("gcr_pertime.npz", cr.shape, np.abs(cr).max(), flags.sum(axis=1))


if __name__ == "__main__":
    main()
