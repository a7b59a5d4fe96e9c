"""DPSS (Slepian) foreground modes (reference: hydra_pspec/dpss.py).

The Gibbs hot path takes its foreground basis ``fgmodes`` (Nfreqs, Nmodes) as an input; the reference
ships a DPSS generator + fitter for it (``dpss_fit_modes``, dpss.py:6-95).  Both live on the host: the
basis is computed once per baseline before the chain is loaded (it then stays resident on the device as
part of ``[Q | F]`` and of the Gram matrix), and the fit is a 2 Nmodes-parameter weighted least-squares
problem.  The reference minimises its quadratic log-likelihood numerically (L-BFGS-B from zero,
dpss.py:82-93); the minimiser is the normal-equation solution, which is what is returned here.
"""
import numpy as np
from scipy.signal.windows import dpss


def dpss_modes(nfreqs, nmodes=10, alpha=1.0):
    """The DPSS basis of dpss.py:70-73, shape (nmodes, nfreqs) like the reference's ``dpss_modes``;
    pass ``dpss_modes(...).T`` as ``fgmodes``."""
    return dpss(nfreqs, NW=alpha, Kmax=nmodes, sym=False)


def dpss_fit_modes(d, w, freqs, cov, nmodes=10, alpha=1.0, minimize_method=None, taper=None):
    """Weighted DPSS fit to masked complex 1-D data (dpss.py:6-95).

    Returns ``(dpss_modes, amps)`` with ``amps`` the interleaved (re, im) coefficients, like the
    reference.  ``minimize_method`` is accepted for signature compatibility: the quadratic
    ``x^H C^-1 x``, ``x = taper * w * (d - A . modes)``, is minimised exactly.
    """
    d = np.asarray(d)
    w = np.asarray(w, dtype=float)
    freqs = np.asarray(freqs)
    cov = np.asarray(cov)
    assert d.size == cov.shape[0] == cov.shape[1] == freqs.size == w.size, \
        "Data, flags, covariance, and freqs arrays must have same number of channels"
    if taper is None:
        taper = 1.0
    else:
        taper = np.asarray(taper)
        assert taper.size == freqs.size, "'taper' must be evaluated at locations given in 'freqs'"
    modes = dpss_modes(freqs.size, nmodes=nmodes, alpha=alpha)
    invcov = np.linalg.inv(cov)
    tw = taper * w
    B = (modes * tw).T                      # x = tw d - B a,  a complex (nmodes,)
    BhC = B.conj().T @ invcov
    a = np.linalg.lstsq(BhC @ B, BhC @ (tw * d), rcond=None)[0]
    amps = np.empty(2 * nmodes)
    amps[0::2], amps[1::2] = a.real, a.imag
    return modes, amps
