"""The ``hydra_pspec.utils`` helpers used by the Gibbs hot path (reference: hydra_pspec/utils.py).

The uvh5 / pyuvdata helpers of the reference's utils module are outside the hot path and are
not duplicated here.
"""
from pathlib import Path

import numpy as np


def fourier_operator(n):
    """Shifted DFT matrix ``F[k, x] = exp(-2 pi i (k - n//2)(x - n//2) / n)``.

    Same operator as hydra_pspec/utils.py:14-40 (equivalent to ifftshift -> fft -> fftshift);
    computed on the device by ``hp_fourier_operator``.
    """
    from . import _lib
    out = np.empty((n, n), dtype=np.complex128)
    _lib.check(_lib.lib().hp_fourier_operator(0, int(n), _lib.ptr(out)))
    return out


def write_numpy_files(fp, signal_cr, signal_S, signal_ps, fg_amps, chisq, ln_post):
    """Write the sample arrays with the reference's file names (hydra_pspec/utils.py:269-312)."""
    fp = Path(fp)
    np.save(fp / "gcr-eor.npy", signal_cr)
    np.save(fp / "cov-eor.npy", signal_S)
    np.save(fp / "dps-eor.npy", signal_ps)
    np.save(fp / "fg-amps.npy", fg_amps)
    np.save(fp / "chisq.npy", chisq)
    np.save(fp / "ln-post.npy", ln_post)
