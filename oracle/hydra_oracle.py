"""CPU oracle (numpy/scipy restatement) of hydra-pspec's per-baseline Gibbs hot path.

TEST INFRASTRUCTURE ONLY.  The product (``hydra_pspec_b200``) never imports this
module; it is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (I/O-only dependencies stubbed) and stores its
outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function here against those vectors.

Each function cites the reference lines it restates (paths relative to the
reference checkout, ``hydra_pspec/pspec.py`` unless noted).  The random draws are
*arguments* here (``omega_a``, ``omega_b``, ``u``): the reference draws them from
numpy's global MT19937 stream, and :func:`reference_gcr_draws` /
:class:`ReferenceSDraws` reproduce that stream so that a chain can be replayed
draw-for-draw.

Two solver modes are provided for the GCR step:

``solver="cg"``      what the reference does (pspec.py:228): preconditioned CG with
                     ``M = pinv(A)``, ``rtol=1e-8, atol=1e-6``.  Because M A = I the
                     iterates are scalar multiples of ``A^-1 b`` and CG stops a few
                     1e-9 (relative) short of it.
``solver="direct"``  ``A^-1 b`` exactly (LU), the quantity CG approximates.

and an extension the reference does not have (it asserts ``flags.shape ==
(Nfreqs,)``, pspec.py:428): ``flags`` of shape ``(Ntimes, Nfreqs)`` applies the
same algebra with a separate flag vector per time.
"""
import numpy as np
import scipy.linalg
import scipy.sparse.linalg
import scipy.special
from scipy.interpolate import interp1d

GCR_SEED_BASE = 912983  # pspec.py:153 (multiprocess_seed default)


# ----------------------------------------------------------------------------
# utils.py:14-40
def fourier_operator(n):
    """Shifted DFT matrix, F[k, x] = exp(-2 pi i (k - n//2)(x - n//2) / n)."""
    idx = np.arange(n) - n // 2
    return np.exp(-2j * np.pi * np.outer(idx, idx) / n)


# pspec.py:313-322
def covariance_from_pspec(ps, fourier_op):
    """C = F^H diag(ps) F."""
    return fourier_op.conj().T @ (np.asarray(ps, dtype=complex)[:, None] * fourier_op)


# pspec.py:130-148
def sprior(signals, bins, factor):
    nobs, nfreq = signals.shape
    sk = np.fft.fft(signals, axis=-1)
    ds = np.sum(np.abs(sk) ** 2, axis=0)
    prior = np.zeros((2, nfreq))
    prior[0] = ds * factor
    prior[1] = ds / factor
    prior[:, bins + 1:-bins] = 0
    return prior / (nobs / 2 - 1)


# ----------------------------------------------------------------------------
# Random draws, exactly as the reference consumes numpy's global stream.
def reference_gcr_draws(ntimes, nfreqs, seed_base=GCR_SEED_BASE):
    """(omega_a, omega_b), each (Ntimes, Nfreqs) complex.

    pspec.py:195-217 -- every call of gcr_fgmodes_1d reseeds with
    ``multiprocess_seed + idx`` and draws four (Nfreqs, 1) normal vectors, so the
    realisation for time ``idx`` is the same in every Gibbs iteration.
    """
    oma = np.empty((ntimes, nfreqs), dtype=complex)
    omb = np.empty((ntimes, nfreqs), dtype=complex)
    for idx in range(ntimes):
        rs = np.random.RandomState(seed_base + idx)
        omi, omj = rs.randn(nfreqs), rs.randn(nfreqs)
        omk, oml = rs.randn(nfreqs), rs.randn(nfreqs)
        oma[idx] = (omi + 1j * omj) / 2 ** 0.5
        omb[idx] = (omk + 1j * oml) / 2 ** 0.5
    return oma, omb


class ReferenceSDraws:
    """The parent-process uniform stream used by sample_S.

    pspec.py:577 seeds numpy's global generator once per chain; sample_S then takes
    exactly one uniform per delay bin per iteration, in bin order: explicitly
    (pspec.py:58) for prior-bounded bins, and inside ``invgamma.rvs`` for the
    others (scipy's invgamma has no ``_rvs``; the generic ``rvs`` evaluates
    ``ppf(uniform())``).  The GCR draws happen in a worker process
    (multiprocess.Pool, pspec.py:287) and do not advance this stream.
    """

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def next(self, nfreqs):
        return self.rs.uniform(size=nfreqs)


# ----------------------------------------------------------------------------
# pspec.py:11-64
def inversion_sample_invgamma(alpha, beta, prior_min, prior_max, u, ngrid=1000):
    if prior_min <= 0:
        raise ValueError("prior_min must be greater than zero")
    if prior_max <= 0:
        raise ValueError("prior_max must be greater than zero")
    if not np.isfinite(prior_max):
        raise ValueError("prior_max must be finite")
    if prior_max <= prior_min:
        raise ValueError("prior_max must be greater than prior_min")
    x = np.logspace(np.log10(prior_min), np.log10(prior_max), ngrid)
    # invgamma.cdf(x, a, scale=beta) = Q(a, beta / x)
    cdf = scipy.special.gammaincc(alpha, beta / x)
    cdf = cdf - cdf.min()
    cdf = cdf / cdf.max()
    cdf_unique, idx = np.unique(cdf, return_index=True)
    return float(interp1d(cdf_unique, x[idx], kind="linear")(u))


# pspec.py:67-127
def sample_S(s=None, sk=None, prior=None, u=None):
    """One draw of the delay power spectrum given the signal realisations.

    ``u``: (Nfreqs,) uniforms, one per bin (see ReferenceSDraws).
    Returns (ps_sample, beta).
    """
    if s is None and sk is None:
        raise ValueError("Must pass in s (real space) or sk (Fourier space) vector.")
    if sk is None:
        sk = np.fft.fftshift(np.fft.fft(np.fft.ifftshift(s, axes=1), axis=1), axes=1)
    nobs, nfreqs = sk.shape
    if prior is None:
        prior = np.zeros((2, nfreqs))
    beta = np.sum(np.abs(sk) ** 2, axis=0)
    alpha = nobs - 1.0
    x = np.zeros(nfreqs)
    for i in range(nfreqs):
        if np.any(prior[:, i] > 0):
            x[i] = inversion_sample_invgamma(alpha + 1, beta[i], prior[1, i], prior[0, i], u[i])
        else:
            # invgamma.rvs(a) == invgamma.ppf(U, a) == 1 / gammainccinv(a, U)
            x[i] = beta[i] / scipy.special.gammainccinv(alpha, u[i])
    return x, beta


# ----------------------------------------------------------------------------
# pspec.py:325-374
def build_matrices(flags, signal_S, Ninv, fgmodes, need_pinv=True, symmetric_flags=False):
    """Operators of the GCR system for one flag vector.

    Returns dict(Sh, S, Ni, Nih, A, Ai).  ``Ni[i, j] = Ninv[i, j] * flags[j]``
    (pspec.py:361 multiplies by the 1-D flag vector twice along the last axis).
    """
    nfreqs = signal_S.shape[0]
    nmodes = fgmodes.shape[1]
    fl = np.asarray(flags).astype(float)
    Sh = scipy.linalg.sqrtm(signal_S).astype(complex)
    S = np.array(signal_S, dtype=complex)
    Ni = (fl * Ninv * fl).astype(complex)
    if symmetric_flags:
        # flags on rows and columns (what the CUDA path does for a non-diagonal Ninv; identical to the
        # reference's column-only masking when Ninv is diagonal or nothing is flagged)
        Ni = (fl[:, None] * Ninv * fl[None, :]).astype(complex)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Nih = scipy.linalg.sqrtm(Ni).astype(complex)
    if symmetric_flags:
        # Schur-based sqrtm loses half the digits on a singular matrix (flagged rows/columns are zero):
        # take the Hermitian PSD square root from the eigendecomposition instead.
        ev, V = np.linalg.eigh(0.5 * (Ni + Ni.conj().T))
        Nih = (V * np.sqrt(np.clip(ev, 0.0, None))) @ V.conj().T
    if not np.all(np.isfinite(Nih)) and np.count_nonzero(Ni - np.diag(np.diagonal(Ni))) == 0:
        # scipy >= 1.13 returns NaN from sqrtm for a diagonal matrix with two zero eigenvalues in the
        # same block (two flagged channels close together), which turns the reference's whole chain
        # into NaN.  The principal square root of a diagonal PSD matrix is the element-wise one.
        Nih = np.diag(np.sqrt(np.diagonal(Ni))).astype(complex)
    A = np.zeros((nfreqs + nmodes,) * 2, dtype=complex)
    SNi = S @ Ni
    FhNi = fgmodes.conj().T @ Ni
    A[:nfreqs, :nfreqs] = np.eye(nfreqs) + SNi
    A[:nfreqs, nfreqs:] = SNi @ fgmodes
    A[nfreqs:, :nfreqs] = FhNi
    A[nfreqs:, nfreqs:] = FhNi @ fgmodes
    Ai = np.linalg.pinv(A) if need_pinv else None
    return dict(Sh=Sh, S=S, Ni=Ni, Nih=Nih, A=A, Ai=Ai)


# pspec.py:219-222
def gcr_rhs(d, w, mats, fgmodes, oma, omb):
    """RHS vector b for one time sample (1-D arrays in, 1-D out)."""
    wd = (w * d)
    noise_term = mats["Ni"] @ wd + mats["Nih"] @ omb
    top = mats["S"] @ noise_term + mats["Sh"] @ oma
    return np.concatenate([top, fgmodes.conj().T @ noise_term])


# pspec.py:224-235
def gcr_solve_1d(d, w, mats, fgmodes, oma, omb, solver="cg"):
    b = gcr_rhs(d, w, mats, fgmodes, oma, omb)
    if solver == "cg":
        x, info = scipy.sparse.linalg.cg(
            mats["A"], b.reshape(-1, 1), maxiter=int(1e5), rtol=1e-8, atol=1e-6, M=mats["Ai"]
        )
        return x
    if solver == "direct":
        return np.linalg.solve(mats["A"], b)
    if solver == "theta":
        # direct solve scaled by the scalar CG model: what the CUDA path's cg_compat mode computes
        x = np.linalg.solve(mats["A"], b)
        return cg_theta(np.vdot(b, x), np.linalg.norm(b)) * x
    raise ValueError(solver)


# pspec.py:238-310
def gcr_fgmodes(vis, w, mats, fgmodes, oma=None, omb=None, map_estimate=False, solver="cg"):
    """All times.  ``w``/``mats`` may be per-time lists for the 2-D flag extension."""
    ntimes, nfreqs = vis.shape
    if map_estimate:
        oma = np.zeros((ntimes, nfreqs), dtype=complex)
        omb = np.zeros((ntimes, nfreqs), dtype=complex)
    per_time = isinstance(mats, (list, tuple))
    out = np.zeros((ntimes, nfreqs + fgmodes.shape[1]), dtype=complex)
    if solver == "direct" and not per_time:
        # the same LAPACK LU solve as np.linalg.solve, factored once instead of once per time (all times share A)
        lu = scipy.linalg.lu_factor(mats["A"])
        for t in range(ntimes):
            out[t] = scipy.linalg.lu_solve(lu, gcr_rhs(vis[t], w, mats, fgmodes, oma[t], omb[t]))
        return out
    for t in range(ntimes):
        m = mats[t] if per_time else mats
        wt = w[t] if per_time else w
        out[t] = gcr_solve_1d(vis[t], wt, m, fgmodes, oma[t], omb[t], solver=solver)
    return out


# pspec.py:377-490
def gibbs_step_fgmodes(vis, flags, signal_S, fgmodes, Ninv, ps_prior, oma, omb, u,
                       map_estimate=False, solver="cg", symmetric_flags=False):
    """One Gibbs iteration.  ``vis`` is already multiplied by the flags (pspec.py:613).

    Returns signal_cr, S_sample, ps_sample, fg_amps, chisq, ln_post.
    """
    ntimes, nfreqs = vis.shape
    flags = np.asarray(flags, dtype=bool)
    per_time = flags.ndim == 2
    fop = fourier_operator(nfreqs)
    need_pinv = solver == "cg"
    if per_time:
        cache = {}
        mats = []
        for t in range(ntimes):
            key = flags[t].tobytes()
            if key not in cache:
                cache[key] = build_matrices(flags[t], signal_S, Ninv, fgmodes, need_pinv, symmetric_flags)
            mats.append(cache[key])
        w = [flags[t] for t in range(ntimes)]
    else:
        mats = build_matrices(flags, signal_S, Ninv, fgmodes, need_pinv, symmetric_flags)
        w = flags
    cr = gcr_fgmodes(vis, w, mats, fgmodes, oma, omb, map_estimate=map_estimate, solver=solver)
    signal_cr = cr[:, :nfreqs]
    fg_amps = cr[:, nfreqs:]
    model = signal_cr + fg_amps @ fgmodes.T
    resid = vis - model
    chisq = np.abs(resid) ** 2 * np.real(np.diagonal(Ninv))[None, :]
    ps_sample, _ = sample_S(s=signal_cr, prior=ps_prior, u=u)
    S_sample = covariance_from_pspec(ps_sample / nfreqs ** 2, fop)
    Sinv = np.linalg.inv(S_sample)
    # pspec.py:472-485: only the diagonal (same-time) terms of the quadratic forms.
    ln_post = 0.0
    for t in range(ntimes):
        f = flags[t] if per_time else flags
        r = resid[t, f]
        s = signal_cr[t, f]
        ln_post += -(r.conj() @ Ninv[f][:, f] @ r) - (s.conj() @ Sinv[f][:, f] @ s)
    return signal_cr, S_sample, ps_sample, fg_amps, chisq, float(np.real(ln_post))


# pspec.py:493-658
def gibbs_sample_with_fg(vis, flags, S_initial, fgmodes, Ninv, ps_prior, Niter=100, seed=None,
                         map_estimate=False, solver="cg", draws=None, symmetric_flags=False):
    """Chain for one baseline, reference draw sequence (or ``draws=(oma, omb, u[Niter, Nfreqs])``).

    Returns signal_cr[Niter], signal_S (last), signal_ps[Niter], fg_amps[Niter],
    chisq[Niter], ln_post[Niter].
    """
    if map_estimate:
        Niter = 1
    ntimes, nfreqs = vis.shape
    nmodes = fgmodes.shape[1]
    flags = np.asarray(flags, dtype=bool)
    if draws is None:
        oma, omb = reference_gcr_draws(ntimes, nfreqs)
        sd = ReferenceSDraws(seed)
        u_all = None
    else:
        oma, omb, u_all = draws
    signal_cr = np.zeros((Niter, ntimes, nfreqs), dtype=complex)
    signal_ps = np.zeros((Niter, nfreqs))
    fg_amps = np.zeros((Niter, ntimes, nmodes), dtype=complex)
    chisq = np.zeros((Niter, ntimes, nfreqs))
    ln_post = np.zeros(Niter)
    signal_S = np.array(S_initial, dtype=complex)
    visf = vis * flags
    for i in range(Niter):
        u = sd.next(nfreqs) if u_all is None else u_all[i]
        signal_cr[i], signal_S, signal_ps[i], fg_amps[i], chisq[i], ln_post[i] = gibbs_step_fgmodes(
            visf, flags, signal_S, fgmodes, Ninv, ps_prior, oma, omb, u,
            map_estimate=map_estimate, solver=solver, symmetric_flags=symmetric_flags)
    return signal_cr, signal_S, signal_ps, fg_amps, chisq, ln_post


# ----------------------------------------------------------------------------
# Scalar model of the reference's CG call (pspec.py:228 -> scipy.sparse.linalg.cg).
def cg_theta(c, bnorm, rtol=1e-8, atol=1e-6, maxiter=100000):
    """theta such that the reference's CG returns theta * A^-1 b.

    With M = pinv(A) = A^-1, z = M r is parallel to x* = A^-1 b whenever r is
    parallel to b, so x = xi x*, r = rho b, p = pi x* throughout and scipy's
    recursion reduces to scalars driven by c = vdot(b, x*):

        rho_cur = |rho|^2 c ;  alpha = rho_cur / (|pi|^2 conj(c)) ;
        xi += alpha pi ; rho -= alpha pi ;  stop when |rho| ||b|| < max(atol, rtol ||b||).

    For Hermitian A, c is real and theta = 1 after one step; the reference's A is
    not Hermitian (the S-block row is multiplied through by S) so CG ends 1e-9-ish
    away from x*.
    """
    if bnorm == 0:
        return 0.0 + 0.0j
    tol = max(float(atol), float(rtol) * float(bnorm))
    xi, rho, pi, rho_prev = 0.0 + 0.0j, 1.0 + 0.0j, 0.0 + 0.0j, None
    for it in range(maxiter):
        if abs(rho) * bnorm < tol:
            break
        rho_cur = (abs(rho) ** 2) * c
        pi = rho if it == 0 else rho + (rho_cur / rho_prev) * pi
        alpha = rho_cur / ((abs(pi) ** 2) * np.conj(c))
        xi += alpha * pi
        rho -= alpha * pi
        rho_prev = rho_cur
    return xi
