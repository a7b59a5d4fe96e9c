"""CPU check of the algebra behind the low-rank form of the per-time solve (hydra_pspec_b200/csrc/hp_ptlow.cu).

No GPU: the identities the kernels rely on are verified with numpy against the direct per-time systems that the oracle
(and the reference's gcr_fgmodes_1d, pspec.py:151-235, called with one time's operators) solves:
  * x_t = x0_t + R_f K_t^-1 (R^H b_t)_f   reproduces  M_t^-1 b_t;
  * M_0^-1 + R_f K_t^-1 R_f^H = M_t^-1    (the covariance the device draws must have: xi through M_0, zeta through K_t);
  * K_t = I - P_ff is positive definite with cond(K_t) <= cond(M_t).
"""
import numpy as np
import pytest

from oracle import hydra_oracle as ho


def crandn(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)


def systems(nt, nf, nm, frac, seed, general_basis=False):
    rng = np.random.default_rng(seed)
    F = np.linalg.qr(crandn(rng, nf, nm))[0] if nm else np.zeros((nf, 0), dtype=complex)
    if general_basis:
        Q = np.linalg.qr(crandn(rng, nf, nf))[0]
    else:
        Q = (ho.fourier_operator(nf) / np.sqrt(nf)).conj().T
    lam = np.sqrt(0.2 + rng.random(nf))
    ninv = 1.0 / (0.3 + rng.random(nf)) ** 2
    flags = rng.random((nt, nf)) > frac
    flags[0] = True
    flags[:, 3] = False                      # one channel flagged at all times
    B = np.hstack([Q, F])
    D = np.concatenate([lam, np.ones(nm)])
    J = np.diag(np.concatenate([np.ones(nf), np.zeros(nm)]))
    return rng, B, D, J, ninv, flags


@pytest.mark.parametrize("nt,nf,nm,frac,general", [(6, 32, 4, 0.1, False), (5, 64, 0, 0.3, False), (7, 96, 9, 0.15, True)])
def test_low_rank_form_equals_the_per_time_systems(nt, nf, nm, frac, general):
    rng, B, D, J, ninv, flags = systems(nt, nf, nm, frac, 11 + nf, general)
    N = nf + nm
    wbar = flags.any(axis=0)
    A = (D[:, None] * B.conj().T) * np.sqrt(ninv * wbar)[None, :]          # N x n
    M0 = J + A @ A.conj().T
    M0inv = np.linalg.inv(M0)
    R = M0inv @ A
    P = A.conj().T @ R
    for t in range(nt):
        wt = flags[t]
        Mt = J + (D[:, None] * B.conj().T * (wt * ninv)[None, :]) @ (B * D[None, :])
        b = crandn(rng, N)
        f = np.flatnonzero(wbar & ~wt)
        K = np.eye(f.size) - P[np.ix_(f, f)]
        x0 = M0inv @ b
        x = x0 + R[:, f] @ np.linalg.solve(K, (R.conj().T @ b)[f]) if f.size else x0
        want = np.linalg.solve(Mt, b)
        assert np.max(np.abs(x - want)) / np.max(np.abs(want)) < 1e-11
        cov = M0inv + (R[:, f] @ np.linalg.solve(K, R[:, f].conj().T) if f.size else 0.0)
        Mtinv = np.linalg.inv(Mt)
        assert np.max(np.abs(cov - Mtinv)) / np.max(np.abs(Mtinv)) < 1e-10
        if f.size:
            ev = np.linalg.eigvalsh(0.5 * (K + K.conj().T))
            assert ev.min() > 0 and ev.max() <= 1.0 + 1e-12
            assert ev.max() / ev.min() <= np.linalg.cond(Mt) * (1 + 1e-9)


def test_device_draw_construction_has_the_per_time_covariance():
    """x_t = W^H (W r + xi) + R_f L_K^-H (L_K^-1 (R^H r)_f + zeta) with independent unit complex Gaussians xi (N), zeta (k): the
    linear map applied to (xi, zeta) has covariance M_t^-1 exactly, and the mean is M_t^-1 r."""
    rng, B, D, J, ninv, flags = systems(4, 48, 5, 0.2, 3)
    N = 53
    wbar = flags.any(axis=0)
    A = (D[:, None] * B.conj().T) * np.sqrt(ninv * wbar)[None, :]
    M0 = J + A @ A.conj().T
    L0 = np.linalg.cholesky(M0)
    W = np.linalg.inv(L0)
    R = W.conj().T @ W @ A
    P = A.conj().T @ R
    t = 2
    f = np.flatnonzero(wbar & ~flags[t])
    assert f.size > 0
    K = np.eye(f.size) - P[np.ix_(f, f)]
    LK = np.linalg.cholesky(K)
    T_xi = W.conj().T                                        # x depends on xi through W^H
    T_zeta = R[:, f] @ np.linalg.inv(LK).conj().T            # and on zeta through R_f L_K^-H
    cov = T_xi @ T_xi.conj().T + T_zeta @ T_zeta.conj().T
    Mt = J + (D[:, None] * B.conj().T * (flags[t] * ninv)[None, :]) @ (B * D[None, :])
    Mtinv = np.linalg.inv(Mt)
    assert np.max(np.abs(cov - Mtinv)) / np.max(np.abs(Mtinv)) < 1e-11
    r = crandn(rng, N)
    mean = W.conj().T @ (W @ r) + T_zeta @ np.linalg.solve(LK, (R.conj().T @ r)[f])
    assert np.max(np.abs(mean - Mtinv @ r)) / np.max(np.abs(Mtinv @ r)) < 1e-11
