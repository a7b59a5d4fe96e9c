"""The C-ABI library loads and exports every symbol include/hydra_pspec_b200.h declares; host-side
logic of the drop-in module (validation, draw streams, covariance analysis).  CPU only -- no
compute entry point is called."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    txt = (ROOT / "include" / "hydra_pspec_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hp_[a-z0-9_A-Z]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from hydra_pspec_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"
    for s in _lib.EXPORTED_SYMBOLS:
        assert s in syms, f"{s} bound by _lib.py but not declared in the header"
    assert L.hp_version().startswith(b"hydra_pspec_b200")


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hydra_pspec_b200 import _lib, pspec, utils
    with pytest.raises(_lib.HydraLibError, match="no CPU fallback"):
        utils.fourier_operator(8)
    with pytest.raises(_lib.HydraLibError, match="no CPU fallback"):
        pspec.gibbs_sample_with_fg(np.ones((4, 8), complex), np.ones(8, bool), np.eye(8), np.ones((8, 2)), np.eye(8),
                                   np.zeros((2, 8)), Niter=1, verbose=False)


def test_product_does_not_import_oracle():
    for f in (ROOT / "hydra_pspec_b200").rglob("*.py"):
        assert "oracle" not in f.read_text(), f
    for f in (ROOT / "hydra_pspec_b200" / "csrc").glob("*.cu*"):
        assert "oracle/" not in f.read_text().replace("oracle/hydra_oracle.py:cg_theta for the derivation", ""), f


def test_reference_draw_streams(golden_dir):
    from hydra_pspec_b200 import pspec
    fn = np.load(golden_dir / "functions.npz")
    oma, omb = pspec._reference_gcr_draws(6, 6)
    for idx in (0, 5):
        assert np.array_equal(oma[idx], fn[f"gcrdraw_a_{idx}"])
        assert np.array_equal(omb[idx], fn[f"gcrdraw_b_{idx}"])


def test_s_draw_transformation():
    import scipy.stats
    from hydra_pspec_b200 import pspec
    u = np.random.default_rng(0).uniform(size=(3, 10))
    prior = np.zeros((2, 10))
    prior[:, 4] = [2.0, 0.1]
    d = pspec._s_draws_from_uniforms(u, prior, 25)
    assert np.array_equal(d[:, 4], u[:, 4])
    want = scipy.stats.invgamma.ppf(u, a=24.0)  # what invgamma.rvs(a) evaluates for the same uniform
    mask = np.ones(10, bool)
    mask[4] = False
    assert np.allclose(d[:, mask], want[:, mask], rtol=1e-13)


def test_prior_validation_errors():
    from hydra_pspec_b200 import pspec
    for bad, msg in [([[1.0], [0.0]], "prior_min must be greater than zero"),
                     ([[np.inf], [1.0]], "prior_max must be finite"),
                     ([[1.0], [2.0]], "prior_max must be greater than prior_min")]:
        with pytest.raises(ValueError, match=msg):
            pspec._check_prior(np.array(bad), 1)
    assert pspec._check_prior(None, 5).shape == (2, 5)


def test_signal_cov_analysis():
    from hydra_pspec_b200 import pspec
    n = 12
    idx = np.arange(n) - n // 2
    fop = np.exp(-2j * np.pi * np.outer(idx, idx) / n)
    p = np.random.default_rng(1).random(n) + 0.2
    S = fop.conj().T @ np.diag(p / n ** 2) @ fop
    b0, lam = pspec._analyse_signal_cov(S)
    assert b0 is None and np.allclose(lam, p / n)      # Lambda = n q = ps / n
    b0, lam = pspec._analyse_signal_cov(np.eye(n))
    assert b0 is None and np.allclose(lam, 1.0)
    X = np.random.default_rng(2).standard_normal((n, n))
    # a general covariance goes to the batched device eigensolver (tests/test_gpu_kernels.py); here only the host logic,
    # with numpy standing in for it
    host_eigh = lambda m: np.linalg.eigh(m)
    b0, lam = pspec._analyse_signal_cov(X @ X.T + np.eye(n), eigh=host_eigh)
    assert b0 is not None and np.allclose((b0 * lam) @ b0.conj().T, X @ X.T + np.eye(n))
    # almost delay-diagonal is not delay-diagonal
    S2 = S.copy()
    S2[1, 5] += 1e-9 * np.abs(S).max()
    S2[5, 1] += 1e-9 * np.abs(S).max()
    assert pspec._analyse_signal_cov(S2, eigh=host_eigh)[0] is not None


def test_host_helpers_match_reference(golden_dir):
    from hydra_pspec_b200 import pspec
    fn = np.load(golden_dir / "functions.npz")
    idx = np.arange(12) - 6
    fop = np.exp(-2j * np.pi * np.outer(idx, idx) / 12)
    assert np.allclose(pspec.covariance_from_pspec(fn["cov_ps"], fop), fn["cov_out"], rtol=1e-13, atol=1e-15)
    assert np.allclose(pspec.sprior(fn["sprior_in"], 2, 10.0), fn["sprior_out"], rtol=1e-13)
    for k, ((alpha, beta, lo, hi), want) in enumerate(zip(fn["invsamp_cases"], fn["invsamp_out"])):
        np.random.seed(1000 + 10 * (k // 4) + k % 4)
        assert abs(pspec.inversion_sample_invgamma(alpha, beta, lo, hi) - want) < 1e-13 * want


def test_noise_model_split():
    from hydra_pspec_b200 import pspec
    d, dense, nih = pspec._noise_model(np.diag([1.0, 2.0, 3.0]), np.ones(3, bool), 3, True)
    assert dense is None and nih is None and np.array_equal(d, [1.0, 2.0, 3.0])
    X = np.random.default_rng(0).standard_normal((4, 9))
    N = X @ X.T
    fl = np.array([True, False, True, True])
    d, dense, nih = pspec._noise_model(N, fl, 4, True)
    Ni = fl[:, None] * N * fl[None, :]
    assert np.allclose(nih @ nih, Ni) and np.allclose(nih, nih.conj().T) and np.allclose(d, np.diag(N))
    assert pspec._noise_model(N, fl, 4, False)[2] is None
    with pytest.raises(NotImplementedError):
        pspec._noise_model(np.zeros((5, 4, 4)), fl, 4, False)


def test_numa_binding_is_best_effort():
    """No GPU / no topology here: the helper must return None and leave the affinity alone."""
    import os
    from hydra_pspec_b200 import _lib
    before = os.sched_getaffinity(0)
    assert _lib.bind_to_device_numa(0) is None
    assert os.sched_getaffinity(0) == before


def test_header_compiles_as_c_and_links(tmp_path):
    """include/hydra_pspec_b200.h is a plain C header: a C translation unit that takes the address of every declared
    entry point must compile with gcc and link against the shared library (no compute call is made)."""
    import re
    import shutil
    import subprocess
    from pathlib import Path
    from hydra_pspec_b200 import _lib
    if shutil.which("gcc") is None or not _lib.LIB_PATH.exists():
        import pytest
        pytest.skip("gcc or the built library is missing")
    root = Path(__file__).resolve().parent.parent
    header = (root / "include" / "hydra_pspec_b200.h").read_text()
    names = sorted(set(re.findall(r"\b(hp_[a-z_0-9]+)\s*\(", header)))
    assert "hp_engine_create" in names and "hp_release_cached_memory" in names and len(names) >= 20
    src = tmp_path / "abi.c"
    src.write_text('#include "hydra_pspec_b200.h"\n#include <stdio.h>\nint main(void) {\n  void* p[] = {' +
                   ", ".join(f"(void*){n}" for n in names) +
                   '};\n  printf("%d %s\\n", (int)(sizeof p / sizeof p[0]), hp_version());\n  return 0;\n}\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(root / "include"), str(src), "-o", str(exe),
                    "-L", str(_lib.LIB_PATH.parent), "-lhydra_pspec_b200", f"-Wl,-rpath,{_lib.LIB_PATH.parent}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == len(names) and "hydra_pspec_b200" in " ".join(out[1:])


@pytest.mark.parametrize("nblk", [1, 2, 5, 13, 14])
def test_solve3_schedule_covers_every_strip_once(nblk):
    """k_solve3's static work schedule (hp_solve3.cu: longest-first over the warps): in both passes every 16-row strip of the
    system is owned by exactly one warp, and at the headline size no warp carries more than 1.25 x the mean load."""
    import ctypes as C
    from hydra_pspec_b200 import _lib
    L = _lib.lib()
    n = np.zeros(2 * 64, dtype=np.uint8)
    strips = np.zeros(2 * 64 * 16, dtype=np.uint8)
    nw, mp = C.c_int(), C.c_int()
    assert L.hp_test_solve3_schedule(nblk, _lib.ptr(n), _lib.ptr(strips), C.byref(nw), C.byref(mp)) == 0
    nw, mp = nw.value, mp.value
    n = n[:2 * nw].reshape(2, nw)
    strips = strips[:2 * nw * mp].reshape(2, nw, mp)
    for p in range(2):
        owned = [int(strips[p, w, e]) for w in range(nw) for e in range(n[p, w])]
        assert sorted(owned) == list(range(2 * nblk))
        length = (lambda s: 4 * (s + 1)) if p == 0 else (lambda s: 8 * nblk - 4 * s)
        loads = np.array([sum(length(int(strips[p, w, e])) for e in range(n[p, w])) for w in range(nw)], dtype=float)
        # the warps of every scheduler (warp id mod 4) together carry about a quarter of the pass
        sched = np.array([loads[s::4].sum() for s in range(4)])
        if nblk == 13:
            assert sched.max() <= 1.03 * sched.mean(), sched
            assert loads.max() <= 1.25 * loads.mean(), loads


def test_scatter_chunk_threads_match_a_plain_copy(monkeypatch):
    """pspec._scatter_chunk: iteration-major staging -> chain-major destinations, split over the copy threads."""
    from hydra_pspec_b200 import pspec
    rng = np.random.default_rng(0)
    nch, niter, nt, nf, nm, c, done = 37, 5, 64, 512, 3, 2, 3        # signal_cr chunk: 37 x 2 x 64 x 512 x 16 B = 38.8 MB (threaded)
    stage = {"signal_cr": rng.standard_normal((c, nch, nt, nf)) + 1j * rng.standard_normal((c, nch, nt, nf)),
             "fg_amps": rng.standard_normal((c, nch, nt, nm)) + 0j, "chisq": rng.standard_normal((c, nch, nt, nf)),
             "signal_ps": rng.standard_normal((nch, c, nf)), "ln_post": rng.standard_normal((nch, c))}
    dest = {"signal_cr": np.zeros((nch, niter, nt, nf), dtype=complex), "fg_amps": np.zeros((nch, niter, nt, nm), dtype=complex),
            "chisq": np.zeros((nch, niter, nt, nf)), "signal_ps": np.zeros((nch, niter, nf)), "ln_post": None}
    monkeypatch.setattr(pspec, "_COPY_THREADS", 4)
    pspec._scatter_chunk(dest, stage, done, c)
    for k in ("signal_cr", "fg_amps", "chisq"):
        assert np.array_equal(dest[k][:, done:done + c], np.swapaxes(stage[k], 0, 1))
        assert not dest[k][:, :done].any()
    assert np.array_equal(dest["signal_ps"][:, done:done + c], stage["signal_ps"])
