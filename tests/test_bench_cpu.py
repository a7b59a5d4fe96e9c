"""Host side of bench.py that needs no GPU: synthetic inputs, and the reference arm on the smallest configuration."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def test_synthetic_inputs_are_deterministic_and_shaped():
    a = bench.make_baseline(5, 16, 32, 4)
    b = bench.make_baseline(5, 16, 32, 4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    vis, flags, F, nd, l0 = a
    assert vis.shape == (16, 32) and flags.shape == (32,) and F.shape == (32, 4) and not flags.all()
    assert np.allclose(F.conj().T @ F, np.eye(4), atol=1e-12)
    assert bench.make_baseline(5, 16, 32, 4, flagged=False)[1].all()
    pt = bench.per_time_flags(5, flags, 16)
    assert pt.shape == (16, 32) and not np.any(pt[:, ~flags]) and len({r.tobytes() for r in pt}) > 1
    Ni = bench.dense_ninv(24)
    assert np.allclose(Ni, Ni.conj().T) and np.all(np.linalg.eigvalsh(Ni) > 0)


def test_every_config_names_its_metric_and_workload():
    for k in (0, 1, 2, 3, 4):
        assert "baseline-Gibbs-iterations/sec" in bench.metric_name(k)
        assert f"configs[{k}]" in bench.workload_name(k, 3)
    assert bench.metric_name(3) == bench.METRIC


def test_cpu_arm_runs_on_the_smallest_config():
    """One Gibbs iteration of configs[1] through the CPU arm's worker: the unmodified reference when baseline/_ref is
    installed (pip install --target), else the oracle port."""
    dt, kind = bench._cpu_worker((1, 3, 16, 1))
    assert dt > 0 and kind.split(" ")[0] in ("reference", "port")
    if (bench.ROOT / "baseline" / "_ref" / "hydra_pspec" / "pspec.py").is_file():
        assert kind.startswith("reference")
