"""BASELINE.json configs[0]: the reference's own test_data run (vis-eor-fgs.uvh5, config.yaml parameters, Niter=1000,
seed 7123689) through the driver, reference random streams (--rng numpy) and device draws (--rng philox)."""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import run_hydra_pspec_b200 as drv  # noqa: E402
from make_golden_testdata import driver_argv  # noqa: E402

from testdata_fixture import materialize  # noqa: E402
td = materialize(tempfile.mkdtemp())
for rng in ("numpy", "philox"):
    for rep in range(2):
        with tempfile.TemporaryDirectory() as out:
            argv = driver_argv(td, out)
            argv[argv.index("--Niter") + 1] = "1000"
            argv += ["--rng", rng]
            t0 = time.perf_counter()
            sys.stdout = open("/dev/null", "w")
            rc = drv.main(argv)
            sys.stdout = sys.__stdout__
            dt = time.perf_counter() - t0
            print(f"--rng {rng} rep {rep}: rc={rc}  1000 Gibbs iterations of the 203 x 120 x 12 test baseline incl. uvh5 read and "
                  f".npy output: {dt:.2f} s")
