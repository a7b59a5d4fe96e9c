#!/usr/bin/env python
"""Golden output of BASELINE.json configs[0]: the UNMODIFIED reference sampler (`/root/reference`) on
its own `test_data` (vis-eor-fgs.uvh5 + 0-1/*.npy, config.yaml parameters), 3 iterations.

pyuvdata is not installed here, so the visibilities are read with this repo's uvh5 reader and
assembled by this repo's driver functions (the part of run-hydra-pspec.py that precedes the hot
path); everything from `gibbs_sample_with_fg` down is the reference's code.  Build container only.
Output: tests/golden/chain_T_testdata_driver.npz.  The inputs under tests/golden/testdata/ are the reference's
test_data files (data, not source), stored compressed; tests/golden/testdata_fixture.py unpacks them.
"""
import sys
import warnings
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))

from make_golden import load_reference  # noqa: E402
import run_hydra_pspec_b200 as drv  # noqa: E402

NITER = 3


def driver_argv(td, out_dir):
    return ["--ant_str", "0_1", "--seed", "7123689", "--Niter", str(NITER), "--ps_prior_lo", "0.1", "--ps_prior_hi", "2",
            "--n_ps_prior_bins", "3", "--dirname", "results", "--out_dir", str(out_dir), "--clobber",
            "--sigcov0", str(td), "--sigcov0_file", "eor-cov.npy", "--Nfgmodes", "12", "--fgmodes", str(td),
            "--fgmodes_file", "fgmodes.npy", "--noise", str(td), "--noise_file", "noise.npy", "--noise_cov", str(td),
            "--noise_cov_file", "noise-cov.npy", str(td / "vis-eor-fgs.uvh5")]


def main():
    import tempfile
    from testdata_fixture import materialize
    td = materialize(tempfile.mkdtemp())
    _, args = drv.parse_args(driver_argv(td, "/tmp/unused"))
    antpairs, freqs, get = drv.read_visibilities([Path(p) for p in args.file_paths], args.ant_str, args.freq_range)
    bls = drv.assemble_baselines(args, antpairs, freqs, get, Path("/tmp/unused"))
    assert len(bls) == 1 and bls[0]["antpair"] == (0, 1)
    b = bls[0]
    w = drv.time_invariant_flags(~np.asarray(b["w"], dtype=bool))
    pr = drv.ps_prior_for(args, b["d"].shape[1])
    hp = load_reference()  # after the driver part: it installs empty pyuvdata / astropy stub modules
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cr, S, ps, fg, chisq, lnp, _ = hp.pspec.gibbs_sample_with_fg(
            b["d"].copy(), w.copy(), b["S_initial"].copy(), b["fgmodes"].copy(), b["Ninv"].copy(), pr.copy(),
            Niter=NITER, seed=args.seed, verbose=False, nproc=1)
    np.savez_compressed(HERE / "chain_T_testdata_driver.npz", signal_ps=ps, ln_post=lnp, fg_amps=fg,
                        signal_cr_last=cr[-1], chisq_last=chisq[-1], signal_S=S, vis_checksum=np.sum(b["d"]),
                        vis_row0=b["d"][0])
    print("ps[:, 58:62] =", ps[:, 58:62], "ln_post =", lnp)


if __name__ == "__main__":
    main()
