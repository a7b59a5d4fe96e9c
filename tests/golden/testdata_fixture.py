"""The reference's `test_data` inputs (BASELINE.json configs[0]) as test fixtures.

`testdata/vis-eor-fgs.uvh5.gz` is the reference's test_data/vis-eor-fgs.uvh5 (data, not source), gzip-compressed;
`testdata/inputs_0-1.npz` holds its test_data/0-1/{eor-cov,fgmodes,noise-cov,noise}.npy.  `materialize(dst)`
recreates the reference's directory layout under `dst`:

    dst/vis-eor-fgs.uvh5
    dst/0-1/eor-cov.npy  fgmodes.npy  noise-cov.npy  noise.npy
"""
import gzip
import shutil
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent / "testdata"


def materialize(dst):
    dst = Path(dst)
    (dst / "0-1").mkdir(parents=True, exist_ok=True)
    if not (dst / "vis-eor-fgs.uvh5").exists():
        with gzip.open(HERE / "vis-eor-fgs.uvh5.gz", "rb") as f, open(dst / "vis-eor-fgs.uvh5", "wb") as g:
            shutil.copyfileobj(f, g)
    z = np.load(HERE / "inputs_0-1.npz")
    for k in z.files:
        if not (dst / "0-1" / f"{k}.npy").exists():
            np.save(dst / "0-1" / f"{k}.npy", z[k])
    return dst
