"""Multi-GPU check of the baseline sharding + NCCL gather (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        profiles/scripts/multi_gpu_check.py

Every rank runs its share of 5 small baselines (numpy-stream draws, so chains are deterministic), the
sample arrays are gathered over NCCL, and rank 0 compares them bit for bit with a single-GPU run of all baselines
(Philox draws depend on (seed, global baseline index, iteration) only)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from hydra_pspec_b200 import driver, pspec  # noqa: E402
from bench import make_baseline  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nt, nf, nm, nbl, niter = 32, 96, 6, 5, 6
bls = []
for i in range(nbl):
    vis, flags, F, nd, l0 = make_baseline(500 + i, nt, nf, nm)
    bls.append(dict(vis=vis, flags=flags, fgmodes=F, ninv_diag=nd, lam0sq=l0))


ps, lp = driver.run_baselines(bls, Niter=niter, seed=7, rng="philox", device=local)
assert ps.shape == (nbl, niter, nf) and lp.shape == (nbl, niter), (ps.shape, lp.shape)
ok = True
if rank == 0:
    # single-GPU run of ALL baselines in one engine: same key, chain id = global baseline index -> bit-identical
    eng = pspec.GibbsEngine(nbl, nt, nf, nm, max_iters=niter, rng="philox", keep=(), seed=7, device=local)
    for c, b in enumerate(bls):
        eng.load_chain(c, b["vis"], b["flags"], b["fgmodes"], b["ninv_diag"], b["lam0sq"])
    eng.run(niter)
    ref_ps = np.stack([eng.signal_ps(c) for c in range(nbl)])
    eng.close()
    ok = bool(np.array_equal(ps, ref_ps)) and bool(np.all(np.isfinite(lp)))
    print(f"multi_gpu_check: world={world} gathered {ps.shape}, bit-identical to the single-GPU run: {ok}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
