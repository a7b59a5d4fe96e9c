"""Baseline sharding across GPUs (reference: the MPI layout of run-hydra-pspec.py:268-287, 482-557).

Baselines are independent Gibbs chains.  The reference scatters them over MPI ranks and each rank
loops over its share; here one process per GPU holds its whole share resident in one
:class:`~hydra_pspec_b200.pspec.GibbsEngine` and advances all of it together.  There is no
collective on the hot path; `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) is
used only for the final gather of the sample arrays.
"""
import numpy as np


def split_data_for_scatter(data, n_ranks):
    """Split a list into ``n_ranks`` contiguous sub-lists, the first ``len(data) % n_ranks`` one longer
    (run-hydra-pspec.py:268-287).  The reference aborts the MPI job when there are fewer baselines
    than ranks; here that is a ``ValueError``."""
    data_length = len(data)
    quot, rem = divmod(data_length, n_ranks)
    if quot == 0:
        raise ValueError(f"Number of baselines ({data_length}) should be >= number of ranks ({n_ranks})!")
    counts = [quot + 1 if n < rem else quot for n in range(n_ranks)]
    starts = [sum(counts[:n]) for n in range(n_ranks)]
    ends = [sum(counts[:n + 1]) for n in range(n_ranks)]
    return [data[starts[n]:ends[n]] for n in range(n_ranks)]


def shard_counts(n_items, n_ranks):
    return [len(x) for x in split_data_for_scatter(list(range(n_items)), n_ranks)]


def gather_samples(local, n_total, group=None):
    """All ranks call this with their ``[n_local, ...]`` float64 array (same trailing shape); every
    rank gets the ``[n_total, ...]`` array in global baseline order.  Uses all_gather on padded
    blocks (counts follow :func:`split_data_for_scatter`)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_total
        return np.asarray(local)
    world = dist.get_world_size(group)
    counts = shard_counts(n_total, world)
    rank = dist.get_rank(group)
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    pad = max(counts)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=torch.float64, device=dev)
    buf[:counts[rank]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


def run_baselines(baselines, Niter, seed=0, rng="philox", keep=(), device=None, gather=True, engine_factory=None):
    """Run ``Niter`` Gibbs iterations for this rank's share of ``baselines``.

    ``baselines``: the full list (every rank passes the same list, as after an MPI broadcast) of dicts
    with keys ``vis, flags, fgmodes, ninv_diag, lam0sq`` and optional ``ps_prior``; all of identical
    shape.  Returns ``(signal_ps, ln_post)`` with shapes ``[n, Niter, Nfreqs]`` / ``[n, Niter]``,
    ``n`` = all baselines when ``gather`` else this rank's share.
    """
    import torch.distributed as dist
    from . import pspec
    dist_on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if dist_on else 1
    rank = dist.get_rank() if dist_on else 0
    mine = split_data_for_scatter(list(range(len(baselines))), world)[rank]
    b0 = baselines[mine[0]]
    ntimes, nfreqs = np.asarray(b0["vis"]).shape
    nmodes = np.asarray(b0["fgmodes"]).shape[1]
    if device is None:
        device = rank
    make = engine_factory or pspec.GibbsEngine
    # The device draws of a chain depend on (seed, chain id, iteration) only: every rank uses the same key and the
    # global baseline index as chain id, so the samples do not depend on how the baselines are sharded over GPUs.
    eng = make(len(mine), ntimes, nfreqs, nmodes, max_iters=Niter, rng=rng, keep=keep,
               seed=int(seed) & 0xFFFFFFFFFFFFFFFF, device=device)
    if hasattr(eng, "set_chain_ids"):
        eng.set_chain_ids(np.asarray(mine, dtype=np.int32))
    try:
        for c, gi in enumerate(mine):
            b = baselines[gi]
            eng.load_chain(c, b["vis"], b["flags"], b["fgmodes"], b["ninv_diag"], b["lam0sq"], ps_prior=b.get("ps_prior"))
        eng.run(Niter)
        ps = np.stack([eng.signal_ps(c) for c in range(len(mine))])
        lp = np.stack([eng.ln_post(c) for c in range(len(mine))])
    finally:
        eng.close()
    if gather and world > 1:
        ps = gather_samples(ps, len(baselines))
        lp = gather_samples(lp, len(baselines))
    return ps, lp
