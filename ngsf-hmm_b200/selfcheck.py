"""Multi-rank self-check: a small EM run sharded over the ranks of the current torch.distributed job must
give the numbers of one rank owning everything (individuals sharded for the recursions, sites sharded for the
frequency update, posteriors and emission ratios crossing ranks in between).  Used by bench.py before it times
a multi-GPU run and by tests/multi_gpu_check.py; compares this library with itself only."""
from __future__ import annotations

import numpy as np

from . import api, sim
from .em import EmRank

N, S, ITERS = 11, 30000, 3                      # N not divisible by the world size on purpose


def _run(gl, dist_mb, device, n_ranks, rank, direct, group=None):
    ctx = api.Context(N, S, device=device, n_ranks=n_ranks, rank=rank)
    ctx.upload_gl(np.ascontiguousarray(gl[ctx.site_begin:ctx.site_begin + ctx.sites_owned]))
    ctx.upload_pos_dist(dist_mb)
    ctx.set_freq(np.full(ctx.sites_owned, 0.1))
    n = ctx.n_ind_owned
    F = np.full(n, 0.1); a = np.full(n, 0.2)
    ctx.set_ind_params(F, a)
    runner = EmRank(ctx, freq_est=1, group=group)
    if direct:
        runner.enable_peer_direct(posteriors=direct != "mixed")
    runner.refresh_emissions()
    lks = []
    fr = None
    for _ in range(ITERS):
        lk, fr = runner.iteration(F, a)
        lks.append(lk)
    fr = np.array(fr) if fr is not None else np.empty(0)
    runner.refresh_emissions(with_e0=True)
    ctx.set_ind_params(F, a)
    path = ctx.viterbi()
    post = ctx.get_posterior()
    out = dict(F=F, a=a, lk=np.stack(lks) if n else np.empty((ITERS, 0)), freq=fr, path=path, post=post)
    if n_ranks > 1:
        import torch.distributed as dist
        dist.barrier(group=group)           # nobody closes its windows while a peer may still read them
    ctx.close()
    return out


def multi_rank_check(local_device: int, direct=True, group=None) -> dict | None:
    """Collective over the current process group; direct = True (kernels store posteriors and emission ratios into
    peer windows), "mixed" (emission ratios by peer stores, posteriors by all-to-all) or False (all-to-alls).
    Returns the comparison on rank 0 (None elsewhere):
    max deviations from the single-rank run and ``ok`` = far inside the parity tolerances
    (lkl 1e-9 relative, F / alpha / freq 1e-6, posterior 1e-8, identical Viterbi paths)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    d = sim.simulate(N, S, seed=2024, freq=(0.05, 0.5), indF=(0.0, 0.5), alpha=0.02)
    gl = d.log_gl - np.log(np.exp(d.log_gl).sum(-1, keepdims=True))
    gl = gl - np.log(np.exp(gl).sum(-1, keepdims=True))
    one = _run(gl, d.dist_mb, local_device, 1, 0, False) if rank == 0 else None
    mine = _run(gl, d.dist_mb, local_device, world, rank, direct, group)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    if rank != 0:
        return None
    F = np.concatenate([g["F"] for g in gathered]); a = np.concatenate([g["a"] for g in gathered])
    lk = np.concatenate([g["lk"] for g in gathered], axis=1)
    freq = np.concatenate([g["freq"] for g in gathered])
    path = np.concatenate([g["path"] for g in gathered]); post = np.concatenate([g["post"] for g in gathered])
    dpost = np.abs(post - one["post"])
    res = {
        "case": f"{N} individuals x {S} sites, {ITERS} EM iterations + Viterbi, {world} ranks vs 1 rank",
        "exchange": {True: "fused peer stores", False: "NCCL all-to-all"}.get(direct, "mixed: posteriors by all-to-all, emission ratios by peer stores"),
        "max_abs_dF": float(np.abs(F - one["F"]).max()), "max_abs_dalpha": float(np.abs(a - one["a"]).max()),
        "max_abs_dfreq": float(np.abs(freq - one["freq"]).max()),
        "max_rel_dlkl": float((np.abs(lk - one["lk"]) / np.abs(one["lk"])).max()),
        "posterior_over_1e-8": int((dpost > 1e-8).sum()), "max_abs_dposterior": float(dpost.max()),
        "paths_identical": bool(np.array_equal(path, one["path"])),
    }
    # Not bitwise: the site-block size (hence tile boundaries and the order in which sum log e0 is accumulated)
    # depends on the number of ranks.
    res["ok"] = bool(res["max_abs_dF"] < 1e-7 and res["max_abs_dalpha"] < 1e-7 and res["max_abs_dfreq"] < 1e-9
                     and res["max_rel_dlkl"] < 1e-11 and res["paths_identical"] and res["posterior_over_1e-8"] <= 2
                     and res["max_abs_dposterior"] < 1.1e-5)
    return res
