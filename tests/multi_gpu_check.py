"""Run under torchrun with >= 2 GPUs: the multi-rank EM (individuals sharded for the recursions,
sites sharded for the frequency update, all-to-all in between) must give the same numbers as
one rank owning everything.  Prints MULTI_GPU_OK on success (rank 0)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ngsf_hmm_b200 as nfh  # noqa: E402
from ngsf_hmm_b200 import sim  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    N, S, iters = 11, 30000, 3                      # N not divisible by the world size on purpose
    d = sim.simulate(N, S, seed=2024, freq=(0.05, 0.5), indF=(0.0, 0.5), alpha=0.02)
    gl = d.log_gl - np.log(np.exp(d.log_gl).sum(-1, keepdims=True))
    gl = gl - np.log(np.exp(gl).sum(-1, keepdims=True))

    def run(n_ranks, r, direct=False):
        ctx = nfh.Context(N, S, device=lr, n_ranks=n_ranks, rank=r)
        ctx.upload_gl(np.ascontiguousarray(gl[ctx.site_begin:ctx.site_begin + ctx.sites_owned]))
        ctx.upload_pos_dist(d.dist_mb)
        ctx.set_freq(np.full(ctx.sites_owned, 0.1))
        n = ctx.n_ind_owned
        F = np.full(n, 0.1); a = np.full(n, 0.2)
        ctx.set_ind_params(F, a)
        runner = nfh.EmRank(ctx, freq_est=1)
        if direct:
            runner.enable_peer_direct()
        runner.refresh_emissions()
        lks = []
        for _ in range(iters):
            lk, fr = runner.iteration(F, a)
            lks.append(lk)
        runner.refresh_emissions(with_e0=True)
        ctx.set_ind_params(F, a)
        path = ctx.viterbi()
        post = ctx.get_posterior()
        out = dict(F=F, a=a, lk=np.stack(lks), freq=fr, path=path, post=post, ind_begin=ctx.ind_begin,
                   site_begin=ctx.site_begin)
        ctx.close()
        return out

    ok = True
    one = run(1, 0) if rank == 0 else None
    for direct in (False, True):
        mine = run(world, rank, direct)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank != 0:
            continue
        print("exchange:", "fused peer stores over NVLink" if direct else "NCCL all-to-all", flush=True)
        F = np.concatenate([g["F"] for g in gathered]); a = np.concatenate([g["a"] for g in gathered])
        lk = np.concatenate([g["lk"] for g in gathered], axis=1)
        freq = np.concatenate([g["freq"] for g in gathered])
        path = np.concatenate([g["path"] for g in gathered]); post = np.concatenate([g["post"] for g in gathered])
        checks = {"F": np.array_equal(F, one["F"]), "alpha": np.array_equal(a, one["a"]),
                  "lkl": np.array_equal(lk, one["lk"]), "freq": np.array_equal(freq, one["freq"]),
                  "path": np.array_equal(path, one["path"]), "posterior": np.array_equal(post, one["post"])}
        # Not bitwise: the site-block size (hence tile boundaries and the order in which sum log e0 is
        # accumulated) depends on the number of ranks.  The runs must agree far inside the parity
        # tolerances (lkl 1e-9 relative, F/alpha/freq 1e-6, posterior 1e-8, identical paths).
        print("bitwise identical to the single-rank run:", checks, flush=True)
        dF = np.abs(F - one["F"]).max(); da = np.abs(a - one["a"]).max(); dfr = np.abs(freq - one["freq"]).max()
        dlk = (np.abs(lk - one["lk"]) / np.abs(one["lk"])).max(); dpost = np.abs(post - one["post"])
        print(f"max |dF| {dF:.3e} |dalpha| {da:.3e} |dfreq| {dfr:.3e} rel |dlkl| {dlk:.3e} "
              f"posterior > 1e-8: {(dpost > 1e-8).sum()} of {dpost.size}", flush=True)
        ok = ok and bool(dF < 1e-7 and da < 1e-7 and dfr < 1e-9 and dlk < 1e-11 and checks["path"]
                         and (dpost > 1e-8).sum() <= 2 and dpost.max() < 1.1e-5)
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_MISMATCH", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ok = bool(flag.item())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
