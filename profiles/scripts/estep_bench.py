#!/usr/bin/env python
"""Stand-alone timing of the fused E-step (and optionally one objective round) on synthetic emission ratios.

    python profiles/scripts/estep_bench.py --n_ind 100 --n_sites 1000000 --alpha 0.01,0.2

The emission-ratio window is filled on the device (log-normal ratios), distances follow the simulator's
Normal(1e5, 1e5/3) bp, so no GL is generated: only the recursion kernels run.  Prints one JSON line per
alpha: device time of nfh_estep (CUDA events on the context stream, family timers) and the achieved
fraction of the measured HBM peak at the contract figure of 24 B per individual-site.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n_ind", type=int, default=100)
    ap.add_argument("--n_sites", type=int, default=1_000_000)
    ap.add_argument("--alpha", default="0.01,0.2")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--lkl", action="store_true", help="also time one 5-point objective round per individual")
    ap.add_argument("--breaks", type=int, default=0, help="number of chromosome starts (d = +inf)")
    args = ap.parse_args()

    import torch
    import ngsf_hmm_b200 as nfh
    from ngsf_hmm_b200 import api

    N, S = args.n_ind, args.n_sites
    peak = 6548.8
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    rng = np.random.default_rng(7)
    dist = np.maximum(1.0, np.floor(rng.normal(1e5, 1e5 / 3, S))) / 1e6
    if args.breaks:
        dist[rng.choice(S, args.breaks, replace=False)] = np.inf
    with nfh.Context(N, S) as ctx:
        ctx.upload_pos_dist(dist)
        ptr, nbytes, _ = ctx.window(api.WIN_EMIS_RECV)

        class _W:
            pass
        w = _W()
        w.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        emis = torch.as_tensor(w, device="cuda").view(ctx.n_ind_local, ctx.site_block)
        g = torch.Generator(device="cuda").manual_seed(3)
        emis[:, :S] = torch.exp(torch.randn((ctx.n_ind_local, S), generator=g, device="cuda", dtype=torch.float64))
        torch.cuda.synchronize()
        F = rng.uniform(0.05, 0.5, N)
        for alpha in [float(x) for x in args.alpha.split(",")]:
            a = np.full(N, alpha)
            ctx.set_ind_params(F, a)
            for _ in range(3):
                lk = ctx.estep()
            ctx.timing(True)
            ctx.timing_read(reset=True)
            for _ in range(args.reps):
                ctx.estep_async()
            ms = ctx.timing_read(reset=True)["estep"][0] / args.reps
            out = {"kernel": "estep", "single_launch": os.environ.get("NFH_ESTEP_FUSED", "0") not in ("", "0"), "n_ind": N, "n_sites": S,
                   "alpha": alpha, "ms": ms, "gbs_at_24B": 24.0 * N * S / (ms * 1e-3) / 1e9,
                   "frac_of_hbm_peak": 24.0 * N * S / (ms * 1e-3) / 1e9 / peak, "lkl0": float(lk[0])}
            if args.lkl:
                eh = 4e-6
                ind = np.repeat(np.arange(N), 5)
                Fq = np.repeat(F, 5) + np.tile([0, eh, -eh, 0, 0], N)
                aq = np.repeat(a, 5) + np.tile([0, 0, 0, eh, -eh], N)
                for _ in range(2):
                    ctx.lkl_batch(ind, Fq, aq)
                ctx.timing_read(reset=True)
                for _ in range(args.reps):
                    ctx.lkl_batch(ind, Fq, aq)
                out["lkl_round_ms"] = ctx.timing_read(reset=True)["lkl_batch"][0] / args.reps
            ctx.timing(False)
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
