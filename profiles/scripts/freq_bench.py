#!/usr/bin/env python
"""Stand-alone timing of the frequency EM + emission refresh (nfh_freq_update, method 1) against the number of
individuals a site has, at a fixed number of individual-sites.

    python profiles/scripts/freq_bench.py --n_ind 100,125,200,400,800,1000 --ind_sites 5e7

Synthetic depth-2 GL generated on the GPU (the bench's generator), posterior = a smooth pseudo-random field, start
frequency as est_maf's (0.01).  One JSON line per n_ind: ms per call, picoseconds per individual-pass (passes are
counted by the kernel) and the fraction of the DFMA probe peak at 13.75 flop per individual-pass."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n_ind", default="100,125,200,400,800,1000")
    ap.add_argument("--ind_sites", type=float, default=5e7)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import ngsf_hmm_b200 as nfh
    from ngsf_hmm_b200 import api, sim

    dev = torch.device("cuda", 0)
    for N in [int(x) for x in args.n_ind.split(",")]:
        S = int(args.ind_sites // N)
        with nfh.Context(N, S) as ctx:
            chunk = 1 << 16
            while chunk > 4096 and chunk * N > 2.2e8:
                chunk >>= 1
            buf = torch.empty((chunk, N, 3), dtype=torch.float64, device=dev)
            for lo in range(0, S, chunk):
                hi = min(lo + chunk, S)
                sim.simulate_torch(N, S, device=dev, seed=1002, site_chunk=chunk, site_begin=lo, site_end=hi, out=buf[:hi - lo])
                ctx.upload_gl(buf[:hi - lo], first_site=lo)
            del buf
            # posterior window: something between 0 and 1 that varies by individual and site
            ptr, nbytes, _ = ctx.window(api.WIN_POST_RECV)

            class _W:
                pass
            w = _W()
            w.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 2}
            post = torch.as_tensor(w, device="cuda")
            g = torch.Generator(device="cuda").manual_seed(5)
            post.copy_(torch.rand(post.shape, generator=g, device="cuda", dtype=torch.float64) * 0.6)
            torch.cuda.synchronize()
            ctx.set_freq(np.full(S, 0.1))
            probe = ctx.probe_fp64()
            ctx.freq_update(want_freq=False)
            ctx.sync()
            ctx.timing(True)
            ctx.timing_read(reset=True)
            ctx.freq_passes(reset=True)
            for _ in range(args.reps):
                ctx.freq_update(want_freq=False)
            ctx.sync()
            ms = ctx.timing_read(reset=True)["freq"][0] / args.reps
            passes = ctx.freq_passes(reset=True) / args.reps
            ind_passes = passes * N
            print(json.dumps({"n_ind": N, "n_sites": S, "ms": ms, "passes_per_site": passes / S,
                              "ps_per_ind_pass": ms * 1e-3 / ind_passes * 1e12,
                              "frac_of_dfma_probe": 13.75 * ind_passes / (ms * 1e-3) / probe,
                              "force_g": os.environ.get("NFH_FREQ_G")}), flush=True)


if __name__ == "__main__":
    main()
