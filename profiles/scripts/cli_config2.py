#!/usr/bin/env python
"""Wall time of the drop-in binary on BASELINE configs[1] (100 individuals x 1,000,000 sites, raw-double log GL,
--freq_est 1): writes the synthetic input to a scratch directory, runs ngsf-hmm_b200/ngsF-HMM on it and prints
the phases.  Usage (on a B200): python profiles/scripts/cli_config2.py [n_ind] [n_sites] [max_iters]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ngsf_hmm_b200  # noqa: E402,F401
from ngsf_hmm_b200 import sim  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 20
tmp = tempfile.mkdtemp(prefix="nfh_cli_")
t0 = time.perf_counter()
d = sim.simulate_torch(N, S, device=torch.device("cuda", 0), seed=1002)
gl = d["log_gl"].cpu().numpy()
gl.tofile(os.path.join(tmp, "in.glf"))
pos = np.cumsum(np.rint(d["dist_mb"] * 1e6).astype(np.int64))
sim.write_pos(os.path.join(tmp, "in.pos"), pos)
t1 = time.perf_counter()
print(f"input: {gl.nbytes / 1e9:.2f} GB written in {t1 - t0:.1f} s", flush=True)
del gl, d
cmd = [os.path.join(ROOT, "ngsf-hmm_b200", "ngsF-HMM"), "--geno", "in.glf", "--loglkl", "--n_ind", str(N), "--n_sites",
       str(S), "--pos", "in.pos", "--freq", "0.1", "--indF", "0.1,0.2", "--freq_est", "1", "--min_iters", "10",
       "--max_iters", str(ITERS), "--out", "run", "--verbose", "1"]
t2 = time.perf_counter()
p = subprocess.Popen(cmd, cwd=tmp, stdout=subprocess.PIPE, text=True)
marks = {}
for line in p.stdout:
    now = time.perf_counter() - t2
    for key in ("==> Calculating initial emission", "Iteration 1:", "==> Decoding most probable path", "Printing final results",
                "Freeing memory", "Final logLkl", "WARN"):
        if line.startswith(key) and key not in marks:
            marks[key] = now
            print(f"[{now:8.2f} s] {line.strip()}", flush=True)
p.wait()
total = time.perf_counter() - t2
n_it = ITERS
print(f"exit {p.returncode}; total {total:.1f} s: read+upload {marks.get('Iteration 1:', 0):.1f} s, "
      f"EM {marks.get('==> Decoding most probable path', 0) - marks.get('Iteration 1:', 0):.2f} s, "
      f"write {marks.get('Freeing memory', total) - marks.get('Printing final results', 0):.1f} s")
for f in os.listdir(tmp):
    print(f, os.path.getsize(os.path.join(tmp, f)))
    os.remove(os.path.join(tmp, f))
os.rmdir(tmp)
