#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run here (where /root/reference exists):   python tests/golden/make_golden.py
It builds oracle/_ref from the reference sources (oracle/Makefile), drives the
reference's own init / iter_EM / EM / viterbi through oracle/ref_harness.cpp on
small seeded inputs and stores inputs + outputs as .npz.  The GPU box has no
/root/reference: tests there read these files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ngsf_hmm_b200  # noqa: E402,F401
from ngsf_hmm_b200 import sim  # noqa: E402
from _oracle import Ref  # noqa: E402


def one_iteration_case(ref, name, N, S, seed, *, breaks=(), call_geno=False, freq0=0.1, F0=0.1, a0=0.2,
                       sim_kw=None):
    d = sim.simulate(N, S, seed=seed, **(sim_kw or dict(freq=(0.05, 0.5), indF=(0.0, 0.5))))
    dist = d.dist_mb.copy()
    for b in breaks:
        dist[b] = np.inf
    st = ref.state(d.log_gl, dist, freq0, F0, a0, freq_est=1, call_geno=call_geno)
    init = st.get()
    st.iter_EM()
    after = st.get()
    st.viterbi()
    vit = st.get()
    st.close()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        log_gl=d.log_gl, dist_mb=dist, freq0=np.float64(freq0), F0=np.float64(F0), a0=np.float64(a0),
        call_geno=np.bool_(call_geno),
        gl_norm=init["gl_norm"], e_prob0=init["e_prob"],
        marg1=after["marg1"], ind_lkl=after["ind_lkl"], indF=after["indF"], alpha=after["alpha"],
        freq=after["freq"], e_prob1=after["e_prob"], path=vit["path"])
    print(name, "ok", N, S)


def full_em_case(ref, name, N, S, seed, min_iters=10, max_iters=30):
    d = sim.simulate(N, S, seed=seed, freq=0.2, indF=0.5, alpha=0.01, depth=2.0)
    st = ref.state(d.log_gl, d.dist_mb, 0.1, 0.1, 0.2, freq_est=1, n_threads=4,
                   out_prefix=os.path.join("/tmp", "golden_" + name))
    st.run_EM(min_iters, max_iters, 1e-5)
    out = st.get()
    st.close()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        log_gl=d.log_gl, dist_mb=d.dist_mb, min_iters=min_iters, max_iters=max_iters,
        indF=out["indF"], alpha=out["alpha"], freq=out["freq"], ind_lkl=out["ind_lkl"], marg1=out["marg1"],
        path=out["path"], tot_lkl=np.float64(out["tot_lkl"]))
    print(name, "ok", N, S, "tot_lkl", out["tot_lkl"])


if __name__ == "__main__":
    ref = Ref()
    one_iteration_case(ref, "iter_small", 5, 300, 101)
    one_iteration_case(ref, "iter_breaks", 4, 400, 102, breaks=(0, 120, 121, 333), F0=0.3, a0=1.5, freq0=0.25)
    one_iteration_case(ref, "iter_called", 3, 250, 103, call_geno=True,
                       sim_kw=dict(freq=(0.1, 0.5), indF=(0.1, 0.6), depth=8.0))
    one_iteration_case(ref, "iter_manyind", 37, 120, 104)
    full_em_case(ref, "em_full", 6, 1500, 12345)
