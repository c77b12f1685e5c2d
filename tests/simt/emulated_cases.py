"""Cases for the emulated product only (TEST INFRASTRUCTURE; collected when tests/test_gpu_suite_emulated_subset_cpu.py or
tests/simt/run_gpu_suite_emulated.py pass this file to pytest): the multi-rank test of tests/test_gpu_group.py at a size
an emulator can afford in the default CPU run."""
import numpy as np
import pytest

import test_gpu_group as tg
from ngsf_hmm_b200 import sim

pytestmark = pytest.mark.gpu


def test_small_multi_rank_geometry_matches_one_rank():
    """2 ranks with kernels storing into the peers' windows and 3 ranks with block copies between the windows
    (individuals and sites not divisible by the rank counts; rank 1 of 2 owns 176 sites) against one rank."""
    keep = (tg.N, tg.S, tg.ITERS)
    try:
        tg.N, tg.S, tg.ITERS = 5, 4400, 1
        _run_small()
    finally:
        tg.N, tg.S, tg.ITERS = keep


def _run_small():
    d = sim.simulate(tg.N, tg.S, seed=2024, freq=(0.05, 0.5), indF=(0.0, 0.5), alpha=0.02)
    gl = d.log_gl - np.log(np.exp(d.log_gl).sum(-1, keepdims=True))
    gl = gl - np.log(np.exp(gl).sum(-1, keepdims=True))
    d.dist_mb[4300] = np.inf                      # a chromosome start inside the second site block
    data = (d, np.ascontiguousarray(gl))
    one = tg._run(data, [0], True)
    tg._check(tg._run(data, [0, 0], True), one)
    tg._check(tg._run(data, [0, 0, 0], False), one)
