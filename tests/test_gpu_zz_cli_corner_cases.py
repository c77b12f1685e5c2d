"""GPU: corner cases of the drop-in binary found by reading the reference's start-up code against ours on the CPU
(this file sorts last on purpose: these combinations were written after the round's last GPU session)."""
import numpy as np
import pytest

import ngsf_hmm_b200  # noqa: F401
from ngsf_hmm_b200 import sim
from test_gpu_cli import OURS, REF, _compare, _parse_indF, _run

pytestmark = [pytest.mark.gpu, pytest.mark.ref]


def test_freq_e_with_fixed_frequencies_estimates_only_the_first_site(tmp_path):
    """--freq e --freq_est 0: init_output estimates a site only `if(freq_est == 1 || s == 1)` (parse_args.cpp:316-318),
    so site 1 gets est_maf with F = 0 and every other site keeps 0.01 for the whole run."""
    N, S = 6, 2500
    d = sim.simulate(N, S, seed=2718, freq=(0.1, 0.5), indF=(0.1, 0.6), alpha=0.02, depth=3.0)
    sim.write_binary_gl(str(tmp_path / "in.glf"), d.log_gl)
    sim.write_pos(str(tmp_path / "in.pos"), d.pos_bp)
    common = ["--geno", "in.glf", "--loglkl", "--n_ind", str(N), "--n_sites", str(S), "--pos", "in.pos", "--freq", "e",
              "--freq_est", "0", "--indF", "0.3,0.05", "--indF_fixed", "--alpha_fixed", "--min_iters", "1", "--max_iters", "2",
              "--verbose", "0"]                      # fixed F / alpha: nothing chaotic between the quirk and the files
    _run(REF, common + ["--out", "ref"], str(tmp_path))
    _run(OURS, common + ["--out", "ours"], str(tmp_path))
    _, _, _, fr = _parse_indF(str(tmp_path / "ours.indF"), N)
    assert fr[0] > 0.011 and np.all(fr[1:] == 0.01)
    _compare(tmp_path, N, S, f_tol=1e-9)
