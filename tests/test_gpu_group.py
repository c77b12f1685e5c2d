"""GPU (one device is enough): the multi-rank geometry driven by ONE process (nfh_group_*, the C++ host
library's multi-GPU driver) equals the single-rank run.  Ranks that share a device exercise everything the
8-GPU run does - individuals sharded for the recursions, sites sharded for the frequency EM, E-step and
frequency kernels storing into the peer windows (or block copies between the windows), the reduction of
sum log e0 - so the driver's single-GPU test run covers the sharded path; with two or more GPUs the same
test also runs with one rank per device over NVLink peer memory."""
import numpy as np
import pytest

import ngsf_hmm_b200 as nfh
from ngsf_hmm_b200 import sim

pytestmark = pytest.mark.gpu

N, S, ITERS = 11, 30000, 3                      # N not divisible by the rank counts on purpose


@pytest.fixture(scope="module")
def data():
    d = sim.simulate(N, S, seed=2024, freq=(0.05, 0.5), indF=(0.0, 0.5), alpha=0.02)
    gl = d.log_gl - np.log(np.exp(d.log_gl).sum(-1, keepdims=True))
    gl = gl - np.log(np.exp(gl).sum(-1, keepdims=True))
    d.dist_mb[S // 2] = np.inf                   # a chromosome start inside the second site block
    return d, np.ascontiguousarray(gl)


def _run(data, devices, fused):
    d, gl = data
    with nfh.Group(N, S, devices, fused_exchange=fused) as g:
        g.upload_gl(gl)
        g.upload_pos_dist(d.dist_mb)
        g.set_freq(np.full(S, 0.1))
        F = np.full(N, 0.1); a = np.full(N, 0.2)
        g.set_ind_params(F, a)
        g.refresh_emissions()
        lks = []
        for _ in range(ITERS):
            lk, fr = g.iteration(F, a)
            lks.append(lk)
        g.refresh_emissions(with_e0=True)
        g.set_ind_params(F, a)
        path = g.viterbi()
        post = g.get_posterior()
        geno = g.geno_posterior(path)
        return dict(F=F, a=a, lk=np.stack(lks), freq=fr, path=path, post=post, geno=geno)


@pytest.fixture(scope="module")
def one_rank(data):
    return _run(data, [0], True)


def _check(got, one):
    # Not bitwise: the site-block size (hence tile boundaries and the order in which sum log e0 is
    # accumulated) depends on the number of ranks.  The runs must agree far inside the parity tolerances
    # (lkl 1e-9 relative, F/alpha/freq 1e-6, posterior 1e-8, identical paths).
    assert np.abs(got["F"] - one["F"]).max() < 1e-7
    assert np.abs(got["a"] - one["a"]).max() < 1e-7
    assert np.abs(got["freq"] - one["freq"]).max() < 1e-9
    assert (np.abs(got["lk"] - one["lk"]) / np.abs(one["lk"])).max() < 1e-11
    assert np.array_equal(got["path"], one["path"])
    dpost = np.abs(got["post"] - one["post"])
    assert (dpost > 1e-8).sum() <= 2 and dpost.max() < 1.1e-5
    assert np.abs(got["geno"] - one["geno"]).max() < 1e-9


@pytest.mark.parametrize("n_ranks,fused", [(2, True), (2, False), (3, True), (8, True)])
def test_ranks_on_one_device_match_one_rank(data, one_rank, n_ranks, fused):
    _check(_run(data, [0] * n_ranks, fused), one_rank)


def test_one_rank_per_device_matches_one_rank(data, one_rank):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the single-device emulation above covers the sharded path)")
    _check(_run(data, list(range(min(n, 8))), True), one_rank)
    _check(_run(data, list(range(min(n, 8))), False), one_rank)


def test_group_matches_single_context_api(data, one_rank):
    """The group with one rank is the plain Context / EmRank path."""
    d, gl = data
    with nfh.Context(N, S) as ctx:
        ctx.upload_gl(gl); ctx.upload_pos_dist(d.dist_mb)
        ctx.set_freq(np.full(S, 0.1))
        F = np.full(N, 0.1); a = np.full(N, 0.2)
        ctx.set_ind_params(F, a)
        runner = nfh.EmRank(ctx, freq_est=1)
        runner.refresh_emissions()
        for _ in range(ITERS):
            lk, fr = runner.iteration(F, a)
        np.testing.assert_array_equal(F, one_rank["F"])
        np.testing.assert_array_equal(fr, one_rank["freq"])
        np.testing.assert_array_equal(lk, one_rank["lk"][-1])
