"""GPU: the CUDA path through the C ABI against (a) the committed golden fixtures generated from the
unmodified reference and (b) the reference itself (prebuilt oracle/_ref) on a full EM run of the
config-1 shape.  Tolerances: the north-star's (lkl 1e-9 rel, posterior 1e-8 abs, F/alpha/freq 1e-6,
Viterbi identical)."""
import glob
import os

import numpy as np
import pytest

import ngsf_hmm_b200 as nfh
from ngsf_hmm_b200 import sim

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ITER_CASES = sorted(glob.glob(os.path.join(HERE, "golden", "iter_*.npz")))


def _ctx_from_gl_norm(gl_norm, dist, freq, F, a):
    N, S, _ = gl_norm.shape
    ctx = nfh.Context(N, S)
    ctx.upload_gl(np.ascontiguousarray(np.transpose(gl_norm, (1, 0, 2))))
    ctx.upload_pos_dist(dist)
    ctx.set_freq(np.broadcast_to(freq, (S,)).astype(np.float64))
    ctx.set_ind_params(np.broadcast_to(F, (N,)).astype(np.float64), np.broadcast_to(a, (N,)).astype(np.float64))
    ctx.emission_refresh()
    return ctx


def _post_ok(got, want):
    diff = np.abs(got - want)
    bad = diff > 1e-8
    flips = bad & ((want == 0) | (want == 1) | (got == 0) | (got == 1)) & (diff < 1.1e-5)
    return not (bad & ~flips).any() and flips.sum() <= max(2, got.size // 20000)


@pytest.mark.parametrize("path", ITER_CASES, ids=[os.path.basename(p) for p in ITER_CASES])
def test_one_em_iteration_matches_reference_fixture(path):
    g = {k: v for k, v in np.load(path).items()}
    S, N, _ = g["log_gl"].shape
    F0 = np.full(N, float(g["F0"])); a0 = np.full(N, float(g["a0"]))
    with _ctx_from_gl_norm(g["gl_norm"], g["dist_mb"], float(g["freq0"]), F0, a0) as ctx:
        runner = nfh.EmRank(ctx, freq_est=1)
        F, a = F0.copy(), a0.copy()
        lk, fr = runner.iteration(F, a)
        np.testing.assert_allclose(lk, g["ind_lkl"], rtol=1e-9, atol=0)
        assert _post_ok(ctx.get_posterior(), g["marg1"])
        np.testing.assert_allclose(fr, g["freq"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(F, g["indF"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(a, g["alpha"], rtol=0, atol=1e-6)
        # Viterbi with the reference's own parameters after the iteration
        ctx.set_freq(g["freq"]); ctx.set_ind_params(g["indF"], g["alpha"])
        ctx.emission_refresh(with_e0=True)
        assert (ctx.viterbi() == g["path"]).all()


def test_full_em_matches_reference_fixture():
    g = {k: v for k, v in np.load(os.path.join(HERE, "golden", "em_full.npz")).items()}
    S, N, _ = g["log_gl"].shape
    from _oracle import Oracle
    gl = Oracle().normalize_gl(np.transpose(g["log_gl"], (1, 0, 2)))
    with _ctx_from_gl_norm(gl, g["dist_mb"], 0.1, 0.1, 0.2) as ctx:
        runner = nfh.EmRank(ctx, freq_est=1)
        F = np.full(N, 0.1); a = np.full(N, 0.2)
        out = nfh.run_em(runner, F, a, min_iters=int(g["min_iters"]), max_iters=int(g["max_iters"]))
        np.testing.assert_allclose(out["tot_lkl"], float(g["tot_lkl"]), rtol=1e-9)
        np.testing.assert_allclose(F, g["indF"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(a, g["alpha"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(out["freq"], g["freq"], rtol=0, atol=1e-6)
        assert _post_ok(ctx.get_posterior(), g["marg1"])
        assert (out["path"] != g["path"]).sum() == 0


@pytest.mark.ref
@pytest.mark.timeout(900)
def test_full_em_config1_shape_against_reference_binary(ref):
    """BASELINE configs[0] shape: 20 individuals x 10,000 sites, --freq_est 1, --freq 0.1 --indF 0.1,0.2."""
    N, S = 20, 10000
    d = sim.simulate(N, S, seed=12345, freq=0.2, indF=0.5, alpha=0.01, depth=2.0)
    st = ref.state(d.log_gl, d.dist_mb, 0.1, 0.1, 0.2, freq_est=1, n_threads=min(N, os.cpu_count() or 1),
                   out_prefix="/tmp/nfh_cfg1_ref")
    gl = st.get()["gl_norm"]
    st.run_EM(10, 100, 1e-5)
    want = st.get()
    st.close()
    with _ctx_from_gl_norm(gl, d.dist_mb, 0.1, 0.1, 0.2) as ctx:
        runner = nfh.EmRank(ctx, freq_est=1)
        F = np.full(N, 0.1); a = np.full(N, 0.2)
        out = nfh.run_em(runner, F, a, min_iters=10, max_iters=100)
        np.testing.assert_allclose(out["tot_lkl"], want["tot_lkl"], rtol=1e-9)
        np.testing.assert_allclose(F, want["indF"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(a, want["alpha"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(out["freq"], want["freq"], rtol=0, atol=1e-6)
        assert _post_ok(ctx.get_posterior(), want["marg1"])
        assert (out["path"] != want["path"]).sum() == 0
