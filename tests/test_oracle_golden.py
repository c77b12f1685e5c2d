"""CPU: the oracle (our C restatement) against the golden fixtures generated from the unmodified
reference (tests/golden/make_golden.py).  On the same libm the match is bit for bit; the asserted
bound (1e-12) only leaves room for a different libm build."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ITER_CASES = sorted(glob.glob(os.path.join(HERE, "golden", "iter_*.npz")))
TOL = 1e-12


def _load(path):
    return {k: v for k, v in np.load(path).items()}


def test_fixtures_present():
    assert len(ITER_CASES) >= 4
    assert os.path.exists(os.path.join(HERE, "golden", "em_full.npz"))


@pytest.mark.parametrize("path", ITER_CASES, ids=[os.path.basename(p) for p in ITER_CASES])
def test_oracle_reproduces_reference_iteration(oracle, path):
    g = _load(path)
    S, N, _ = g["log_gl"].shape
    gl = oracle.normalize_gl(np.transpose(g["log_gl"], (1, 0, 2)), call_geno=bool(g["call_geno"]))
    finite = np.isfinite(g["gl_norm"]) & (g["gl_norm"] > -1e14)
    np.testing.assert_allclose(gl[finite], g["gl_norm"][finite], rtol=0, atol=TOL)
    gl = g["gl_norm"]                                   # continue from the reference's own values
    freq0 = np.full(S, float(g["freq0"])); F0 = np.full(N, float(g["F0"])); a0 = np.full(N, float(g["a0"]))
    _, e0 = oracle.freq_emission(gl, None, freq0, update_freq=False)
    np.testing.assert_allclose(e0, g["e_prob0"], rtol=0, atol=TOL * 10)
    st, marg1, lk = oracle.estep(g["e_prob0"], g["dist_mb"], F0, a0)
    assert st == 0
    np.testing.assert_allclose(lk, g["ind_lkl"], rtol=1e-14, atol=0)
    np.testing.assert_allclose(marg1, g["marg1"], rtol=0, atol=TOL)
    fr, e1 = oracle.freq_emission(gl, g["marg1"], freq0, update_freq=True)
    np.testing.assert_allclose(fr, g["freq"], rtol=0, atol=TOL)
    np.testing.assert_allclose(e1, g["e_prob1"], rtol=0, atol=1e-9)
    for i in range(N):
        _, p = oracle.viterbi(g["e_prob1"][i], g["dist_mb"], g["indF"][i], g["alpha"][i])
        assert (p == g["path"][i]).all()


@pytest.mark.parametrize("path", ITER_CASES[:2], ids=[os.path.basename(p) for p in ITER_CASES[:2]])
def test_host_optimiser_reproduces_reference_bfgs(oracle, path):
    """F/alpha after one iteration: our L-BFGS-B + numeric gradient on the oracle objective."""
    from _host import minimize
    g = _load(path)
    S, N, _ = g["log_gl"].shape
    for i in range(N):
        x, _, _ = minimize(lambda v: oracle.lkl(g["e_prob0"][i], g["dist_mb"], v[0], v[1]),
                           [float(g["F0"]), float(g["a0"])], [1e-15, 1e-15], [1 - 1e-15, 10.0])
        assert abs(x[0] - g["indF"][i]) <= 1e-12 and abs(x[1] - g["alpha"][i]) <= 1e-12, (i, x, g["indF"][i], g["alpha"][i])


def test_extended_precision_adjudicator_agrees(oracle):
    g = _load(ITER_CASES[0])
    S, N, _ = g["log_gl"].shape
    for i in range(N):
        m, lk = oracle.estep_extended(g["e_prob0"][i], g["dist_mb"], float(g["F0"]), float(g["a0"]))
        assert abs(lk - g["ind_lkl"][i]) < 1e-9 * abs(lk)
        clamped = np.where(m < 1e-5, 0.0, np.where(m > 1 - 1e-5, 1.0, m))
        assert np.abs(clamped - g["marg1"][i]).max() < 1e-8
