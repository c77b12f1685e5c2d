"""CPU, needs oracle/_ref (built here from /root/reference): every oracle function against the
unmodified reference on seeded random inputs, bit for bit."""
import numpy as np
import pytest

import ngsf_hmm_b200  # noqa: F401
from ngsf_hmm_b200 import sim

pytestmark = pytest.mark.ref


@pytest.fixture(scope="module")
def case(ref, oracle):
    d = sim.simulate(7, 1200, seed=77, freq=(0.05, 0.5), indF=(0.0, 0.5))
    d.dist_mb[[0, 400]] = np.inf
    st = ref.state(d.log_gl, d.dist_mb, 0.15, 0.2, 0.7)
    init = st.get()
    st.iter_EM()
    after = st.get()
    st.viterbi()
    path = st.get()["path"]
    st.close()
    return d, init, after, path


def test_scalar_functions(ref, oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        maf = rng.uniform(0, 1); F = rng.choice([0.0, 1.0, rng.uniform(0, 1)])
        for ls in (True, False):
            np.testing.assert_array_equal(oracle.calc_HWE(maf, F, ls), ref.calc_HWE(maf, F, ls))
        gl = np.log(rng.dirichlet([0.3, 0.3, 0.3]) + 1e-300)
        for k in (0, 1):
            assert oracle.calc_emission(gl, maf, k) == ref.calc_emission(gl, maf, k)
    for n in (1, 3, 20, 101):
        gl = np.log(rng.dirichlet([0.5, 0.5, 0.5], size=n))
        F = rng.choice([0.0, 1.0, 0.3, 0.9999], size=n)
        assert oracle.est_maf(gl, F) == ref.est_maf(gl, F)


def test_recursions_bit_identical(ref, oracle, case):
    d, init, after, path = case
    for i in range(d.n_ind):
        e = init["e_prob"][i]
        for F, a in ((0.2, 0.7), (1e-6, 1e-6), (0.999, 9.0)):
            lo, Fo = oracle.forward(e, d.dist_mb, F, a, want_table=True)
            lr, Fr = ref.forward(e, d.dist_mb, F, a, want_table=True)
            assert lo == lr
            np.testing.assert_array_equal(Fo, Fr)
            bo, Bo = oracle.backward(e, d.dist_mb, F, a, want_table=True)
            br, Br = ref.backward(e, d.dist_mb, F, a, want_table=True)
            assert bo == br
            np.testing.assert_array_equal(Bo, Br)
            assert oracle.lkl(e, d.dist_mb, F, a) == ref.lkl(e, d.dist_mb, F, a)
            vo, po = oracle.viterbi(e, d.dist_mb, F, a)
            vr, pr = ref.viterbi(e, d.dist_mb, F, a)
            assert vo == vr and (po == pr).all()


def test_iteration_bit_identical(ref, oracle, case):
    d, init, after, path = case
    N, S = d.n_ind, d.n_sites
    st, marg1, lk = oracle.estep(init["e_prob"], d.dist_mb, np.full(N, 0.2), np.full(N, 0.7))
    assert st == 0
    np.testing.assert_array_equal(lk, after["ind_lkl"])
    np.testing.assert_array_equal(marg1, after["marg1"])
    fr, e1 = oracle.freq_emission(init["gl_norm"], marg1, np.full(S, 0.15), update_freq=True)
    np.testing.assert_array_equal(fr, after["freq"])
    np.testing.assert_array_equal(e1, after["e_prob"])
    for i in range(N):
        _, p = oracle.viterbi(after["e_prob"][i], d.dist_mb, after["indF"][i], after["alpha"][i])
        assert (p == path[i]).all()


def test_fixed_parameter_iteration_changes_nothing_but_the_posterior(ref, oracle):
    """configs[4] semantics (--indF_fixed --alpha_fixed --freq_est 0): the reference's iter_EM then leaves freq, F,
    alpha and every emission untouched (EM.cpp:188-271 is skipped) and only produces ind_lkl and the posterior - which
    is why a multi-GPU run of that configuration has nothing to exchange per iteration (nfh_peer_direct mode 2).  The
    oracle's E-step must reproduce both bit for bit."""
    d = sim.simulate(5, 900, seed=404, freq=(0.05, 0.5), indF=(0.05, 0.5), alpha=0.03)
    freq0 = np.clip(d.true_freq, 0.01, 0.49)
    F0 = np.clip(d.true_F, 1e-3, 1 - 1e-3); a0 = d.true_alpha.copy()
    st = ref.state(d.log_gl, d.dist_mb, freq0, F0, a0, freq_est=0, indF_fixed=True, alpha_fixed=True)
    before = st.get()
    st.iter_EM()
    after = st.get()
    st.iter_EM()
    again = st.get()
    st.close()
    for key in ("freq", "indF", "alpha", "e_prob"):
        np.testing.assert_array_equal(after[key], before[key])
    for key in ("ind_lkl", "marg1"):                      # a second iteration repeats the first
        np.testing.assert_array_equal(again[key], after[key])
    status, marg1, lk = oracle.estep(before["e_prob"], d.dist_mb, F0, a0)
    assert status == 0
    np.testing.assert_array_equal(lk, after["ind_lkl"])
    np.testing.assert_array_equal(marg1, after["marg1"])
