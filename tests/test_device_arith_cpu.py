"""CPU: the ARITHMETIC of the CUDA kernels against the oracle, without a GPU.

The per-thread bodies of the hot kernels are pure functions in headers of ngsf-hmm_b200/csrc (nfh_math.cuh,
nfh_device.cuh, nfh_estep_math.cuh, nfh_freq_math.cuh).  tests/device_arith_host.cpp compiles exactly those headers
with g++ and strings the bodies together sequentially; here the results meet the oracle at the north star's
tolerances (log-likelihood 1e-9 relative, posterior 1e-8 absolute, frequency 1e-11 per update with the same number
of passes).  What the `-m gpu` tests then add is the parallel decomposition (scans, tiles, lanes) and the hardware's
own reciprocal seed.  The harness is built twice - floating-point contraction off and on (-mfma) - because nvcc
contracts a * b + c where the host compiler must not: the device's rounding lies between the two.

Test infrastructure only: nothing under ngsf-hmm_b200/ builds, loads or falls back to this code.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from ngsf_hmm_b200 import sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ngsf-hmm_b200", "csrc")
_dp = C.POINTER(C.c_double)

LKL_RTOL = 1e-9
POST_ATOL = 1e-8


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class DeviceArith:
    def __init__(self, so):
        L = self.lib = C.CDLL(so)
        for name in ("da_expm1_pos", "da_expm1_small", "da_rcp_seed"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double]
        L.da_rcp_pos.restype = C.c_double
        L.da_rcp_pos.argtypes = [C.c_double, C.c_int]
        L.da_constants.argtypes = [_dp]
        L.da_estep.restype = C.c_int
        L.da_estep.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp, _dp, C.POINTER(C.c_int)]
        L.da_neg_lkl.restype = C.c_double
        L.da_neg_lkl.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double, C.c_double]
        L.da_freq_site.restype = C.c_int
        L.da_freq_site.argtypes = [C.c_int, C.c_uint64, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp]
        L.da_freq_site_stream.restype = C.c_int
        L.da_freq_site_stream.argtypes = [C.c_uint64, _dp, _dp, _dp, _dp, _dp]
        L.da_viterbi.restype = None
        L.da_viterbi.argtypes = [C.c_uint64, _dp, _dp, _dp, C.c_double, C.c_double, C.c_void_p]
        c = np.empty(10)
        L.da_constants(_d(c))
        self.big_x, self.fast_x, self.mid_x, self.eps, self.start_freq, self.start_odds, self.min_odds_inv = c[:7]
        self.chunk, self.tile, self.stash = (int(v) for v in c[7:10])

    def estep(self, e_prob, dist, F, alpha):
        """e_prob (S,2) log emissions of one individual -> (status, posterior, lkl_forward, lkl_backward, tiers)."""
        e = np.ascontiguousarray(e_prob, dtype=np.float64)
        ratio = np.exp(e[:, 1] - e[:, 0])
        dist = np.ascontiguousarray(dist, dtype=np.float64)
        S = len(dist)
        post, lk, tiers = np.empty(S), np.empty(2), (C.c_int * 3)()
        st = self.lib.da_estep(S, _d(ratio), _d(dist), float(F), float(alpha), float(e[:, 0].sum()), _d(post), _d(lk),
                               tiers)
        return st, post, lk[0], lk[1], list(tiers)

    def neg_lkl(self, e_prob, dist, F, alpha):
        e = np.ascontiguousarray(e_prob, dtype=np.float64)
        ratio = np.exp(e[:, 1] - e[:, 0])
        dist = np.ascontiguousarray(dist, dtype=np.float64)
        return self.lib.da_neg_lkl(len(dist), _d(ratio), _d(dist), float(F), float(alpha), float(e[:, 0].sum()))

    def viterbi(self, e_prob, dist, F, alpha):
        e = np.ascontiguousarray(e_prob, dtype=np.float64)
        ratio = np.exp(e[:, 1] - e[:, 0]); e0 = np.exp(e[:, 0])
        dist = np.ascontiguousarray(dist, dtype=np.float64)
        path = np.zeros(len(dist), dtype=np.uint8)
        self.lib.da_viterbi(len(dist), _d(ratio), _d(e0), _d(dist), float(F), float(alpha), path.ctypes.data)
        return path.astype(np.int8)

    def freq_site(self, shape, gl_log, post, freq=None):
        """gl_log (n_ind,3) normalised log GL of one site -> (freq, passes, e1/e0, e0).  freq given: refresh only."""
        L = np.exp(np.ascontiguousarray(gl_log, dtype=np.float64))          # what the gl_ingest kernel stores
        L0, L1, L2 = (np.ascontiguousarray(L[:, k]) for k in range(3))
        p = np.ascontiguousarray(post, dtype=np.float64) if post is not None else None
        n = len(L0)
        f = C.c_double(0.0 if freq is None else float(freq))
        ratio, e0 = np.empty(n), np.empty(n)
        passes = self.lib.da_freq_site(shape, n, _d(L0), _d(L1), _d(L2), _d(p), int(freq is None),
                                       C.cast(C.byref(f), _dp), _d(ratio), _d(e0))
        assert passes >= 0, "shape too small for n_ind"
        return f.value, passes, ratio, e0

    def freq_site_stream(self, gl_log, post):
        L = np.exp(np.ascontiguousarray(gl_log, dtype=np.float64))
        L0, L1, L2 = (np.ascontiguousarray(L[:, k]) for k in range(3))
        p = np.ascontiguousarray(post, dtype=np.float64) if post is not None else None
        f = C.c_double()
        passes = self.lib.da_freq_site_stream(len(L0), _d(L0), _d(L1), _d(L2), _d(p), C.cast(C.byref(f), _dp))
        return f.value, passes


@pytest.fixture(scope="module", params=["contract-off", "contract-fma"])
def dev(request, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("device_arith") / f"libdevice_arith_{request.param}.so")
    flags = ["-ffp-contract=off"] if request.param == "contract-off" else ["-ffp-contract=fast", "-mfma"]
    if request.param == "contract-fma" and "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU has no FMA instructions")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas"] + flags +
                          ["-I", CSRC, "-I", os.path.join(ROOT, "include"), "-o", out,
                           os.path.join(ROOT, "tests", "device_arith_host.cpp")])
    return DeviceArith(out)


def _case(oracle, N, S, seed, freq0=0.1, **simkw):
    d = sim.simulate(N, S, seed=seed, **simkw)
    gl_ind = oracle.normalize_gl(np.transpose(d.log_gl, (1, 0, 2)))          # (N,S,3)
    freq = np.broadcast_to(np.asarray(freq0, dtype=np.float64), (S,)).copy()
    _, e = oracle.freq_emission(gl_ind, None, freq, update_freq=False)       # (N,S,2) log emissions
    return d, gl_ind, freq, e


def _posterior_check(got, want):
    """1e-8 absolute; a site whose value sits on a clamp threshold may flip (counted), as in test_gpu_parity.py."""
    diff = np.abs(got - want)
    bad = diff > POST_ATOL
    if bad.any():
        near = (np.abs(got - 1e-5) < 1e-7) | (np.abs(got - (1 - 1e-5)) < 1e-7) | \
               (np.abs(want - 1e-5) < 1e-7) | (np.abs(want - (1 - 1e-5)) < 1e-7)
        flips = bad & ((want == 0) | (want == 1) | (got == 0) | (got == 1)) & (diff < 1.1e-5)
        assert not (bad & ~flips & ~near).any(), f"max posterior diff {diff.max()}"
        assert flips.sum() <= max(3, got.size // 20000), f"{flips.sum()} clamp flips"


# ---------------------------------------------------------------------------------------------------------------
# nfh_math.cuh
# ---------------------------------------------------------------------------------------------------------------
def _ulps(got, want_ld):
    want = np.asarray(want_ld, dtype=np.float64)
    return np.abs((np.asarray(got, dtype=np.longdouble) - want_ld) / np.spacing(np.abs(want)).astype(np.longdouble))


def test_expm1_pos_is_accurate_over_its_whole_range(dev):
    """kappa = e^x - 1 for x = alpha d in [0, 110 ln 2]: about one ulp of e^x (nfh_math.cuh), i.e. a relative error of
    kappa no worse than ~1 ulp * (1 + 1/kappa)."""
    assert np.finfo(np.longdouble).nmant >= 63, "needs x87 long double"
    rng = np.random.default_rng(1)
    xs = np.concatenate([10.0 ** rng.uniform(-300, -3, 4000), rng.uniform(0, dev.fast_x, 4000),
                         rng.uniform(dev.fast_x, 1.0, 4000), rng.uniform(1.0, dev.big_x, 4000),
                         [0.0, dev.fast_x, dev.mid_x, dev.big_x, np.log(2) / 128, np.log(2) / 64, np.log(2)]])
    got = np.array([dev.lib.da_expm1_pos(float(x)) for x in xs])
    want = np.expm1(xs.astype(np.longdouble))
    err_in_ulps_of_exp = np.abs(got.astype(np.longdouble) - want) / np.spacing(np.exp(xs)).astype(np.longdouble)
    assert err_in_ulps_of_exp.max() <= 1.5, err_in_ulps_of_exp.max()
    # k = 0 (x < ln2/128): the degree-5 polynomial alone - rounding plus its truncation x^6/720, i.e. relative to
    # kappa ~ x at most 2 ulp + x^5/720 (6e-15 at the top of the range, still 0.2 ulp of e^x)
    small = (xs < np.log(2) / 128) & (xs > 0)
    rel = np.abs((got[small].astype(np.longdouble) - want[small]) / want[small]).astype(np.float64)
    assert (rel <= 4.5e-16 + xs[small] ** 5 / 700).all(), (rel - xs[small] ** 5 / 700).max()
    assert dev.lib.da_expm1_pos(0.0) == 0.0
    assert dev.lib.da_expm1_pos(dev.big_x) == pytest.approx(2.0 ** 110, rel=2e-14)     # kBigX is 110 ln 2 rounded


def test_expm1_small_is_expm1_pos_bit_for_bit_in_the_fast_tier(dev):
    """kTierFast replaces expm1_pos() by its polynomial for x < 0.0054: same bits, so the tier choice (made per
    (individual, tile) from alpha * largest distance) can never change a result."""
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(0, dev.fast_x, 20000), 10.0 ** rng.uniform(-30, np.log10(dev.fast_x), 5000),
                         [0.0, np.nextafter(dev.fast_x, 0)]])
    for x in xs:
        assert dev.lib.da_expm1_small(float(x)) == dev.lib.da_expm1_pos(float(x))


def test_rcp_pos_from_a_pessimistic_seed(dev):
    """One cubic Newton step from a seed good to ~2^-20 (worse than MUFU.RCP64H): <= 2 ulp; refined: <= 1 ulp."""
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(1, 2, 5000), 2.0 ** rng.uniform(-200, 200, 5000), [1.0, 1.5, 3.0, 1e-300, 1e300]])
    want = 1 / xs.astype(np.longdouble)
    seed = np.array([dev.lib.da_rcp_seed(float(x)) for x in xs])
    assert 1e-8 < np.abs(seed * xs - 1).max() < 2.0 ** -18       # the stand-in really is a rough seed
    one = np.array([dev.lib.da_rcp_pos(float(x), 0) for x in xs])
    two = np.array([dev.lib.da_rcp_pos(float(x), 1) for x in xs])
    assert _ulps(one, want).max() <= 2.0
    assert _ulps(two, want).max() <= 1.0


def test_layout_constants(dev):
    assert (dev.chunk, dev.tile, dev.stash) == (33, 4224, 8)
    assert dev.eps == 1e-5 and dev.start_freq == 0.01                   # EPSILON gen_func.hpp:16; gen_func.cpp:980
    assert dev.big_x == pytest.approx(110 * np.log(2), rel=1e-15)


# ---------------------------------------------------------------------------------------------------------------
# E-step: products_chunk<TIER>() + apply_chunk<TIER>() against forward / backward / posterior of the oracle
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,S,seed", [(6, 3000, 1), (4, 10000, 12345), (3, 4224, 5), (3, 4225, 6), (2, 17, 7),
                                      (2, 1, 8), (2, 33, 9), (2, 8449, 10)])
def test_estep_bodies_match_oracle(oracle, dev, N, S, seed):
    d, gl_ind, freq, e = _case(oracle, N, S, seed, freq=(0.05, 0.5), indF=(0.0, 0.5))
    rng = np.random.default_rng(seed)
    F = rng.uniform(0.01, 0.6, N); a = rng.uniform(0.005, 2.0, N)
    st, marg1, lk_o = oracle.estep(e, d.dist_mb, F, a)
    assert st == 0
    for i in range(N):
        rc, post, lf, lb, tiers = dev.estep(e[i], d.dist_mb, F[i], a[i])
        assert rc == 0
        assert abs(lf - lk_o[i]) <= LKL_RTOL * abs(lk_o[i])
        assert abs(lf - lb) <= 1e-9 * abs(lf)                               # EM.cpp:166 allows 1e-3
        _posterior_check(post, marg1[i])


def test_estep_bodies_chromosome_breaks_and_extreme_parameters(oracle, dev):
    """Chromosome starts (d = +inf, read_data.cpp:207-209) and alpha d > 1 run the clamp tier (kappa at 110 ln 2); F and
    alpha on the optimiser's bounds (EM.cpp:425-426).  First the case of test_gpu_parity.py, then a longer one.
    (alpha = 1e-15 only on the short case: there 1 - exp(-alpha d) of the reference is quantised to 0 or 1.1e-16, which
    no longer matters over 5,000 sites but costs the REFERENCE 0.29 log units against its own long-double restatement
    on the longer one - the device arithmetic, with kappa from expm1, stays within 1e-4 of the latter.)"""
    for N, S, breaks, F, a in [
            (4, 5000, [0, 1000, 2500, 4999], [1e-6, 1 - 1e-6, 0.3, 1e-15], [1e-6, 10.0, 1e-15, 0.5]),
            (6, 9000, [0, 1000, 2500, 4999, 8999], [1e-6, 1 - 1e-6, 0.3, 1e-15, 0.2, 1 - 1e-15],
             [1e-6, 10.0, 1e-4, 0.5, 0.01, 3.0])]:
        d, gl_ind, freq, e = _case(oracle, N, S, 11, freq0=0.25, freq=(0.05, 0.5), indF=(0.0, 0.5))
        d.dist_mb[breaks] = np.inf
        F = np.array(F); a = np.array(a)
        st, marg1, lk_o = oracle.estep(e, d.dist_mb, F, a)
        assert st == 0
        for i in range(N):
            rc, post, lf, lb, tiers = dev.estep(e[i], d.dist_mb, F[i], a[i])
            assert rc == 0 and tiers[0] == tiers[1] == 0                    # every tile has a break: clamp tier
            assert abs(lf - lk_o[i]) <= LKL_RTOL * abs(lk_o[i]), (N, i)
            assert abs(lf - lb) <= 1e-9 * abs(lf)
            _posterior_check(post, marg1[i])


def test_estep_bodies_same_results_whichever_tier_evaluates_kappa(oracle, dev):
    """A break in the LAST tile only: the first two tiles run the fast tier with alpha = 0.01 and the mid tier with
    alpha = 1; the oracle does not know about tiers, so agreement across alphas pins all of them."""
    N, S = 3, 3 * 4224
    d, gl_ind, freq, e = _case(oracle, N, S, 12, freq=(0.05, 0.5), indF=(0.0, 0.5))
    d.dist_mb[2 * 4224 + 7] = np.inf
    for alpha, want_tiers in [(0.01, [2, 0, 1]), (1.0, [0, 2, 1]), (9.0, [0, 0, 3])]:
        F = np.array([0.05, 0.3, 0.7]); a = np.full(N, alpha)
        st, marg1, lk_o = oracle.estep(e, d.dist_mb, F, a)
        for i in range(N):
            rc, post, lf, lb, tiers = dev.estep(e[i], d.dist_mb, F[i], a[i])
            assert tiers == want_tiers, (alpha, tiers, d.dist_mb[:8448].max())
            assert rc == 0 and abs(lf - lk_o[i]) <= LKL_RTOL * abs(lk_o[i])
            _posterior_check(post, marg1[i])


def test_estep_bodies_long_sequence_against_extended_precision(oracle, dev):
    """60,000 sites: adjudicated by the long-double restatement (the reference's log-space recursion is itself the
    noisy side at this length, SURVEY.md finding 5); log-likelihood still 1e-9 against the reference arithmetic."""
    N, S = 2, 60000
    d, gl_ind, freq, e = _case(oracle, N, S, 13, freq=(0.05, 0.5), indF=(0.0, 0.5))
    F = np.array([0.1, 0.4]); a = np.array([0.02, 0.3])
    for i in range(N):
        rc, post, lf, lb, _ = dev.estep(e[i], d.dist_mb, F[i], a[i])
        m_ext, lk_ext = oracle.estep_extended(e[i], d.dist_mb, F[i], a[i])
        assert rc == 0
        assert abs(lf - lk_ext) <= 1e-11 * abs(lk_ext)
        assert abs(lf - oracle.lkl(e[i], d.dist_mb, F[i], a[i]) * -1) <= LKL_RTOL * abs(lf)
        clamped = np.where(m_ext < 1e-5, 0.0, np.where(m_ext > 1 - 1e-5, 1.0, m_ext))
        _posterior_check(post, clamped)
        inner = (m_ext > 2e-5) & (m_ext < 1 - 2e-5)
        assert np.abs(post - m_ext)[inner].max() < 1e-11                    # far tighter than the reference itself


def test_estep_bodies_flag_nan(oracle, dev):
    """A NaN emission ratio reaches the status word the host maps to the reference's error() (HMM.cpp / EM.cpp:180)."""
    d, gl_ind, freq, e = _case(oracle, 1, 500, 14, freq=(0.05, 0.5), indF=(0.0, 0.5))
    e = e.copy(); e[0, 77, 1] = np.nan
    rc, post, lf, lb, _ = dev.estep(e[0], d.dist_mb, 0.2, 0.1)
    assert rc & 1


def test_objective_body_matches_oracle(oracle, dev):
    """lkl() (EM.cpp:449-464) the way nfh_lkl.cu walks a chunk (apply_site with kappa q, clamp tier)."""
    N, S = 4, 6000
    d, gl_ind, freq, e = _case(oracle, N, S, 51, freq0=0.2, freq=(0.05, 0.5), indF=(0.0, 0.5))
    d.dist_mb[[0, 3000]] = np.inf
    rng = np.random.default_rng(51)
    for i in range(N):
        x = rng.uniform(0.01, 0.9); y = rng.uniform(0.001, 3.0); h = 4.4e-6
        for ff, aa in [(x, y), (x - h, y), (x + h, y), (x, y - h), (x, y + h), (1e-15, 10.0), (1 - 1e-15, 1e-6)]:
            got = dev.neg_lkl(e[i], d.dist_mb, ff, aa)
            want = oracle.lkl(e[i], d.dist_mb, ff, aa)
            assert abs(got - want) <= LKL_RTOL * abs(want)
    assert dev.neg_lkl(e[0], d.dist_mb, np.nan, 0.1) == -1e15              # EM.cpp:454-456


# ---------------------------------------------------------------------------------------------------------------
# frequency EM + emission refresh: make_coef / pass_denominators / pass_sums / reciprocals / emissions
# ---------------------------------------------------------------------------------------------------------------
SHAPES = {0: (8, 13), 1: (4, 16), 2: (16, 8), 3: (32, 16), 4: (1, 16)}


@pytest.mark.parametrize("shape,N,S,seed", [(0, 100, 160, 21), (0, 20, 400, 12345), (1, 6, 300, 22), (1, 64, 120, 23),
                                            (2, 125, 120, 24), (3, 400, 40, 25), (4, 13, 300, 26), (4, 1, 50, 27)])
def test_frequency_pass_bodies_match_est_maf(oracle, dev, shape, N, S, seed):
    """Allele odds, running (den - num), four-way reciprocals, lanes summed by butterfly: the same fixed point as the
    reference's log-space est_maf (gen_func.cpp:974-1009) to 1e-11 after the same number of passes; then the
    emission ratio and e0 of calc_emission (HMM.cpp:144-154) from the same coefficients."""
    d, gl_ind, freq, e = _case(oracle, N, S, seed, freq=(0.02, 0.5), indF=(0.0, 0.5))
    rng = np.random.default_rng(seed)
    post = np.where(rng.random((N, S)) < 0.3, rng.choice([0.0, 1.0], (N, S)), rng.random((N, S)))   # incl. clamped
    n_diff = 0
    for s in range(S):
        f_o, n_o = oracle.est_maf_counted(gl_ind[:, s, :], post[:, s])
        f, n, ratio, e0 = dev.freq_site(shape, gl_ind[:, s, :], post[:, s])
        assert abs(f - f_o) <= 1e-11, (s, f, f_o)
        n_diff += n != n_o
        for i in range(0, N, max(1, N // 7)):
            le0 = oracle.calc_emission(gl_ind[i, s], f_o, 0); le1 = oracle.calc_emission(gl_ind[i, s], f_o, 1)
            assert abs(np.log(e0[i]) - le0) <= 1e-12 + 2e-11 / max(f_o, 1e-3)        # d log e / d f <= ~2 / f
            assert abs(np.log(ratio[i]) - (le1 - le0)) <= 1e-12 + 4e-11 / max(f_o, 1e-3)
    assert n_diff == 0, f"{n_diff} of {S} sites needed a different number of passes"


def test_frequency_zero_posterior_start_estimate(oracle, dev):
    """--freq e: est_maf with F = 0 for everybody (parse_args.cpp:316-318) = a NULL posterior plane."""
    N, S = 10, 200
    d, gl_ind, freq, e = _case(oracle, N, S, 41, freq=(0.05, 0.5), indF=(0.0, 0.5))
    for s in range(S):
        f_o, n_o = oracle.est_maf_counted(gl_ind[:, s, :], np.zeros(N))
        f, n, _, _ = dev.freq_site(0, gl_ind[:, s, :], None)
        assert abs(f - f_o) <= 1e-11 and n == n_o


def test_frequency_hard_calls_and_monomorphic_sites(oracle, dev):
    """Called genotypes (GL exactly 0 / 1): sites fixed for either allele drive the odds to 0 or to the 1e35 cap and
    the frequency to exactly 0 / 1; all-heterozygous sites sit at 0.5; nothing overflows in the four-way reciprocals."""
    rng = np.random.default_rng(5)
    N, S = 12, 96
    geno = rng.integers(0, 3, size=(N, S))
    geno[:, 0] = 0; geno[:, 1] = 2; geno[:, 2] = 1; geno[:, 3] = 2; geno[:6, 4] = 0; geno[6:, 4] = 2
    geno[:, 40:48] = 2; geno[:, 60:64] = 0
    gl = np.full((N, S, 3), -np.inf)
    np.put_along_axis(gl, geno[:, :, None], 0.0, axis=2)
    gl_ind = oracle.normalize_gl(gl)
    post = np.where(rng.random((N, S)) < 0.5, 0.0, rng.random((N, S)))
    post[geno == 1] = 0.0                      # a hard heterozygote has e1 = 0: its posterior is 0 (nfh_freq_math.cuh)
    for s in range(S):
        f_o, n_o = oracle.est_maf_counted(gl_ind[:, s, :], post[:, s])
        for shape in (0, 1):
            f, n, ratio, e0 = dev.freq_site(shape, gl_ind[:, s, :], post[:, s])
            assert np.isfinite(f) and np.isfinite(ratio).all() and np.isfinite(e0).all()
            assert abs(f - f_o) <= 1e-11 and n == n_o, (s, f, f_o, n, n_o)
    f0 = dev.freq_site(0, gl_ind[:, 0, :], post[:, 0])[0]; f1 = dev.freq_site(0, gl_ind[:, 1, :], post[:, 1])[0]
    f2 = dev.freq_site(0, gl_ind[:, 2, :], post[:, 2])[0]
    assert f0 == 0.0 and abs(f1 - 1.0) < 1e-15 and abs(f2 - 0.5) < 1e-15


def test_frequency_streaming_form_matches_est_maf(oracle, dev):
    """freq_emission_stream (any number of individuals): u / v / a form, individuals in the reference's order."""
    N, S = 30, 150
    d, gl_ind, freq, e = _case(oracle, N, S, 61, freq=(0.02, 0.5), indF=(0.0, 0.5))
    post = np.random.default_rng(61).random((N, S))
    for s in range(S):
        f_o, n_o = oracle.est_maf_counted(gl_ind[:, s, :], post[:, s])
        f, n = dev.freq_site_stream(gl_ind[:, s, :], post[:, s])
        assert abs(f - f_o) <= 1e-11 and n == n_o


def test_emission_refresh_only(oracle, dev):
    """--freq_est 0: the frequency stays, ratio and e0 are refreshed from it."""
    N, S = 9, 120
    d, gl_ind, _, _ = _case(oracle, N, S, 71, freq=(0.05, 0.5), indF=(0.0, 0.5))
    fr = np.random.default_rng(71).uniform(0.01, 0.99, S)
    for s in range(S):
        f, n, ratio, e0 = dev.freq_site(1, gl_ind[:, s, :], None, freq=fr[s])
        assert f == fr[s] and n == 0
        for i in range(N):
            le0 = oracle.calc_emission(gl_ind[i, s], fr[s], 0); le1 = oracle.calc_emission(gl_ind[i, s], fr[s], 1)
            assert abs(np.log(e0[i]) - le0) <= 1e-13 and abs(np.log(ratio[i]) - (le1 - le0)) <= 1e-12


# ---------------------------------------------------------------------------------------------------------------
# Viterbi: site_q / trop_apply / map_compose / vit_chunk_trace + the two restated per-site loops
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,S,seed", [(6, 3000, 61), (8, 10000, 12345), (4, 9, 62), (3, 33, 63), (3, 34, 64),
                                      (2, 4225, 65)])
def test_viterbi_bodies_decode_the_reference_path(oracle, dev, N, S, seed):
    """(max, x) semiring in linear space with the in-place quirk (state 1 competes against the state-0 score that
    already carries this site's emission and transition) and strict comparisons: the same tracts as the reference's
    log-space viterbi() (HMM.cpp:98-125), site for site."""
    d, gl_ind, freq, e = _case(oracle, N, S, seed, freq0=0.15, freq=(0.05, 0.5), indF=(0.0, 0.5))
    rng = np.random.default_rng(seed)
    F = rng.uniform(0.02, 0.6, N); a = rng.uniform(0.005, 1.0, N)
    mism = 0
    for i in range(N):
        _, want = oracle.viterbi(e[i], d.dist_mb, F[i], a[i])
        got = dev.viterbi(e[i], d.dist_mb, F[i], a[i])
        mism += int((got != want).sum())
    assert mism == 0


def test_viterbi_bodies_chromosome_breaks_hard_calls_and_bounds(oracle, dev):
    """Chromosome starts (the chain restarts from q: t01 = q1), F / alpha on their bounds, and called genotypes, whose
    hard heterozygotes give e1 = 0 exactly (score ties at zero must keep the k = 0 predecessor, HMM.cpp:112)."""
    N, S = 5, 6000
    d, gl_ind, freq, e = _case(oracle, N, S, 66, freq0=0.2, freq=(0.05, 0.5), indF=(0.0, 0.5))
    d.dist_mb[[0, 777, 3000, 5999]] = np.inf
    F = np.array([1e-6, 1 - 1e-6, 0.3, 0.05, 0.5]); a = np.array([1e-3, 10.0, 0.05, 0.5, 2.0])
    for i in range(N):
        _, want = oracle.viterbi(e[i], d.dist_mb, F[i], a[i])
        assert (dev.viterbi(e[i], d.dist_mb, F[i], a[i]) != want).sum() == 0
    rng = np.random.default_rng(66)
    geno = rng.integers(0, 3, size=(3, 2000))
    gl = np.full((3, 2000, 3), -np.inf)
    np.put_along_axis(gl, geno[:, :, None], 0.0, axis=2)
    gl_ind = oracle.normalize_gl(gl)
    _, e = oracle.freq_emission(gl_ind, None, np.full(2000, 0.3), update_freq=False)
    dist = d.dist_mb[:2000].copy()
    for i in range(3):
        _, want = oracle.viterbi(e[i], dist, 0.3, 0.2)
        got = dev.viterbi(e[i], dist, 0.3, 0.2)
        assert (got != want).sum() == 0
        assert (got[geno[i] == 1] == 0).all()                               # a heterozygote is never IBD


# ---------------------------------------------------------------------------------------------------------------
# What the harness restates instead of including must still be what the kernels say
# ---------------------------------------------------------------------------------------------------------------
def _squash(text):
    text = re.sub(r"//[^\n]*", "", text)
    return re.sub(r"\s+", "", text)


def test_restated_scalar_update_is_the_kernels_text():
    """The six statements after the lane reduction of a frequency pass live inline in the kernels (as a function they
    changed the register allocation of the widest instantiations); device_arith_host.cpp restates them.  Pin the
    restatement to the kernel source: same statements, same order, in every register kernel."""
    kernel = _squash(open(os.path.join(CSRC, "nfh_freq.cu")).read())
    harness = _squash(open(os.path.join(ROOT, "tests", "device_arith_host.cpp")).read())
    update = _squash("""num = fma(odds, X, num);
                        const double dmn = fmax(fma(odds, Z, dmn_next), num * kMinOddsInv);
                        odds = num * rcp_pos(dmn);""")
    assert kernel.count(update) == 3                                        # warp, hybrid, team
    assert harness.count(update.replace("X,", "X[0],").replace("Z,", "Z[0],")) == 1
    for stmt in ("dmn_next = dmn + g_sum;", "num * rcp_pos<true>(num + dmn);", "> kEps) && (passes <= 100);",
                 "double num = 0.0, dmn_next = g_sum"):
        assert kernel.count(_squash(stmt)) >= 3, stmt
    for stmt in ("dmn_next = dmn + gs;", "const double now = num * rcp_pos<true>(num + dmn);",
                 "active = active && (fabs(prev - now) > kEps) && (passes <= 100);",
                 "double num = 0.0, dmn_next = gs;", "double odds = kStartOdds, prev = kStartFreq;"):
        assert harness.count(_squash(stmt)) == 1, stmt
    assert kernel.count(_squash("odds = kStartOdds, prev = kStartFreq;")) == 3


def _loop_after(text, marker):
    """Squashed text of the first `for (...) {...}` block after `marker`."""
    i = text.index(marker)
    i = text.index("for (int j", i)
    depth, k = 0, text.index("{", i)
    for k in range(k, len(text)):
        depth += text[k] == "{"
        depth -= text[k] == "}"
        if depth == 0:
            break
    return _squash(text[i:k + 1])


def test_restated_viterbi_loops_are_the_kernels_text():
    """viterbi_chunk_products' and viterbi_chunk_pointers' per-site loops stay inline in the kernels (as functions they
    compiled to different code); device_arith_host.cpp restates them.  Same text, comments and the shared-memory
    prefix of the table aside."""
    kernel = open(os.path.join(CSRC, "nfh_viterbi.cu")).read().replace("sm.tab", "tab")
    harness = open(os.path.join(ROOT, "tests", "device_arith_host.cpp")).read()
    assert _loop_after(kernel, "viterbi_chunk_products(ViterbiArgs A)") == _loop_after(harness, "// ---- viterbi_chunk_products")
    assert _loop_after(kernel, "unsigned chunk_map = 2u;") == _loop_after(harness, "// ---- viterbi_chunk_pointers")


def test_kernels_call_the_bodies_the_harness_compiles():
    """The headers hold the bodies; the kernels must be their only users' counterpart (no second copy drifting)."""
    estep = open(os.path.join(CSRC, "nfh_estep.cu")).read()
    freq = open(os.path.join(CSRC, "nfh_freq.cu")).read()
    vit = open(os.path.join(CSRC, "nfh_viterbi.cu")).read()
    assert '#include "nfh_estep_math.cuh"' in estep and '#include "nfh_freq_math.cuh"' in freq
    assert '#include "nfh_viterbi_math.cuh"' in vit and "vit_chunk_trace(" in vit and "NFH_DEV" not in vit
    for name in ("products_chunk<kTierFast>", "products_chunk<kTierMid>", "products_chunk<kTierSlow>",
                 "apply_chunk<kTierFast>", "apply_chunk<kTierMid>", "apply_chunk<kTierSlow>"):
        assert name in estep, name
    for name in ("make_coef(", "pass_denominators<", "pass_sums<", "emissions("):
        assert name in freq, name
    for name in ("NFH_DEV void products_chunk", "NFH_DEV bool apply_chunk", "NFH_DEV IndCoef make_coef",
                 "NFH_DEV void pass_sums", "NFH_DEV void reciprocals"):
        assert name not in estep and name not in freq                        # defined once, in the headers


def test_product_sources_do_not_reach_the_harness():
    """No CPU fallback: nothing the product builds or loads mentions the host build of its arithmetic."""
    for base, _, files in os.walk(os.path.join(ROOT, "ngsf-hmm_b200")):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".h", ".cu", ".cuh", "Makefile")):
                text = open(os.path.join(base, f), errors="replace").read()
                code = re.sub(r"//[^\n]*|#error[^\n]*", "", text)
                assert "device_arith_host" not in code, os.path.join(base, f)
                assert "nfh_host_rcp_seed" not in code or f == "nfh_math.cuh", os.path.join(base, f)
