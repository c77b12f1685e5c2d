"""CPU: whole EM iterations and whole EM runs of the PRODUCT - its host library compiled as it is (host/lbfgsb.cpp,
bfgs_driver.cpp, host_api.cpp) over a fake device that computes with the KERNELS' OWN ARITHMETIC
(tests/fake_device_arith.cpp on tests/device_arith_host.cpp: the per-thread bodies of the CUDA kernels compiled for
the host) and over one that runs the KERNELS THEMSELVES under the SIMT emulator (tests/fake_device_simt.cpp on
tests/simt_kernels_host.cpp: global functions, launchers, launch geometry) - against the golden fixtures generated from the unmodified reference (tests/golden/make_golden*.py) and
against the reference itself (oracle/_ref), at the north star's tolerances: log-likelihood 1e-9 relative, posterior
1e-8 absolute, F / alpha / frequencies 1e-6 after full EM, identical Viterbi tracts.

These are the assertions of tests/test_gpu_golden.py, word for word, on a machine without a GPU.  What the GPU run of
the same file adds is the parallel decomposition of the kernels and the hardware's reciprocal seed.  The build is
test infrastructure: the product has no CPU path and nothing under ngsf-hmm_b200/ knows about this one.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

import ngsf_hmm_b200  # noqa: F401
from ngsf_hmm_b200 import sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(ROOT, "ngsf-hmm_b200", "host")
CSRC = os.path.join(ROOT, "ngsf-hmm_b200", "csrc")
dp = C.POINTER(C.c_double)
ITER_CASES = sorted(glob.glob(os.path.join(HERE, "golden", "iter_*.npz")))


@pytest.fixture(scope="module", params=["kernel-arithmetic", "kernels-under-emulator"])
def lib(request, tmp_path_factory):
    """The product's host library over (a) tests/fake_device_arith.cpp: the kernels' per-thread arithmetic strung
    together sequentially, or (b) tests/fake_device_simt.cpp: the kernels themselves under the SIMT emulator."""
    d = str(tmp_path_factory.mktemp("host_on_device_arith"))
    inc = ["-I", os.path.join(ROOT, "include"), "-I", HOST, "-I", CSRC]
    jobs = []                                                                  # compiled side by side
    for src in ("lbfgsb.cpp", "bfgs_driver.cpp", "host_api.cpp"):              # the flags of host/Makefile
        o = os.path.join(d, src + ".o")
        jobs.append((subprocess.Popen(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off"] + inc +
                                      ["-c", os.path.join(HOST, src), "-o", o]), o))
    o = os.path.join(d, "device_arith_host.cpp.o")
    jobs.append((subprocess.Popen(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas"] + inc +
                                  ["-c", os.path.join(ROOT, "tests", "device_arith_host.cpp"), "-o", o]), o))
    objs = []
    for pr, o in jobs:
        assert pr.wait() == 0, o
        objs.append(o)
    if request.param == "kernel-arithmetic":
        o = os.path.join(d, "fake_device_arith.cpp.o")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off"] + inc +
                              ["-c", os.path.join(ROOT, "tests", "fake_device_arith.cpp"), "-o", o])
        so = os.path.join(d, "libhost_on_device_arith.so")
        subprocess.check_call(["g++", "-shared", "-o", so] + objs + [o, "-Wl,--no-undefined"])
    else:
        import _simt_build
        so = _simt_build.build(os.path.join(d, "simt"), extra_sources=[os.path.join(ROOT, "tests", "fake_device_simt.cpp")],
                               extra_objects=objs, name="libhost_on_simt_kernels.so")
    L = C.CDLL(so)
    L.kind = request.param
    L.fake_ctx_create.restype = C.c_void_p
    L.fake_ctx_create.argtypes = [C.c_uint64, C.c_uint64, dp, dp, dp] + ([C.c_int] if L.kind != "kernel-arithmetic" else [])
    L.fake_ctx_destroy.argtypes = [C.c_void_p]
    L.fake_ctx_get.argtypes = [C.c_void_p, dp, dp]
    L.fake_ctx_set_freq.argtypes = [C.c_void_p, dp]
    L.fake_ctx_viterbi.argtypes = [C.c_void_p, C.c_void_p]
    L.nfh_set_ind_params.argtypes = [C.c_void_p, dp, dp]
    L.nfh_host_em_iteration.restype = C.c_int
    L.nfh_host_em_iteration.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_uint64)]
    return L


def _p(a):
    return a.ctypes.data_as(dp)


class FakeRun:
    """The calls tests/test_gpu_golden.py makes on Context / EmRank / run_em, on the fake device."""

    def __init__(self, lib, gl_norm, dist, freq, freq_est=1, freq_kernels=False):
        """freq_kernels (emulator only): the frequency EM runs on the emulated kernels too - minutes per thousand
        sites, so only the one-iteration fixtures ask for it; else on the kernels' arithmetic compiled for the host."""
        self.L = lib
        self.N, self.S, _ = gl_norm.shape
        gl_site = np.ascontiguousarray(np.transpose(gl_norm, (1, 0, 2)), dtype=np.float64)
        self.dist = np.ascontiguousarray(dist, dtype=np.float64)
        f = np.broadcast_to(np.asarray(freq, dtype=np.float64), (self.S,)).copy()
        extra = [int(freq_kernels)] if lib.kind != "kernel-arithmetic" else []
        self.h = lib.fake_ctx_create(self.N, self.S, _p(gl_site), _p(self.dist), _p(f), *extra)
        assert self.h
        self.freq_est = freq_est

    def close(self):
        self.L.fake_ctx_destroy(self.h)

    def iteration(self, F, a):
        lk, fr = np.empty(self.N), np.empty(self.S)
        stats = (C.c_uint64 * 3)()
        rc = self.L.nfh_host_em_iteration(self.h, _p(F), _p(a), 0, 0, self.freq_est, _p(lk), _p(fr), stats)
        assert rc == 0, rc
        return lk, fr

    def posterior(self):
        m = np.empty((self.N, self.S))
        self.L.fake_ctx_get(self.h, _p(m), None)
        return m

    def viterbi(self, F, a, freq=None):
        if freq is not None:
            self.L.fake_ctx_set_freq(self.h, _p(np.ascontiguousarray(freq, dtype=np.float64)))
        self.L.nfh_set_ind_params(self.h, _p(np.ascontiguousarray(F)), _p(np.ascontiguousarray(a)))
        path = np.zeros((self.N, self.S), dtype=np.uint8)
        self.L.fake_ctx_viterbi(self.h, path.ctypes.data)
        return path.astype(np.int8)

    def run_em(self, F, a, min_iters=10, max_iters=100, min_epsilon=1e-5):
        """EM() (EM.cpp:27-135) as ngsf_hmm_b200.em.run_em runs it."""
        it, prev_tot, tot, max_eps = 0, 0.0, 0.0, -np.inf
        prev_ind = np.full(self.N, -np.inf)
        fr = None
        while ((prev_tot - tot > min_epsilon) or (max_eps > min_epsilon) or it < min_iters) and it < max_iters:
            it += 1
            lk, fr = self.iteration(F, a)
            prev_tot, tot = tot, 0.0
            for v in lk:
                tot += float(v)
            with np.errstate(invalid="ignore", divide="ignore"):
                eps = (lk - prev_ind) / np.abs(prev_ind)
            best, max_eps = 0, -np.inf
            for i, e in enumerate(eps):
                if e > max_eps:
                    best, max_eps = i, e
            max_eps = eps[best]
            prev_ind = lk.copy()
        return dict(iterations=it, tot_lkl=tot, ind_lkl=lk, freq=fr.copy(), path=self.viterbi(F, a))


def _post_ok(got, want):
    diff = np.abs(got - want)
    bad = diff > 1e-8
    flips = bad & ((want == 0) | (want == 1) | (got == 0) | (got == 1)) & (diff < 1.1e-5)
    return not (bad & ~flips).any() and flips.sum() <= max(2, got.size // 20000)


@pytest.mark.parametrize("path", ITER_CASES, ids=[os.path.basename(p) for p in ITER_CASES])
def test_one_em_iteration_matches_reference_fixture(lib, path):
    g = {k: v for k, v in np.load(path).items()}
    S, N, _ = g["log_gl"].shape
    F = np.full(N, float(g["F0"])); a = np.full(N, float(g["a0"]))
    run = FakeRun(lib, g["gl_norm"], g["dist_mb"], float(g["freq0"]), freq_kernels=True)
    try:
        lk, fr = run.iteration(F, a)
        np.testing.assert_allclose(lk, g["ind_lkl"], rtol=1e-9, atol=0)
        assert _post_ok(run.posterior(), g["marg1"])
        np.testing.assert_allclose(fr, g["freq"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(F, g["indF"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(a, g["alpha"], rtol=1e-6, atol=1e-6)
        assert (run.viterbi(g["indF"], g["alpha"], freq=g["freq"]) == g["path"]).all()
    finally:
        run.close()


def test_full_em_matches_reference_fixture(lib, oracle):
    g = {k: v for k, v in np.load(os.path.join(HERE, "golden", "em_full.npz")).items()}
    S, N, _ = g["log_gl"].shape
    gl = oracle.normalize_gl(np.transpose(g["log_gl"], (1, 0, 2)))
    run = FakeRun(lib, gl, g["dist_mb"], 0.1)
    try:
        F = np.full(N, 0.1); a = np.full(N, 0.2)
        out = run.run_em(F, a, min_iters=int(g["min_iters"]), max_iters=int(g["max_iters"]))
        np.testing.assert_allclose(out["tot_lkl"], float(g["tot_lkl"]), rtol=1e-9)
        np.testing.assert_allclose(F, g["indF"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(a, g["alpha"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(out["freq"], g["freq"], rtol=0, atol=1e-6)
        assert _post_ok(run.posterior(), g["marg1"])
        assert (out["path"] != g["path"]).sum() == 0
    finally:
        run.close()


def test_full_em_config1_shape_with_adjudication(lib, oracle):
    """BASELINE configs[0] shape (20 individuals x 10,000 sites, --freq_est 1, --freq 0.1 --indF 0.1,0.2, EM to
    convergence) with the rule of tests/test_gpu_golden.py: alpha and frequencies 1e-6 against the reference outright;
    at least 16 of 20 F within 1e-6 of the reference, every exception within 1e-6 of the EM rerun with a long-double
    objective AND at a long-double likelihood no lower than the reference's."""
    if lib.kind == "kernel-arithmetic":
        pytest.skip("run once, on the stronger of the two fake devices (the kernels under the emulator); the sequential "
                    "arithmetic gives the same 16 / 20 split in 17 s")
    g = {k: v for k, v in np.load(os.path.join(HERE, "golden", "em_cfg1_adjudication.npz")).items()}
    N, S = int(g["n_ind"]), int(g["n_sites"])
    d = sim.simulate(N, S, seed=int(g["seed"]), freq=0.2, indF=0.5, alpha=0.01, depth=2.0)
    gl = oracle.normalize_gl(np.transpose(d.log_gl, (1, 0, 2)))
    run = FakeRun(lib, gl, d.dist_mb, 0.1)
    try:
        F = np.full(N, 0.1); a = np.full(N, 0.2)
        out = run.run_em(F, a, min_iters=10, max_iters=100)
    finally:
        run.close()
    assert out["iterations"] == int(g["iters_ext"])
    np.testing.assert_allclose(out["tot_lkl"], float(g["tot_ref"]), rtol=1e-9)
    np.testing.assert_allclose(a, g["a_ref"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["freq"], g["freq_ref"], rtol=0, atol=1e-6)
    near_ref = np.abs(F - g["F_ref"]) <= 1e-6
    near_ext = np.abs(F - g["F_ext"]) <= 1e-6
    print(f"F: {near_ref.sum()}/{N} within 1e-6 of the reference, {near_ext.sum()}/{N} of the extended-precision EM; "
          f"max |F - F_ref| = {np.abs(F - g['F_ref']).max():.2e}")
    assert near_ref.sum() >= 16, (F - g["F_ref"])
    _, e = oracle.freq_emission(gl, None, g["freq_ref"], update_freq=False)
    for i in np.nonzero(~near_ref)[0]:
        assert near_ext[i], (i, F[i], g["F_ref"][i], g["F_ext"][i])
        lk_ours = oracle.estep_extended(e[i], d.dist_mb, F[i], a[i])[1]
        lk_ref = oracle.estep_extended(e[i], d.dist_mb, g["F_ref"][i], g["a_ref"][i])[1]
        assert lk_ours >= lk_ref - 1e-9 * abs(lk_ref), (i, lk_ours, lk_ref)
    p_ref = np.unpackbits(g["path_ref"], axis=1)[:, :S]; p_ext = np.unpackbits(g["path_ext"], axis=1)[:, :S]
    for i in range(N):
        assert (out["path"][i] == p_ref[i]).all() or (out["path"][i] == p_ext[i]).all()


@pytest.mark.ref
def test_em_iterations_track_the_reference_in_process(lib, oracle, ref):
    """Four iterations side by side with the unmodified reference (oracle/_ref, in-process iter_EM) on a case with a
    chromosome break: every iteration's likelihoods to 1e-9, parameters and frequencies to 1e-6, posteriors to 1e-8
    while both sides still hold the same parameters (first iteration)."""
    N, S, ITERS = 6, 1500, 4
    d = sim.simulate(N, S, seed=808, freq=(0.05, 0.5), indF=(0.05, 0.6), alpha=0.02, depth=3.0)
    d.dist_mb[S // 3] = np.inf
    F0 = np.linspace(0.05, 0.4, N); a0 = np.linspace(0.1, 0.9, N); f0 = np.full(S, 0.15)
    st = ref.state(d.log_gl, d.dist_mb, f0, F0, a0, freq_est=1, indF_fixed=False, alpha_fixed=False, n_threads=2)
    gl = oracle.normalize_gl(np.transpose(d.log_gl, (1, 0, 2)))
    run = FakeRun(lib, gl, d.dist_mb, f0)
    try:
        F, a = F0.copy(), a0.copy()
        for it in range(ITERS):
            st.iter_EM()
            w = st.get()
            lk, fr = run.iteration(F, a)
            np.testing.assert_allclose(lk, w["ind_lkl"], rtol=1e-9, atol=0, err_msg=f"iteration {it + 1}")
            if it == 0:
                assert _post_ok(run.posterior(), w["marg1"])                # same parameters on both sides
            else:       # parameters agree to 1e-6 only from here on; the posterior follows them (clamp flips aside)
                diff = np.abs(run.posterior() - w["marg1"])
                assert np.quantile(diff, 0.999) < 1e-6 and diff.max() < 1.1e-5, (it, diff.max())
            np.testing.assert_allclose(fr, w["freq"], rtol=0, atol=1e-6)
            np.testing.assert_allclose(F, w["indF"], rtol=0, atol=1e-6)
            np.testing.assert_allclose(a, w["alpha"], rtol=1e-6, atol=1e-6)
    finally:
        st.close()
        run.close()
