"""CPU: the PRODUCT's host library (host/lbfgsb.cpp, bfgs_driver.cpp, host_api.cpp, compiled as they are) driving a
fake device whose arithmetic is the oracle's (tests/fake_device_oracle.c; the oracle equals the reference bit for
bit).  nfh_host_em_iteration must then reproduce the reference's iter_EM (EM.cpp:139-289) TO THE LAST BIT over
several EM iterations - F, alpha, allele frequencies, per-individual likelihoods, posteriors, emissions - for free
parameters, either one fixed, both fixed, and fixed frequencies.  This pins everything the host side contributes:
the lockstep L-BFGS-B (which individual asks for which points in which round, who has stopped), the E-step riding
on the first round, and the order E-step -> F/alpha update on the OLD emissions -> frequency update with the NEW
posterior.  What the GPU tests then have to show is only that the kernels compute these functions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ngsf_hmm_b200  # noqa: F401
from ngsf_hmm_b200 import sim

pytestmark = pytest.mark.ref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ngsf-hmm_b200", "host")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("host_on_oracle"))
    inc = ["-I", os.path.join(ROOT, "include"), "-I", HOST, "-I", os.path.join(ROOT, "oracle")]
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    objs = []
    for src in ("lbfgsb.cpp", "bfgs_driver.cpp", "host_api.cpp"):
        o = os.path.join(d, src + ".o")
        # the flags of host/Makefile
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off"] + inc +
                              ["-c", os.path.join(HOST, src), "-o", o])
        objs.append(o)
    o = os.path.join(d, "fake_device_oracle.o")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-ffp-contract=off"] + inc +
                          ["-c", os.path.join(ROOT, "tests", "fake_device_oracle.c"), "-o", o])
    so = os.path.join(d, "libhost_on_oracle.so")
    subprocess.check_call(["g++", "-shared", "-o", so] + objs + [o, "-L", os.path.join(ROOT, "oracle"), "-loracle",
                                                                  "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
                                                                  "-Wl,--no-undefined"])
    L = C.CDLL(so)
    L.fake_ctx_create.restype = C.c_void_p
    L.fake_ctx_create.argtypes = [C.c_uint64, C.c_uint64, dp, dp, dp]
    L.fake_ctx_destroy.argtypes = [C.c_void_p]
    L.fake_ctx_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.fake_ctx_get.argtypes = [C.c_void_p, dp, dp]
    L.nfh_host_em_iteration.restype = C.c_int
    L.nfh_host_em_iteration.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_uint64)]
    return L


def _p(a):
    return a.ctypes.data_as(dp)


CASES = {
    "free": dict(F_fixed=False, a_fixed=False, freq_est=1),
    "F_fixed": dict(F_fixed=True, a_fixed=False, freq_est=1),
    "alpha_fixed": dict(F_fixed=False, a_fixed=True, freq_est=1),
    "both_fixed": dict(F_fixed=True, a_fixed=True, freq_est=1),
    "freq_fixed": dict(F_fixed=False, a_fixed=False, freq_est=0),
    "all_fixed": dict(F_fixed=True, a_fixed=True, freq_est=0),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("shape", ["inbred", "outbred_on_the_lower_bound", "alpha_starts_above_the_box"])
def test_host_iteration_equals_reference_iter_EM(ref, lib, case, shape):
    cfg = CASES[case]
    N, S, ITERS = 6, 1500, 4
    if shape != "outbred_on_the_lower_bound":
        d = sim.simulate(N, S, seed=808, freq=(0.05, 0.5), indF=(0.05, 0.6), alpha=0.02, depth=3.0)
    else:   # F = 0 for everyone: the optimum of F sits on its lower bound and variables enter / leave the free set
        d = sim.simulate(N, S, seed=809, freq=(0.05, 0.5), indF=(0.0, 0.0), alpha=0.02, depth=1.0)
    d.dist_mb[S // 3] = np.inf
    F0 = np.linspace(0.05, 0.4, N); a0 = np.linspace(0.1, 0.9, N); f0 = np.full(S, 0.15)
    if shape == "alpha_starts_above_the_box":
        # alpha = 12 > 10: the reference runs forward / backward at 12 and only the optimiser projects to 10
        # (EM.cpp:151-185 vs setulb_); the E-step then cannot ride on the first round (bfgs_driver.cpp)
        a0[2] = 12.0

    st = ref.state(d.log_gl, d.dist_mb, f0, F0, a0, freq_est=cfg["freq_est"], indF_fixed=cfg["F_fixed"],
                   alpha_fixed=cfg["a_fixed"], n_threads=2)
    want = []
    for _ in range(ITERS):
        st.iter_EM()
        want.append(st.get())
    st.close()

    gl = np.ascontiguousarray(d.log_gl, dtype=np.float64)
    dist = np.ascontiguousarray(d.dist_mb, dtype=np.float64)
    ctx = lib.fake_ctx_create(N, S, _p(gl), _p(dist), _p(f0.copy()))
    try:
        F = F0.copy(); a = a0.copy()
        lk = np.empty(N); fr = f0.copy()
        stats = (C.c_uint64 * 3)()
        marg = np.empty((N, S)); e = np.empty((N, S, 2))
        for it in range(ITERS):
            rc = lib.nfh_host_em_iteration(ctx, _p(F), _p(a), int(cfg["F_fixed"]), int(cfg["a_fixed"]), cfg["freq_est"],
                                           _p(lk), _p(fr), stats)
            assert rc == 0
            w = want[it]
            np.testing.assert_array_equal(lk, w["ind_lkl"], err_msg=f"ind_lkl, iteration {it + 1}")
            np.testing.assert_array_equal(F, w["indF"], err_msg=f"indF, iteration {it + 1}")
            np.testing.assert_array_equal(a, w["alpha"], err_msg=f"alpha, iteration {it + 1}")
            if cfg["freq_est"]:
                np.testing.assert_array_equal(fr, w["freq"], err_msg=f"freq, iteration {it + 1}")
            lib.fake_ctx_get(ctx, _p(marg), _p(e))
            np.testing.assert_array_equal(marg, w["marg1"])
            np.testing.assert_array_equal(e, w["e_prob"])
        counts = (C.c_uint64 * 3)()
        lib.fake_ctx_counts(ctx, counts)
        assert counts[0] == ITERS                                  # one E-step per iteration
        assert counts[2] == (ITERS if cfg["freq_est"] else 0)
        if cfg["F_fixed"] and cfg["a_fixed"]:
            assert counts[1] == 0                                  # no objective evaluation at all
        else:
            assert stats[0] >= 1 and counts[1] >= ITERS
    finally:
        lib.fake_ctx_destroy(ctx)
