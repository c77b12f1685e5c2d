"""CPU: the drop-in binary's output writer (host/cli/report.cpp) against the reference's print_iter
(EM.cpp:293-380 through oracle/_ref): <out>.indF and <out>.ibd BYTE FOR BYTE on states full of awkward values -
F within 1e-5 of 0 and 1 (alpha printed as NA), exact 0 / 1 posteriors, decimal ties of %f and %.5f, tiny and
large log-likelihoods - written with one and with several host threads.  (<out>.geno needs the device: GPU tests.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.ref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ngsf-hmm_b200", "host", "cli")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("report") / "cli_report_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           "-I", CLI, "-o", exe, os.path.join(ROOT, "tests", "cli_report_check.cpp"),
                           os.path.join(CLI, "options.cpp"), os.path.join(CLI, "report.cpp"), "-lz", "-lpthread"])
    return exe


def _awkward_unit_values(rng, n):
    v = rng.random(n)
    r = rng.random(n)
    v[r < 0.10] = 0.0
    v[(r >= 0.10) & (r < 0.20)] = 1.0
    k = (r >= 0.20) & (r < 0.35)                       # exact binary ties of the sixth decimal: (2j+1)/2^7 * 1e-?..
    v[k] = rng.integers(0, 128, k.sum()) / 128.0 + 0.0
    k = (r >= 0.35) & (r < 0.45)                       # decimal half-way points k*1e-6 + 5e-7 and their neighbours
    base = rng.integers(0, 10 ** 6, k.sum()) * 1e-6 + 5e-7
    v[k] = np.nextafter(base, rng.choice([0.0, 1.0], k.sum()))
    k = (r >= 0.45) & (r < 0.50)
    v[k] = rng.choice([1e-5, 1 - 1e-5, 9.99999e-6, 1e-7, 4.9999999e-7, 5e-7, 0.9999995, 0.99999949], k.sum())
    return v


@pytest.mark.parametrize("threads", [1, 5])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_indF_and_ibd_files_byte_for_byte(ref, checker, tmp_path, seed, threads):
    rng = np.random.default_rng(seed)
    N, S = 7, 1500
    tmp = str(tmp_path)
    F = _awkward_unit_values(rng, N); F[0] = 5e-6; F[1] = 1 - 5e-6; F[2] = 1e-5; F[3] = 0.123455
    alpha = rng.choice([1e-6, 0.0123455, 0.2, 9.9999995, 10.0], N)
    freq = np.clip(_awkward_unit_values(rng, S), 0.0, 0.5)
    ind_lkl = -rng.random(N) * rng.choice([1e-3, 1.0, 1e4, 1e7], N)
    tot = float(ind_lkl.sum())
    marg = _awkward_unit_values(rng, N * S).reshape(N, S)
    path = rng.integers(0, 2, size=(N, S)).astype(np.int8)
    gl = np.full((S, N, 3), np.log(1 / 3))
    for name, arr in (("tot", np.array([tot])), ("indF", F), ("alpha", alpha), ("ind_lkl", ind_lkl), ("freq", freq),
                      ("marg", marg), ("path", path)):
        np.ascontiguousarray(arr).tofile(os.path.join(tmp, "in." + name))
    dp = C.POINTER(C.c_double)
    fn = ref.lib.ref_print_iter
    fn.restype = None
    fn.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_double, dp, dp, dp, dp, C.c_char_p, dp, dp]
    keep = [np.ascontiguousarray(x, dtype=np.float64) for x in (F, alpha, freq, ind_lkl, marg, gl)]
    fn(os.path.join(tmp, "ref").encode(), N, S, tot, keep[0].ctypes.data_as(dp), keep[1].ctypes.data_as(dp),
       keep[2].ctypes.data_as(dp), keep[3].ctypes.data_as(dp), path.tobytes(), keep[4].ctypes.data_as(dp),
       keep[5].ctypes.data_as(dp))
    p = subprocess.run([checker, os.path.join(tmp, "in"), os.path.join(tmp, "ours"), str(N), str(S), str(threads)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    for ext in (".indF", ".ibd"):
        a = open(os.path.join(tmp, "ref" + ext), "rb").read()
        b = open(os.path.join(tmp, "ours" + ext), "rb").read()
        assert len(a) > 0 and a == b, ext
    assert b"NA" in open(os.path.join(tmp, "ours.indF"), "rb").read()
