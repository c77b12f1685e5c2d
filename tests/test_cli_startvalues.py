"""CPU: the drop-in binary's start values (host/cli/startvalues.cpp, device calls stubbed) against the reference's
init_output (parse_args.cpp:229-419 through oracle/_ref): --indF / --freq given as numbers, as files and as "r"
(random, several seeds), with the reference's clamps, separators and file-format errors."""
import ctypes as C
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.ref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ngsf-hmm_b200", "host", "cli")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sv") / "cli_startvalues_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           "-I", CLI, "-o", exe, os.path.join(ROOT, "tests", "cli_startvalues_check.cpp"),
                           os.path.join(CLI, "options.cpp"), os.path.join(CLI, "startvalues.cpp"), "-lz", "-lpthread"])
    return exe


def _ours(checker, tmp, indF_arg, freq_arg, seed, N, S, freq_est=1):
    out = os.path.join(tmp, "ours")
    p = subprocess.run([checker, indF_arg, freq_arg, str(seed), str(N), str(S), str(freq_est), out],
                       capture_output=True, text=True, timeout=120)
    err = [ln for ln in p.stderr.splitlines() if ln.startswith("ERROR:")]
    if p.returncode != 0:
        return p.returncode, err, None
    return 0, err, tuple(np.fromfile(out + ext) for ext in (".indF", ".alpha", ".freq"))


_REF_CODE = """
import sys, ctypes as C
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np
from _oracle import Ref
lib = Ref().lib
dp = C.POINTER(C.c_double)
lib.ref_init_start_values.restype = None
lib.ref_init_start_values.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_uint64, C.c_uint64, C.c_int, dp, dp, dp]
N, S = {N}, {S}
F = np.empty(N); a = np.empty(N); fr = np.empty(S)
lib.ref_init_start_values({indF!r}.encode(), {freq!r}.encode(), {seed}, N, S, {freq_est}, F.ctypes.data_as(dp),
                          a.ctypes.data_as(dp), fr.ctypes.data_as(dp))
F.tofile({out!r} + ".indF"); a.tofile({out!r} + ".alpha"); fr.tofile({out!r} + ".freq")
"""


def _ref(tmp, indF_arg, freq_arg, seed, N, S, freq_est=1):
    """In a child process: the reference leaves through exit(-1) on a format error."""
    out = os.path.join(tmp, "ref")
    code = _REF_CODE.format(root=ROOT, tests=os.path.join(ROOT, "tests"), N=N, S=S, indF=indF_arg, freq=freq_arg,
                            seed=seed, freq_est=freq_est, out=out)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    err = [ln for ln in p.stderr.splitlines() if ln.startswith("ERROR:")]
    if p.returncode != 0:
        return p.returncode, err, None
    return 0, err, tuple(np.fromfile(out + ext) for ext in (".indF", ".alpha", ".freq"))


def _same(checker, tmp, *args, **kw):
    rc_o, err_o, got = _ours(checker, tmp, *args, **kw)
    rc_r, err_r, want = _ref(tmp, *args, **kw)
    assert (rc_o == 0) == (rc_r == 0), (rc_o, rc_r, err_o, err_r)
    assert err_o == err_r
    if rc_r == 0:
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)
    return rc_r


@pytest.mark.parametrize("indF_arg,freq_arg", [
    ("0.1,0.2", "0.1"), ("0.1-0.2", "0.3"), ("0.01-0.001", "0.05"),       # the default --indF is "0.01-0.001"
    ("0,5", "0.7"), ("1.5,1e3", "0.001"), ("2e-7,0.5", "-3"),              # clamps: [1e-6, 1-1e-6] and [0.01, 0.49]
    ("0.3,0.4", "abc"), ("0.3,0.4", "0.2x"),                               # atof() of a non-number
    ("0.3", "0.1"), ("0.1,0.2,0.3", "0.1"), ("x,0.2", "0.1"), ("1e-3,0.2", "0.1"),   # != 2 numbers -> error
])
def test_numbers_on_the_command_line(checker, tmp_path, indF_arg, freq_arg):
    _same(checker, str(tmp_path), indF_arg, freq_arg, 1, 5, 9)


@pytest.mark.parametrize("seed", [0, 1, 2, 12345, 4294967295])
def test_random_start_values(checker, tmp_path, seed):
    """--indF r / --freq r: one Tausworthe stream (the oracle build's gsl_rng_taus shim), F and alpha interleaved per
    individual, then the frequencies."""
    assert _same(checker, str(tmp_path), "r", "r", seed, 7, 50) == 0
    assert _same(checker, str(tmp_path), "0.1,0.2", "r", seed, 7, 50) == 0
    assert _same(checker, str(tmp_path), "r", "0.2", seed, 7, 50) == 0


@pytest.mark.parametrize("case", ["plain", "gz", "separators", "clamped", "blank_lines", "freq_header", "crlf",
                                  "indF_three_columns", "freq_two_columns", "freq_too_many"])
def test_start_values_from_files(checker, tmp_path, case):
    tmp = str(tmp_path)
    N, S = 4, 6
    rng = np.random.default_rng(11)
    F = rng.uniform(0.01, 0.9, N); a = rng.uniform(0.001, 0.5, N); fr = rng.uniform(0.02, 0.45, S)
    indF_file = os.path.join(tmp, "start.indF" + (".gz" if case == "gz" else ""))
    freq_file = os.path.join(tmp, "start.freq" + (".gz" if case == "gz" else ""))
    sep = "\t"
    eol = "\r\n" if case == "crlf" else "\n"
    ind_lines = [f"{F[i]!r}{sep}{a[i]!r}" for i in range(N)]
    freq_lines = [repr(float(v)) for v in fr]
    if case == "separators":
        ind_lines = [f"{F[0]!r},{a[0]!r}", f"{F[1]!r} {a[1]!r}", f"{F[2]!r}-{a[2]!r}", f"{F[3]!r}\t\t{a[3]!r}"]
    elif case == "clamped":
        ind_lines[0] = "0\t0"; ind_lines[1] = "1\t25"; freq_lines[0] = "0"; freq_lines[1] = "0.9"
    elif case == "blank_lines":
        ind_lines.insert(2, ""); freq_lines.insert(3, "")
    elif case == "freq_header":
        freq_lines.insert(0, "freq")
    elif case == "indF_three_columns":
        ind_lines[2] += "\t0.5"
    elif case == "freq_two_columns":
        freq_lines[2] += "\t0.5"
    elif case == "freq_too_many":
        freq_lines.append("0.3")
    opener = gzip.open if case == "gz" else open
    with opener(indF_file, "wt", newline="") as fh:
        fh.write(eol.join(ind_lines) + eol)
    with opener(freq_file, "wt", newline="") as fh:
        fh.write(eol.join(freq_lines) + eol)
    _same(checker, tmp, indF_file, "0.1", 1, N, S)
    _same(checker, tmp, "0.1,0.2", freq_file, 1, N, S)
