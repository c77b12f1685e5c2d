"""CPU: the drop-in binary's command line against the reference binary's (parse_args.cpp:5-225 and the first
checks of main(), ngsF-HMM.cpp:27-66): same echo of the arguments, same messages, same order of the checks and the
same exit status for every way of stopping before any data is processed.  No device is touched: every case ends
in an error raised before the context would be created."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.ref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "ngsf-hmm_b200", "ngsF-HMM")
REF = os.path.join(ROOT, "oracle", "_ref", "ngsF-HMM")

BASE = ["--geno", "in.glf", "--pos", "in.pos", "--n_ind", "5", "--n_sites", "30", "--out", "o"]
CASES = {
    "nothing": [],
    "no_pos": ["--geno", "in.glf"],
    "no_n_ind": ["--geno", "in.glf", "--pos", "in.pos"],
    "no_n_sites": ["--geno", "in.glf", "--pos", "in.pos", "--n_ind", "5"],
    "no_out": ["--geno", "in.glf", "--pos", "in.pos", "--n_ind", "5", "--n_sites", "30"],
    "call_geno_without_lkl": BASE + ["--call_geno"],
    "bad_freq_est": BASE + ["--freq_est", "3"],
    "bad_e_prob": BASE + ["--e_prob", "3"],
    "min_iters_not_below_max": BASE + ["--min_iters", "5", "--max_iters", "5"],
    "zero_iters": BASE + ["--min_iters", "0"],
    "zero_threads": BASE + ["--n_threads", "0"],
    "unknown_flag": BASE + ["--bogus", "1"],
    "single_dash_flags": ["-geno", "in.glf", "-pos", "in.pos", "-n_ind", "5", "-n_sites", "30", "-freq_est", "7"],
    "missing_geno_file": ["--geno", "nope.glf", "--pos", "in.pos", "--n_ind", "5", "--n_sites", "30", "--out", "o"],
    "binary_size_mismatch": ["--geno", "in.glf", "--pos", "in.pos", "--n_ind", "5", "--n_sites", "31", "--out", "o"],
    "all_flags_echoed": BASE + ["--lkl", "--loglkl", "--freq", "0.3", "--freq_est", "0", "--indF", "0.2,0.1",
                                "--indF_fixed", "--alpha_fixed", "--log", "3", "--min_iters", "4", "--max_iters", "9",
                                "--min_epsilon", "1e-7", "--n_threads", "3", "--seed", "77", "--n_sites", "29"],
    "log_bin_echoed": BASE + ["--log_bin", "2", "--n_sites", "29", "--seed", "5"],
    "more_threads_than_individuals": BASE + ["--n_threads", "9", "--n_sites", "29", "--seed", "5"],
}


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("opts")
    np.full(30 * 5 * 3, np.log(1 / 3)).tofile(str(d / "in.glf"))
    with open(str(d / "in.pos"), "w") as fh:
        fh.write("".join(f"chr1\t{100 * (s + 1)}\n" for s in range(30)))
    return str(d)


def _run(binary, args, cwd):
    p = subprocess.run([binary] + args + ["--verbose", "1"], cwd=cwd, capture_output=True, text=True, timeout=120)
    out = [ln for ln in p.stdout.splitlines() if not ln.startswith("\tversion:")]
    err = [ln.replace(binary, "ngsF-HMM") for ln in p.stderr.splitlines()]
    return p.returncode, out, err


@pytest.mark.parametrize("case", sorted(CASES))
def test_same_echo_messages_and_status(workdir, case):
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("binaries not built")
    rc_r, out_r, err_r = _run(REF, CASES[case], workdir)
    rc_o, out_o, err_o = _run(OURS, CASES[case], workdir)
    assert rc_r != 0, "every case must stop before the data is processed"
    assert rc_o == rc_r
    assert out_o == out_r
    assert err_o == err_r
