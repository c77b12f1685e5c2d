"""CPU: the drop-in binary's whole host side against the reference binary, BYTE FOR BYTE.

The PRODUCT's command-line sources (host/cli/*.cpp) and host library sources (lbfgsb, bfgs_driver, host_api, group),
compiled as they are, are linked against a fake device whose arithmetic is the oracle's
(tests/fake_device_ranks_oracle.c: the context-level C ABI with its multi-rank geometry, exchange windows and
peer-direct stores; the oracle equals the reference bit for bit) instead of the CUDA library.  In that build
only the device arithmetic is replaced - by the reference's own - so `.indF`, `.ibd`, `.geno` and the progress
output must equal the unmodified reference binary's exactly, for every input mode and flag combination: readers,
start values (numbers, files, random, estimated), iteration control and stop rule, the lockstep optimiser, Viterbi
hand-over, output formats, --log dumps, replicates - and for --n_gpus 2 / 3 / 8, where est_maf still sums over the
individuals in index order at the owner of a site, so the sharded run must not differ in a single bit.  (That the
kernels compute the oracle's functions is what the GPU tests show.)"""
import os
import subprocess

import numpy as np
import pytest

import ngsf_hmm_b200  # noqa: F401
from ngsf_hmm_b200 import sim

pytestmark = pytest.mark.ref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ngsf-hmm_b200", "host")
REF = os.path.join(ROOT, "oracle", "_ref", "ngsF-HMM")


@pytest.fixture(scope="module")
def ours(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("cli_on_oracle"))
    inc = ["-I", os.path.join(ROOT, "include"), "-I", HOST, "-I", os.path.join(HOST, "cli"), "-I",
           os.path.join(ROOT, "oracle")]
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    objs = []
    srcs = [os.path.join(HOST, f) for f in ("lbfgsb.cpp", "bfgs_driver.cpp", "host_api.cpp", "group.cpp")]
    srcs += [os.path.join(HOST, "cli", f) for f in ("options.cpp", "ingest.cpp", "startvalues.cpp", "em_loop.cpp",
                                                     "report.cpp", "ngsfhmm_main.cpp")]
    procs = []
    for src in srcs:
        o = os.path.join(d, os.path.basename(src) + ".o")
        procs.append(subprocess.Popen(["g++", "-O2", "-std=c++17", "-ffp-contract=off"] + inc + ["-c", src, "-o", o]))
        objs.append(o)
    for name in ("fake_device_ranks_oracle",):
        o = os.path.join(d, name + ".o")
        procs.append(subprocess.Popen(["gcc", "-O2", "-std=c11", "-ffp-contract=off"] + inc +
                                      ["-c", os.path.join(ROOT, "tests", name + ".c"), "-o", o]))
        objs.append(o)
    assert all(p.wait() == 0 for p in procs)
    exe = os.path.join(d, "ngsF-HMM_on_oracle")
    subprocess.check_call(["g++", "-o", exe] + objs + ["-L", os.path.join(ROOT, "oracle"), "-loracle",
                                                       "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lz", "-lpthread"])
    return exe


def _filtered(text):
    # not compared: build stamp, output prefix, our optimiser statistics (--verbose 3) and the reference's per-stage
    # wall-clock block of --verbose 3 (EM.cpp:273-281), which has no counterpart on the device path
    skip = ("\tversion:", "\tout:", "\tBFGS:", "Fw: ", "Bw: ", "MP: ", "indF: ", "freqs: ")
    out = []
    for ln in text.splitlines():
        if ln.startswith(skip) or not ln:
            continue
        if "\ttime: " in ln:
            ln = ln[:ln.index("\ttime: ")]
        out.append(ln)
    return out


def _both(ours, tmp, args, our_extra=(), expect_ok=True):
    pr = subprocess.run([REF] + args + ["--out", "ref"], cwd=tmp, capture_output=True, text=True, timeout=900)
    po = subprocess.run([ours] + args + list(our_extra) + ["--out", "ours"], cwd=tmp, capture_output=True, text=True,
                        timeout=900)
    assert (pr.returncode == 0) == expect_ok, pr.stderr[-2000:]
    assert po.returncode == pr.returncode, po.stderr[-2000:]
    return pr, po


def _same_files(tmp):
    for ext in (".indF", ".ibd", ".geno"):
        a = open(os.path.join(tmp, "ref" + ext), "rb").read()
        b = open(os.path.join(tmp, "ours" + ext), "rb").read()
        assert len(a) > 0 and a == b, ext


def _two_chromosome_pos(path, pos_bp, cut):
    with open(path, "w") as fh:
        for s in range(len(pos_bp)):
            chrom = "chr1" if s < cut else "chr2"
            pos = int(pos_bp[s]) if s < cut else int(pos_bp[s] - pos_bp[cut - 1])
            fh.write(f"{chrom}\t{pos}\n")


@pytest.mark.parametrize("n_gpus", [1, 2, 3, 8])
def test_beagle_likelihoods_free_parameters(ours, tmp_path, n_gpus):
    """BASELINE configs[0] style input; also the progress output, line for line; one rank and sharded."""
    tmp = str(tmp_path)
    N, S = 8, 3000
    d = sim.simulate(N, S, seed=4242, freq=0.2, indF=0.5, alpha=0.01, depth=2.0)
    sim.write_beagle_gz(os.path.join(tmp, "in.beagle.gz"), d.log_gl, d.pos_bp)
    sim.write_pos(os.path.join(tmp, "in.pos"), d.pos_bp)
    args = ["--geno", "in.beagle.gz", "--lkl", "--n_ind", str(N), "--n_sites", str(S), "--pos", "in.pos", "--freq", "0.1",
            "--indF", "0.1,0.2", "--min_iters", "3", "--max_iters", "8", "--verbose", "3", "--n_threads", "4"]
    pr, po = _both(ours, tmp, args, our_extra=["--n_gpus", str(n_gpus)])
    _same_files(tmp)
    assert _filtered(po.stdout) == _filtered(pr.stdout)


@pytest.mark.parametrize("fixed", [["--indF_fixed", "--alpha_fixed"], ["--indF_fixed"], ["--alpha_fixed"], []],
                         ids=["both-fixed", "F-fixed", "alpha-fixed", "free"])
@pytest.mark.parametrize("freq_est", ["0", "1"])
def test_binary_input_start_values_from_files(ours, tmp_path, fixed, freq_est):
    tmp = str(tmp_path)
    N, S = 5, 2000
    d = sim.simulate(N, S, seed=777, freq=(0.05, 0.5), indF=(0.1, 0.6), alpha=0.05)
    sim.write_binary_gl(os.path.join(tmp, "in.glf"), d.log_gl)
    _two_chromosome_pos(os.path.join(tmp, "in.pos"), d.pos_bp, 900)
    np.savetxt(os.path.join(tmp, "freq.txt"), np.clip(d.true_freq, 0.01, 0.49), fmt="%.6f")
    with open(os.path.join(tmp, "indF.txt"), "w") as fh:
        for i in range(N):
            fh.write(f"{max(d.true_F[i], 1e-3):.6f}\t{d.true_alpha[i]:.6f}\n")
    args = ["--geno", "in.glf", "--loglkl", "--n_ind", str(N), "--n_sites", str(S), "--pos", "in.pos", "--freq", "freq.txt",
            "--freq_est", freq_est, "--indF", "indF.txt", "--min_iters", "2", "--max_iters", "4", "--verbose", "1"] + fixed
    pr, po = _both(ours, tmp, args, our_extra=["--n_gpus", "2"] if len(fixed) != 1 else [])
    _same_files(tmp)
    assert _filtered(po.stdout) == _filtered(pr.stdout)


@pytest.mark.parametrize("n_gpus", [1, 3])
@pytest.mark.parametrize("freq_est", ["0", "1"])
def test_called_genotypes_and_estimated_start_frequencies(ours, tmp_path, freq_est, n_gpus):
    """--freq e; with --freq_est 0 only the first site is estimated (parse_args.cpp:316-318)."""
    tmp = str(tmp_path)
    N, S = 6, 1500
    d = sim.simulate(N, S, seed=31337, freq=(0.1, 0.5), indF=(0.1, 0.6), alpha=0.02, depth=20.0)
    sim.write_geno_gz(os.path.join(tmp, "in.geno.gz"), d.geno)
    sim.write_pos(os.path.join(tmp, "in.pos"), d.pos_bp)
    args = ["--geno", "in.geno.gz", "--n_ind", str(N), "--n_sites", str(S), "--pos", "in.pos", "--freq", "e", "--freq_est",
            freq_est, "--indF", "0.1,0.2", "--min_iters", "2", "--max_iters", "4", "--verbose", "0"]
    _both(ours, tmp, args, our_extra=["--n_gpus", str(n_gpus)])
    _same_files(tmp)


@pytest.mark.parametrize("seed", ["7", "8"])
def test_random_start_values_and_genotype_calling(ours, tmp_path, seed):
    tmp = str(tmp_path)
    N, S = 6, 1500
    d = sim.simulate(N, S, seed=99, freq=(0.05, 0.5), indF=(0.0, 0.7), alpha=0.03, depth=4.0)
    sim.write_binary_gl(os.path.join(tmp, "in.glf"), d.log_gl)
    sim.write_pos(os.path.join(tmp, "in.pos"), d.pos_bp)
    args = ["--geno", "in.glf", "--loglkl", "--call_geno", "--n_ind", str(N), "--n_sites", str(S), "--pos", "in.pos",
            "--freq", "r", "--indF", "r", "--seed", seed, "--min_iters", "2", "--max_iters", "5", "--verbose", "1"]
    pr, po = _both(ours, tmp, args)
    _same_files(tmp)
    assert _filtered(po.stdout) == _filtered(pr.stdout)


def test_runs_to_convergence_with_the_same_number_of_iterations(ours, tmp_path):
    """The stop rule (EM.cpp:56): default --min_iters 10 / --max_iters 100 / --min_epsilon 1e-5; F = 0 for everyone, so the
    optimum of F sits on its bound and the output carries NA for alpha."""
    tmp = str(tmp_path)
    N, S = 5, 1200
    d = sim.simulate(N, S, seed=5150, freq=(0.05, 0.5), indF=(0.0, 0.0), alpha=0.02, depth=3.0)
    sim.write_binary_gl(os.path.join(tmp, "in.glf"), d.log_gl)
    sim.write_pos(os.path.join(tmp, "in.pos"), d.pos_bp)
    args = ["--geno", "in.glf", "--loglkl", "--n_ind", str(N), "--n_sites", str(S), "--pos", "in.pos", "--freq", "0.1",
            "--indF", "0.1,0.2", "--verbose", "1", "--log", "4"]
    pr, po = _both(ours, tmp, args)
    _same_files(tmp)
    assert _filtered(po.stdout) == _filtered(pr.stdout)
    assert sum(ln.startswith("Iteration ") for ln in po.stdout.splitlines()) >= 10
    assert b"NA" in open(os.path.join(tmp, "ours.indF"), "rb").read()


def test_max_iters_warning_and_a_single_individual_single_chromosome(ours, tmp_path):
    tmp = str(tmp_path)
    N, S = 1, 300
    d = sim.simulate(N, S, seed=11, freq=(0.1, 0.5), indF=(0.3, 0.3), alpha=0.05, depth=5.0)
    sim.write_binary_gl(os.path.join(tmp, "in.glf"), d.log_gl)
    sim.write_pos(os.path.join(tmp, "in.pos"), d.pos_bp)
    args = ["--geno", "in.glf", "--loglkl", "--n_ind", "1", "--n_sites", str(S), "--pos", "in.pos", "--freq", "0.2",
            "--indF", "0.2-0.3", "--min_iters", "1", "--max_iters", "2", "--verbose", "1"]
    pr, po = _both(ours, tmp, args)
    _same_files(tmp)
    assert "Maximum number of iterations reached" in po.stdout
    assert _filtered(po.stdout) == _filtered(pr.stdout)


def test_replicates_write_the_best_single_run(ours, tmp_path):
    """--n_rep R (extension) = what ngsF-HMM.sh does with R processes (ngsF-HMM.sh:83-116): runs from --seed, --seed+1, ...
    and the files of the highest final logLkl - here against R runs of the reference binary."""
    tmp = str(tmp_path)
    N, S = 4, 800
    d = sim.simulate(N, S, seed=2112, freq=(0.05, 0.5), indF=(0.0, 0.6), alpha=0.03, depth=2.0)
    sim.write_binary_gl(os.path.join(tmp, "in.glf"), d.log_gl)
    sim.write_pos(os.path.join(tmp, "in.pos"), d.pos_bp)
    common = ["--geno", "in.glf", "--loglkl", "--n_ind", str(N), "--n_sites", str(S), "--pos", "in.pos", "--freq", "r",
              "--indF", "r", "--min_iters", "2", "--max_iters", "4", "--verbose", "0"]
    best = None
    for seed in (20, 21, 22):
        subprocess.run([REF] + common + ["--seed", str(seed), "--out", f"ref{seed}"], cwd=tmp, check=True,
                       capture_output=True, timeout=600)
        tot = float(open(os.path.join(tmp, f"ref{seed}.indF")).readline())
        if best is None or tot > best[0]:
            best = (tot, seed)
    subprocess.run([ours] + common + ["--seed", "20", "--n_rep", "3", "--out", "ours"], cwd=tmp, check=True,
                   capture_output=True, timeout=600)
    for ext in (".indF", ".ibd", ".geno"):
        assert open(os.path.join(tmp, f"ref{best[1]}" + ext), "rb").read() == open(os.path.join(tmp, "ours" + ext), "rb").read()
