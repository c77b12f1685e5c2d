"""CPU: the drop-in binary's input readers (host/cli/ingest.cpp) against the reference's own readers
(read_data.cpp:13-218 + the ingest part of main(), ngsF-HMM.cpp:47-117, through oracle/_ref): same normalised
log-GL and same distances, bit for bit, for every input mode of examples/test.sh (called genotypes, likelihoods,
log-likelihoods as gz text; raw-double binary) with BEAGLE-style leading columns, a header line, blank lines,
zero likelihoods, missing calls, --call_geno with ties, chromosome changes and '#' comments in the POS file; and
the same error text and exit status as the reference binary on malformed input."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.ref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ngsf-hmm_b200", "host", "cli")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ngsF-HMM")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("ingest") / "cli_ingest_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           "-I", CLI, "-o", exe, os.path.join(ROOT, "tests", "cli_ingest_check.cpp"),
                           os.path.join(CLI, "options.cpp"), os.path.join(CLI, "ingest.cpp"), "-lz", "-lpthread"])
    return exe


def _ref_read(ref, geno, pos, N, S, lkl, loglkl, call):
    gl = np.empty((S, N, 3)); dist = np.empty(S)
    dp = C.POINTER(C.c_double)
    ref.lib.ref_read_inputs.restype = None
    ref.lib.ref_read_inputs.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, dp, dp]
    ref.lib.ref_read_inputs(geno.encode(), pos.encode(), N, S, int(lkl), int(loglkl), int(call),
                            gl.ctypes.data_as(dp), dist.ctypes.data_as(dp))
    return gl, dist


def _ours(checker, tmp, geno, pos, N, S, flags, threads=3):
    out = os.path.join(tmp, "ours")
    cmd = [checker, "--geno", geno, "--pos", pos, "--n_ind", str(N), "--n_sites", str(S), "--out", out, "--verbose", "0",
           "--n_threads", str(threads)] + flags
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return (np.fromfile(out + ".gl").reshape(S, N, 3), np.fromfile(out + ".dist"))


def _write_pos(path, S, rng, gz=False, comments=True, extra_col=False):
    lines = []
    chrom, pos = 1, 0
    for s in range(S):
        if s and rng.random() < 0.05:
            chrom += 1; pos = 0                       # chromosome change -> +inf distance
        pos += int(rng.integers(1, 200000))
        lines.append(f"chr{chrom}\t{pos}" + ("\tx" if extra_col else ""))
        if comments and rng.random() < 0.03:
            lines.append("# a comment line")
    text = "\n".join(lines) + "\n"
    if gz:
        with gzip.open(path, "wt") as fh:
            fh.write(text)
    else:
        with open(path, "w") as fh:
            fh.write(text)


def _lik(rng, S, N):
    """Linear likelihoods with awkward entries: exact zeros, ties, all-equal triples, tiny and huge scales."""
    g = rng.dirichlet([0.6, 0.6, 0.6], size=(S, N))
    r = rng.random((S, N))
    g[r < 0.05] = [1 / 3, 1 / 3, 1 / 3]
    g[(r >= 0.05) & (r < 0.10)] = [0.0, 0.5, 0.5]
    g[(r >= 0.10) & (r < 0.13)] = [1.0, 0.0, 0.0]
    g[(r >= 0.13) & (r < 0.16)] *= 1e-300
    g[(r >= 0.16) & (r < 0.19)] *= 1e6
    return g


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("mode", ["lkl_beagle", "loglkl_text", "calls", "bin_linear", "bin_log", "lkl_call_geno",
                                  "bin_call_geno", "lkl_crlf_shifted"])
def test_readers_match_the_reference(ref, checker, tmp_path, mode):
    rng = np.random.default_rng(int(os.environ.get("NFH_INGEST_SEED", "7")) + sum(map(ord, mode)))
    N, S = 7, 400
    tmp = str(tmp_path)
    pos = os.path.join(tmp, "in.pos.gz" if mode in ("calls", "bin_log") else "in.pos")
    _write_pos(pos, S, rng, gz=pos.endswith(".gz"), extra_col=mode == "loglkl_text")
    g = _lik(rng, S, N)
    flags, lkl, loglkl, call = [], False, False, False
    if mode in ("lkl_beagle", "lkl_call_geno"):
        geno = os.path.join(tmp, "in.beagle.gz")
        with gzip.open(geno, "wt") as fh:
            fh.write("marker\tallele1\tallele2\t" + "\t".join(f"Ind{i}\tInd{i}\tInd{i}" for i in range(N)) + "\n")
            for s in range(S):
                if s == 17:
                    fh.write("\n")                     # an empty line: the site keeps its initial value
                    continue
                fh.write(f"chr1_{s}\tA\tC\t" + "\t".join(repr(float(v)) for v in g[s].ravel()) + "\n")
        flags, lkl = ["--lkl"], True
        if mode == "lkl_call_geno":
            flags.append("--call_geno"); call = True
    elif mode == "lkl_crlf_shifted":
        # CRLF lines: the reference's chomp removes one character, the '\r' stays glued to the last likelihood and
        # split() drops that token, so the LAST n_ind*3 numeric fields start one column early (here: at the numeric
        # position column).  Garbage in the reference - and the same garbage here.
        geno = os.path.join(tmp, "in.crlf.gz")
        with gzip.open(geno, "wt", newline="") as fh:
            for s in range(S):
                fh.write(f"{s + 1}\t" + "\t".join(repr(float(v)) for v in g[s].ravel()) + "\r\n")
        flags, lkl = ["--lkl"], True
    elif mode == "loglkl_text":
        geno = os.path.join(tmp, "in.glf.gz")
        with np.errstate(divide="ignore"):
            lg = np.log(g)
        lg[np.isneginf(lg)] = -1e15
        with gzip.open(geno, "wt") as fh:
            for s in range(S):
                fh.write(" ".join(repr(float(v)) for v in lg[s].ravel()) + "\n")
        flags, lkl, loglkl = ["--loglkl"], True, True
    elif mode == "calls":
        geno = os.path.join(tmp, "in.geno.gz")
        calls = rng.integers(-1, 3, size=(S, N))
        with gzip.open(geno, "wt") as fh:
            for s in range(S):
                fh.write(f"chr1\t{s + 1}\t" + "\t".join(str(int(v)) for v in calls[s]) + "\n")
    else:
        geno = os.path.join(tmp, "in.glf")
        if mode == "bin_log":
            with np.errstate(divide="ignore"):
                lg = np.log(g)
            lg.tofile(geno)                            # -inf stays in the file: the reader's logsum copes
            flags, loglkl = ["--loglkl"], True
        else:
            g.tofile(geno)
            flags = ["--lkl"]
        lkl = True
        if mode == "bin_call_geno":
            flags.append("--call_geno"); call = True
    gl_ref, dist_ref = _ref_read(ref, geno, pos, N, S, lkl, loglkl, call)
    gl, dist = _ours(checker, tmp, geno, pos, N, S, flags)
    assert _same(dist, dist_ref)
    assert np.isinf(dist).sum() >= 1 and dist[0] > 0
    assert _same(gl, gl_ref), np.argwhere(~((gl == gl_ref) | (np.isnan(gl) & np.isnan(gl_ref))))[:5]


def _run(exe, args):
    p = subprocess.run([exe] + args, capture_output=True, text=True, timeout=120)
    err = [ln for ln in p.stderr.splitlines() if ln.startswith("ERROR:")]
    return p.returncode, err


@pytest.mark.parametrize("case", ["short_line", "bad_call", "premature_eof", "not_at_eof", "bin_size", "pos_lines",
                                  "pos_backwards", "pos_one_column", "pos_no_final_newline", "pos_ragged_columns",
                                  "pos_empty", "pos_header", "geno_crlf"])
def test_malformed_input_same_error_as_the_reference_binary(checker, tmp_path, case):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/ngsF-HMM not built")
    tmp = str(tmp_path)
    N, S = 3, 6
    pos = os.path.join(tmp, "in.pos")
    pos_lines = [f"chr1\t{100 * (s + 1)}" for s in range(S)]
    geno = os.path.join(tmp, "in.geno.gz")
    rows = [["0.2", "0.3", "0.5"] * N for _ in range(S)]
    flags = ["--lkl"]
    if case == "short_line":
        rows[3] = rows[3][:-2]
    elif case == "bad_call":
        rows = [["0", "1", "3"] for _ in range(S)]; flags = []
    elif case == "premature_eof":
        rows = rows[:-2]
    elif case == "not_at_eof":
        rows = rows + rows[:2]
    elif case == "bin_size":
        geno = os.path.join(tmp, "in.glf")
    elif case == "pos_lines":
        pos_lines = pos_lines[:-1]
    elif case == "pos_backwards":
        pos_lines[3] = "chr1\t150"
    elif case == "pos_one_column":
        pos_lines = [ln.split("\t")[1] for ln in pos_lines]
    pos_text = "\n".join(pos_lines) + "\n"
    eol = "\n"
    if case == "pos_no_final_newline":
        pos_text = pos_text[:-1]                  # read_file tests for EOF after the read: the last line is lost
    elif case == "pos_ragged_columns":
        pos_text = pos_text.replace("chr1\t300\n", "chr1\t300\tx\n")
    elif case == "pos_empty":
        pos_text = "# nothing but a comment\n"
    elif case == "pos_header":
        pos_text = "chr\tpos\n" + pos_text
    elif case == "geno_crlf":
        eol = "\r\n"                             # the '\r' hides the last field of every line
    with open(pos, "w") as fh:
        fh.write(pos_text)
    if geno.endswith(".gz"):
        with gzip.open(geno, "wt", newline="") as fh:
            fh.write("".join("\t".join(r) + eol for r in rows))
    else:
        np.zeros(S * N * 3 - 5).tofile(geno)
    args = ["--geno", geno, "--pos", pos, "--n_ind", str(N), "--n_sites", str(S), "--out", os.path.join(tmp, "o"),
            "--freq", "0.1", "--indF", "0.1,0.2", "--verbose", "0", "--n_threads", "1"] + flags
    rc_ref, err_ref = _run(REF_BIN, args)
    rc, err = _run(checker, args)
    assert rc_ref != 0 and rc == rc_ref, (rc, rc_ref)
    assert err == err_ref and len(err) == 1, (err, err_ref)
