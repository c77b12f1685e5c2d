#!/usr/bin/env python3
"""Pre-flight for the round-end GPU run, on a machine WITHOUT a GPU (TEST INFRASTRUCTURE, run by hand):

    python tests/simt/run_gpu_suite_emulated.py [pytest args, e.g. -k "estep or viterbi"]

builds the whole product for the SIMT emulator into a scratch directory (tests/_simt_build.py::build_whole_product:
kernels + launchers + nfh_ctx.cu from mechanically rewritten copies, the host library and the ngsF-HMM binary from the
product's host sources as they are) and runs the `-m gpu` tests against it through the same ctypes bindings and the
same command line - minus the cases an emulator cannot afford (1e5 - 1e8 individual-sites take hours at ~1e4
individual-site-passes per second), that read device memory through torch or that need a second real device.  What passes here has exercised the C ABI
layer, the launch geometry, every kernel and the host side at HEAD; what only the hardware can show (real concurrency,
the MUFU seed, NVLink peers, performance) is what the GPU run is for.  Not part of the default test run, not a
backend: the package never loads this build.
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

# too large for an emulator, or reading device memory through torch / needing real devices
TOO_BIG = ["full_size_properties", "stream", "emission_matches_oracle", "two_ranks_match_one_rank",
           "one_rank_per_device"]


def main():
    import _simt_build
    scratch = os.environ.get("NFH_EMULATED_DIR") or tempfile.mkdtemp(prefix="nfh_emulated_")
    print(f"building the emulated product in {scratch} ...", flush=True)
    _simt_build.build_whole_product(scratch)
    env = dict(os.environ, NFH_EMULATED_DIR=scratch, PYTHONPATH=os.pathsep.join(
        [os.path.join(ROOT, "tests", "simt"), os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")]))
    args = sys.argv[1:]
    if not any(a == "-k" for a in args):
        args = ["-k", " and ".join(f"not {t}" for t in TOO_BIG)] + args
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-o", "python_files=test_*.py emulated_cases.py",
           "-m", "gpu", "-p", "emulated_plugin", "-q",
           "--durations=15"] + args
    return subprocess.call(cmd, env=env, cwd=ROOT)


if __name__ == "__main__":
    sys.exit(main())
