"""CPU: a slice of the `-m gpu` test files, unchanged, against the WHOLE product built for the SIMT emulator
(tests/_simt_build.py::build_whole_product: kernels, launchers, the C ABI layer nfh_ctx.cu over an emulated CUDA runtime
slice, the host library and the ngsF-HMM binary from their sources as they are).  A pytest plugin
(tests/simt/emulated_plugin.py) points the ctypes bindings and the CLI tests at that build, in a subprocess so that this
process keeps the real library.  tests/simt/run_gpu_suite_emulated.py runs every affordable GPU case this way by hand
(12 minutes; profiles/r02/emulated_gpu_suite_preflight_r02.txt); here, in the default CPU run, the cases that cover the
layers nothing else on the CPU reaches: the C ABI layer, the multi-rank driver with kernels storing into the peers'
windows, the drop-in binary end to end.  A checker, not a backend: the package never loads this build.
"""
import os
import subprocess
import sys

import pytest

import _simt_build

ROOT = _simt_build.ROOT
SUBSET = ("tiny_and_tile_boundary or estep_with_batch_equals or error_statuses or outside_the_optimiser_box "
          "or small_multi_rank_geometry or fixed_parameters_two_chromosomes or corner_cases or single_launch_estep")


@pytest.mark.ref
def test_gpu_test_files_pass_against_the_emulated_product(tmp_path_factory):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ngsF-HMM")):
        pytest.skip("oracle/_ref not built (the CLI cases compare with the reference binary)")
    scratch = str(tmp_path_factory.mktemp("emulated_product"))
    _simt_build.build_whole_product(scratch, freq_opt="-O0")       # this slice has little frequency work: compile fast
    env = dict(os.environ, NFH_EMULATED_DIR=scratch, PYTHONPATH=os.pathsep.join(
        [os.path.join(ROOT, "tests", "simt"), os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")]))
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-o",
                        "python_files=test_*.py emulated_cases.py", "-m", "gpu", "-p", "emulated_plugin",
                        "-q", "-x", "-k", SUBSET, "-p", "no:cacheprovider"],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = p.stdout[-3000:] + p.stderr[-2000:]
    assert p.returncode == 0, tail
    last = [ln for ln in p.stdout.splitlines() if " passed" in ln][-1]
    assert "failed" not in last and "error" not in last, tail
    assert int(last.split(" passed")[0].split()[-1]) >= 16, tail           # the slice really ran
