"""pytest plugin of tests/simt/run_gpu_suite_emulated.py (TEST INFRASTRUCTURE): points the ctypes bindings and the CLI
tests at the emulated build in $NFH_EMULATED_DIR.  Loaded only when that script passes `-p emulated_plugin`."""
import os
import sys

DIR = os.environ["NFH_EMULATED_DIR"]


def pytest_configure(config):
    import ngsf_hmm_b200 as nfh
    from ngsf_hmm_b200 import api, em
    api.library_path = lambda: os.path.join(DIR, "libngsfhmm_b200.so")
    em._HERE = DIR
    assert api._lib is None and em._hostlib is None


def pytest_collection_finish(session):
    for name in ("test_gpu_cli", "test_gpu_zz_cli_corner_cases"):
        m = sys.modules.get(name)
        if m is not None and hasattr(m, "OURS"):
            m.OURS = os.path.join(DIR, "ngsF-HMM")
        if m is not None and hasattr(m, "PATCHED") and os.path.exists(os.path.join(DIR, "ngsF-HMM_b200patch")):
            m.PATCHED = os.path.join(DIR, "ngsF-HMM_b200patch")
