"""CPU: the C-ABI libraries load and export every symbol their headers declare; without a GPU the
product path fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

import ngsf_hmm_b200 as nfh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nfh_[a-z0-9_]+)\s*\(", src)))


def test_cuda_library_exports_every_declared_symbol():
    lib = nfh.load_library()
    names = declared("ngsfhmm_b200.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(nfh.api.EXPORTS) == [n for n in names if n not in ("nfh_ctx",)]
    assert b"sm_100a" in lib.nfh_build_info()


def test_host_library_exports_every_declared_symbol():
    nfh.load_host_library()
    lib = C.CDLL(os.path.join(ROOT, "ngsf-hmm_b200", "libngsfhmm_host.so"))
    for n in declared("ngsfhmm_host.h"):
        assert hasattr(lib, n), n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nfh.NfhError) as e:
        nfh.Context(4, 100)
    assert e.value.status == 6 and "no CPU fallback" in str(e.value)


def test_status_strings_mirror_reference_errors():
    lib = nfh.load_library()
    assert lib.nfh_strerror(3) == b"invalid Lkl found!"                 # HMM.cpp:20
    assert lib.nfh_strerror(4) == b"Fw and Bw lkl do not match!"        # EM.cpp:169


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may touch oracle/."""
    pkg = os.path.join(ROOT, "ngsf-hmm_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), os.path.join(base, f)
