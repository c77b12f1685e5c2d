"""ctypes binding of the C ABI declared in include/ngsfhmm_b200.h.

``Context`` mirrors the reference's ``params`` state for the hot path
(ngsF-HMM.hpp:13-52): upload GL / distances / start values once, then call
``estep`` / ``lkl_batch`` / ``freq_update`` per EM iteration exactly where
``iter_EM`` (EM.cpp:139-289) ran forward/backward, the BFGS objective and the
frequency loop.  All arrays are numpy float64, host memory.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libngsfhmm_b200.so"

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

WIN_POST_SEND, WIN_POST_RECV, WIN_EMIS_SEND, WIN_EMIS_RECV, WIN_E0_SEND, WIN_E0_RECV, WIN_LOGE0_SUM = range(7)

EXPORTS = [
    "nfh_strerror", "nfh_last_error", "nfh_build_info", "nfh_kernel_launches", "nfh_ctx_create", "nfh_ctx_destroy",
    "nfh_n_ind_local", "nfh_n_ind_owned", "nfh_ind_begin", "nfh_site_block", "nfh_site_begin", "nfh_sites_owned",
    "nfh_upload_gl", "nfh_upload_pos_dist", "nfh_set_freq", "nfh_get_freq", "nfh_set_ind_params",
    "nfh_emission_refresh", "nfh_estep", "nfh_lkl_batch", "nfh_freq_update", "nfh_viterbi", "nfh_get_posterior",
    "nfh_geno_posterior", "nfh_exchange_window", "nfh_peer_export", "nfh_peer_import", "nfh_peer_direct", "nfh_sync", "nfh_stream", "nfh_probe_fp64", "nfh_timing",
    "nfh_timing_read", "nfh_freq_passes", "nfh_host_register", "nfh_host_unregister", "nfh_estep_with_batch",
    "nfh_device_count", "nfh_peer_set", "nfh_window_copy_block", "nfh_window_read", "nfh_window_write",
    "nfh_estep_schedule_item",
]


class NfhError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[nfh status {status}] {msg}")
        self.status = status


def library_path() -> str:
    return os.path.join(_HERE, _LIB_NAME)


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return library_path()


_lib = None


def load_library():
    """Load the compiled library or fail loudly - there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build it with __graft_entry__.build() or `make -C ngsf-hmm_b200/csrc` "
            "(the hot path has no CPU fallback)")
    L = C.CDLL(path)
    u64, i32, cint = C.c_uint64, C.c_int32, C.c_int
    L.nfh_strerror.restype = C.c_char_p; L.nfh_strerror.argtypes = [cint]
    L.nfh_last_error.restype = C.c_char_p; L.nfh_last_error.argtypes = [_vp]
    L.nfh_build_info.restype = C.c_char_p
    L.nfh_kernel_launches.restype = u64; L.nfh_kernel_launches.argtypes = [_vp]
    L.nfh_device_count.restype = cint; L.nfh_device_count.argtypes = []
    L.nfh_ctx_create.restype = cint; L.nfh_ctx_create.argtypes = [C.POINTER(_vp), cint, u64, u64, cint, cint]
    L.nfh_ctx_destroy.restype = None; L.nfh_ctx_destroy.argtypes = [_vp]
    for n in ("nfh_n_ind_local", "nfh_n_ind_owned", "nfh_ind_begin", "nfh_site_block", "nfh_site_begin",
              "nfh_sites_owned"):
        getattr(L, n).restype = u64; getattr(L, n).argtypes = [_vp]
    L.nfh_upload_gl.restype = cint; L.nfh_upload_gl.argtypes = [_vp, _vp, u64, u64]
    L.nfh_upload_pos_dist.restype = cint; L.nfh_upload_pos_dist.argtypes = [_vp, _dp]
    L.nfh_set_freq.restype = cint; L.nfh_set_freq.argtypes = [_vp, _dp]
    L.nfh_get_freq.restype = cint; L.nfh_get_freq.argtypes = [_vp, _dp]
    L.nfh_set_ind_params.restype = cint; L.nfh_set_ind_params.argtypes = [_vp, _dp, _dp]
    L.nfh_emission_refresh.restype = cint; L.nfh_emission_refresh.argtypes = [_vp, cint]
    L.nfh_estep.restype = cint; L.nfh_estep.argtypes = [_vp, _dp]
    L.nfh_lkl_batch.restype = cint; L.nfh_lkl_batch.argtypes = [_vp, u64, C.POINTER(i32), _dp, _dp, _dp]
    L.nfh_estep_with_batch.restype = cint
    L.nfh_estep_with_batch.argtypes = [_vp, u64, C.POINTER(i32), _dp, _dp, _dp, _dp]
    L.nfh_freq_update.restype = cint; L.nfh_freq_update.argtypes = [_vp, cint, cint, _dp]
    L.nfh_viterbi.restype = cint; L.nfh_viterbi.argtypes = [_vp, _vp]
    L.nfh_get_posterior.restype = cint; L.nfh_get_posterior.argtypes = [_vp, _dp]
    L.nfh_geno_posterior.restype = cint; L.nfh_geno_posterior.argtypes = [_vp, _vp, _dp]
    L.nfh_exchange_window.restype = cint
    L.nfh_exchange_window.argtypes = [_vp, cint, C.POINTER(_vp), C.POINTER(u64), C.POINTER(u64)]
    L.nfh_peer_export.restype = cint; L.nfh_peer_export.argtypes = [_vp, cint, C.c_char_p]
    L.nfh_peer_import.restype = cint; L.nfh_peer_import.argtypes = [_vp, cint, cint, C.c_char_p]
    L.nfh_peer_direct.restype = cint; L.nfh_peer_direct.argtypes = [_vp, cint]
    L.nfh_peer_set.restype = cint; L.nfh_peer_set.argtypes = [_vp, cint, cint, _vp]
    L.nfh_window_copy_block.restype = cint; L.nfh_window_copy_block.argtypes = [_vp, cint, cint, _vp, cint, cint]
    L.nfh_window_read.restype = cint; L.nfh_window_read.argtypes = [_vp, cint, u64, u64, _vp]
    L.nfh_window_write.restype = cint; L.nfh_window_write.argtypes = [_vp, cint, u64, u64, _vp]
    L.nfh_sync.restype = cint; L.nfh_sync.argtypes = [_vp]
    L.nfh_stream.restype = _vp; L.nfh_stream.argtypes = [_vp]
    L.nfh_probe_fp64.restype = cint; L.nfh_probe_fp64.argtypes = [_vp, _dp]
    L.nfh_timing.restype = cint; L.nfh_timing.argtypes = [_vp, cint]
    L.nfh_timing_read.restype = cint; L.nfh_timing_read.argtypes = [_vp, _dp, C.POINTER(u64), cint]
    L.nfh_freq_passes.restype = cint; L.nfh_freq_passes.argtypes = [_vp, C.POINTER(u64), cint]
    L.nfh_host_register.restype = cint; L.nfh_host_register.argtypes = [_vp, _vp, u64]
    L.nfh_host_unregister.restype = cint; L.nfh_host_unregister.argtypes = [_vp, _vp]
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


TIMING_FAMILIES = ("estep", "lkl_batch", "freq", "viterbi", "emission", "ingest", "_6", "_7")


class Context:
    """Device-resident EM state of one rank (see include/ngsfhmm_b200.h)."""

    def __init__(self, n_ind_total: int, n_sites: int, device: int = 0, n_ranks: int = 1, rank: int = 0):
        self.L = load_library()
        h = _vp()
        rc = self.L.nfh_ctx_create(C.byref(h), device, n_ind_total, n_sites, n_ranks, rank)
        if rc != 0:
            raise NfhError(rc, self.L.nfh_last_error(None).decode())
        self.h = h
        self._pinned = []
        self.n_ind_total, self.n_sites, self.n_ranks, self.rank = n_ind_total, n_sites, n_ranks, rank
        self.n_ind_local = self.L.nfh_n_ind_local(h)
        self.n_ind_owned = self.L.nfh_n_ind_owned(h)
        self.ind_begin = self.L.nfh_ind_begin(h)
        self.site_block = self.L.nfh_site_block(h)
        self.site_begin = self.L.nfh_site_begin(h)
        self.sites_owned = self.L.nfh_sites_owned(h)

    # -- plumbing
    def _chk(self, rc):
        if rc != 0:
            raise NfhError(rc, self.L.nfh_last_error(self.h).decode() or self.L.nfh_strerror(rc).decode())

    def close(self):
        if getattr(self, "h", None):
            for a in getattr(self, "_pinned", []):
                self.L.nfh_host_unregister(self.h, a.ctypes.data)
            self._pinned = []
            self.L.nfh_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- uploads
    def upload_gl(self, log_gl_site_major, first_site=None):
        """log_gl: (n, n_ind_total, 3) natural-log normalised GL for this rank's sites (numpy or pinned torch)."""
        if hasattr(log_gl_site_major, "data_ptr"):   # torch tensor (pinned host memory)
            t = log_gl_site_major
            assert t.is_contiguous() and t.dtype.is_floating_point and t.element_size() == 8
            n = t.shape[0]; ptr = t.data_ptr()
        else:
            g = _f64(log_gl_site_major)
            self._keep = g
            n = g.shape[0]; ptr = g.ctypes.data
        first = self.site_begin if first_site is None else first_site
        self._chk(self.L.nfh_upload_gl(self.h, ptr, first, n))

    def upload_pos_dist(self, dist_mb):
        d = _f64(dist_mb)
        assert d.shape == (self.n_sites,)
        self._chk(self.L.nfh_upload_pos_dist(self.h, _p(d)))

    def set_freq(self, freq):
        f = _f64(np.broadcast_to(freq, (self.sites_owned,)))
        self._chk(self.L.nfh_set_freq(self.h, _p(f)))

    def get_freq(self):
        f = np.empty(self.sites_owned)
        self._chk(self.L.nfh_get_freq(self.h, _p(f)))
        return f

    def set_ind_params(self, indF, alpha):
        F = _f64(np.broadcast_to(indF, (self.n_ind_owned,)))
        a = _f64(np.broadcast_to(alpha, (self.n_ind_owned,)))
        self._chk(self.L.nfh_set_ind_params(self.h, _p(F), _p(a)))

    # -- hot path
    def emission_refresh(self, with_e0=False):
        self._chk(self.L.nfh_emission_refresh(self.h, int(with_e0)))

    def estep(self):
        lk = np.empty(self.n_ind_owned)
        self._chk(self.L.nfh_estep(self.h, _p(lk)))
        return lk

    def estep_async(self):
        self._chk(self.L.nfh_estep(self.h, None))

    def lkl_batch(self, ind, F, alpha):
        ind = np.ascontiguousarray(ind, dtype=np.int32); F = _f64(F); a = _f64(alpha)
        out = np.empty(len(ind))
        self._chk(self.L.nfh_lkl_batch(self.h, len(ind), ind.ctypes.data_as(C.POINTER(C.c_int32)), _p(F), _p(a),
                                       _p(out)))
        return out

    def estep_with_batch(self, ind, F, alpha):
        """E-step + first objective batch in one call; returns (neg_lkl of the requests, ind_lkl)."""
        ind = np.ascontiguousarray(ind, dtype=np.int32); F = _f64(F); a = _f64(alpha)
        out = np.empty(len(ind)); lk = np.empty(self.n_ind_owned)
        self._chk(self.L.nfh_estep_with_batch(self.h, len(ind), ind.ctypes.data_as(C.POINTER(C.c_int32)), _p(F), _p(a),
                                              _p(out), _p(lk)))
        return out, lk

    def pinned_empty(self, n):
        """float64 host array of n elements, page-locked until the context is closed (full-speed copies)."""
        a = np.empty(max(int(n), 1))
        self._chk(self.L.nfh_host_register(self.h, a.ctypes.data, a.nbytes))
        self._pinned.append(a)
        return a[:int(n)]

    def freq_update(self, method=1, posterior_is_zero=False, want_freq=True, out=None):
        f = (out if out is not None else np.empty(self.sites_owned)) if want_freq else None
        self._chk(self.L.nfh_freq_update(self.h, method, int(posterior_is_zero), _p(f) if want_freq else None))
        return f

    def viterbi(self):
        path = np.zeros((self.n_ind_owned, self.n_sites), dtype=np.int8)
        self._chk(self.L.nfh_viterbi(self.h, path.ctypes.data))
        return path

    def get_posterior(self):
        m = np.empty((self.n_ind_owned, self.n_sites))
        self._chk(self.L.nfh_get_posterior(self.h, _p(m)))
        return m

    def geno_posterior(self, path_all):
        p = np.ascontiguousarray(path_all, dtype=np.int8)
        assert p.shape == (self.n_ind_total, self.sites_owned)
        out = np.empty((self.sites_owned, self.n_ind_total, 3))
        self._chk(self.L.nfh_geno_posterior(self.h, p.ctypes.data, _p(out)))
        return out

    # -- multi-rank plumbing / diagnostics
    def window(self, which):
        ptr, nbytes, per = _vp(), C.c_uint64(), C.c_uint64()
        self._chk(self.L.nfh_exchange_window(self.h, which, C.byref(ptr), C.byref(nbytes), C.byref(per)))
        return ptr.value, nbytes.value, per.value

    def sync(self):
        self._chk(self.L.nfh_sync(self.h))

    def peer_export(self, which):
        buf = C.create_string_buffer(64)
        self._chk(self.L.nfh_peer_export(self.h, which, buf))
        return buf.raw

    def peer_import(self, which, peer_rank, handle):
        self._chk(self.L.nfh_peer_import(self.h, which, peer_rank, handle))

    def peer_set(self, which, peer_rank, peer_ctx):
        """Same-process fused exchange: map rank peer_rank's receive window directly (no IPC)."""
        self._chk(self.L.nfh_peer_set(self.h, which, peer_rank, peer_ctx.h))

    def peer_direct(self, enable=True):
        self._chk(self.L.nfh_peer_direct(self.h, int(enable)))

    @property
    def stream(self):
        return self.L.nfh_stream(self.h)

    @property
    def kernel_launches(self):
        return self.L.nfh_kernel_launches(self.h)

    def probe_fp64(self):
        v = C.c_double()
        self._chk(self.L.nfh_probe_fp64(self.h, C.cast(C.byref(v), _dp)))
        return v.value

    def timing(self, enable=True):
        self._chk(self.L.nfh_timing(self.h, int(enable)))

    def timing_read(self, reset=True):
        ms = np.zeros(8); n = np.zeros(8, dtype=np.uint64)
        self._chk(self.L.nfh_timing_read(self.h, _p(ms), n.ctypes.data_as(C.POINTER(C.c_uint64)), int(reset)))
        return {TIMING_FAMILIES[i]: (float(ms[i]), int(n[i])) for i in range(6)}

    def freq_passes(self, reset=True):
        """Sum over this rank's sites of the est_maf passes run since the last reset."""
        v = C.c_uint64()
        self._chk(self.L.nfh_freq_passes(self.h, C.byref(v), int(reset)))
        return int(v.value)
