"""ctypes windows onto the CPU checkers (test infrastructure only).

``Oracle``  -> oracle/liboracle.so        (our C restatement, always buildable)
``Ref``     -> oracle/_ref/libngsfhmm_ref.so (the unmodified reference; prebuilt
               here from /root/reference, travels to the GPU box as a binary)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libngsfhmm_ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "ngsF-HMM")

_dp = C.POINTER(C.c_double)
_cp = C.c_char_p


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "ngsfhmm_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return ORACLE_SO


def build_ref() -> str | None:
    """Build oracle/_ref when the reference sources are present; else use the prebuilt file."""
    if os.path.isdir("/root/reference") and not os.path.exists(REF_SO):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])
    return REF_SO if os.path.exists(REF_SO) else None


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_logsum.restype = C.c_double
        L.orc_logsum.argtypes = [_dp, C.c_uint64]
        L.orc_calc_trans.restype = C.c_double
        L.orc_calc_trans.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_calc_HWE.argtypes = [_dp, C.c_double, C.c_double, C.c_int]
        L.orc_post_prob.argtypes = [_dp, _dp, _dp]
        L.orc_calc_emission.restype = C.c_double
        L.orc_calc_emission.argtypes = [_dp, C.c_double, C.c_int]
        L.orc_est_maf.restype = C.c_double
        L.orc_est_maf.argtypes = [C.c_uint64, _dp, _dp]
        L.orc_est_maf_counted.restype = C.c_double
        L.orc_est_maf_counted.argtypes = [C.c_uint64, _dp, _dp, C.POINTER(C.c_int)]
        for name in ("orc_forward", "orc_backward"):
            fn = getattr(L, name)
            fn.restype = C.c_double
            fn.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double, _dp]
        L.orc_viterbi.restype = C.c_double
        L.orc_viterbi.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double, C.c_void_p]
        L.orc_lkl.restype = C.c_double
        L.orc_lkl.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double]
        L.orc_estep.restype = C.c_int
        L.orc_estep.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_freq_emission.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, C.c_int, _dp, _dp]
        L.orc_normalize_gl.argtypes = [C.c_uint64, _dp]
        L.orc_call_geno.argtypes = [C.c_uint64, _dp]
        L.orc_estep_extended.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double, _dp, _dp]

    # --- scalar helpers
    def calc_emission(self, gl3, maf, k):
        g = f64(gl3)
        return self.lib.orc_calc_emission(_d(g), float(maf), int(k))

    def est_maf(self, gl, indF):
        g = f64(gl); F = f64(indF)
        return self.lib.orc_est_maf(len(F), _d(g), _d(F))

    def est_maf_counted(self, gl, indF):
        """(freq, passes) of one site."""
        g = f64(gl); F = f64(indF)
        n = C.c_int(0)
        f = self.lib.orc_est_maf_counted(len(F), _d(g), _d(F), C.byref(n))
        return f, n.value

    def calc_HWE(self, maf, F, log_scale=True):
        out = np.empty(3)
        self.lib.orc_calc_HWE(_d(out), float(maf), float(F), int(log_scale))
        return out

    # --- per-individual recursions; e_prob (S,2), dist (S,)
    def forward(self, e_prob, dist, F, alpha, want_table=False):
        e = f64(e_prob); d = f64(dist); S = len(d)
        Fw = np.empty((S + 1, 2)) if want_table else None
        v = self.lib.orc_forward(S, _d(e), _d(d), float(F), float(alpha), _d(Fw))
        return (v, Fw) if want_table else v

    def backward(self, e_prob, dist, F, alpha, want_table=False):
        e = f64(e_prob); d = f64(dist); S = len(d)
        Bw = np.empty((S + 1, 2)) if want_table else None
        v = self.lib.orc_backward(S, _d(e), _d(d), float(F), float(alpha), _d(Bw))
        return (v, Bw) if want_table else v

    def viterbi(self, e_prob, dist, F, alpha):
        e = f64(e_prob); d = f64(dist); S = len(d)
        path = np.zeros(S, dtype=np.int8)
        v = self.lib.orc_viterbi(S, _d(e), _d(d), float(F), float(alpha), path.ctypes.data)
        return v, path

    def lkl(self, e_prob, dist, F, alpha):
        e = f64(e_prob); d = f64(dist)
        return self.lib.orc_lkl(len(d), _d(e), _d(d), float(F), float(alpha))

    # --- all individuals; e_prob (N,S,2)
    def estep(self, e_prob, dist, F, alpha):
        e = f64(e_prob); d = f64(dist); F = f64(F); a = f64(alpha)
        N, S = e.shape[0], e.shape[1]
        marg1 = np.empty((N, S)); lk = np.empty(N)
        st = self.lib.orc_estep(N, S, _d(e), _d(d), _d(F), _d(a), _d(marg1), _d(lk))
        return st, marg1, lk

    def freq_emission(self, gl_ind_major, marg1, freq, update_freq=True):
        g = f64(gl_ind_major); N, S = g.shape[0], g.shape[1]
        m = f64(marg1) if marg1 is not None else np.zeros((N, S))
        fr = f64(freq).copy()
        e = np.empty((N, S, 2))
        self.lib.orc_freq_emission(N, S, _d(g), _d(m), int(update_freq), _d(fr), _d(e))
        return fr, e

    def normalize_gl(self, gl, call_geno=False):
        """What read_geno + main() do to raw log GL: normalise, optionally call genotypes, normalise."""
        g = f64(gl).copy()
        n = g.size // 3
        if call_geno:
            # read_geno normalises once, main() calls genotypes, then normalises again
            one = Oracle._norm_once(g)
            self.lib.orc_call_geno(n, _d(one))
            return Oracle._norm_once(one)
        self.lib.orc_normalize_gl(n, _d(g))
        return g

    @staticmethod
    def _norm_once(g):
        m = g.max(axis=-1, keepdims=True)
        with np.errstate(divide="ignore"):
            return g - (np.log(np.exp(g - m).sum(axis=-1, keepdims=True)) + m)

    def estep_extended(self, e_prob, dist, F, alpha):
        e = f64(e_prob); d = f64(dist); S = len(d)
        m = np.empty(S); lk = C.c_double()
        self.lib.orc_estep_extended(S, _d(e), _d(d), float(F), float(alpha), _d(m), C.cast(C.byref(lk), _dp))
        return m, lk.value


OBJECTIVE = C.CFUNCTYPE(C.c_double, _dp, C.c_void_p)


class Ref:
    """The unmodified reference behind oracle/ref_harness.cpp."""

    def __init__(self):
        so = build_ref()
        if so is None:
            raise FileNotFoundError("oracle/_ref/libngsfhmm_ref.so missing and /root/reference absent")
        self.lib = C.CDLL(so)
        L = self.lib
        for name in ("ref_forward", "ref_backward"):
            fn = getattr(L, name)
            fn.restype = C.c_double
            fn.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double, _dp]
        L.ref_viterbi.restype = C.c_double
        L.ref_viterbi.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double, C.c_void_p]
        L.ref_calc_emission.restype = C.c_double
        L.ref_calc_emission.argtypes = [_dp, C.c_double, C.c_uint64]
        L.ref_calc_HWE.argtypes = [_dp, C.c_double, C.c_double, C.c_int]
        L.ref_post_prob.argtypes = [_dp, _dp, _dp]
        L.ref_est_maf.restype = C.c_double
        L.ref_est_maf.argtypes = [C.c_uint64, _dp, _dp]
        L.ref_lkl.restype = C.c_double
        L.ref_lkl.argtypes = [C.c_uint64, _dp, _dp, C.c_double, C.c_double]
        L.ref_bfgs_individual.argtypes = [C.c_uint64, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.ref_findmax_bfgs.restype = C.c_double
        L.ref_findmax_bfgs.argtypes = [C.c_int, _dp, OBJECTIVE, C.c_void_p, _dp, _dp]
        L.ref_state_create.restype = C.c_void_p
        L.ref_state_create.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_uint, _cp]
        L.ref_state_iter_EM.argtypes = [C.c_void_p]
        L.ref_state_run_EM.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_double]
        L.ref_state_viterbi.argtypes = [C.c_void_p]
        L.ref_state_get.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, C.c_void_p, _dp, _dp]
        L.ref_state_set.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.ref_state_destroy.argtypes = [C.c_void_p]

    def calc_emission(self, gl3, maf, k):
        g = f64(gl3)
        return self.lib.ref_calc_emission(_d(g), float(maf), int(k))

    def calc_HWE(self, maf, F, log_scale=True):
        out = np.empty(3)
        self.lib.ref_calc_HWE(_d(out), float(maf), float(F), int(log_scale))
        return out

    def est_maf(self, gl, indF):
        g = f64(gl); F = f64(indF)
        return self.lib.ref_est_maf(len(F), _d(g), _d(F))

    def forward(self, e_prob, dist, F, alpha, want_table=False):
        e = f64(e_prob); d = f64(dist); S = len(d)
        Fw = np.empty((S + 1, 2)) if want_table else None
        v = self.lib.ref_forward(S, _d(e), _d(d), float(F), float(alpha), _d(Fw))
        return (v, Fw) if want_table else v

    def backward(self, e_prob, dist, F, alpha, want_table=False):
        e = f64(e_prob); d = f64(dist); S = len(d)
        Bw = np.empty((S + 1, 2)) if want_table else None
        v = self.lib.ref_backward(S, _d(e), _d(d), float(F), float(alpha), _d(Bw))
        return (v, Bw) if want_table else v

    def viterbi(self, e_prob, dist, F, alpha):
        e = f64(e_prob); d = f64(dist); S = len(d)
        path = np.zeros(S, dtype=np.int8)
        v = self.lib.ref_viterbi(S, _d(e), _d(d), float(F), float(alpha), path.ctypes.data)
        return v, path

    def lkl(self, e_prob, dist, F, alpha):
        e = f64(e_prob); d = f64(dist)
        return self.lib.ref_lkl(len(d), _d(e), _d(d), float(F), float(alpha))

    def bfgs_individual(self, e_prob, dist, F, alpha, F_fixed=False, alpha_fixed=False):
        e = f64(e_prob); d = f64(dist)
        Fv = C.c_double(F); av = C.c_double(alpha); n = C.c_uint64(0)
        self.lib.ref_bfgs_individual(len(d), _d(e), _d(d), C.cast(C.byref(Fv), _dp), C.cast(C.byref(av), _dp),
                                     int(F_fixed), int(alpha_fixed), C.byref(n))
        return Fv.value, av.value, n.value

    def findmax_bfgs(self, x0, fun, lb, ub):
        """fun(np.ndarray) -> float.  Returns (x, trace of evaluation points)."""
        x = f64(x0).copy(); n = len(x)
        lo = f64(lb).copy(); hi = f64(ub).copy()
        trace = []

        def cb(px, _):
            v = np.array([px[i] for i in range(n)])
            trace.append(v)
            return float(fun(v))

        self.lib.ref_findmax_bfgs(n, _d(x), OBJECTIVE(cb), None, _d(lo), _d(hi))
        return x, trace

    def state(self, log_gl_site_major, dist, freq, indF, alpha, *, freq_est=1, indF_fixed=False,
              alpha_fixed=False, call_geno=False, n_threads=1, out_prefix=None):
        return RefState(self, log_gl_site_major, dist, freq, indF, alpha, freq_est, indF_fixed, alpha_fixed,
                        call_geno, n_threads, out_prefix)


class RefState:
    def __init__(self, ref, gl, dist, freq, indF, alpha, freq_est, indF_fixed, alpha_fixed, call_geno, n_threads,
                 out_prefix):
        self.ref = ref
        g = f64(gl); S, N = g.shape[0], g.shape[1]
        self.N, self.S = N, S
        d = f64(dist); fr = f64(np.broadcast_to(freq, (S,))); F = f64(np.broadcast_to(indF, (N,)))
        a = f64(np.broadcast_to(alpha, (N,)))
        pre = (out_prefix or f"/tmp/ngsfhmm_ref_{os.getpid()}").encode()
        self.h = ref.lib.ref_state_create(N, S, _d(g), _d(d), _d(fr), _d(F), _d(a), int(freq_est), int(indF_fixed),
                                          int(alpha_fixed), int(call_geno), int(n_threads), pre)

    def iter_EM(self):
        self.ref.lib.ref_state_iter_EM(self.h)

    def run_EM(self, min_iters=10, max_iters=100, min_epsilon=1e-5):
        self.ref.lib.ref_state_run_EM(self.h, min_iters, max_iters, min_epsilon)

    def viterbi(self):
        self.ref.lib.ref_state_viterbi(self.h)

    def get(self):
        N, S = self.N, self.S
        out = dict(freq=np.empty(S), indF=np.empty(N), alpha=np.empty(N), ind_lkl=np.empty(N),
                   e_prob=np.empty((N, S, 2)), marg1=np.empty((N, S)), path=np.zeros((N, S), dtype=np.int8),
                   gl_norm=np.empty((N, S, 3)))
        tot = C.c_double()
        self.ref.lib.ref_state_get(self.h, _d(out["freq"]), _d(out["indF"]), _d(out["alpha"]), _d(out["ind_lkl"]),
                                   _d(out["e_prob"]), _d(out["marg1"]), out["path"].ctypes.data, _d(out["gl_norm"]),
                                   C.cast(C.byref(tot), _dp))
        out["tot_lkl"] = tot.value
        return out

    def set(self, freq=None, indF=None, alpha=None, e_prob=None):
        a = [f64(x) if x is not None else None for x in (freq, indF, alpha, e_prob)]
        self.ref.lib.ref_state_set(self.h, *[_d(x) for x in a])

    def close(self):
        if self.h:
            self.ref.lib.ref_state_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
