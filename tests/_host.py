"""ctypes window onto the host-side library (libngsfhmm_host.so)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "ngsf-hmm_b200", "libngsfhmm_host.so")
_dp = C.POINTER(C.c_double)
OBJECTIVE = C.CFUNCTYPE(C.c_double, _dp, C.c_void_p)


def load_host():
    L = C.CDLL(HOST_SO)
    L.nfh_host_minimize.restype = C.c_double
    L.nfh_host_minimize.argtypes = [C.c_int, _dp, OBJECTIVE, C.c_void_p, _dp, _dp, C.POINTER(C.c_int)]
    return L


def minimize(fun, x0, lb, ub):
    """Our findmax_bfgs equivalent; returns (x, list of evaluation points, n_evals)."""
    L = load_host()
    x = np.ascontiguousarray(x0, dtype=np.float64).copy(); n = len(x)
    lo = np.ascontiguousarray(lb, dtype=np.float64); hi = np.ascontiguousarray(ub, dtype=np.float64)
    trace = []

    def cb(px, _):
        v = np.array([px[i] for i in range(n)])
        trace.append(v)
        return float(fun(v))

    ne = C.c_int(0)
    L.nfh_host_minimize(n, x.ctypes.data_as(_dp), OBJECTIVE(cb), None, lo.ctypes.data_as(_dp), hi.ctypes.data_as(_dp),
                        C.byref(ne))
    return x, trace, ne.value
