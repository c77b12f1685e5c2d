"""Host-side mirror of the reference's EM iteration for Python callers.

``EmRank`` plays the role of ``iter_EM(params*)`` (EM.cpp:139-289) for one
rank: E-step, F/alpha update (host L-BFGS-B bookkeeping in
libngsfhmm_host.so around batched objective launches), frequency update with
emission refresh.  With more than one rank the posteriors travel from the
individual-sharded recursion side to the site-sharded frequency side and the
refreshed emissions travel back - two equal-split all-to-alls plus one small
all-reduce per iteration, moved by torch.distributed (NCCL over NVLink on the
GPU box; gloo in the CPU tests of this plumbing).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import api

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_HOOK = C.CFUNCTYPE(None, C.c_void_p)
_hostlib = None


def load_host_library():
    global _hostlib
    if _hostlib is None:
        api.load_library()
        path = os.path.join(_HERE, "libngsfhmm_host.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run __graft_entry__.build()")
        L = C.CDLL(path)
        u64p = C.POINTER(C.c_uint64)
        L.nfh_host_bfgs_update.restype = C.c_int
        L.nfh_host_bfgs_update.argtypes = [C.c_void_p, C.c_uint64, _dp, _dp, C.c_int, C.c_int, u64p]
        L.nfh_host_estep_bfgs_update.restype = C.c_int
        L.nfh_host_estep_bfgs_update.argtypes = [C.c_void_p, C.c_uint64, _dp, _dp, C.c_int, C.c_int, _dp, u64p]
        L.nfh_host_em_iteration.restype = C.c_int
        L.nfh_host_em_iteration.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, u64p]
        _hostlib = L
    return _hostlib


def exchange_all_to_all(send, recv, group=None):
    """Equal-split all-to-all of a [n_ranks, n_ind_local, site_block] window (any device/backend)."""
    import torch.distributed as dist
    if send.data_ptr() == recv.data_ptr():
        return
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)


def blocked_owner_layout(n_ranks, n_ind_local, site_block):
    """Shape of every exchange window: [destination or source rank][local individual][site in block]."""
    return (n_ranks, n_ind_local, site_block)


class _DevWindow:
    """Zero-copy torch view of a device window exported by the C ABI."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 2}


class EmRank:
    def __init__(self, ctx: api.Context, *, indF_fixed=False, alpha_fixed=False, freq_est=1, group=None):
        self.ctx = ctx
        self.H = load_host_library()
        self.indF_fixed, self.alpha_fixed, self.freq_est = indF_fixed, alpha_fixed, freq_est
        self.group = group
        self.stats = np.zeros(3, dtype=np.uint64)
        self.total_evals = 0
        self.total_rounds = 0
        self._freq_host = None        # page-locked home of the per-iteration frequency download
        self._win = {}
        self._ext_stream = None
        self._side = None
        self._post_done = None
        self.direct = False           # posterior tiles AND emission ratios stored into peer windows by the kernels
        self.mixed = False            # emission ratios by peer stores, posteriors by all-to-all behind the BFGS rounds
        self._token = None
        self._hook = None
        self.trace = None             # list of per-iteration CUDA-event tuples when tracing (bench.py --trace)

    # -- multi-rank plumbing ------------------------------------------------
    def _tensor(self, which):
        import torch
        if which not in self._win:
            ptr, nbytes, _ = self.ctx.window(which)
            t = torch.as_tensor(_DevWindow(ptr, nbytes), device=f"cuda:{torch.cuda.current_device()}")
            if which != api.WIN_LOGE0_SUM:
                t = t.view(*blocked_owner_layout(self.ctx.n_ranks, self.ctx.n_ind_local, self.ctx.site_block))
            self._win[which] = t
        return self._win[which]

    def _stream_ctx(self):
        import torch
        if self._ext_stream is None:
            self._ext_stream = torch.cuda.ExternalStream(self.ctx.stream)
        return torch.cuda.stream(self._ext_stream)

    def exchange_posteriors(self):
        if self.ctx.n_ranks == 1:
            return
        with self._stream_ctx():
            exchange_all_to_all(self._tensor(api.WIN_POST_SEND), self._tensor(api.WIN_POST_RECV), self.group)

    def exchange_posteriors_begin(self):
        """Start the posterior all-to-all on a side stream (ordered after everything queued on the context
        stream so far) so that it overlaps the F/alpha update, which only reads the emission window."""
        if self.ctx.n_ranks == 1:
            return
        import torch
        if self._side is None:
            self._side = torch.cuda.Stream()
            self._ext_stream = self._ext_stream or torch.cuda.ExternalStream(self.ctx.stream)
        ready = torch.cuda.Event()
        ready.record(self._ext_stream)
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            exchange_all_to_all(self._tensor(api.WIN_POST_SEND), self._tensor(api.WIN_POST_RECV), self.group)
            self._post_done = torch.cuda.Event()
            self._post_done.record(self._side)

    def exchange_posteriors_end(self):
        if self.ctx.n_ranks == 1 or self._post_done is None:
            return
        self._ext_stream.wait_event(self._post_done)      # the frequency kernel is queued behind the exchange
        self._post_done = None

    def exchange_emissions(self, with_e0=False):
        if self.ctx.n_ranks == 1:
            return
        import torch.distributed as dist
        with self._stream_ctx():
            exchange_all_to_all(self._tensor(api.WIN_EMIS_SEND), self._tensor(api.WIN_EMIS_RECV), self.group)
            if with_e0:
                exchange_all_to_all(self._tensor(api.WIN_E0_SEND), self._tensor(api.WIN_E0_RECV), self.group)
            dist.all_reduce(self._tensor(api.WIN_LOGE0_SUM), group=self.group)

    def enable_peer_direct(self, posteriors=True):
        """Fused exchange: all-gather the CUDA IPC handles of every rank's receive windows and let the
        kernels store straight into them (nfh_peer_export / nfh_peer_import / nfh_peer_direct).

        posteriors=False ("mixed"): only the emission ratios go by peer stores; the posteriors stay local
        and travel by an all-to-all that is started as soon as the E-step is complete and runs behind the
        remaining BFGS rounds.  On 8 GPUs the posterior tiles of all ranks cross NVLink in the same
        millisecond when the kernels store them (E-step + BFGS phase 4.2-6.4 ms per rank against 4.0-4.3
        with the all-to-all, profiles/r02/bench_r02t_*), while the emission ratios are spread over the 14 ms
        of the frequency kernel, where peer stores beat a separate all-to-all (0.6 against 1.6 ms)."""
        import torch
        import torch.distributed as dist
        if self.ctx.n_ranks == 1:
            return False
        for which in (api.WIN_POST_RECV, api.WIN_EMIS_RECV):
            mine = self.ctx.peer_export(which)
            handles = [None] * self.ctx.n_ranks
            dist.all_gather_object(handles, mine, group=self.group)
            for r, h in enumerate(handles):
                self.ctx.peer_import(which, r, h)
        # fixed frequencies (--freq_est 0): nothing on the frequency side reads the posteriors - keep them local
        self.ctx.peer_direct(1 if (self.freq_est and posteriors) else 2)
        self.direct = bool(posteriors or not self.freq_est)
        self.mixed = not self.direct
        self._token = torch.zeros(1, device=f"cuda:{torch.cuda.current_device()}")
        return True

    def _rank_fence(self):
        """Order the stages across ranks on the context stream: completes only after every rank has
        reached it in its own stream, i.e. after the kernels it queued before (1-element all-reduce)."""
        import torch.distributed as dist
        with self._stream_ctx():
            dist.all_reduce(self._token, group=self.group)

    def _allreduce_loge0(self):
        import torch.distributed as dist
        with self._stream_ctx():
            dist.all_reduce(self._tensor(api.WIN_LOGE0_SUM), group=self.group)

    # -- iteration ----------------------------------------------------------
    def refresh_emissions(self, with_e0=False):
        self.ctx.emission_refresh(with_e0)
        if self.direct or self.mixed:
            # ratios were stored into the owners' windows by the kernel; e0 (Viterbi only) still travels by NCCL
            if with_e0:
                with self._stream_ctx():
                    exchange_all_to_all(self._tensor(api.WIN_E0_SEND), self._tensor(api.WIN_E0_RECV), self.group)
            self._allreduce_loge0()                      # doubles as the cross-rank fence for the stores
        else:
            self.exchange_emissions(with_e0)

    def bfgs_update(self, indF, alpha):
        n = self.ctx.n_ind_owned
        rc = self.H.nfh_host_bfgs_update(self.ctx.h, n, indF.ctypes.data_as(_dp), alpha.ctypes.data_as(_dp),
                                         int(self.indF_fixed), int(self.alpha_fixed),
                                         self.stats.ctypes.data_as(C.POINTER(C.c_uint64)))
        self.ctx._chk(rc)
        self.total_rounds += int(self.stats[0]); self.total_evals += int(self.stats[1])

    def estep_bfgs_update(self, indF, alpha, posterior_ready=None):
        """E-step + F / alpha update sharing one forward pass; returns ind_lkl.  posterior_ready() is called as
        soon as the posteriors are complete (after the first round), see nfh_host_estep_bfgs_update_hook."""
        n = self.ctx.n_ind_owned
        lk = np.empty(n)
        if posterior_ready is not None:
            if self._hook is None:
                fn = self.H.nfh_host_estep_bfgs_update_hook
                fn.restype = C.c_int
                fn.argtypes = [C.c_void_p, C.c_uint64, _dp, _dp, C.c_int, C.c_int, _dp, C.POINTER(C.c_uint64),
                               _HOOK, C.c_void_p]
                self._hook = _HOOK(lambda _user: self._hook_target())
            self._hook_target = posterior_ready
            rc = self.H.nfh_host_estep_bfgs_update_hook(
                self.ctx.h, n, indF.ctypes.data_as(_dp), alpha.ctypes.data_as(_dp), int(self.indF_fixed),
                int(self.alpha_fixed), lk.ctypes.data_as(_dp), self.stats.ctypes.data_as(C.POINTER(C.c_uint64)),
                self._hook, None)
            self.ctx._chk(rc)
            self.total_rounds += int(self.stats[0]); self.total_evals += int(self.stats[1])
            return lk
        rc = self.H.nfh_host_estep_bfgs_update(self.ctx.h, n, indF.ctypes.data_as(_dp), alpha.ctypes.data_as(_dp),
                                               int(self.indF_fixed), int(self.alpha_fixed), lk.ctypes.data_as(_dp),
                                               self.stats.ctypes.data_as(C.POINTER(C.c_uint64)))
        self.ctx._chk(rc)
        self.total_rounds += int(self.stats[0]); self.total_evals += int(self.stats[1])
        return lk

    def iteration(self, indF, alpha, want_freq=True):
        """One EM iteration; indF/alpha (float64, n_ind_owned) are updated in place.
        Returns (ind_lkl, freq_of_this_rank's_sites or None); the frequency array is one page-locked
        buffer that every call overwrites (copy it to keep an iteration's values)."""
        ctx = self.ctx
        if want_freq and self.freq_est and self._freq_host is None:
            self._freq_host = ctx.pinned_empty(ctx.sites_owned)    # crosses PCIe every iteration: page-locked
        if ctx.n_ranks == 1:
            lk = np.empty(ctx.n_ind_owned)
            fr = self._freq_host if (want_freq and self.freq_est) else None
            rc = self.H.nfh_host_em_iteration(ctx.h, indF.ctypes.data_as(_dp), alpha.ctypes.data_as(_dp),
                                              int(self.indF_fixed), int(self.alpha_fixed), int(self.freq_est),
                                              lk.ctypes.data_as(_dp), fr.ctypes.data_as(_dp) if fr is not None else None,
                                              self.stats.ctypes.data_as(C.POINTER(C.c_uint64)))
            ctx._chk(rc)
            self.total_rounds += int(self.stats[0]); self.total_evals += int(self.stats[1])
            return lk, fr
        marks = [self._mark()] if self.trace is not None else None
        ctx.set_ind_params(indF, alpha)
        if self.mixed:
            # posteriors leave by all-to-all as soon as they are complete, behind the remaining BFGS rounds
            lk = self.estep_bfgs_update(indF, alpha, posterior_ready=self.exchange_posteriors_begin)
        elif self.direct or not self.freq_est:
            lk = self.estep_bfgs_update(indF, alpha)     # posterior tiles go to their owners from the kernel
        else:
            lk = ctx.estep()
            self.exchange_posteriors_begin()             # overlaps the host/device BFGS rounds below
            self.bfgs_update(indF, alpha)
        if marks is not None:
            marks.append(self._mark())                   # E-step + BFGS rounds of this rank done
        fr = None
        if self.freq_est:
            if self.mixed:
                self.exchange_posteriors_end()
                self._rank_fence()                       # nobody still reads the emissions the kernel overwrites
            elif self.direct:
                self._rank_fence()                       # every rank's E-step stores have landed
            else:
                self.exchange_posteriors_end()
            if marks is not None:
                marks.append(self._mark())               # ... and everybody else's
            fr = ctx.freq_update(1, want_freq=want_freq, out=self._freq_host)
            if marks is not None:
                marks.append(self._mark())               # frequency kernel
            if self.direct or self.mixed:
                self._allreduce_loge0()                  # + fence: every rank's emission stores have landed
            else:
                self.exchange_emissions()
            if marks is not None:
                marks.append(self._mark())               # emissions back with their individuals
                self.trace.append(marks)
        return lk, fr

    def _mark(self):
        import torch
        if self._ext_stream is None:
            self._ext_stream = torch.cuda.ExternalStream(self.ctx.stream)
        e = torch.cuda.Event(enable_timing=True)
        e.record(self._ext_stream)
        return e

    def trace_summary(self):
        """Mean milliseconds on the context stream between the marks of iteration(): E-step + BFGS of this rank,
        wait for the other ranks (fence / posterior exchange), frequency kernel, emission exchange / fence."""
        import torch
        torch.cuda.synchronize()
        if not self.trace:
            return None
        names = ["estep_bfgs", "wait_ranks", "freq", "exchange_back"]
        out = {n: 0.0 for n in names}
        for m in self.trace:
            for k, n in enumerate(names):
                out[n] += m[k].elapsed_time(m[k + 1]) / len(self.trace)
        return out


def run_em(runner: EmRank, indF, alpha, *, min_iters=10, max_iters=100, min_epsilon=1e-5, on_iteration=None):
    """The reference's EM() loop (EM.cpp:27-135) on one rank: iterate until the stop rule of EM.cpp:56
    fails, then decode the Viterbi path with the final parameters (EM.cpp:110-116).

    indF / alpha: float64 arrays (n_ind_owned), updated in place.
    Returns dict(iterations, tot_lkl, ind_lkl, freq, path)."""
    n = len(indF)
    it = 0
    prev_tot = 0.0
    tot = 0.0
    max_eps = -np.inf
    prev_ind = np.full(n, -np.inf)
    lk = np.full(n, -np.inf)
    fr = None
    while ((prev_tot - tot > min_epsilon) or (max_eps > min_epsilon) or it < min_iters) and it < max_iters:
        it += 1
        lk, fr_new = runner.iteration(indF, alpha)
        if fr_new is not None:
            fr = fr_new
        prev_tot = tot
        tot = 0.0
        for v in lk:                       # same left-to-right sum as EM.cpp:77
            tot += float(v)
        with np.errstate(invalid="ignore", divide="ignore"):
            eps = (lk - prev_ind) / np.abs(prev_ind)
        # array_max_pos (gen_func.cpp:73-84): first strictly larger than -inf; NaN never wins
        best, max_eps = 0, -np.inf
        for i, e in enumerate(eps):
            if e > max_eps:
                best, max_eps = i, e
        max_eps = eps[best]
        prev_ind = lk.copy()
        if on_iteration:
            on_iteration(it, tot, max_eps)
    runner.refresh_emissions(with_e0=True)
    runner.ctx.set_ind_params(indF, alpha)
    path = runner.ctx.viterbi()
    return dict(iterations=it, tot_lkl=tot, ind_lkl=lk, freq=fr, path=path)


class Group:
    """One process driving several device contexts (libngsfhmm_host.so, nfh_group_*): the multi-GPU form of
    the reference's main() -> EM() -> iter_EM() (ngsF-HMM.cpp:27-171, EM.cpp:27-135, EM.cpp:139-289).

    ``devices[r]`` is the CUDA ordinal of rank r; the same ordinal may repeat, which runs the multi-rank
    geometry (individuals sharded for the recursions, sites sharded for the frequency EM, posteriors and
    emission ratios crossing between the two) on one GPU.  All arrays are global: indF / alpha / ind_lkl
    [n_ind], freq [n_sites], path / posterior [n_ind, n_sites].
    """

    def __init__(self, n_ind, n_sites, devices, fused_exchange=True, *, indF_fixed=False, alpha_fixed=False,
                 freq_est=1):
        H = load_host_library()
        u64, cint, vp = C.c_uint64, C.c_int, C.c_void_p
        u64p = C.POINTER(u64)
        sig = {
            "nfh_group_create": [C.POINTER(vp), cint, C.POINTER(cint), u64, u64, cint],
            "nfh_group_upload_gl": [vp, vp, u64, u64], "nfh_group_upload_pos_dist": [vp, _dp],
            "nfh_group_set_freq": [vp, _dp], "nfh_group_set_ind_params": [vp, _dp, _dp],
            "nfh_group_refresh_emissions": [vp, cint], "nfh_group_freq_init": [vp, _dp],
            "nfh_group_em_iteration": [vp, _dp, _dp, cint, cint, cint, _dp, _dp, u64p],
            "nfh_group_estep": [vp, _dp], "nfh_group_viterbi": [vp, vp], "nfh_group_get_posterior": [vp, _dp],
            "nfh_group_get_freq": [vp, _dp], "nfh_group_geno_posterior": [vp, vp, _dp], "nfh_group_size": [vp],
        }
        for name, argtypes in sig.items():
            fn = getattr(H, name); fn.restype = cint; fn.argtypes = argtypes
        H.nfh_group_destroy.restype = None; H.nfh_group_destroy.argtypes = [vp]
        H.nfh_group_last_error.restype = C.c_char_p; H.nfh_group_last_error.argtypes = [vp]
        H.nfh_group_ctx.restype = vp; H.nfh_group_ctx.argtypes = [vp, cint]
        self.H = H
        self.n_ind, self.n_sites, self.n_ranks = int(n_ind), int(n_sites), len(devices)
        self.indF_fixed, self.alpha_fixed, self.freq_est = indF_fixed, alpha_fixed, freq_est
        self.stats = np.zeros(3, dtype=np.uint64)
        self.total_evals = 0
        self.total_rounds = 0
        dev = (cint * len(devices))(*devices)
        h = vp()
        rc = H.nfh_group_create(C.byref(h), len(devices), dev, self.n_ind, self.n_sites, int(fused_exchange))
        self.h = h
        self._chk(rc)

    def _chk(self, rc):
        if rc != 0:
            msg = self.H.nfh_group_last_error(self.h).decode() if self.h else ""
            raise api.NfhError(rc, msg or api.load_library().nfh_strerror(rc).decode())

    def close(self):
        if getattr(self, "h", None):
            self.H.nfh_group_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def upload_gl(self, log_gl_site_major, first_site=0):
        g = np.ascontiguousarray(log_gl_site_major, dtype=np.float64)
        self._chk(self.H.nfh_group_upload_gl(self.h, g.ctypes.data, first_site, g.shape[0]))

    def upload_pos_dist(self, dist_mb):
        d = np.ascontiguousarray(dist_mb, dtype=np.float64)
        assert d.shape == (self.n_sites,)
        self._chk(self.H.nfh_group_upload_pos_dist(self.h, d.ctypes.data_as(_dp)))

    def set_freq(self, freq):
        f = np.ascontiguousarray(np.broadcast_to(freq, (self.n_sites,)), dtype=np.float64)
        self._chk(self.H.nfh_group_set_freq(self.h, f.ctypes.data_as(_dp)))

    def set_ind_params(self, indF, alpha):
        F = np.ascontiguousarray(np.broadcast_to(indF, (self.n_ind,)), dtype=np.float64)
        a = np.ascontiguousarray(np.broadcast_to(alpha, (self.n_ind,)), dtype=np.float64)
        self._chk(self.H.nfh_group_set_ind_params(self.h, F.ctypes.data_as(_dp), a.ctypes.data_as(_dp)))

    def refresh_emissions(self, with_e0=False):
        self._chk(self.H.nfh_group_refresh_emissions(self.h, int(with_e0)))

    def freq_init(self):
        f = np.empty(self.n_sites)
        self._chk(self.H.nfh_group_freq_init(self.h, f.ctypes.data_as(_dp)))
        return f

    def iteration(self, indF, alpha, want_freq=True):
        """One EM iteration over all ranks; indF / alpha (float64, n_ind) are updated in place."""
        lk = np.empty(self.n_ind)
        fr = np.empty(self.n_sites) if (want_freq and self.freq_est) else None
        rc = self.H.nfh_group_em_iteration(self.h, indF.ctypes.data_as(_dp), alpha.ctypes.data_as(_dp),
                                           int(self.indF_fixed), int(self.alpha_fixed), int(self.freq_est),
                                           lk.ctypes.data_as(_dp), fr.ctypes.data_as(_dp) if fr is not None else None,
                                           self.stats.ctypes.data_as(C.POINTER(C.c_uint64)))
        self._chk(rc)
        self.total_rounds += int(self.stats[0]); self.total_evals += int(self.stats[1])
        return lk, fr

    def estep(self):
        lk = np.empty(self.n_ind)
        self._chk(self.H.nfh_group_estep(self.h, lk.ctypes.data_as(_dp)))
        return lk

    def viterbi(self):
        path = np.zeros((self.n_ind, self.n_sites), dtype=np.int8)
        self._chk(self.H.nfh_group_viterbi(self.h, path.ctypes.data))
        return path

    def get_posterior(self):
        m = np.empty((self.n_ind, self.n_sites))
        self._chk(self.H.nfh_group_get_posterior(self.h, m.ctypes.data_as(_dp)))
        return m

    def get_freq(self):
        f = np.empty(self.n_sites)
        self._chk(self.H.nfh_group_get_freq(self.h, f.ctypes.data_as(_dp)))
        return f

    def geno_posterior(self, path_all):
        p = np.ascontiguousarray(path_all, dtype=np.int8)
        out = np.empty((self.n_sites, self.n_ind, 3))
        self._chk(self.H.nfh_group_geno_posterior(self.h, p.ctypes.data, out.ctypes.data_as(_dp)))
        return out

    # run_em() treats a Group like an EmRank
    @property
    def ctx(self):
        return self
