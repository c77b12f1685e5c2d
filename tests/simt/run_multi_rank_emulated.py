#!/usr/bin/env python3
"""Pre-flight of the MULTI-RANK path of bench.py on a machine without a GPU (TEST INFRASTRUCTURE, run by hand):

    python tests/simt/run_multi_rank_emulated.py [world:mode ...]      # default: 2:mixed 4:mixed 8:nccl 8:mixed 8:direct

One process per rank under torch.distributed (gloo), each with the whole product built for the SIMT emulator
(tests/_simt_build.py::build_whole_product) and "device" memory in POSIX shared memory (SIMT_IPC=1), so that
nfh_peer_export / nfh_peer_import hand real cross-process windows to the kernels.  Every rank runs
ngsf_hmm_b200.selfcheck.multi_rank_check - the N-ranks-against-one-rank check bench.py executes before it times a
multi-GPU run - through ngsf_hmm_b200/em.py (EmRank) as it is; only what is CUDA-specific in em.py's torch plumbing is
replaced (device tensors over the exchange windows -> CPU tensors over the same memory, the side stream of the
posterior all-to-all -> an immediate all-to-all).  What this shows: the order of stages, hooks, fences and collectives
of every exchange mode is consistent on every rank (no rank skips a collective - the hang of the first 8-rank
`mixed` run), ranks that own no individual included, and the sharded numbers equal the single-rank ones.  What it
cannot show: NCCL, NVLink, timing.
"""
import contextlib
import ctypes as C
import json
import os
import socket
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _patch_for_cpu(scratch):
    import torch
    from ngsf_hmm_b200 import api, em
    api.library_path = lambda: os.path.join(scratch, "libngsfhmm_b200.so")
    em._HERE = scratch
    torch.cuda.current_device = lambda: 0
    zeros = torch.zeros
    torch.zeros = lambda *a, device=None, **k: zeros(*a, **k)

    def _tensor(self, which):
        if which not in self._win:
            ptr, nbytes, _ = self.ctx.window(which)
            t = torch.frombuffer((C.c_char * nbytes).from_address(ptr), dtype=torch.float64)
            if which != api.WIN_LOGE0_SUM:
                t = t.view(*em.blocked_owner_layout(self.ctx.n_ranks, self.ctx.n_ind_local, self.ctx.site_block))
            self._win[which] = t
        return self._win[which]

    def begin(self):                       # the side stream of the real thing: here the all-to-all happens at once
        if self.ctx.n_ranks == 1:
            return
        self.ctx.sync()
        em.exchange_all_to_all(self._tensor(api.WIN_POST_SEND), self._tensor(api.WIN_POST_RECV), self.group)
        self._post_done = True

    def end(self):
        self._post_done = None

    em.EmRank._tensor = _tensor
    em.EmRank._stream_ctx = lambda self: contextlib.nullcontext()
    em.EmRank.exchange_posteriors_begin = begin
    em.EmRank.exchange_posteriors_end = end


def _worker(rank, world, port, mode, scratch, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SIMT_IPC="1")
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import torch.distributed as dist
    _patch_for_cpu(scratch)
    from ngsf_hmm_b200 import selfcheck
    import datetime
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=900))
    res = selfcheck.multi_rank_check(0, direct={"direct": True, "mixed": "mixed", "nccl": False}[mode])
    if rank == 0:
        out["res"] = res
    dist.barrier()
    dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp
    import _simt_build
    cases = sys.argv[1:] or ["2:mixed", "4:mixed", "8:nccl", "8:mixed", "8:direct"]
    scratch = os.environ.get("NFH_EMULATED_DIR") or tempfile.mkdtemp(prefix="nfh_emulated_")
    print(f"building the emulated product in {scratch} ...", flush=True)
    _simt_build.build_whole_product(scratch)
    ok = True
    for case in cases:
        world, mode = case.split(":")
        world = int(world)
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), mode, scratch, out), nprocs=world, join=True)
        res = dict(out["res"])
        print(json.dumps({"world": world, "mode": mode, **res}), flush=True)
        ok = ok and res["ok"]
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
