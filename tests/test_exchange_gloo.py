"""CPU, world_size 2 over gloo: the multi-rank plumbing.  Rank r owns individuals
[r*n_loc, (r+1)*n_loc) for the recursions and site block r for the frequency update; the
posterior window travels [dest rank][local individual][site in block] -> the owner of the site
block, which must then see row = global individual, column = its own sites."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_loc, sb, out):
    sys.path.insert(0, ROOT)
    import ngsf_hmm_b200 as nfh
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape = nfh.em.blocked_owner_layout(world, n_loc, sb)
    # recursion side of rank r: element (block b, local i, off) encodes global individual and global site
    b = torch.arange(world).view(world, 1, 1); i = torch.arange(n_loc).view(1, n_loc, 1); o = torch.arange(sb).view(1, 1, sb)
    send = ((rank * n_loc + i) * 1_000_000 + (b * sb + o)).to(torch.float64).expand(shape).contiguous()
    recv = torch.zeros(shape, dtype=torch.float64)
    nfh.em.exchange_all_to_all(send, recv)
    # frequency side of rank r: rows are global individuals (source rank major), columns its site block
    want = ((b * n_loc + i) * 1_000_000 + (rank * sb + o)).to(torch.float64).expand(shape)
    ok1 = bool(torch.equal(recv, want))
    # and back: emissions produced on the frequency side return to the owners of the individuals
    back = torch.zeros(shape, dtype=torch.float64)
    nfh.em.exchange_all_to_all(recv, back)
    ok2 = bool(torch.equal(back, send))
    # per-individual partial sums of log e0 are all-reduced
    part = torch.full((world * n_loc,), float(rank + 1), dtype=torch.float64)
    dist.all_reduce(part)
    ok3 = bool((part == sum(range(1, world + 1))).all())
    # aliasing windows (n_ranks == 1) are a no-op
    nfh.em.exchange_all_to_all(send, send)
    out[rank] = ok1 and ok2 and ok3
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_to_all_layout_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 3, 8, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_geometry_partition_covers_everything():
    """Mirror of the context's geometry (nfh_ctx_create): padded blocks cover all individuals and sites."""
    TILE = 4224
    for N, S, G in [(100, 1_000_000, 8), (20, 10_000, 2), (1000, 10_000_000, 8), (7, 5000, 4)]:
        n_loc = -(-N // G)
        sb = -(-(-(-S // G)) // TILE) * TILE
        owned_i = [max(0, min(n_loc, N - r * n_loc)) for r in range(G)]
        owned_s = [max(0, min(sb, S - r * sb)) for r in range(G)]
        assert sum(owned_i) == N and sum(owned_s) == S and sb % TILE == 0
