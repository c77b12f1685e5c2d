"""Run under torchrun with >= 2 GPUs: the multi-rank EM (individuals sharded for the recursions,
sites sharded for the frequency update, all-to-all in between) must give the same numbers as
one rank owning everything, with both exchange modes.  Prints MULTI_GPU_OK on success (rank 0)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ngsf_hmm_b200  # noqa: E402,F401
from ngsf_hmm_b200 import selfcheck  # noqa: E402


def main():
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ok = True
    for direct in (False, True, "mixed"):
        res = selfcheck.multi_rank_check(lr, direct=direct)
        if dist.get_rank() == 0:
            print(res, flush=True)
            ok = ok and res["ok"]
    if dist.get_rank() == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_MISMATCH", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ok = bool(flag.item())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
