#!/usr/bin/env python
"""torchrun check of the C-ABI collective (needs >= 2 GPUs):  mm_nccl_allreduce vs torch.distributed.all_reduce.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/check_nccl.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from macaw_llm_b200 import dist as D

    assert D.init_nccl(dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for dtype, n in ((torch.bfloat16, 1 << 24), (torch.float32, 12345)):
        a = torch.randn(n, device=dev, generator=g).to(dtype)
        b = a.clone()
        D.nccl_allreduce_(a, average=True)
        dist.all_reduce(b, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        assert torch.equal(a, b), (dtype, float((a.float() - b.float()).abs().max()))
    D.destroy_nccl()
    dist.barrier()
    if rank == 0:
        print(f"[check_nccl] OK: mm_nccl_allreduce == torch.distributed all_reduce (AVG) on {world} ranks, bf16 and fp32")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
