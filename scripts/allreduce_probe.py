"""How long does the gradient all-reduce take by itself?  (developer probe, torchrun, NCCL)"""
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
for mb in (3.2, 12.8, 25.6, 64.0, 77.0, 154.0):
    t = torch.ones(int(mb * 1e6 / 4), device=dev)
    for _ in range(5):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if rank == 0:
        print(f"world {world} all_reduce {mb:6.1f} MB: {ms * 1e3:7.1f} us  algbw {mb / ms:6.1f} GB/s  "
              f"[{os.environ.get('NCCL_ALGO', 'default')}/{os.environ.get('NCCL_PROTO', 'default')}]", flush=True)
dist.destroy_process_group()
