"""Time the in-switch all-reduce kernel alone (barrier + kernel + barrier), sweeping the CTA count."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

sys.path.insert(0, ".")
from cirkit_b200 import _lib as L  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
lib = L.load()
n = int(77e6 / 4) // 4 * 4
buf = symm.empty(n, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
mc = hdl.multicast_ptr + (buf.data_ptr() - hdl.buffer_ptrs[rank])
stream = torch.cuda.current_stream().cuda_stream
buf.fill_(1.0)
torch.cuda.synchronize()
dist.barrier()
hdl.barrier(channel=0)
L.check(lib.ckb_nvls_allreduce(mc, n, rank, world, 0, stream), "nvls")
hdl.barrier(channel=1)
torch.cuda.synchronize()
ok = bool((buf == float(world)).all())
for ctas in (16, 32, 64, 148, 296):
    for with_barriers in (True, False):
        buf.fill_(1.0)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            if with_barriers:
                hdl.barrier(channel=0)
            L.check(lib.ckb_nvls_allreduce(mc, n, rank, world, ctas, stream), "nvls")
            if with_barriers:
                hdl.barrier(channel=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if rank == 0:
            print(f"world {world} nvls 77 MB ctas {ctas:3d} unroll {os.environ.get('CKB_NVLS_UNROLL', '4')} "
                  f"{'with' if with_barriers else 'no  '} barriers: {ms * 1e3:7.1f} us  algbw {77.0 / ms:6.1f} GB/s  "
                  f"correct={ok}", flush=True)
dist.destroy_process_group()
