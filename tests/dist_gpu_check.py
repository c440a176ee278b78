"""Multi-GPU parity check, launched with torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Every rank evaluates its row block of the golden fixtures' batch with the CUDA library; the
all-gathered log-likelihoods must match the reference's outputs and the all-reduced gradients the
reference's gradients of -mean(ll), with the tolerances of tests/test_gpu_parity.py.
"""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from helpers import Golden, grad_tolerance  # noqa: E402

from cirkit_b200 import B200Circuit  # noqa: E402
from cirkit_b200.distributed import BatchShardedCircuit  # noqa: E402


def overlap_check(dev, rank: int, world: int) -> None:
    """The staged backward pass with its per-stage asynchronous all-reduces
    (`overlap_gradient_sync`) against one all-reduce after the backward pass, same data: equal up
    to the summation order of the collective (bit-equal on two ranks)."""
    import dataclasses

    for name, units, rows in (("qt8_cp_k4", 64, 640), ("pd32_cp_k4", 16, 96), ("qt8_cp_k4", 4, 40)):
        g = Golden(name)
        plan = dataclasses.replace(g.plan, meta={"units": 4}).with_units(units)
        x = torch.randint(0, 256, (world * rows, plan.num_variables),
                          generator=torch.Generator().manual_seed(9))
        grads = []
        for overlap in (False, True, "nvls"):
            cc = B200Circuit(plan, seed=21).to(dev)
            sharded = BatchShardedCircuit(cc)
            if overlap == "nvls":
                # in-switch sum over NVLink multicast (csrc/nvls_allreduce.cu) instead of NCCL
                if not sharded.nvls_gradient_sync():
                    if rank == 0:
                        print(f"dist nvls SKIPPED {name}: no NVLink multicast support", flush=True)
                    continue
            elif overlap:
                assert sharded.overlap_gradient_sync(4, bucket_bytes=1 << 16)
            for _ in range(2):  # the second pass re-uses buffers the first one's collectives read
                for p in cc.leaves:
                    p.grad = None
                sharded.loss(sharded.shard(x).to(dev), x.shape[0]).backward()
                nbytes = sharded.sync_gradients()
            assert nbytes >= sum(4 * p.numel() for p in cc.leaves)
            grads.append([p.grad.clone() for p in cc.leaves])
        for other in grads[1:]:
          for i, (a, b) in enumerate(zip(grads[0], other)):
            # equal up to summation order: a fold chunk sums its split-K slabs in a different
            # grouping, host-side (PyTorch) parameter ops scatter-add pointer slices with atomics,
            # and beyond two ranks the collective's own order depends on the message size
            err, tol = float((a - b).abs().max()), 1e-5 * float(a.abs().max()) + 1e-12
            assert err <= tol, f"{name} leaf {i}: overlapped / nvls != serial ({err:.3e} > {tol:.3e})"
        if rank == 0:
            print(f"dist overlap ok {name} K={units}: world {world}, {nbytes} bytes reduced in stages; "
                  f"{len(grads)} reduction modes agree", flush=True)


def main() -> None:
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    for name in ("qt8_cp_k4", "qg8_cp_k4", "qt8_tucker_k4", "qt8x4_cpt_k5", "rbt12_gaussian_k5"):
        g = Golden(name)
        cc = B200Circuit(g.plan).to(dev)
        with torch.no_grad():
            for p, v in zip(cc.leaves, g.leaves(torch.float32)):
                p.copy_(v)
        sharded = BatchShardedCircuit(cc)
        x, y = g.x(), g.y()
        n = x.shape[0]
        x_local = sharded.shard(x).to(dev)
        ll = sharded.log_likelihoods(x_local, n)
        err = (ll.double().cpu() - y).abs().max().item()
        tol = (5e-7 * y.abs() + 1e-5).min().item()
        assert ll.shape == y.shape and err <= tol, f"{name}: forward err {err:.3e} > {tol:.3e}"
        if x_local.shape[0] > 0:
            sharded.loss(x_local, n).backward()
        sharded.sync_gradients()
        for i, (p, gr) in enumerate(zip(cc.leaves, g.grads())):
            e = (p.grad.double().cpu() - gr).abs().max().item()
            assert e <= grad_tolerance(gr), f"{name} leaf {i}: grad err {e:.3e} > {grad_tolerance(gr):.3e}"
        if rank == 0:
            print(f"dist ok {name}: world {world}, batch {n}, forward err {err:.2e}", flush=True)
    overlap_check(dev, rank, world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
