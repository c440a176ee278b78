"""Batch-sharded evaluation on two gloo ranks, CPU only.

The sharding helper is backend-agnostic; here it wraps the oracle circuit (test infrastructure) so
that the host-side logic — row blocks, the single all-gather of the root log-densities, the
gradient all-reduce — is checked against a single-process evaluation of the whole batch.
"""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cirkit_b200.distributed import BatchShardedCircuit, all_gather_rows, shard_rows
from helpers import Golden
from oracle.reference_eval import OracleCircuit, make_inputs


def test_shard_rows_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 1000, 2049):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def test_single_process_is_identity():
    ll = torch.arange(6.0).reshape(6, 1, 1)
    assert all_gather_rows(ll, 6) is ll
    with pytest.raises(ValueError):
        all_gather_rows(ll, 7)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, name: str, batch: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        g = Golden(name)
        oc = OracleCircuit(g.plan, dtype=torch.float64)
        with torch.no_grad():
            for p, v in zip(oc.leaves, g.leaves(torch.float64)):
                # rank 1 starts from different values: broadcast_parameters must fix that
                p.copy_(v if rank == 0 else v + 1.0)
        sharded = BatchShardedCircuit(oc)
        sharded.broadcast_parameters(src=0)
        x = make_inputs(g.plan, batch, seed=7)  # same global batch on every rank
        x_local = sharded.shard(x)
        ll_all = sharded.log_likelihoods(x_local, batch)
        sharded.loss(x_local, batch).backward()
        nbytes = sharded.sync_gradients()
        torch.save(
            {"ll": ll_all, "grads": [p.grad for p in oc.leaves], "rows": sharded.local_rows(batch),
             "nbytes": nbytes},
            os.path.join(out_dir, f"rank{rank}.pt"),
        )
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,batch", [("qt8_cp_k4", 16), ("qg8_cp_k4", 9), ("gmm1d_k8", 5)])
def test_two_ranks_match_single_process(tmp_path, name, batch):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, batch, str(tmp_path)), nprocs=world, join=True)
    g = Golden(name)
    oc = OracleCircuit(g.plan, dtype=torch.float64)
    with torch.no_grad():
        for p, v in zip(oc.leaves, g.leaves(torch.float64)):
            p.copy_(v)
    x = make_inputs(g.plan, batch, seed=7)
    ll = oc(x)
    (-ll.mean()).backward()
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    assert outs[0]["rows"][1] == outs[1]["rows"][0] and outs[1]["rows"][1] == batch
    for o in outs:
        # every rank sees the full batch of root log-densities, in row order
        assert o["ll"].shape == ll.shape
        torch.testing.assert_close(o["ll"], ll.detach(), rtol=0, atol=1e-12)
        # summed replica gradients = gradient of the global-batch mean NLL (fp64: round-off only)
        for gr, p in zip(o["grads"], oc.leaves):
            if p.requires_grad:
                torch.testing.assert_close(gr, p.grad, rtol=1e-10, atol=1e-12)
        assert o["nbytes"] == sum(p.numel() * 8 for p in oc.leaves if p.requires_grad)


def test_flat_gradient_detection():
    """The CUDA runtime hands out parameter gradients as views of one flat buffer; the sharding
    helper sums that buffer with ONE collective when (and only when) every gradient lives in it."""
    from cirkit_b200.distributed import _flat_gradient, all_reduce_gradients

    shapes = [(3, 4, 5), (7,), (2, 6)]
    sizes = [-(-torch.Size(s).numel() // 4) * 4 for s in shapes]
    flat = torch.arange(float(sum(sizes)))
    params, off = [], 0
    for s, sz in zip(shapes, sizes):
        p = torch.nn.Parameter(torch.zeros(s, dtype=flat.dtype))
        p.grad = flat[off : off + torch.Size(s).numel()].view(s)
        off += sz
        params.append(p)
    assert _flat_gradient(params, flat) is flat
    assert _flat_gradient(params, None) is None
    assert _flat_gradient(params, torch.zeros(4)) is None  # some other buffer
    before = [p.grad.clone() for p in params]
    assert all_reduce_gradients(params, flat=flat) == flat.numel() * flat.element_size()  # world 1: no-op
    for p, b in zip(params, before):
        assert torch.equal(p.grad, b)
    # a gradient that was re-allocated (e.g. accumulated out of place) disables the short cut
    params[1].grad = params[1].grad.clone()
    assert _flat_gradient(params, flat) is None
    # ... and so does a buffer that holds more than these parameters' gradients
    params[1].grad = flat[sizes[0] : sizes[0] + 7]
    assert _flat_gradient(params[:2], flat) is None
