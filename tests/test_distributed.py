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


def _reducer_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    from cirkit_b200.distributed import OverlappedGradientReducer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat = torch.arange(40.0, dtype=torch.float64) * (rank + 1)
        red = OverlappedGradientReducer(average=False)
        red([flat[24:40]])  # stage 0: the tail of the buffer is final first
        red([flat[0:8], flat[16:24]])
        red([flat[8:16], flat[0:0]])  # an empty piece is skipped on every rank alike
        nbytes = red.finish()
        avg = torch.ones(6) * (rank + 1)
        red2 = OverlappedGradientReducer(average=True)
        red2([avg])
        red2.finish()
        torch.save({"flat": flat, "nbytes": nbytes, "avg": avg}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_overlapped_reducer_sums_pieces_stage_by_stage(tmp_path):
    """The reducer the staged CUDA backward calls once per stage: every piece is summed over the
    ranks by its own asynchronous all-reduce, `finish()` waits for all of them and returns the
    bytes reduced (2 gloo ranks)."""
    world = 2
    mp.spawn(_reducer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = torch.arange(40.0, dtype=torch.float64) * 3
    for r in range(world):
        o = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(o["flat"], want)
        assert o["nbytes"] == 40 * 8
        assert torch.equal(o["avg"], torch.full((6,), 1.5))


def test_gradient_stages_partition_the_backward_pass():
    """`PlanRuntime.enable_gradient_stages` (host logic only, no device): the staged execution
    plan runs every step exactly once, its parameter-op ranges partition the op list, and the
    gradient pieces of the stages tile the flat gradient buffer without gaps or overlaps."""
    import numpy as np

    from cirkit_b200.runtime import STEP_TABLE_DENSE, PlanRuntime, _stage_pieces

    g = Golden("qt28_cp_k64")
    rt = PlanRuntime(g.plan)
    assert rt.enable_gradient_stages(4, chunk_steps=True)
    fused, sync = rt.exec_plans["fused"], rt.exec_plans["fused_sync"]
    td = [es for es in fused if es.kind == STEP_TABLE_DENSE]
    assert len(td) == 1 and len(sync) == len(fused) + 3
    F = g.plan.steps[td[0].out_sid].num_folds
    chunks = [es for es in sync if es.folds is not None]
    assert [es.folds for es in chunks] == [(c * F // 4, (c + 1) * F // 4) for c in range(4)]
    # step ranges: disjoint, cover the list; backward order = inner layers first, then the chunks
    stages = rt.grad_stages["fused"]
    covered = sorted(i for st in stages for i in range(*st.steps))
    assert covered == list(range(len(sync)))
    assert stages[0].steps == (4, len(sync)) and len(stages) == 5
    ops_cov = sorted(i for st in stages for i in range(*st.ops))
    assert ops_cov == list(range(len(rt.sync_ops["fused_sync"])))
    # alias slots point into the base buffers at the fold offset
    V, K = 256, 64
    for es in chunks:
        f0 = es.folds[0]
        for slot, per_fold in zip(es.slots, (V * K, K * K, V * K)):
            if f0 == 0:
                assert slot not in rt.aliases
            else:
                assert rt.aliases[slot][1] == f0 * per_fold
    # the pieces of all stages tile the flat buffer (as laid out by _grad_table)
    sizes = [-(-int(np.prod(b.src_shape)) // 4) * 4 for b in rt.bindings]
    offs = list(np.cumsum([0] + sizes[:-1]))
    flat = torch.zeros(sum(sizes))
    for st in stages:
        for piece in _stage_pieces(rt, st, flat, offs):
            piece += 1
    assert torch.equal(flat, torch.ones_like(flat))
    # the bulk of the bytes is in the chunk stages, the inner layers' weights come first
    first = sum(p.numel() for p in _stage_pieces(rt, stages[0], flat, offs))
    assert first < 0.2 * flat.numel()


def test_gradient_stages_default_cuts_only_the_parameter_ops():
    """Default staging: the layers' own backward kernels stay whole launches; only the parameter
    ops of the large input tensors (Categorical table, first sum weights) go out in fold ranges."""
    import numpy as np

    from cirkit_b200.runtime import PlanRuntime, _stage_pieces

    g = Golden("qt28_cp_k64")
    rt = PlanRuntime(g.plan)
    assert rt.enable_gradient_stages(4)
    fused, sync, stages = rt.exec_plans["fused"], rt.exec_plans["fused_sync"], rt.grad_stages["fused"]
    assert len(sync) == len(fused) and all(es.folds is None for es in sync)
    assert [st.steps for st in stages] == [(1, len(sync)), (0, 1)] + [(0, 0)] * 4
    assert sorted(i for st in stages for i in range(*st.ops)) == list(range(len(rt.sync_ops["fused_sync"])))
    assert all(len(st.pieces) == 2 for st in stages[2:])  # table rows + sum-weight rows of a fold range
    sizes = [-(-int(np.prod(b.src_shape)) // 4) * 4 for b in rt.bindings]
    offs = list(np.cumsum([0] + sizes[:-1]))
    flat = torch.zeros(sum(sizes))
    for st in stages:
        for piece in _stage_pieces(rt, st, flat, offs):
            piece += 1
    assert torch.equal(flat, torch.ones_like(flat))


@pytest.mark.parametrize("name,which", [("pd32_cp_k4", "plain"), ("qt28_cp_k64", "plain"),
                                        ("qg8_cp_k4", "plain"), ("rbt12_gaussian_k5", "plain")])
def test_gradient_stages_of_other_plans(name, which):
    """Bucketed stages for plans without a fused input step (PoonDomingos: the Categorical table
    feeds a Hadamard layer) and for the plain plan of small batches: steps and ops are partitioned,
    the pieces tile the flat buffer, inner layers are cut into several stages when large."""
    import numpy as np

    from cirkit_b200.runtime import PlanRuntime, _stage_pieces

    g = Golden(name)
    import dataclasses

    plan = dataclasses.replace(g.plan, meta={"units": 4}).with_units(32) if name == "pd32_cp_k4" else g.plan
    rt = PlanRuntime(plan)
    assert rt.enable_gradient_stages(4, bucket_bytes=1 << 20 if name != "qg8_cp_k4" else 1 << 10,
                                     chunk_steps=name == "pd32_cp_k4")
    stages, sync = rt.grad_stages[which], rt.exec_plans[which + "_sync"]
    assert sorted(i for st in stages for i in range(*st.steps)) == list(range(len(sync)))
    assert sorted(i for st in stages for i in range(*st.ops)) == list(range(len(rt.sync_ops[which + "_sync"])))
    # backward order: every stage lies below the previous one in the step list
    his = [st.steps[1] for st in stages]
    assert his == sorted(his, reverse=True)
    sizes = [-(-int(np.prod(b.src_shape)) // 4) * 4 for b in rt.bindings]
    offs = list(np.cumsum([0] + sizes[:-1]))
    flat = torch.zeros(sum(sizes))
    for st in stages:
        for piece in _stage_pieces(rt, st, flat, offs):
            piece += 1
    assert torch.equal(flat, torch.ones_like(flat))
    if name == "pd32_cp_k4":
        assert len(stages) > 5 and any(es.folds is not None for es in sync)


def test_nvls_reducer_is_a_no_op_in_a_single_process():
    """World size 1: the in-switch reducer owns no buffer (the runtime keeps writing into its own
    flat tensor) and `finish` only reports the bytes it was told about."""
    from cirkit_b200.distributed import NvlsGradientReducer

    red = NvlsGradientReducer()
    assert red.alloc(1000, torch.device("cpu")) is None
    red([torch.zeros(10), torch.zeros(6)])
    assert red.finish() == 64
    assert red.finish() == 0


def test_graph_policy_follows_the_arena_size():
    """CUDA-graph replay is on where the host is the bottleneck (activation arenas up to 256 MB)
    and off for plans with per-call PyTorch inputs (host logic only, no device)."""
    from cirkit_b200.runtime import PlanRuntime

    rt = PlanRuntime(Golden("qt28_cp_k64").plan)
    assert rt.use_graphs == "auto"
    assert rt.graphs_for(256) and not rt.graphs_for(2048)  # 154 MB / 1.23 GB of activations
    rt.use_graphs = False
    assert not rt.graphs_for(8)
    rt.use_graphs = True
    assert rt.graphs_for(8192)
