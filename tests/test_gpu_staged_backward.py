"""The staged backward pass (data-parallel overlap) on one GPU.

With a gradient-sync hook installed, the backward pass runs the inner layers first and the fused
input step in fold chunks, calling the hook after each stage with the slices of the flat gradient
buffer that have become final.  Same kernels on the same data: the gradients equal the unstaged
pass up to the order in which a launch sums its split-K partials (the number of partial slabs of
a dW kernel follows the fold count of the launch, and a fold chunk is a smaller launch), and every
slice must already hold its final values when the hook sees it."""
import dataclasses

import pytest
import torch

from helpers import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


class _Recorder:
    """Stands in for OverlappedGradientReducer: snapshots every piece when it is announced."""

    average = False

    def __init__(self):
        self.calls, self.snaps, self.bytes = [], [], 0

    def __call__(self, pieces):
        pieces = list(pieces)
        self.calls.append([(p.data_ptr(), p.numel()) for p in pieces])
        self.snaps.append([(p, p.clone()) for p in pieces])
        self.bytes += sum(p.numel() * 4 for p in pieces)

    def finish(self):
        n, self.bytes = self.bytes, 0
        return n


@pytest.mark.parametrize("chunk_steps", [False, True])
@pytest.mark.parametrize("name,units,batch,chunks,bucket", [
    ("qt28_cp_k64", None, 256, 4, 8 << 20),
    ("qt28_cp_k64", None, 64, 4, 2 << 20),  # plain plan (batch < V / 2): chunked Categorical step
    ("qt8_cp_k4", None, 300, 3, 1 << 10),
    ("qt8_cp_k4", 64, 700, 5, 1 << 18),
    ("pd32_cp_k4", 16, 200, 4, 1 << 18),  # no fused input step: buckets over the inner layers
    ("qg8_cp_k4", None, 150, 2, 1 << 8),
])
def test_staged_backward_matches_unstaged(dev, name, units, batch, chunks, bucket, chunk_steps):
    from cirkit_b200 import B200Circuit

    g = Golden(name)
    plan = g.plan if units is None else dataclasses.replace(g.plan, meta={"units": 4}).with_units(units)
    V = max(int(s.config.get("num_categories", 0)) for s in plan.steps)
    x = torch.randint(0, V, (batch, plan.num_variables), generator=torch.Generator().manual_seed(5)).to(dev)

    def run(stage):
        cc = B200Circuit(plan, seed=11).to(dev)
        rec = None
        if stage:
            assert cc.runtime.enable_gradient_stages(chunks, bucket_bytes=bucket, chunk_steps=chunk_steps)
            rec = cc.runtime.grad_sync = _Recorder()
        y = cc(x)
        (-y.mean()).backward()
        return cc, y.detach().clone(), [p.grad.clone() for p in cc.leaves], rec

    _, y0, g0, _ = run(False)
    cc, y1, g1, rec = run(True)
    assert torch.equal(y0, y1)
    for i, (a, b) in enumerate(zip(g0, g1)):
        err, tol = float((a - b).abs().max()), 2e-6 * float(a.abs().max()) + 1e-12
        assert err <= tol, f"leaf {i}: staged backward differs, max {err:.3e} > {tol:.3e}"
    n_calls = len(rec.calls)
    # a second staged pass on the same data is bit-equal to the first (no atomics anywhere)
    for p in cc.leaves:
        p.grad = None
    (-cc(x).mean()).backward()
    for i, (a, p) in enumerate(zip(g1, cc.leaves)):
        assert torch.equal(a, p.grad), f"leaf {i}: staged backward is not reproducible"
    rt = cc.runtime
    which = rt.choose_plan(batch, False)
    assert n_calls == len(rt.grad_stages[which]) >= 2
    # the pieces tile the flat buffer exactly once ...
    flat = rt.last_flat_grad
    spans = sorted((p - flat.data_ptr(), n) for call in rec.calls[n_calls:] for p, n in call)
    pos = 0
    for off, n in spans:
        assert off == 4 * pos
        pos += n
    assert pos == flat.numel() and rt.last_synced_bytes == 4 * flat.numel()
    # ... and each one was final when announced (later stages do not touch it)
    for call in rec.snaps[n_calls:]:
        for piece, snap in call:
            assert torch.equal(piece, snap)


def test_hook_without_stages_syncs_once(dev):
    """A hook on a runtime whose stages were never prepared is called once, with the whole flat
    buffer, after the backward pass."""
    from cirkit_b200 import B200Circuit

    g = Golden("qt8_cp_k4")
    cc = B200Circuit(g.plan, seed=3).to(dev)
    rec = cc.runtime.grad_sync = _Recorder()
    x = torch.randint(0, 256, (16, g.plan.num_variables)).to(dev)
    (-cc(x).mean()).backward()
    assert len(rec.calls) == 1 and rec.calls[0][0][1] == cc.runtime.last_flat_grad.numel()
