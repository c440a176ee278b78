"""CUDA-graph replay of repeated plan calls (CKB_USE_GRAPHS): a training loop whose buffers come
back at the same addresses runs its forward and backward launches as one graph launch each from
the third step on; results must be bit-equal to the eager path, and a change of any argument
(new batch size, new parameter storage) must fall back to eager launches."""
import pytest
import torch

from helpers import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _loop(cc, xs, steps):
    outs = []
    for i in range(steps):
        for p in cc.leaves:
            p.grad = None
        y = cc(xs[i % len(xs)])
        (-y.mean()).backward()
        outs.append((y.detach().clone(), [p.grad.clone() for p in cc.leaves]))
    return outs


@pytest.mark.parametrize("name,batch", [("qt28_cp_k64", 256), ("qt8_cp_k4", 300), ("qt8_tucker_k4", 64),
                                        ("rbt12_gaussian_k5", 40), ("pd6_cp_k3_unopt", 33)])
def test_graph_replay_is_bit_equal_to_eager(dev, name, batch):
    from cirkit_b200 import B200Circuit
    from oracle.reference_eval import make_inputs

    g = Golden(name)
    xs = [make_inputs(g.plan, batch, seed=s).to(dev) for s in (1, 2)]
    res = {}
    for graphs in (False, True):
        cc = B200Circuit(g.plan, seed=9).to(dev)
        cc.runtime.use_graphs = graphs
        res[graphs] = _loop(cc, xs, 6)
    for (y0, g0), (y1, g1) in zip(res[False], res[True]):
        assert torch.equal(y0, y1)
        for a, b in zip(g0, g1):
            assert torch.equal(a, b)
    # steps with the same input agree among themselves too (replays are not stale)
    assert torch.equal(res[True][0][0], res[True][2][0]) and torch.equal(res[True][1][0], res[True][3][0])


def test_changed_arguments_fall_back_to_eager(dev):
    from cirkit_b200 import B200Circuit
    from oracle import OracleCircuit
    from oracle.reference_eval import make_inputs

    g = Golden("qt8_cp_k4")
    cc = B200Circuit(g.plan, seed=4).to(dev)
    assert cc.runtime.use_graphs
    oc = OracleCircuit(g.plan, dtype=torch.float64)
    with torch.no_grad():
        for p, v in zip(oc.leaves, cc.leaves):
            p.copy_(v.double().cpu())
    for batch in (200, 200, 200, 130, 200, 130, 130, 130):  # shapes alternate: new keys, then replays
        x = make_inputs(g.plan, batch, seed=batch)
        with torch.no_grad():
            y = cc(x.to(dev))
            yo = oc(x)
        assert (y.double().cpu() - yo).abs().max().item() <= 5e-7 * yo.abs().max().item() + 1e-5
    # an optimiser step changes the parameter VALUES (same storage): replays must see them
    opt = torch.optim.SGD(cc.parameters(), lr=0.5)
    x = make_inputs(g.plan, 200, seed=1).to(dev)
    prev = None
    for _ in range(5):
        opt.zero_grad()
        y = cc(x)
        (-y.mean()).backward()
        opt.step()
        if prev is not None:
            assert float(y.mean()) > prev  # the likelihood of the batch goes up step after step
        prev = float(y.mean())
