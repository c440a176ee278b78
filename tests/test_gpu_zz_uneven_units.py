"""Sum layers whose input and output widths differ (Ki != Ko), through the generic FP32 kernels.

The circuits of BASELINE.json use one width K everywhere (plus the Ko = 1 root), but the reference
accepts any pair, e.g. `num_classes` = 10 outputs over K = 64 sum units
(cirkit/templates/data_modalities.py `image_data(num_classes=...)`;  TorchSumLayer
layers/inner.py:266-273, TorchCPTLayer layers/optimized.py:171-178).  The forward kernel sizes its
register tile by Ko, so every case with Ki > 32 * ceil(Ko / 32) has to take its per-sample gather:
round 1 found the batched gather silently dropping the tail of the reduction there.
"""
import numpy as np
import pytest
import torch

from helpers import grad_tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def uneven_plan(kind, H, K, Ko, V=16):
    from cirkit_b200.plan import CircuitPlan, LeafSpec, ParamSpec, StepSpec

    kred = H * K if kind == "sum" else K
    steps = [
        StepSpec("categorical", H, 1, 1, K, params={"probs": ParamSpec(0, [("softmax", {"dim": 1})], (H, K, V))},
                 scope_idx=np.arange(H, dtype=np.int32), config={"num_categories": V}),
        StepSpec(kind, 1, H, K, Ko, params={"weight": ParamSpec(1, [("softmax", {"dim": 1})], (1, Ko, kred))},
                 in_step=np.zeros((1, H), np.int32), in_fold=np.arange(H, dtype=np.int32).reshape(1, H)),
    ]
    return CircuitPlan(steps, [LeafSpec((H, K, V)), LeafSpec((1, Ko, kred))], np.array([1], np.int32),
                       np.array([0], np.int32), H, tuple(range(H)))


CASES = [("sum", 1, 64, 10), ("sum", 1, 48, 16), ("sum", 1, 128, 32), ("sum", 1, 10, 64),
         ("cpt", 2, 64, 10), ("cpt", 2, 100, 40), ("cpt", 3, 33, 32), ("cpt", 2, 128, 64),
         ("sum", 3, 40, 33), ("sum", 2, 64, 2)]


@pytest.mark.parametrize("kind,H,K,Ko", CASES)
def test_uneven_sum_layers_vs_oracle(kind, H, K, Ko, dev):
    from cirkit_b200 import B200Circuit
    from oracle import OracleCircuit

    plan = uneven_plan(kind, H, K, Ko)
    gen = torch.Generator().manual_seed(1000 * H + 10 * K + Ko)
    vals = [torch.randn(l.shape, generator=gen) * (3.0 if i == 0 else 1.0) for i, l in enumerate(plan.leaves)]
    cc, oc = B200Circuit(plan), OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for p, q, v in zip(cc.leaves, oc.leaves, vals):
            p.copy_(v)
            q.copy_(v.double())
    cc = cc.to(dev)
    for batch in (1, 37, 300):
        x = torch.randint(0, 16, (batch, H), generator=gen)
        for p, q in zip(cc.leaves, oc.leaves):
            p.grad = q.grad = None
        y, yo = cc(x.to(dev)), oc(x)
        assert y.shape == yo.shape
        err = (y.detach().double().cpu() - yo.detach()).abs().max().item()
        assert err <= 5e-7 * yo.abs().max().item() + 1e-5, f"B={batch}: {err:.3e}"
        # weight every output differently so that a dropped column cannot cancel
        w = torch.linspace(0.5, 1.5, y.numel(), dtype=torch.float64).reshape(yo.shape)
        (y * w.to(dev, y.dtype)).sum().backward()
        (yo * w).sum().backward()
        for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
            gerr = (p.grad.double().cpu() - q.grad).abs().max().item()
            tol = grad_tolerance(q.grad, gout_l1=float(w.abs().sum()))
            assert gerr <= tol, f"B={batch} leaf {i}: {gerr:.3e} > {tol:.3e}"


@pytest.mark.parametrize("units", [6, 8, 11, 12])
def test_tucker_mid_sizes_vs_oracle(units, dev):
    """Tucker layers between the fixture size (K = 4) and the tcgen05 size (K = 64): the Kronecker
    buffer + generic sum kernel route with a reduction length K*K in (32, 128] (K = 12: 144, the any-shape kernels) and K outputs, i.e.
    again beyond the Ko-sized tile of the batched gather (TorchTuckerLayer,
    layers/optimized.py:89-103)."""
    import dataclasses

    from cirkit_b200 import B200Circuit
    from cirkit_b200.plan import seeded_leaves
    from helpers import Golden
    from oracle import OracleCircuit
    from oracle.reference_eval import make_inputs

    g = Golden("qt8_tucker_k4")
    plan = dataclasses.replace(g.plan, meta={"units": 4}).with_units(units)
    cc = B200Circuit(plan, seed=9).to(dev)
    oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for q, v in zip(oc.leaves, seeded_leaves(plan, 9)):
            q.copy_(v)
    batch = 53
    x = make_inputs(plan, batch, seed=units)
    y, yo = cc(x.to(dev)), oc(x)
    assert y.shape == yo.shape
    err = (y.detach().double().cpu() - yo.detach()).abs().max().item()
    assert err <= 5e-7 * yo.abs().max().item() + 1e-5, f"{err:.3e}"
    w = torch.randn(batch, 1, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    (y * w.to(dev, torch.float32)).sum().backward()
    (yo * w).sum().backward()
    ll_max = float(yo.detach().abs().max())
    for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
        gerr = (p.grad.double().cpu() - q.grad).abs().max().item()
        tol = grad_tolerance(q.grad, gout_l1=float(w.abs().sum()), ll_max=ll_max)
        assert gerr <= tol, f"leaf {i}: {gerr:.3e} > {tol:.3e}"
