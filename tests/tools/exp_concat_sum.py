"""Focused check of concatenating sum layers (TorchSumLayer, arity H > 1) against the oracle."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
from cirkit_b200 import B200Circuit
from cirkit_b200.plan import CircuitPlan, StepSpec, ParamSpec, LeafSpec
from oracle import OracleCircuit

dev = torch.device("cuda:0")
def make(H, K, Ko, spread):
    steps = [StepSpec("categorical", H, 1, 1, K, params={"probs": ParamSpec(0, [("softmax", {"dim": 1})], (H, K, 16))},
                      scope_idx=np.arange(H, dtype=np.int32), config={"num_categories": 16}),
             StepSpec("sum", 1, H, K, Ko, params={"weight": ParamSpec(1, [("softmax", {"dim": 1})], (1, Ko, H * K))},
                      in_step=np.zeros((1, H), np.int32), in_fold=np.arange(H, dtype=np.int32).reshape(1, H))]
    leaves = [LeafSpec((H, K, 16)), LeafSpec((1, Ko, H * K))]
    plan = CircuitPlan(steps, leaves, np.array([1], np.int32), np.array([0], np.int32), H, tuple(range(H)))
    plan.validate()
    return plan
for H, K, Ko, spread in [(2, 4, 4, 1.0), (8, 4, 4, 1.0), (9, 4, 4, 1.0), (13, 4, 4, 1.0), (13, 4, 4, 300.0), (13, 4, 4, 3000.0), (3, 40, 4, 300.0)]:
    plan = make(H, K, Ko, spread)
    g = torch.Generator().manual_seed(H)
    vals = [torch.randn(l.shape, generator=g) for l in plan.leaves]
    vals[0] = vals[0] * spread   # widely different magnitudes between the concatenated inputs
    cc = B200Circuit(plan); oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for p, q, v in zip(cc.leaves, oc.leaves, vals):
            p.copy_(v); q.copy_(v.double())
    cc = cc.to(dev)
    x = torch.randint(0, 16, (37, H), generator=g)
    y = cc(x.to(dev)); yo = oc(x)
    (-y.mean()).backward(); (-yo.mean()).backward()
    err = (y.detach().double().cpu() - yo.detach()).abs().max().item()
    gerr = max((p.grad.double().cpu() - q.grad).abs().max().item() for p, q in zip(cc.leaves, oc.leaves))
    print(f"H={H} K={K} Ko={Ko} spread={spread}: |y| {yo.abs().max().item():.1f} fwd err {err:.3e} grad err {gerr:.3e}")
