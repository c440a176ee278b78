"""tcgen05 kernels for Ki = Ko = 128 (csrc/dense128_tc.cu; BASELINE.json configs[3]), selected by
bit 9 of CKB_OPT_TC_FAST_MATH (default on since round 2, first validated on the B200 in round 2):
forward and gradients against the float64 oracle, and against the FP32 SIMT route (bit 9 off)."""
import dataclasses
import os

import pytest
import torch

from helpers import Golden

pytestmark = pytest.mark.gpu
OPT_TC_FAST_MATH = 1


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name,batch", [("qt8_cp_k4", 1), ("qt8_cp_k4", 128), ("qt8_cp_k4", 300),
                                        ("qg8_cp_k4_densemix", 257), ("pd6_cp_k3_unopt", 45)])
def test_dense128_vs_simt_and_oracle(name, batch, dev):
    from cirkit_b200 import B200Circuit, _lib
    from helpers import grad_tolerance
    from cirkit_b200.plan import seeded_leaves
    from oracle import OracleCircuit
    from oracle.reference_eval import make_inputs

    g = Golden(name)
    k0 = g.plan.steps[0].num_output_units
    plan = dataclasses.replace(g.plan, meta={"units": k0}).with_units(128)
    assert any(s.kind in ("cpt", "sum") and s.num_input_units == 128 and s.num_output_units == 128
               for s in plan.steps)
    cc = B200Circuit(plan, seed=5).to(dev)
    oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for q, v in zip(oc.leaves, seeded_leaves(plan, 5)):
            q.copy_(v)
    x = make_inputs(plan, batch, seed=batch)
    yo = oc(x)
    w = torch.randn(batch, 1, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    (yo * w).sum().backward()
    yo = yo.detach()
    lib = _lib.load()
    res = {}
    try:
        # the tcgen05 route first: the workspace is sized when a batch size is first seen
        for bits in (3 | 512, 3):
            assert lib.ckb_set_option(OPT_TC_FAST_MATH, bits) == 0
            for p in cc.leaves:
                p.grad = None
            y = cc(x.to(dev))
            (y * w.to(dev, torch.float32)).sum().backward()
            res[bits] = (y.detach().double().cpu(), [p.grad.double().cpu() for p in cc.leaves])
    finally:
        lib.ckb_set_option(OPT_TC_FAST_MATH, 3 | 512)
    tol = 5e-7 * yo.abs().max().item() + 1e-5
    assert torch.isfinite(res[3 | 512][0]).all()
    assert (res[3][0] - yo).abs().max().item() <= tol  # the SIMT route (sanity)
    err = (res[3 | 512][0] - yo).abs().max().item()
    assert err <= tol, f"tcgen05 K=128 forward vs oracle: {err:.3e} > {tol:.3e}"
    ll_max = float(yo.abs().max())
    for i, (got, q) in enumerate(zip(res[3 | 512][1], oc.leaves)):
        gr = torch.zeros_like(q) if q.grad is None else q.grad
        gerr = (got - gr).abs().max().item()
        gtol = grad_tolerance(gr, gout_l1=float(w.abs().sum()), ll_max=ll_max)
        assert gerr <= gtol, f"tcgen05 K=128 backward, leaf {i}: {gerr:.3e} > {gtol:.3e}"
