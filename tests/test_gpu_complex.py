"""'complex-lse-sum' circuits through the plan executor (BASELINE.json
configs[4]): forward values and gradients against the reference's own outputs (fixtures of kind
"complex", generated from cirkit's torch backend in float64 / complex128), and against the oracle
on larger seeded batches at K = 64.

Complex logarithms are compared through exp(y - y_ref) = 1 (the imaginary part is a phase, defined
up to 2 pi: tests/backend/torch/test_compile_circuit_operators.py:236-238 does the same)."""
import dataclasses

import pytest
import torch

from helpers import Golden, golden_names

pytestmark = pytest.mark.gpu
COMPLEX = golden_names("complex")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _close_log(y, y_ref, tol=2e-5):
    y, y_ref = y.detach().cpu().to(torch.complex128), y_ref.detach().to(torch.complex128)
    assert y.shape == y_ref.shape
    err = (torch.exp(y - y_ref) - 1).abs().max().item()
    assert err <= tol, f"exp(y - y_ref) - 1 = {err:.3e}"


def _close_grad(g, g_ref, what):
    g_ref = g_ref.to(torch.complex128 if g_ref.is_complex() else torch.float64)
    g = g.detach().cpu().to(g_ref.dtype)
    err = (g - g_ref).abs().max().item()
    tol = max(2e-6, 1e-4 * g_ref.abs().max().item())
    assert err <= tol, f"{what}: {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("name", COMPLEX)
def test_complex_fixtures_vs_reference(name, dev):
    from cirkit_b200 import B200Circuit

    g = Golden(name)
    assert g.plan.semiring == "complex-lse-sum"
    cc = B200Circuit(g.plan)
    with torch.no_grad():
        for p, v in zip(cc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    cc = cc.to(dev)
    y = cc(g.x().to(dev))
    assert y.is_complex()
    _close_log(y, g.y())
    # the SoS loss differentiates the real part (notebooks/sum-of-squares-circuits.ipynb cell 32)
    (-y.real.mean()).backward()
    for i, (p, gr) in enumerate(zip(cc.leaves, g.grads())):
        _close_grad(torch.zeros_like(p) if p.grad is None else p.grad, gr, f"leaf {i}")


@pytest.mark.parametrize("name,batch", [("rbt16_cpt_k4_complex", 300), ("rbt16_cpt_k4_complex_conj", 129),
                                        ("rbt12_cp_k3_complex_unopt", 70)])
def test_complex_k64_vs_oracle(name, batch, dev):
    """The configs[4] shape: the reference structures resized to K = 64 units, seeded complex
    leaves, ragged batches, against the complex128 oracle; loss 2 Re c(x) as in the SoS objective."""
    from cirkit_b200 import B200Circuit
    from cirkit_b200.plan import seeded_leaves
    from oracle import OracleCircuit

    g = Golden(name)
    k0 = g.plan.steps[0].num_output_units
    plan = dataclasses.replace(g.plan, meta={"units": k0}).with_units(64)
    vals = seeded_leaves(plan, 3)
    cc, oc = B200Circuit(plan), OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for p, q, v in zip(cc.leaves, oc.leaves, vals):
            p.copy_(v)
            q.copy_(v.to(q.dtype))
    cc = cc.to(dev)
    x = torch.randint(0, 16, (batch, plan.num_variables), generator=torch.Generator().manual_seed(batch))
    y, yo = cc(x.to(dev)), oc(x)
    _close_log(y, yo, tol=1e-4)
    (-2 * y.real.mean()).backward()
    (-2 * yo.real.mean()).backward()
    for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
        _close_grad(p.grad, q.grad, f"leaf {i}")
    # determinism: no atomics on this path
    g1 = [p.grad.clone() for p in cc.leaves]
    for p in cc.leaves:
        p.grad = None
    (-2 * cc(x.to(dev)).real.mean()).backward()
    assert all(torch.equal(a, p.grad) for a, p in zip(g1, cc.leaves))
