"""GPU parity of the Tucker tcgen05 kernels (Ki = Ko = 64; `cirkit_b200/csrc/tucker_tc.cu`) against
the float64 oracle and against the FP32 SIMT route, through the C ABI.  Reference layer:
TorchTuckerLayer, cirkit/backend/torch/layers/optimized.py:89-103.

Tolerances: the ones of test_gpu_parity.py (forward 5e-7*|ll| + 1e-5, gradients
helpers.grad_tolerance)."""
import dataclasses

import pytest
import torch

from helpers import Golden, grad_tolerance

pytestmark = pytest.mark.gpu
FWD_RTOL, FWD_ATOL = 5e-7, 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def plan64():
    """QuadTree 8x8 Tucker structure of the reference fixture, resized to K = 64."""
    g = Golden("qt8_tucker_k4")
    return dataclasses.replace(g.plan, meta={"units": 4}).with_units(64)


def _pair(plan, dev):
    from cirkit_b200 import B200Circuit
    from cirkit_b200.plan import seeded_leaves
    from oracle import OracleCircuit

    cc = B200Circuit(plan, seed=1).to(dev)
    oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for q, v in zip(oc.leaves, seeded_leaves(plan, 1)):
            q.copy_(v)
    return cc, oc


@pytest.mark.parametrize("batch", [5, 256, 300, 700])
def test_tucker_tc_vs_oracle(batch, plan64, dev):
    from oracle.reference_eval import make_inputs

    cc, oc = _pair(plan64, dev)
    x = make_inputs(plan64, batch, seed=batch)
    y = cc(x.to(dev))
    yo = oc(x)
    err = (y.detach().double().cpu() - yo.detach()).abs()
    tol = FWD_RTOL * yo.detach().abs() + FWD_ATOL
    assert bool((err <= tol).all()), f"forward err {err.max().item():.3e}"
    w = torch.randn(batch, 1, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    (y * w.to(dev, torch.float32)).sum().backward()
    (yo * w).sum().backward()
    ll_max = float(yo.detach().abs().max())
    for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
        e = (p.grad.double().cpu() - q.grad).abs().max().item()
        tol = grad_tolerance(q.grad, gout_l1=float(w.abs().sum()), ll_max=ll_max,
                             w_max=float(torch.softmax(q.detach(), dim=-1).max()))
        assert e <= tol, f"leaf {i}: {e:.3e} > {tol:.3e}"


def test_tucker_tc_matches_simt(plan64, dev):
    """Same circuit on the FP32 SIMT route (Kronecker scratch + dense block)."""
    from cirkit_b200 import _lib
    from oracle.reference_eval import make_inputs

    cc, _ = _pair(plan64, dev)
    x = make_inputs(plan64, 130, seed=9).to(dev)
    lib = _lib.load()
    res = []
    try:
        for on in (1, 0):
            assert lib.ckb_set_option(_lib.OPT_TENSOR_CORES, on) == 0
            for p in cc.leaves:
                p.grad = None
            y = cc(x)
            (-y.mean()).backward()
            res.append((y.detach().clone(), [p.grad.clone() for p in cc.leaves]))
    finally:
        lib.ckb_set_option(_lib.OPT_TENSOR_CORES, 1)
    (yt, gt), (ys, gs) = res
    assert torch.isfinite(yt).all()
    assert (yt.double() - ys.double()).abs().max().item() <= FWD_RTOL * ys.abs().max().item() + FWD_ATOL
    for i, (a, b) in enumerate(zip(gt, gs)):
        e = (a.double() - b.double()).abs().max().item()
        assert e <= grad_tolerance(b, ll_max=float(ys.abs().max())), f"leaf {i}: {e:.3e}"


def test_tucker_determinism_and_row_independence(plan64, dev):
    from oracle.reference_eval import make_inputs

    cc, _ = _pair(plan64, dev)
    x = make_inputs(plan64, 513, seed=2).to(dev)
    with torch.no_grad():
        y = cc(x)
        assert torch.equal(cc(x), y)
        assert torch.equal(cc(x[:7]), y[:7])
        assert torch.equal(cc(x[256:300]), y[256:300])


def test_full_size_properties(dev):
    """BASELINE.json configs[2] (QuadTree 28x28, Tucker, K=64, B=2048): size-independent properties
    plus the reference's own values on the fixture rows embedded in the batch."""
    from cirkit_b200 import B200Circuit, IntegrateQuery

    g = Golden("qt28_tucker_k64")
    cc = B200Circuit(g.plan)
    with torch.no_grad():
        for p, v in zip(cc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    cc = cc.to(dev)
    B = 2048
    x = torch.randint(0, 256, (B, 784), generator=torch.Generator().manual_seed(0))
    n_fix = g.x().shape[0]
    x[:n_fix] = g.x()
    xd = x.to(dev)
    y = cc(xd)
    # (1) reference values (float64) of the fixture rows
    ref = g.y()
    err = (y[:n_fix].detach().double().cpu() - ref).abs()
    assert bool((err <= FWD_RTOL * ref.abs() + FWD_ATOL).all()), f"forward err {err.max().item():.3e}"
    # (2) run-to-run determinism
    assert torch.equal(cc(xd), y)
    # (3) a normalised circuit integrates to one
    with torch.no_grad():
        z = IntegrateQuery(cc)(xd[:64], integrate_vars=torch.ones(1, 784, dtype=torch.bool))
    assert z.abs().max().item() < 1e-3
    # (4) every row of d(theta) of a softmax-parameterised weight sums to zero
    (-y.mean()).backward()
    for p in cc.leaves:
        assert torch.isfinite(p.grad).all()
        assert p.grad.double().sum(dim=-1).abs().max().item() < 2e-6
    # (5) the gradient of the mean log-likelihood is linear in the batch
    grads = [p.grad.clone() for p in cc.leaves]
    halves = []
    for sl in (slice(0, B // 2), slice(B // 2, B)):
        for p in cc.leaves:
            p.grad = None
        (-cc(xd[sl]).mean()).backward()
        halves.append([p.grad.clone() for p in cc.leaves])
    for ga, h0, h1 in zip(grads, *halves):
        half = 0.5 * (h0.double() + h1.double())
        assert (half - ga.double()).abs().max().item() <= max(2e-6, 1e-4 * ga.abs().max().item())
