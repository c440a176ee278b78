"""GPU parity: the CUDA path (through the C ABI) against the reference's outputs (golden fixtures,
float64) and against the CPU oracle on the same seeded inputs.

Stated tolerances (derived from the reference's own fp32-vs-fp64 gap):
  forward    |ll - ll_ref64| <= 5e-7 * |ll_ref64| + 1e-5        (2.2e-3 at ll = -4357)
  gradients  per parameter tensor  max|g - g_ref64| <= max(2e-6, 1e-4 * max|g_ref64|)
             (see helpers.grad_tolerance for why the floor is 2e-6)
"""
import math

import pytest
import torch

from helpers import Golden, golden_names, grad_tolerance

pytestmark = pytest.mark.gpu

FULL = golden_names("full")
SEEDED = golden_names("seeded")
FWD_RTOL, FWD_ATOL = 5e-7, 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _circuit(g: Golden, dev):
    from cirkit_b200 import B200Circuit

    cc = B200Circuit(g.plan)
    with torch.no_grad():
        for p, v in zip(cc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    return cc.to(dev)


def _check_forward(y, y_ref):
    y = y.detach().double().cpu()
    assert y.shape == y_ref.shape
    err = (y - y_ref).abs()
    tol = FWD_RTOL * y_ref.abs() + FWD_ATOL
    assert bool((err <= tol).all()), f"max err {err.max().item():.3e} (tol {tol.max().item():.3e})"


@pytest.mark.parametrize("name", FULL)
def test_forward_backward_vs_reference(name, dev):
    g = Golden(name)
    cc = _circuit(g, dev)
    y = cc(g.x().to(dev))
    _check_forward(y, g.y())
    (-y.mean()).backward()
    for i, (p, gr) in enumerate(zip(cc.leaves, g.grads())):
        got = torch.zeros_like(p) if p.grad is None else p.grad
        err = (got.double().cpu() - gr).abs().max().item()
        assert err <= grad_tolerance(gr), f"leaf {i}: grad err {err:.3e} > {grad_tolerance(gr):.3e}"


@pytest.mark.parametrize("name", [n for n in FULL if Golden(n).mask()[0] is not None])
def test_integrate_query_vs_reference(name, dev):
    from cirkit_b200 import IntegrateQuery

    g = Golden(name)
    cc = _circuit(g, dev)
    mask, y_mask = g.mask()
    with torch.no_grad():
        y = IntegrateQuery(cc)(g.x().to(dev), integrate_vars=mask.to(dev))
    _check_forward(y, y_mask)


def test_known_answers(dev):
    """The reference's hand-computed values, tests/symbolic/test_utils.py:411-417 and :497-503."""
    from cirkit_b200 import IntegrateQuery

    g = Golden("ka_categorical_cpt")
    cc = _circuit(g, dev)
    with torch.no_grad():
        y = cc(g.x().to(dev)).reshape(-1).double().cpu()
        for bits, val in g.meta["evi"].items():
            assert math.isclose(y[int(bits, 2)].exp().item(), val, rel_tol=2e-6)
        assert math.isclose(torch.logsumexp(y, 0).exp().item(), 318.0, rel_tol=2e-6)
        mask, _ = g.mask()
        ym = IntegrateQuery(cc)(g.x().to(dev), integrate_vars=mask.to(dev)).reshape(-1).double().cpu()
        assert math.isclose(ym[int("10110", 2)].exp().item(), 16.845, rel_tol=2e-6)
    gz = Golden("ka_categorical_cpt_Z")
    cz = _circuit(gz, dev)
    with torch.no_grad():
        z = cz()
    assert z.shape == gz.y().shape
    assert math.isclose(z.double().exp().item(), 318.0, rel_tol=2e-6)

    g = Golden("ka_gaussian")
    cc = _circuit(g, dev)
    with torch.no_grad():
        y = cc(g.x().to(dev)).reshape(-1).double().cpu()
        assert math.isclose(y[0].exp().item(), 3.744904862456293, rel_tol=2e-6)
        mask, _ = g.mask()
        ym = IntegrateQuery(cc)(g.x().to(dev), integrate_vars=mask.to(dev)).reshape(-1).double().cpu()
        assert math.isclose(ym[1].exp().item(), 23.528960785605985, rel_tol=2e-6)
        assert math.isclose(ym[3].exp().item(), 44.0, rel_tol=2e-6)


@pytest.mark.parametrize("name", SEEDED)
def test_benchmark_circuits_vs_reference(name, dev):
    """Benchmark-size circuits (QuadTree 28x28, K=32/64): outputs and gradient summaries of the
    real reference (float64) on leaves re-drawn from the fixture seed."""
    g = Golden(name)
    cc = _circuit(g, dev)
    y = cc(g.x().to(dev))
    _check_forward(y, g.y())
    (-y.mean()).backward()
    ll_max = float(g.y().abs().max())
    dense = 0
    for i, p in enumerate(cc.leaves):
        flat = p.grad.double().cpu().reshape(-1)
        gsum, gabs, gmax = g.z[f"gsum_{i}"]
        w_max = float(torch.softmax(p.detach(), dim=-1).max())
        probe = torch.from_numpy(g.z[f"gval_{i}"])
        idx = torch.from_numpy(g.z[f"gidx_{i}"])
        tol = grad_tolerance(torch.tensor(float(gmax)), ll_max=ll_max, w_max=w_max)
        err = (flat[idx] - probe).abs().max().item()
        assert err <= tol, f"leaf {i} probe: {err:.3e} > {tol:.3e}"
        # L1 norm of the whole tensor: relative 2e-3 plus the per-element fp32 rounding floor.  The
        # floor is the one grad_tolerance derives (4 ulp(1) = 5e-7 per element): near the root all
        # units of a layer agree to below fp32 resolution (ulp(4357) = 4.9e-4), so a root-weight
        # gradient whose float64 value is 1.7e-7 per element may legitimately come out as exactly 0.
        # In deep circuits with large |ll| (PoonDomingos 3x32x32: 45 layers, |ll| = 18 000) the
        # per-element bound of grad_tolerance (activation rounding, 8*eps32*|ll|*max W) exceeds that
        # floor; the L1 check follows it up to 5e-6 per element.
        got_abs = flat.abs().sum().item()
        floor = min(max(5e-7, tol), 5e-6)
        assert abs(got_abs - gabs) <= 2e-3 * gabs + flat.numel() * floor, f"leaf {i} abs-sum {got_abs} vs {gabs}"
        # the small weight tensors (top levels of the tree, up to 2^17 elements) are compared
        # DENSELY with the reference's float64 gradient, element by element
        full = g.grad_full(i)
        if full is not None:
            dense += 1
            err = (p.grad.double().cpu() - full).abs().max().item()
            assert err <= tol, f"leaf {i} dense: {err:.3e} > {tol:.3e}"
    assert dense >= 1, "seeded fixtures carry the full gradients of their small leaves"


@pytest.mark.parametrize("name", ["qt8_cp_k4", "qg8_cp_k4", "qt8_tucker_k4", "rbt12_gaussian_k5"])
def test_against_cpu_oracle_fp32(name, dev):
    """Same seeded inputs, larger ragged batch than the fixture holds: CUDA vs the oracle."""
    from oracle import OracleCircuit
    from oracle.reference_eval import make_inputs

    g = Golden(name)
    cc = _circuit(g, dev)
    oc = OracleCircuit(g.plan, dtype=torch.float64)
    with torch.no_grad():
        for p, v in zip(oc.leaves, g.leaves(torch.float64)):
            p.copy_(v)
    x = make_inputs(g.plan, 131, seed=7)
    y = cc(x.to(dev))
    yo = oc(x)
    _check_forward(y, yo.detach())
    w = torch.randn(131, 1, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    (y * w.to(dev, torch.float32)).sum().backward()
    (yo * w).sum().backward()
    for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
        gr = torch.zeros_like(q) if q.grad is None else q.grad
        got = torch.zeros_like(p) if p.grad is None else p.grad
        err = (got.double().cpu() - gr).abs().max().item()
        tol = grad_tolerance(gr, gout_l1=float(w.abs().sum()))
        assert err <= tol, f"leaf {i}: {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("batch", [1, 2, 31, 128, 129, 300])
def test_ragged_batches_and_input_dtypes(batch, dev):
    g = Golden("qt8_cp_k4")
    cc = _circuit(g, dev)
    from oracle.reference_eval import make_inputs

    x = make_inputs(g.plan, batch, seed=batch)
    with torch.no_grad():
        y64 = cc(x.to(dev))
        for dt in (torch.uint8, torch.int32, torch.int16, torch.float32, torch.float64):
            y = cc(x.to(dt).to(dev))
            assert torch.equal(y, y64), f"dtype {dt}"
        # host tensor and extra trailing columns are accepted like in the reference
        y = cc(torch.cat([x, torch.zeros(batch, 3, dtype=x.dtype)], dim=1))
        assert torch.equal(y, y64)
        # row b of a batch does not depend on the other rows
        y1 = cc(x[:1].to(dev))
        assert torch.equal(y1[0], y64[0])


@pytest.mark.parametrize("name", ["qt8_cp_k4", "qg8_cp_k4", "qt8_cp_k6_embedding", "pd6_cp_k3_unopt"])
def test_table_fusion_matches_layer_by_layer_plan(name, dev):
    """The fused input-table plan (TABLE_DENSE) against the plain one-kernel-per-layer plan."""
    from cirkit_b200 import B200Circuit
    from oracle.reference_eval import make_inputs

    g = Golden(name)
    x = make_inputs(g.plan, 300, seed=5).to(dev)
    res = []
    for fuse in (True, False):
        cc = B200Circuit(g.plan, fuse_tables=fuse)
        with torch.no_grad():
            for p, v in zip(cc.leaves, g.leaves(torch.float32)):
                p.copy_(v)
        cc = cc.to(dev)
        # circuits whose input layer is read 1:1 by an arity-1 sum layer get the fused step
        assert (cc.runtime.choose_plan(300, False) == "fused") == (fuse and bool(cc.runtime.table_pairs))
        y = cc(x)
        (-y.mean()).backward()
        res.append((y.detach(), [p.grad for p in cc.leaves]))
    (yf, gf), (yp, gp) = res
    assert (yf - yp).abs().max().item() <= 5e-7 * yp.abs().max().item() + 1e-5
    for a, b in zip(gf, gp):
        assert (a - b).abs().max().item() <= max(2e-6, 1e-4 * b.abs().max().item())


@pytest.mark.parametrize("name,units,batch", [
    ("qt8_cp_k4", None, 300), ("qt8_cp_k4", 64, 128), ("qt8_cp_k4", 64, 333), ("qt8_cp_k4", 32, 257),
    ("qt8_cp_k4", 12, 130), ("qt28_cp_k64", None, 256),
    # batches of at least 2 V: the gather serves the samples from table slices staged in shared memory
    ("qt8_cp_k4", 64, 700), ("qt8_cp_k4", 32, 513), ("qt8_cp_k4", 8, 600), ("qt8_cp_k4", 128, 512),
    ("qt28_cp_k64", None, 1024),
])
def test_table_rows_gathered_by_their_consumer(name, units, batch, dev):
    """CKB_STEP_TABLE_INPUT: the CP-T layer over the fused input pair gathers the table rows itself
    (one (F, B, K) block u = x0 + x1 instead of the pair's (2F, B, K) output).  Against the plan
    that materialises the pair's output (same kernels otherwise) and against the float64 oracle;
    ragged and tile-aligned batches, tcgen05 (K = 64), K = 32 and generic shapes."""
    import dataclasses

    from cirkit_b200 import B200Circuit, _lib
    from oracle import OracleCircuit
    from oracle.reference_eval import make_inputs

    g = Golden(name)
    plan = g.plan if units is None else dataclasses.replace(g.plan, meta={"units": 4}).with_units(units)
    x = make_inputs(plan, batch, seed=6)
    res = []
    for fuse in (True, False):
        cc = B200Circuit(plan, seed=17, fuse_table_inputs=fuse).to(dev)
        steps = cc.runtime.exec_plans["fused"]
        assert any(es.flags & _lib.STEP_TABLE_INPUT for es in steps) == fuse
        y = cc(x.to(dev))
        (-y.mean()).backward()
        res.append((cc, y.detach(), [p.grad for p in cc.leaves]))
    (cc, yf, gf), (_, yp, gp) = res
    assert (yf - yp).abs().max().item() <= 5e-7 * yp.abs().max().item() + 1e-5
    for a, b in zip(gf, gp):
        assert (a - b).abs().max().item() <= max(2e-6, 1e-4 * b.abs().max().item())
    oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for p, v in zip(oc.leaves, cc.leaves):
            p.copy_(v.double().cpu())
    yo = oc(x)
    (-yo.mean()).backward()
    _check_forward(yf, yo.detach())
    for i, (a, p) in enumerate(zip(gf, oc.leaves)):
        err = (a.double().cpu() - p.grad).abs().max().item()
        tol = grad_tolerance(p.grad, ll_max=float(yo.detach().abs().max()))
        assert err <= tol, f"leaf {i}: {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("batch", [8, 128, 200, 1024])
def test_tensor_core_path_matches_simt(batch, dev):
    """K=64 layers run on tcgen05 (3xTF32) by default; the FP32 SIMT kernels are the yardstick."""
    from cirkit_b200 import _lib

    g = Golden("qt28_cp_k64")
    cc = _circuit(g, dev)
    x = torch.randint(0, 256, (batch, 784), generator=torch.Generator().manual_seed(batch)).to(dev)
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
    err = (yt.double() - ys.double()).abs().max().item()
    assert err <= 5e-7 * ys.abs().max().item() + 1e-5, f"forward {err:.3e}"
    ll_max = float(ys.abs().max())
    for i, (a, b, p) in enumerate(zip(gt, gs, cc.leaves)):
        e = (a.double() - b.double()).abs().max().item()
        tol = grad_tolerance(b, ll_max=ll_max, w_max=float(torch.softmax(p.detach(), dim=-1).max()))
        assert e <= tol, f"leaf {i}: {e:.3e} > {tol:.3e}"


def test_errors(dev):
    g = Golden("qt8_cp_k4")
    cc = _circuit(g, dev)
    with pytest.raises(ValueError, match="Expected some input"):
        cc()
    with pytest.raises(ValueError, match="shape"):
        cc(torch.zeros(5, dtype=torch.int64, device=dev))
    with pytest.raises(IndexError):
        cc(torch.zeros(5, 3, dtype=torch.int64, device=dev))
    cpu = _circuit(g, torch.device("cpu"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cpu(g.x())


def test_full_size_properties(dev):
    """BASELINE north-star shape (QuadTree 28x28, K=64, B=2048): size-independent properties."""
    from cirkit_b200 import IntegrateQuery

    g = Golden("qt28_cp_k64")
    cc = _circuit(g, dev)
    B = 2048
    x = torch.randint(0, 256, (B, 784), generator=torch.Generator().manual_seed(0))
    x[:8] = g.x()
    xd = x.to(dev)
    y = cc(xd)
    # (1) the fixture rows embedded in a big batch reproduce the reference values
    _check_forward(y[:8], g.y())
    # (2) run-to-run determinism of the forward pass
    assert torch.equal(cc(xd), y)
    # (3) a normalised circuit integrates to one: log Z = 0 with every variable marginalised
    with torch.no_grad():
        z = IntegrateQuery(cc)(xd[:64], integrate_vars=torch.ones(1, 784, dtype=torch.bool))
    assert z.abs().max().item() < 1e-3
    # (4) marginalising nothing is the plain forward pass
    with torch.no_grad():
        y0 = IntegrateQuery(cc)(xd[:64], integrate_vars=torch.zeros(1, 784, dtype=torch.bool))
    assert torch.equal(y0, y[:64])
    # (5) softmax re-parameterisation: every row of d(theta) sums to zero (up to the fp32
    #     rounding of sum_i W_i = 1 times |sum_i W_i dW_i| ~ 1, i.e. ~1e-7 absolute per row)
    (-y.mean()).backward()
    for p in cc.leaves:
        rows = p.grad.double().sum(dim=-1)
        assert rows.abs().max().item() < 2e-6
    # (6) gradient of the mean log-likelihood is linear in the batch: two halves average
    grads = [p.grad.clone() for p in cc.leaves]
    for p in cc.leaves:
        p.grad = None
    (-cc(xd[: B // 2]).mean()).backward()
    g1 = [p.grad.clone() for p in cc.leaves]
    for p in cc.leaves:
        p.grad = None
    (-cc(xd[B // 2 :]).mean()).backward()
    for ga, gb, p in zip(grads, g1, cc.leaves):
        half = 0.5 * (gb.double() + p.grad.double())
        tol = max(2e-6, 1e-4 * ga.abs().max().item())
        assert (half - ga.double()).abs().max().item() <= tol


def test_state_dict_roundtrip(dev):
    """tests/backend/torch/test_serialization.py:17-32: save -> rebuild -> load -> same scores."""
    from cirkit_b200 import B200Circuit

    g = Golden("qg8_cp_k4")
    cc = _circuit(g, dev)
    sd = {k: v.cpu() for k, v in cc.state_dict().items()}
    other = B200Circuit(g.plan).to(dev)
    x = g.x().to(dev)
    with torch.no_grad():
        assert not torch.equal(other(x), cc(x))
        other.load_state_dict(sd)
        assert torch.equal(other(x), cc(x))


@pytest.mark.parametrize("name,batch", [("qt8_cp_k4", 131), ("qg8_cp_k4", 37), ("qt8x4_cpt_k5", 260)])
def test_k32_kernels_vs_oracle(name, batch, dev):
    """The dedicated Ki = Ko = 32 kernels (dense32_kernels.cu, BASELINE.json configs[1]): reference
    structures resized to 32 units, ragged batches, against the float64 oracle."""
    import dataclasses

    from cirkit_b200 import B200Circuit
    from cirkit_b200.plan import seeded_leaves
    from oracle import OracleCircuit
    from oracle.reference_eval import make_inputs

    g = Golden(name)
    k0 = g.plan.steps[0].num_output_units
    plan = dataclasses.replace(g.plan, meta={"units": k0}).with_units(32)
    cc = B200Circuit(plan, seed=5).to(dev)
    oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for q, v in zip(oc.leaves, seeded_leaves(plan, 5)):
            q.copy_(v)
    x = make_inputs(plan, batch, seed=batch)
    y = cc(x.to(dev))
    yo = oc(x)
    _check_forward(y, yo.detach())
    w = torch.randn(batch, 1, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    (y * w.to(dev, torch.float32)).sum().backward()
    (yo * w).sum().backward()
    ll_max = float(yo.detach().abs().max())
    for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
        gr = torch.zeros_like(q) if q.grad is None else q.grad
        got = torch.zeros_like(p) if p.grad is None else p.grad
        err = (got.double().cpu() - gr).abs().max().item()
        tol = grad_tolerance(gr, gout_l1=float(w.abs().sum()), ll_max=ll_max,
                             w_max=float(torch.softmax(q.detach(), dim=-1).max()))
        assert err <= tol, f"leaf {i}: {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("name,batch", [("qg8_cp_k4_densemix", 70), ("pd6_cp_k3_unopt", 45), ("rbt12_gaussian_k5", 129)])
def test_k128_circuits_vs_oracle(name, batch, dev):
    """K = 128, the unit count of BASELINE.json configs[3] (PoonDomingos, 8 GPUs): the reference's
    QuadGraph (dense sum + mixing), PoonDomingos (Hadamard, mixing of arity up to 10, unoptimised)
    and Gaussian-input structures resized to 128 units, against the float64 oracle.  These shapes
    run on the FP32 SIMT kernels."""
    import dataclasses

    from cirkit_b200 import B200Circuit
    from cirkit_b200.plan import seeded_leaves
    from oracle import OracleCircuit
    from oracle.reference_eval import make_inputs

    g = Golden(name)
    k0 = g.plan.steps[0].num_output_units
    plan = dataclasses.replace(g.plan, meta={"units": k0}).with_units(128)
    cc = B200Circuit(plan, seed=5).to(dev)
    oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for q, v in zip(oc.leaves, seeded_leaves(plan, 5)):
            q.copy_(v)
    x = make_inputs(plan, batch, seed=batch)
    y = cc(x.to(dev))
    yo = oc(x)
    _check_forward(y, yo.detach())
    w = torch.randn(batch, 1, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    (y * w.to(dev, torch.float32)).sum().backward()
    (yo * w).sum().backward()
    ll_max = float(yo.detach().abs().max())
    for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
        gr = torch.zeros_like(q) if q.grad is None else q.grad
        got = torch.zeros_like(p) if p.grad is None else p.grad
        err = (got.double().cpu() - gr).abs().max().item()
        tol = grad_tolerance(gr, gout_l1=float(w.abs().sum()), ll_max=ll_max)
        assert err <= tol, f"leaf {i}: {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("H,K,Ko", [(2, 4, 4), (9, 4, 4), (13, 4, 4), (3, 40, 4), (5, 32, 32), (3, 64, 64)])
def test_concatenating_sum_layers(H, K, Ko, dev):
    """TorchSumLayer with arity H > 1 (cirkit/backend/torch/layers/inner.py:266-273): one LSE over
    the concatenation of the H inputs, reduction lengths on both sides of 32 and of 128 (the
    PoonDomingos circuit of BASELINE.json configs[3] has one of length 13*K)."""
    import numpy as np

    from cirkit_b200 import B200Circuit
    from cirkit_b200.plan import CircuitPlan, LeafSpec, ParamSpec, StepSpec
    from oracle import OracleCircuit

    steps = [
        StepSpec("categorical", H, 1, 1, K, params={"probs": ParamSpec(0, [("softmax", {"dim": 1})], (H, K, 16))},
                 scope_idx=np.arange(H, dtype=np.int32), config={"num_categories": 16}),
        StepSpec("sum", 1, H, K, Ko, params={"weight": ParamSpec(1, [("softmax", {"dim": 1})], (1, Ko, H * K))},
                 in_step=np.zeros((1, H), np.int32), in_fold=np.arange(H, dtype=np.int32).reshape(1, H)),
    ]
    plan = CircuitPlan(steps, [LeafSpec((H, K, 16)), LeafSpec((1, Ko, H * K))], np.array([1], np.int32),
                       np.array([0], np.int32), H, tuple(range(H)))
    gen = torch.Generator().manual_seed(H * 100 + K)
    vals = [torch.randn(l.shape, generator=gen) * (3.0 if i == 0 else 1.0) for i, l in enumerate(plan.leaves)]
    cc, oc = B200Circuit(plan), OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for p, q, v in zip(cc.leaves, oc.leaves, vals):
            p.copy_(v)
            q.copy_(v.double())
    cc = cc.to(dev)
    x = torch.randint(0, 16, (37, H), generator=gen)
    y, yo = cc(x.to(dev)), oc(x)
    assert y.shape == yo.shape
    assert (y.detach().double().cpu() - yo.detach()).abs().max().item() <= 5e-7 * yo.abs().max().item() + 1e-5
    (-y.mean()).backward()
    (-yo.mean()).backward()
    for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
        err = (p.grad.double().cpu() - q.grad).abs().max().item()
        assert err <= grad_tolerance(q.grad), f"leaf {i}: {err:.3e}"
