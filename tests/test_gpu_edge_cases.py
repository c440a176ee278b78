"""GPU parity on the edge cases the reference's semiring is built around:
all-(-inf) rows, zero-probability categories, a subnormal weight, marginals of unnormalised
categoricals -- through the FP32 SIMT kernels and through the tcgen05 kernels (K = 64), and the
K = 64 tensor-core kernels on full 128-sample tiles against the float64 oracle.

Tolerances as in tests/test_gpu_parity.py: forward 5e-7*|y| + 1e-5, gradients
helpers.grad_tolerance."""
import dataclasses

import numpy as np
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


def _pair(plan, vals, dev, **kw):
    from cirkit_b200 import B200Circuit
    from oracle import OracleCircuit

    cc, oc = B200Circuit(plan, **kw), OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for p, q, v in zip(cc.leaves, oc.leaves, vals):
            p.copy_(v.float())
            q.copy_(v.float().double())  # the oracle sees the fp32 values the GPU sees
    return cc.to(dev), oc


def _same(y, yo, what=""):
    """Equal where the oracle is infinite (same sign), within tolerance elsewhere."""
    y = y.detach().double().cpu()
    yo = yo.detach()
    assert y.shape == yo.shape
    inf = torch.isinf(yo)
    assert torch.equal(y[inf], yo[inf]), f"{what}: infinities differ"
    assert not torch.isnan(y).any(), f"{what}: NaN"
    err = (y[~inf] - yo[~inf]).abs()
    tol = FWD_RTOL * yo[~inf].abs() + FWD_ATOL
    assert bool((err <= tol).all()), f"{what}: max err {err.max().item():.3e}"


@pytest.mark.parametrize("K", [3, 32, 64, 128])
def test_semiring_subnormal_weight(K, dev):
    """cirkit tests/backend/torch/test_semiring.py:41-61: x = (-200, -200, -5), w = (1, 2, 1e-38):
    the shifted exponentials of the first two inputs underflow, the sum is the subnormal 1e-38 and
    the result must stay finite (log(1e-38) - 5 = -92.50).  K = 3 / 32 / 128 run on the FP32 SIMT
    kernels, K = 64 on tcgen05."""
    from cirkit_b200.plan import CircuitPlan, LeafSpec, ParamSpec, StepSpec

    steps = [
        StepSpec("categorical", 1, 1, 1, K, params={"logits": ParamSpec(0, [], (1, K, 2))},
                 scope_idx=np.zeros(1, np.int32), config={"num_categories": 2}),
        StepSpec("sum", 1, 1, K, K, params={"weight": ParamSpec(1, [], (1, K, K))},
                 in_step=np.zeros((1, 1), np.int32), in_fold=np.zeros((1, 1), np.int32)),
    ]
    plan = CircuitPlan(steps, [LeafSpec((1, K, 2)), LeafSpec((1, K, K))], np.array([1], np.int32),
                       np.array([0], np.int32), 1, (0,))
    logits = torch.full((1, K, 2), -float("inf"))
    logits[0, :3, 0] = torch.tensor([-200.0, -200.0, -5.0])
    logits[0, :, 1] = torch.linspace(-3.0, 0.0, K)
    gen = torch.Generator().manual_seed(K)
    w = torch.rand(1, K, K, generator=gen) + 0.05
    w[0, 0] = 0.0
    w[0, 0, :3] = torch.tensor([1.0, 2.0, 1e-38])
    cc, oc = _pair(plan, [logits, w], dev)
    x = torch.tensor([[0], [1], [0], [1], [0]] * 40)  # 200 samples: full and partial tiles
    with torch.no_grad():
        y = cc(x.to(dev))
        yo = oc(x)
    assert yo.shape == (200, 1, K)
    assert torch.isfinite(yo[0, 0, 0]) and abs(yo[0, 0, 0].item() + 92.5) < 0.1
    assert torch.isfinite(y[0, 0, 0]), "the subnormal sum was flushed to zero"
    # The FP32 SIMT kernels keep the subnormal product exactly.  The tcgen05 kernels (K = 64, 128)
    # multiply in 3xTF32: w = hi + lo with hi = tf32(w); for w = 1e-38 the correction lo = 1.4e-42
    # lies below tf32's smallest subnormal (2^-136) and is dropped by the tensor core, so that ONE
    # entry carries tf32's precision, 2^-11 relative in linear space -> 1.4e-4 in the log (measured,
    # scripts/micro/umma_probe2.cu test 4: 1 x 1e-38 -> 9.99859e-39).  Every other entry must meet
    # the usual tolerance.
    if K in (64, 128):
        assert abs(y[0, 0, 0].item() - yo[0, 0, 0].item()) <= 2.0 ** -11
        y = y.clone()
        sub = (x[:, 0] == 0).to(dev)  # the samples whose unit-0 sum is the subnormal
        y[sub, 0, 0] = yo[0, 0, 0].float().item()
    _same(y, yo, "subnormal weight")


def _resized(name, units, seed=11):
    from cirkit_b200.plan import seeded_leaves

    g = Golden(name)
    k0 = g.plan.steps[0].num_output_units
    plan = dataclasses.replace(g.plan, meta={"units": k0}).with_units(units)
    return plan, seeded_leaves(plan, seed)


@pytest.mark.parametrize("name,units", [("qt8_cp_k4", 64), ("qt8_tucker_k4", 64), ("qt8_cp_k4", 32),
                                        ("qt8_cp_k4", 4), ("qg8_cp_k4", 64)])
@pytest.mark.parametrize("batch", [40, 300])
def test_minus_inf_rows_and_zero_probabilities(name, units, batch, dev):
    """LSESumSemiring.apply_reduce clamps the row max to the finite range so that an all-(-inf)
    row gives exp(-inf) = 0 and log 0 = -inf instead of NaN (semiring.py:392-399).  A category of
    probability zero in EVERY unit of a fold makes such rows for the samples that observe it; a
    category of probability zero in SOME units makes -inf entries inside otherwise finite rows.
    batch 40 runs the layer-by-layer plan, 300 the fused table+dense plan (2B >= 256 states)."""
    plan, vals = _resized(name, units)
    s0 = plan.steps[0]
    assert s0.kind == "categorical"
    leaf = s0.params["probs"].leaf
    f_all, v_all, f_some, v_some = 3, 7, 10, 9
    vals[leaf][f_all, :, v_all] = -float("inf")
    vals[leaf][f_some, : max(1, units // 3), v_some] = -float("inf")
    cc, oc = _pair(plan, vals, dev)
    gen = torch.Generator().manual_seed(batch)
    x = torch.randint(0, 256, (batch, plan.num_variables), generator=gen)
    var_all, var_some = int(s0.scope_idx[f_all]), int(s0.scope_idx[f_some])
    x[:, var_all] = torch.where(x[:, var_all] == v_all, v_all + 1, x[:, var_all])
    x[::3, var_some] = v_some
    hit = torch.arange(batch) % 5 == 1
    x[hit, var_all] = v_all  # these samples have probability zero
    with torch.no_grad():
        y = cc(x.to(dev))
        yo = oc(x)
    assert torch.isinf(yo[hit]).all() and torch.isfinite(yo[~hit]).all()
    _same(y, yo, f"{name} K={units}")
    # gradients on the samples of non-zero probability (the reference's autograd turns a -inf
    # log-likelihood into NaN gradients everywhere, and a zero probability that is observed into
    # NaN rows of that categorical's gradient: 0 * inf in the backward of log; those entries are
    # excluded, everything else must agree)
    xs = x[~hit]
    ys = cc(xs.to(dev))
    (-ys.mean()).backward()
    yos = oc(xs)
    (-yos.mean()).backward()
    _same(ys, yos, f"{name} K={units} (finite samples)")
    ll_max = float(yos.detach().abs().max())
    for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
        got, ref = p.grad.double().cpu(), q.grad
        ok = torch.isfinite(ref)
        assert ok.float().mean() > 0.9 and torch.isfinite(got).all(), f"leaf {i}"
        w_max = float(torch.softmax(q.detach(), dim=-1).max())
        tol = grad_tolerance(ref[ok], ll_max=ll_max, w_max=w_max)
        err = (got[ok] - ref[ok]).abs().max().item()
        assert err <= tol, f"leaf {i}: {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("name,batch", [("qt8_cp_k4", 300), ("qt8_cp_k4", 1024), ("qg8_cp_k4", 257),
                                        ("qt8x4_cpt_k5", 384)])
def test_k64_tensor_core_kernels_vs_oracle(name, batch, dev):
    """The tcgen05 sum-product kernels (dense_tc.cu) on full 128-sample tiles plus a ragged tail,
    forward and ALL gradients against the float64 oracle (not against the repo's own SIMT
    kernels): reference structures resized to 64 units."""
    from oracle.reference_eval import make_inputs

    plan, vals = _resized(name, 64, seed=5)
    cc, oc = _pair(plan, vals, dev)
    x = make_inputs(plan, batch, seed=batch)
    y = cc(x.to(dev))
    yo = oc(x)
    _same(y, yo, name)
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


@pytest.mark.parametrize("name", ["cat_logits_k3_unfolded", "qt8_cp_k4", "rbt12_gaussian_k5"])
def test_integrate_query_gradients_vs_reference(name, dev):
    """Marginal log-likelihoods AND their gradients against the reference's autograd through
    `torch.where(mask, layer.integrate(), output)` (queries.py:132-143).  `cat_logits_k3_unfolded`
    holds unnormalised categoricals: an integrated variable contributes logsumexp(logits)
    (layers/input.py:414-421, here CKB_POP_LSE_ROWS) and the logits receive softmax * g."""
    from cirkit_b200 import B200Circuit, IntegrateQuery

    g = Golden(name)
    cc = B200Circuit(g.plan)
    with torch.no_grad():
        for p, v in zip(cc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    cc = cc.to(dev)
    mask, y_mask = g.mask()
    y = IntegrateQuery(cc)(g.x().to(dev), integrate_vars=mask.to(dev))
    _same(y, y_mask, name)
    (-y.mean()).backward()
    for i, (p, gr) in enumerate(zip(cc.leaves, g.grads_mask())):
        got = torch.zeros_like(p) if p.grad is None else p.grad
        err = (got.double().cpu() - gr).abs().max().item()
        assert err <= grad_tolerance(gr), f"leaf {i}: grad err {err:.3e} > {grad_tolerance(gr):.3e}"


def test_integrate_unnormalised_logits_folded_vs_oracle(dev):
    """Same on the FOLDED circuit (4 variables in one Categorical layer), where the reference
    itself raises (its logsumexp is (F, K), not (F, 1, K)); the oracle states the intended
    broadcast.  Full, partial and per-sample masks."""
    from cirkit_b200 import B200Circuit, IntegrateQuery
    from oracle import OracleCircuit

    g = Golden("cat_logits_k3")
    cc = B200Circuit(g.plan)
    oc = OracleCircuit(g.plan, dtype=torch.float64)
    with torch.no_grad():
        for p, q, v in zip(cc.leaves, oc.leaves, g.leaves(torch.float64)):
            p.copy_(v.float())
            q.copy_(v.float().double())
    cc = cc.to(dev)
    B = 150
    gen = torch.Generator().manual_seed(2)
    x = torch.randint(0, 5, (B, 4), generator=gen)
    for mask in (torch.rand(B, 4, generator=gen) < 0.5, torch.ones(1, 4, dtype=torch.bool),
                 torch.tensor([[True, False, False, True]])):
        for p in list(cc.leaves) + list(oc.leaves):
            p.grad = None
        y = IntegrateQuery(cc)(x.to(dev), integrate_vars=mask.to(dev))
        yo = oc(x, integrate_mask=mask)
        _same(y, yo, "logits marginal")
        (-y.mean()).backward()
        (-yo.mean()).backward()
        for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
            err = (p.grad.double().cpu() - q.grad).abs().max().item()
            assert err <= grad_tolerance(q.grad), f"leaf {i}: {err:.3e}"


def test_integrating_an_embedding_raises(dev):
    """TorchInputLayer.integrate raises TypeError for Embedding layers (layers/input.py:82-92);
    a mask that touches none of their variables is fine (queries.py:136-139)."""
    from cirkit_b200 import B200Circuit, IntegrateQuery

    g = Golden("qt8_cp_k6_embedding")
    cc = B200Circuit(g.plan)
    with torch.no_grad():
        for p, v in zip(cc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    cc = cc.to(dev)
    x = g.x().to(dev)
    none = torch.zeros(1, 64, dtype=torch.bool)
    with torch.no_grad():
        assert torch.equal(IntegrateQuery(cc)(x, integrate_vars=none), cc(x))
        some = none.clone()
        some[0, 5] = True
        with pytest.raises(TypeError, match="Integration is not supported"):
            IntegrateQuery(cc)(x, integrate_vars=some)


def test_effective_parameters_are_cached_between_no_grad_calls(dev):
    """the reference re-runs every parameter node on every call
    (parameters/parameter.py:180-188); here the second no_grad call on unchanged parameters
    launches no parameter kernels and returns bit-identical values, an in-place update
    invalidates the cache, and calls that record a graph always re-run the ops."""
    from cirkit_b200 import B200Circuit, IntegrateQuery

    g = Golden("qg8_cp_k4")
    cc = B200Circuit(g.plan)
    with torch.no_grad():
        for p, v in zip(cc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    cc = cc.to(dev)
    x = g.x().to(dev)
    rt = cc.runtime
    with torch.no_grad():
        y0 = cc(x)
        n0 = rt.last_launches
        y1 = cc(x)
        n1 = rt.last_launches
        assert torch.equal(y0, y1) and n1 < n0, (n0, n1)
        mask = torch.zeros(1, 64, dtype=torch.bool)
        mask[0, :7] = True
        m0 = IntegrateQuery(cc)(x, integrate_vars=mask)
        assert torch.equal(IntegrateQuery(cc)(x, integrate_vars=mask), m0)
        cc.leaves[1].add_(0.5)  # version bump -> ops re-run
        y2 = cc(x)
        assert rt.last_launches == n0 and not torch.equal(y2, y0)
        rt.cache_parameters = False
        cc(x)
        assert rt.last_launches == n0
        rt.cache_parameters = True
    y3 = cc(x)  # grad mode: never cached
    assert rt.last_launches == n0 and torch.equal(y3.detach(), y2)
    (-y3.mean()).backward()
    assert all(p.grad is not None for p in cc.leaves)


def test_out_of_range_evidence_is_reported_on_request(dev):
    """The kernels clamp a state to [0, V); with `check_evidence` the runtime raises like the
    reference's indexing does (layers/input.py:399-412)."""
    from cirkit_b200 import B200Circuit

    g = Golden("qt8_cp_k4")
    cc = B200Circuit(g.plan, seed=2).to(dev)
    x = torch.randint(0, 256, (8, g.plan.num_variables))
    x[3, 5] = 256
    y = cc(x.to(dev))  # default: clamped, finite
    assert torch.isfinite(y).all()
    cc.runtime.check_evidence = True
    with pytest.raises(IndexError, match="out of range"):
        cc(x.to(dev))
    x[3, 5] = -1
    with pytest.raises(IndexError, match="out of range"):
        cc(x.to(dev))
    x[3, 5] = 255
    assert torch.isfinite(cc(x.to(dev))).all()
