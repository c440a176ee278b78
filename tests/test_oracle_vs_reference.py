"""Pin the oracle (and the adapter's lowering) against the LIVE reference.

Runs only where the reference checkout exists (the build container); on the GPU box the same
guarantee travels as tests/golden/*.npz.
"""
import itertools

import pytest
import torch

from oracle import OracleCircuit

pytestmark = pytest.mark.reference


@pytest.fixture(autouse=True)
def _fp64():
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(42)
    yield
    torch.set_default_dtype(prev)


def _image(reference, shape, rg, spl, K, **kw):
    from cirkit.templates import data_modalities, utils

    return data_modalities.image_data(
        shape, region_graph=rg, input_layer=kw.pop("input_layer", "categorical"),
        num_input_units=K, sum_product_layer=spl, num_sum_units=K,
        sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"), **kw)


def _compare(tc, x, semirings=("lse-sum",)):
    from cirkit_b200.adapter import plan_from_torch

    low = plan_from_torch(tc, semirings=semirings)
    oc = OracleCircuit(low.plan, dtype=torch.float64)
    with torch.no_grad():
        for p, v in zip(oc.leaves, low.leaves):
            p.copy_(v)
    for p in low.leaves:
        p.grad = None
    ext = {k: p() for k, p in low.externals.items()}
    y_ref = tc(x)
    y = oc(x, externals={k: v.detach() for k, v in ext.items()})
    assert y.shape == y_ref.shape
    torch.testing.assert_close(y, y_ref, rtol=1e-10, atol=1e-12)
    if not low.externals:
        (-(y_ref.real if y_ref.is_complex() else y_ref).mean()).backward()
        (-(y.real if y.is_complex() else y).mean()).backward()
        for a, b in zip(oc.leaves, low.leaves):
            ga = torch.zeros_like(a) if a.grad is None else a.grad
            gb = torch.zeros_like(b) if b.grad is None else b.grad
            torch.testing.assert_close(ga, gb, rtol=1e-9, atol=1e-14)
    return low


@pytest.mark.parametrize(
    "rg,spl,fold,optimize",
    [(rg, spl, f, o)
     for rg, spl in [("quad-tree-2", "cp"), ("quad-graph", "cp"), ("quad-tree-2", "tucker"),
                     ("quad-tree-4", "cp-t"), ("poon-domingos", "cp"), ("random-binary-tree", "cp")]
     for f, o in [(True, True), (True, False), (False, False)]],
)
def test_image_circuits(reference, rg, spl, fold, optimize):
    from cirkit.pipeline import PipelineContext

    sc = _image(reference, (1, 4, 4) if not fold else (1, 6, 6), rg, spl, 3)
    ctx = PipelineContext(backend="torch", semiring="lse-sum", fold=fold, optimize=optimize)
    tc = ctx.compile(sc)
    D = 16 if not fold else 36
    with torch.enable_grad():
        _compare(tc, torch.randint(0, 256, (5, D)))


def test_known_answer_circuits(reference):
    """The reference's own ground truths (tests/symbolic/test_utils.py:293-503) through the
    adapter + oracle, for every fold/optimize combination the reference tests
    (tests/backend/torch/test_compile_circuit.py:76-114)."""
    import cirkit.symbolic.functional as SF
    from cirkit.backend.torch.compiler import TorchCompiler
    from cirkit_b200.adapter import plan_from_torch
    from tests.symbolic.test_utils import (
        build_monotonic_bivariate_gaussian_hadamard_dense_pc,
        build_monotonic_structured_categorical_cpt_pc,
    )

    for fold, optimize in itertools.product([False, True], [False, True]):
        compiler = TorchCompiler(fold=fold, optimize=optimize, semiring="lse-sum")
        sc, gt, gt_z = build_monotonic_structured_categorical_cpt_pc(return_ground_truth=True)
        tc = compiler.compile(sc)
        int_tc = compiler.compile(SF.integrate(sc))
        worlds = torch.tensor(list(itertools.product([0, 1], repeat=5)))
        low = plan_from_torch(tc)
        oc = OracleCircuit(low.plan, dtype=torch.float64)
        with torch.no_grad():
            for p, v in zip(oc.leaves, low.leaves):
                p.copy_(v)
            y = oc(worlds).reshape(-1)
            for xs, val in gt["evi"].items():
                idx = int("".join(map(str, xs)), base=2)
                assert abs(y[idx].exp().item() - val) < 1e-9
            assert abs(torch.logsumexp(y, 0).exp().item() - gt_z) < 1e-8
            lowz = plan_from_torch(int_tc)
            oz = OracleCircuit(lowz.plan, dtype=torch.float64)
            for p, v in zip(oz.leaves, lowz.leaves):
                p.copy_(v)
            z = oz(externals={k: p() for k, p in lowz.externals.items()})
            torch.testing.assert_close(z, int_tc(), rtol=1e-10, atol=1e-12)
            assert abs(z.exp().item() - gt_z) < 1e-8

    compiler = TorchCompiler(fold=True, optimize=True, semiring="lse-sum")
    sc, gt, gt_z = build_monotonic_bivariate_gaussian_hadamard_dense_pc(return_ground_truth=True)
    tc = compiler.compile(sc)
    with torch.enable_grad():
        _compare(tc, torch.tensor([[0.3, 1.2], [1.0, -2.0]]))


def test_integrate_query(reference):
    from cirkit.backend.torch.queries import IntegrateQuery
    from cirkit.pipeline import PipelineContext
    from cirkit_b200.adapter import plan_from_torch

    sc = _image(reference, (1, 4, 4), "quad-graph", "cp", 3)
    tc = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True).compile(sc)
    low = plan_from_torch(tc)
    oc = OracleCircuit(low.plan, dtype=torch.float64)
    x = torch.randint(0, 256, (7, 16))
    with torch.no_grad():
        for p, v in zip(oc.leaves, low.leaves):
            p.copy_(v)
        for mask in (torch.rand(7, 16) < 0.5, torch.rand(1, 16) < 0.5,
                     torch.ones(7, 16, dtype=torch.bool), torch.zeros(7, 16, dtype=torch.bool)):
            ref = IntegrateQuery(tc)(x, integrate_vars=mask)
            got = oc(x, integrate_mask=mask)
            torch.testing.assert_close(got, ref, rtol=1e-10, atol=1e-12)
        # all variables integrated out of a normalised circuit: log Z = 0
        got = oc(x, integrate_mask=torch.ones(1, 16, dtype=torch.bool))
    assert abs(got[0].item()) < 1e-9


def test_sum_collapse_matmul_parameter_is_lowered(reference):
    """The SumCollapse rule (cirkit/backend/torch/optimization/layers.py:30-47) turns sum∘sum into
    one sum layer whose weight is `TorchMatMulParameter(W1, mixing(W2))`
    (parameters/nodes.py:786-805).  The adapter lowers that graph into the plan (`matmul` op with
    a second leaf chain) instead of leaving it to the host, so values AND gradients of both leaves
    are the reference's."""
    from cirkit.pipeline import PipelineContext

    found = False
    for shape in [(1, 6, 6), (1, 8, 8), (3, 8, 8)]:
        sc = _image(reference, shape, "poon-domingos", "cp", 3)
        tc = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True).compile(sc)
        D = shape[0] * shape[1] * shape[2]
        with torch.enable_grad():
            low = _compare(tc, torch.randint(0, 256, (4, D)))
        assert not low.externals
        ops = [o for s in low.plan.steps for p in s.params.values() for o, _ in p.ops]
        found |= "matmul" in ops
        # the plan (with the nested right-operand chain) survives serialisation
        again = type(low.plan).load(low.plan.to_bytes())
        assert [[o for o, _ in p.ops] for s in again.steps for p in s.params.values()] == \
               [[o for o, _ in p.ops] for s in low.plan.steps for p in s.params.values()]
    assert found, "none of the PoonDomingos circuits exercised the SumCollapse weight"


@pytest.mark.parametrize("spl,fold,optimize", [(spl, f, o) for spl in ("cp-t", "cp", "tucker")
                                               for f, o in [(True, True), (True, False), (False, False)]])
def test_complex_semiring_circuits(reference, spl, fold, optimize):
    """'complex-lse-sum' (semiring.py:410-476): complex Embedding inputs and sum weights as in
    notebooks/sum-of-squares-circuits.ipynb cell 12, every fold/optimize combination; the circuit
    c, its conjugate (shared leaves through `conj` nodes) and their gradients.  The runtime does not
    execute these plans yet -- the default lowering must keep refusing them."""
    import cirkit.symbolic.functional as SF
    from cirkit.pipeline import PipelineContext
    from cirkit.templates import data_modalities, utils
    from cirkit_b200.adapter import UnsupportedCircuitError, plan_from_torch

    cplx = utils.Parameterization(dtype="complex", initialization="uniform")
    sc = data_modalities.tabular_data(
        "random-binary-tree", num_features=8,
        input_layers={"name": "embedding", "args": {
            "num_states": 7, "weight_factory": utils.parameterization_to_factory(cplx)}},
        num_input_units=3, sum_product_layer=spl, num_sum_units=3, sum_weight_param=cplx)
    ctx = PipelineContext(backend="torch", semiring="complex-lse-sum", fold=fold, optimize=optimize)
    tc = ctx.compile(sc)
    with pytest.raises(UnsupportedCircuitError):
        plan_from_torch(tc)
    both = ("lse-sum", "complex-lse-sum")
    x = torch.randint(0, 7, (6, 8))
    with torch.enable_grad():
        low = _compare(tc, x, both)
        assert low.plan.semiring == "complex-lse-sum" and not low.externals
        lowc = _compare(ctx.compile(SF.conjugate(sc)), x, both)
    assert any(op == "conj" for s in lowc.plan.steps for p in s.params.values() for op, _ in p.ops)
    # the conjugate circuit reads the very same leaf tensors
    assert {id(t) for t in lowc.leaves} == {id(t) for t in low.leaves}


@pytest.mark.parametrize("rg,spl,fold,optimize", [
    ("quad-tree-2", "cp", True, True), ("quad-tree-2", "cp", True, False),
    ("quad-graph", "cp", True, True), ("quad-graph", "cp", False, False),
    ("quad-tree-2", "cp-t", True, True), ("poon-domingos", "cp", True, True),
])
def test_sampling_oracle_draws_the_references_samples(reference, rg, spl, fold, optimize):
    """oracle/sampling.py::reference_sample restates SamplingQuery (queries.py:187-275) call by
    call: under the same torch seed it returns the reference's samples and mixture samples bit for
    bit."""
    from cirkit.backend.torch.queries import SamplingQuery
    from cirkit.pipeline import PipelineContext
    from cirkit_b200.adapter import plan_from_torch
    from oracle.sampling import reference_sample

    sc = _image(reference, (1, 4, 4), rg, spl, 3, num_categories=5) if False else None
    from cirkit.templates import data_modalities, utils

    sc = data_modalities.image_data(
        (1, 4, 4), region_graph=rg, input_layer="categorical", num_input_units=3,
        sum_product_layer=spl, num_sum_units=3, input_params={"probs": utils.Parameterization(
            activation="softmax", initialization="normal")},
        sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"))
    tc = PipelineContext(backend="torch", semiring="lse-sum", fold=fold, optimize=optimize).compile(sc)
    low = plan_from_torch(tc)
    oc = OracleCircuit(low.plan, dtype=torch.float64)
    with torch.no_grad():
        for p, v in zip(oc.leaves, low.leaves):
            p.copy_(v)
    torch.manual_seed(7)
    ref_samples, ref_mix = SamplingQuery(tc)(num_samples=50)
    torch.manual_seed(7)
    samples, mix = reference_sample(oc, 50)
    assert samples.shape == ref_samples.shape == (50, 16)
    assert torch.equal(samples, ref_samples)
    assert len(mix) == len(ref_mix)
    for a, b in zip(mix, ref_mix):
        assert torch.equal(a, b)
