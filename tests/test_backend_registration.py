"""The drop-in boundary on the reference's side: `PipelineContext(backend="b200")`
and `accelerate(tc)`, checked against the live reference up to the point where a CUDA device is
needed (this container has the reference but no GPU; the GPU box has no reference)."""
import pytest
import torch

pytestmark = pytest.mark.reference


def _symbolic(reference, region_graph="quad-graph", **kw):
    from cirkit.templates import data_modalities, utils

    return data_modalities.image_data(
        (1, 4, 4), region_graph=region_graph, input_layer="categorical", num_input_units=3,
        sum_product_layer="cp", num_sum_units=3,
        sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"), **kw)


def test_pipeline_context_backend_b200(reference):
    """cirkit/pipeline.py:30-51, :348-356 and backend/compiler.py:11: the backend switch."""
    import cirkit.backend.compiler as BC
    import cirkit.symbolic.functional as SF
    from cirkit.backend.torch.circuits import TorchCircuit
    from cirkit.pipeline import PipelineContext

    import cirkit_b200

    cirkit_b200.register_backend()
    cirkit_b200.register_backend()  # idempotent
    assert BC.SUPPORTED_BACKENDS.count("b200") == 1
    with pytest.raises(NotImplementedError):  # pipeline.py:41-42 still guards unknown names
        PipelineContext(backend="nope")

    sc = _symbolic(reference)
    torch.manual_seed(7)
    ctx = PipelineContext(backend="b200", semiring="lse-sum", fold=True, optimize=True)
    cc = ctx.compile(sc)
    torch.manual_seed(7)
    ref = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True).compile(sc)

    # same object model: a TorchCircuit subclass with the reference's layers, keys and values
    assert isinstance(cc, TorchCircuit) and type(cc).__name__ == "B200TorchCircuit"
    assert list(cc.state_dict().keys()) == list(ref.state_dict().keys())
    for (k, a), b in zip(cc.state_dict().items(), ref.state_dict().values()):
        assert torch.equal(a, b), k
    assert cc.scope == ref.scope and cc.num_variables == ref.num_variables
    assert [type(l).__name__ for l in cc.layers] == [type(l).__name__ for l in ref.layers]
    # the runtime binds the circuit's OWN nn.Parameter leaves (optimisers, checkpoints and
    # pointer-shared circuits see one storage, parameters/nodes.py:193-201)
    own = {id(p) for p in cc.parameters()}
    assert {id(p) for p in cc._b200_lowered.leaves} <= own
    cc.load_state_dict(ref.state_dict())
    # compiler bookkeeping keeps working (backend/compiler.py:20-37, pipeline.py:168-229)
    assert ctx.get_symbolic_circuit(cc) is sc
    zc = ctx.compile(SF.integrate(sc))
    assert type(zc).__name__ == "B200TorchCircuit" and not zc.scope
    # learnable leaves are shared with c through pointer nodes (rules/parameters.py:111-117);
    # what integration adds (the constant log-partition values) is not learnable
    assert {id(p) for p in zc._b200_lowered.leaves if p.requires_grad} <= own
    assert any(id(p) in own for p in zc._b200_lowered.leaves)

    # error behaviour of TorchCircuit.forward (circuits.py:60-65, :259-260) ...
    with pytest.raises(ValueError, match="Expected some input"):
        cc()
    # ... and no CPU path behind the accelerated circuit
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cc(torch.randint(0, 256, (2, 16)))


def test_accelerate_leaves_unsupported_circuits_on_the_reference_path(reference):
    from cirkit.pipeline import PipelineContext
    from cirkit.templates import data_modalities, utils

    from cirkit_b200 import UnsupportedCircuitError, accelerate

    sc = _symbolic(reference)
    tc = PipelineContext(backend="torch", semiring="sum-product", fold=True, optimize=True).compile(sc)
    cls = type(tc)
    x = torch.randint(0, 256, (3, 16))
    y = tc(x)
    with pytest.raises(UnsupportedCircuitError):
        accelerate(tc, strict=True)
    assert accelerate(tc) is tc and type(tc) is cls  # untouched: the reference evaluates it
    assert "SumProductSemiring" in tc._b200_reason
    assert torch.equal(tc(x), y)


def test_input_layers_without_a_kernel_become_external_steps(reference):
    """layer kinds without a kernel (here Binomial, layers/input.py:437) keep the circuit
    on the CUDA executor: the reference's own module evaluates that ONE layer with PyTorch per call
    and its output enters the plan as a differentiable external tensor."""
    from cirkit.pipeline import PipelineContext
    from cirkit.templates import data_modalities, utils

    from cirkit_b200 import accelerate

    sc = data_modalities.image_data(
        (1, 4, 4), region_graph="quad-tree-2", input_layer="binomial", num_input_units=3,
        sum_product_layer="cp", num_sum_units=3,
        sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"))
    tc = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True).compile(sc)
    cc = accelerate(tc, strict=True)
    plan = cc._b200_lowered.plan
    assert plan.steps[0].kind == "external" and (0, "output") in cc._b200_lowered.externals
    assert type(cc._b200_lowered.externals[(0, "output")]).__name__ == "TorchBinomialLayer"
    assert not cc._b200_runtime.reads_evidence and cc._b200_runtime.needs_batch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cc(torch.randint(0, 256, (3, 16)))


def test_complex_semiring_circuits_are_accelerated(reference):
    """'complex-lse-sum' circuits (squared circuits, semiring.py:410-476) lower to a complex plan;
    layer kinds without a complex kernel (mixing, concatenating sums) stay on the reference path."""
    from cirkit.pipeline import PipelineContext
    from cirkit.templates import data_modalities, utils

    from cirkit_b200 import accelerate

    cplx = utils.Parameterization(dtype="complex", initialization="uniform")
    sc = data_modalities.tabular_data(
        "random-binary-tree", num_features=4,
        input_layers={"name": "embedding", "args": {
            "num_states": 5, "weight_factory": utils.parameterization_to_factory(cplx)}},
        num_input_units=2, sum_product_layer="cp-t", num_sum_units=2, sum_weight_param=cplx)
    tc = PipelineContext(backend="torch", semiring="complex-lse-sum", fold=True, optimize=True).compile(sc)
    cc = accelerate(tc, strict=True)
    assert type(cc).__name__ == "B200TorchCircuit" and cc._b200_runtime.is_complex
    assert all(p.is_complex() for p in cc._b200_lowered.leaves)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cc(torch.randint(0, 5, (3, 4)))


def test_reference_integrate_query_is_routed_to_the_masked_runtime(reference, monkeypatch):
    """`cirkit.backend.torch.queries.IntegrateQuery(circuit)(x, integrate_vars=...)` calls
    `circuit.evaluate(x, module_fn=partial(_layer_fn, integrate_vars_mask=m))` (queries.py:101-107).
    The accelerated circuit recognises that callback and runs its masked kernels instead of the
    reference's layer loop; any other callback stays on the reference's executor.  No GPU here:
    the runtime is replaced by a recorder."""
    from cirkit.backend.torch.queries import IntegrateQuery
    from cirkit.pipeline import PipelineContext
    from cirkit.utils.scope import Scope

    import cirkit_b200

    sc = _symbolic(reference)
    tc = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True).compile(sc)
    x = torch.randint(0, 256, (5, 16))
    mask = torch.rand(5, 16) < 0.5
    want = IntegrateQuery(tc)(x, integrate_vars=mask)  # the reference's answer, (B, O, K)
    plain = tc(x)

    cc = cirkit_b200.accelerate(tc, strict=True)
    calls = []

    def recorder(x_, leaves, ext, integrate_mask=None):
        calls.append(integrate_mask)
        return want  # (B, O, K), what PlanRuntime.evaluate returns

    monkeypatch.setattr(cc._b200_runtime, "evaluate", recorder)
    got = IntegrateQuery(cc)(x, integrate_vars=mask)
    assert len(calls) == 1 and torch.equal(calls[0], mask)
    assert got.shape == want.shape and torch.equal(got, want)
    # scope form: one mask row, broadcast over the batch (queries.py:88-92)
    IntegrateQuery(cc)(x, integrate_vars=Scope([0, 3]))
    assert calls[1].shape == (1, 16) and calls[1].sum() == 2 and calls[1][0, 3]
    # plain evaluate(): no mask
    cc.evaluate(x)
    assert calls[2] is None
    # a foreign per-layer callback is not ours to interpret: reference executor, on CPU tensors here
    seen = []

    def spy(layer, *inputs):
        seen.append(type(layer).__name__)
        return layer(*inputs)

    y = cc.evaluate(x, module_fn=spy)
    assert len(calls) == 3 and len(seen) == len(list(cc.layers))
    assert torch.equal(y.transpose(0, 1), plain)


def test_b200circuit_from_torch(reference, tmp_path):
    """Route B of SURVEY §8(b): post-compile conversion into a stand-alone module."""
    import cirkit.symbolic.functional as SF
    from cirkit.pipeline import PipelineContext

    from cirkit_b200 import B200Circuit, CircuitPlan, UnsupportedCircuitError

    sc = _symbolic(reference)
    ctx = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True)
    tc = ctx.compile(sc)
    shared = B200Circuit.from_torch(tc)
    assert [id(p) for p in shared.parameters()] == [id(p) for p in shared.leaves]
    assert {id(p) for p in shared.leaves} <= {id(p) for p in tc.parameters()}
    assert shared.scope == tuple(sorted(tc.scope)) and len(shared.layers) == len(list(tc.layers))
    copied = B200Circuit.from_torch(tc, share_parameters=False)
    for a, b in zip(copied.leaves, shared.leaves):
        assert a is not b and torch.equal(a, b.to(a.dtype))
    # the plan travels without the reference
    path = tmp_path / "plan.npz"
    shared.plan.save(str(path))
    again = B200Circuit(CircuitPlan.load(str(path)))
    again.load_state_dict(copied.state_dict())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        shared(torch.randint(0, 256, (2, 16)))
    # product circuits carry kron parameter graphs: not expressible as a stand-alone plan
    qt = _symbolic(reference, "quad-tree-2")
    plain = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=False)
    with pytest.raises(UnsupportedCircuitError, match="parameter graph"):
        B200Circuit.from_torch(plain.compile(SF.multiply(qt, qt)))
