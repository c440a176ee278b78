"""Plan / layout host logic (no GPU)."""
import numpy as np
import pytest
import torch

from cirkit_b200.plan import CircuitPlan, build_layout
from helpers import Golden, golden_names


@pytest.mark.parametrize("name", golden_names())
def test_roundtrip_and_layout(name):
    g = Golden(name)
    plan = g.plan
    again = CircuitPlan.load(plan.to_bytes())
    assert [s.kind for s in again.steps] == [s.kind for s in plan.steps]
    lay = build_layout(plan)
    # activation blocks are disjoint, 4-float aligned and cover every step
    ends = 0
    for sid, s in enumerate(plan.steps):
        assert lay.out_off[sid] % 4 == 0 and lay.out_off[sid] >= ends
        ends = lay.out_off[sid] + s.num_folds * s.num_output_units
    assert lay.arena_units >= ends
    # every gathered row points at the start of a producer row
    for sid, s in enumerate(plan.steps):
        if s.is_input:
            continue
        rows = lay.in_rows[sid].reshape(s.num_folds, s.arity)
        for f in range(s.num_folds):
            for h in range(s.arity):
                p, pf = int(s.in_step[f, h]), int(s.in_fold[f, h])
                assert rows[f, h] == lay.out_off[p] + pf * plan.steps[p].num_output_units
    # consumer lists are the transpose of the gathers (+ the circuit outputs)
    total = sum(len(r) for r in lay.cons_rows)
    edges = sum(s.num_folds * s.arity for s in plan.steps if not s.is_input)
    assert total == edges + plan.num_outputs


def test_survey_sizes():
    """A and P of SURVEY §8 for the benchmark shapes."""
    g = Golden("qt28_cp_k64").plan
    assert (g.activation_units(), g.parameter_elements()) == (150401, 19259456)
    assert g.algorithmic_bytes(2048) == 8 * 2048 * 784 + 20 * 150401 * 2048 + 28 * 19259456
    g = Golden("qt28_cp_k32").plan
    assert (g.activation_units(), g.parameter_elements()) == (75201, 8026144)
    g = Golden("qt28_tucker_k64").plan
    assert (g.activation_units(), g.parameter_elements()) == (100225, 217845760)


def test_validate_rejects_bad_gathers():
    g = Golden("qt8_cp_k4").plan
    bad = CircuitPlan.load(g.to_bytes())
    bad.steps[2].in_fold[0, 0] = 10_000
    with pytest.raises(ValueError):
        bad.validate()


def test_complex_plans():
    """The plan format carries 'complex-lse-sum' circuits (complex leaves, `conj` parameter nodes);
    the runtime binds them with complex64 leaves and the fused conjugate op, and refuses the layer
    kinds it has no complex kernel for."""
    from cirkit_b200 import B200Circuit

    plan = Golden("rbt16_cpt_k4_complex_conj").plan
    assert plan.semiring == "complex-lse-sum"
    assert {l.dtype for l in plan.leaves} == {"complex"}
    assert any(op == "conj" for s in plan.steps for p in s.params.values() for op, _ in p.ops)
    again = CircuitPlan.load(plan.to_bytes())
    assert [l.dtype for l in again.leaves] == [l.dtype for l in plan.leaves]
    cc = B200Circuit(plan)
    assert cc.runtime.is_complex and all(p.dtype == torch.complex64 for p in cc.leaves)
    assert all(b.native is not None and b.is_complex for b in cc.runtime.bindings)  # conj is fused
    mixing = CircuitPlan.load(Golden("qg8_cp_k4").plan.to_bytes())
    mixing.semiring = "complex-lse-sum"
    with pytest.raises(NotImplementedError, match="mixing"):
        B200Circuit(mixing)
    # a complex leaf or a conjugation inside a real-valued circuit is a malformed plan
    bad = CircuitPlan.load(plan.to_bytes())
    bad.semiring = "lse-sum"
    with pytest.raises(ValueError):
        bad.validate()
    real = CircuitPlan.load(Golden("qt8_cp_k4").plan.to_bytes())
    real.steps[1].params["weight"].ops.append(("conj", {}))
    with pytest.raises(ValueError, match="conj"):
        real.validate()
    real.semiring = "tropical"
    with pytest.raises(ValueError, match="semiring"):
        real.validate()


def test_gather_fusion_marks_only_the_fused_plan():
    """`PlanRuntime._fuse_table_inputs` (host logic, no device): the first CP-T layer of a QuadTree
    circuit gathers its inputs from the fused table pair (CKB_STEP_TABLE_INPUT) and the pair skips
    its own gather (CKB_STEP_NO_GATHER) -- in the FUSED execution plan only.  The plain plan
    shares step objects with it and must stay untouched (it serves small batches and masks)."""
    from cirkit_b200 import _lib
    from cirkit_b200.runtime import STEP_TABLE_DENSE, PlanRuntime
    from helpers import Golden

    for name in ("qt28_cp_k64", "qt8_cp_k4", "qt28_cp_k32"):
        rt = PlanRuntime(Golden(name).plan)
        fused, plain = rt.exec_plans["fused"], rt.exec_plans["plain"]
        assert fused[0].kind == STEP_TABLE_DENSE and fused[0].flags & _lib.STEP_NO_GATHER
        assert fused[1].flags & _lib.STEP_TABLE_INPUT and fused[1].table_input == (0, 1, fused[0].scratch[0])
        assert all(es.flags == 0 and es.table_input is None for es in plain)
        assert all(es.flags == 0 for es in fused[2:])
        assert rt.exec_plans["masked"] is plain
    # no fusion: the pair's output has other readers / the reader is not a CP-T layer over it
    for name in ("qg8_cp_k4", "pd32_cp_k4", "qt8_tucker_k4", "qt8_cp_k6_embedding"):
        rt = PlanRuntime(Golden(name).plan)
        assert all(es.table_input is None for es in rt.exec_plans["fused"])
    off = PlanRuntime(Golden("qt8_cp_k4").plan, fuse_table_inputs=False)
    assert all(es.flags == 0 for es in off.exec_plans["fused"])
