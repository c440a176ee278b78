"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [--only NAME ...]

For every case the reference's own torch backend (CPU, float64, `PipelineContext(backend="torch",
semiring="lse-sum", fold=..., optimize=...)`) is compiled and evaluated; the compiled circuit is
lowered to a `CircuitPlan` (cirkit_b200.adapter) and stored with the reference's outputs and
autograd gradients.  Two kinds of files are written:

* ``<name>.npz`` "full" fixtures (small circuits): plan, leaf values, inputs, reference outputs
  ``y`` and the gradient of ``-mean(y)`` w.r.t. every leaf, plus optional integrate-query masks
  and their outputs.  The two known-answer circuits of the reference's test-suite
  (`tests/symbolic/test_utils.py:293-503`) also carry the hand-computed numbers.
* ``<name>.npz`` "seeded" fixtures (benchmark-size circuits): plan + the seed the leaves are drawn
  from (`cirkit_b200.plan.seeded_leaves`), inputs, reference outputs, and per-leaf gradient
  summaries (sum, abs-sum and a strided probe), so the 10^7..10^8 parameters never hit the repo.
"""

from __future__ import annotations

import argparse
import itertools
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")  # `cirkit` and the reference's `tests` package

import numpy as np  # noqa: E402
import torch  # noqa: E402

from cirkit.pipeline import PipelineContext  # noqa: E402
from cirkit.backend.torch.queries import IntegrateQuery  # noqa: E402
from cirkit.templates import data_modalities, utils  # noqa: E402
import cirkit.symbolic.functional as SF  # noqa: E402
from cirkit.symbolic.circuit import Circuit  # noqa: E402
from cirkit.symbolic.layers import CategoricalLayer, EmbeddingLayer, GaussianLayer, SumLayer, HadamardLayer  # noqa: E402
from cirkit.symbolic.parameters import Parameter, TensorParameter, SoftmaxParameter  # noqa: E402
from cirkit.symbolic.initializers import NormalInitializer, UniformInitializer  # noqa: E402
from cirkit.utils.scope import Scope  # noqa: E402

from cirkit_b200.adapter import plan_from_torch  # noqa: E402
from cirkit_b200.plan import seeded_leaves  # noqa: E402

PROBE = 64  # gradient entries kept per leaf in seeded fixtures
FULL_GRAD_MAX = 1 << 17  # leaves up to this many elements keep their whole float64 gradient


def compile_ref(sc, fold=True, optimize=True, semiring="lse-sum"):
    ctx = PipelineContext(backend="torch", semiring=semiring, fold=fold, optimize=optimize)
    return ctx, ctx.compile(sc)


def plan_array(plan) -> np.ndarray:
    return np.frombuffer(plan.to_bytes(), dtype=np.uint8)


def ref_forward_backward(tc, leaves, x):
    for p in leaves:
        p.grad = None
    with torch.enable_grad():
        y = tc(x)
        # complex circuits: the gradient of the real part (the SoS loss is 2 Re c(x) - Re Z,
        # notebooks/sum-of-squares-circuits.ipynb cell 32)
        (-(y.real if y.is_complex() else y).mean()).backward()
    return y.detach(), [
        (torch.zeros_like(p) if p.grad is None else p.grad.detach().clone()) for p in leaves
    ]


NAMES: set | None = None  # --names filter: fixtures to (re)write


def write_full(name, tc, x, *, masks=None, extra=None, kind="full"):
    if NAMES is not None and name not in NAMES:
        return
    low = plan_from_torch(tc, allow_external_params=False, semirings=("lse-sum", "complex-lse-sum"))
    y, grads = ref_forward_backward(tc, low.leaves, x)
    out = {"plan": plan_array(low.plan), "x": x.numpy(), "y": y.numpy()}
    for i, (p, g) in enumerate(zip(low.leaves, grads)):
        out[f"leaf_{i}"] = p.detach().numpy()
        out[f"grad_{i}"] = g.numpy()
    if masks is not None:
        q = IntegrateQuery(tc)
        out["mask"] = masks.numpy()
        for p in low.leaves:
            p.grad = None
        with torch.enable_grad():
            ym = q(x, integrate_vars=masks)
            (-ym.mean()).backward()
        out["y_mask"] = ym.detach().numpy()
        # gradient of the marginal log-likelihood (reference: autograd through
        # torch.where(mask, layer.integrate(), output), queries.py:132-143)
        for i, p in enumerate(low.leaves):
            out[f"grad_mask_{i}"] = (torch.zeros_like(p) if p.grad is None else p.grad.detach().clone()).numpy()
    meta = {"kind": kind, "steps": [s.kind for s in low.plan.steps]}
    meta.update(extra or {})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: steps={meta['steps']} B={x.shape[0]} y[0]={y.reshape(-1)[0].item():.6f}")


def write_seeded(name, tc, x, seed, extra=None):
    if NAMES is not None and name not in NAMES:
        return
    low = plan_from_torch(tc, allow_external_params=False)
    vals = seeded_leaves(low.plan, seed)
    with torch.no_grad():
        for p, v in zip(low.leaves, vals):
            p.copy_(v.to(p.dtype))
    y, grads = ref_forward_backward(tc, low.leaves, x)
    out = {"plan": plan_array(low.plan), "x": x.numpy(), "y": y.numpy()}
    for i, g in enumerate(grads):
        flat = g.reshape(-1)
        stride = max(1, flat.numel() // PROBE)
        idx = torch.arange(0, flat.numel(), stride)[:PROBE]
        out[f"gsum_{i}"] = np.array([flat.sum().item(), flat.abs().sum().item(), flat.abs().max().item()])
        out[f"gidx_{i}"] = idx.numpy()
        out[f"gval_{i}"] = flat[idx].numpy()
        if flat.numel() <= FULL_GRAD_MAX:
            # the small (top-of-the-tree) weight tensors are compared densely on the GPU
            out[f"gfull_{i}"] = g.numpy()
    meta = {"kind": "seeded", "seed": seed, "steps": [s.kind for s in low.plan.steps]}
    meta.update(extra or {})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: steps={len(meta['steps'])} B={x.shape[0]} y[0]={y.reshape(-1)[0].item():.6f} "
          f"A={low.plan.activation_units()} P={low.plan.parameter_elements()}")


# --------------------------------------------------------------------------- cases
def image_circuit(shape, rg, spl, K, input_layer="categorical", **kw):
    return data_modalities.image_data(
        shape, region_graph=rg, input_layer=input_layer, num_input_units=K,
        sum_product_layer=spl, num_sum_units=K,
        sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"),
        **kw,
    )


def case_ka_categorical():
    """Hand-computed 5-variable circuit, reference tests/symbolic/test_utils.py:293-417."""
    from tests.symbolic.test_utils import build_monotonic_structured_categorical_cpt_pc

    sc, gt, gt_z = build_monotonic_structured_categorical_cpt_pc(return_ground_truth=True)
    ctx = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True)
    tc = ctx.compile(sc)
    int_tc = ctx.compile(SF.integrate(sc))
    worlds = torch.tensor(list(itertools.product([0, 1], repeat=5)))
    mask = torch.zeros(32, 5, dtype=torch.bool)
    mask[:, 4] = True
    extra = {
        "evi": {"".join(map(str, k)): v for k, v in gt["evi"].items()},
        "mar": {"10110": 16.845},
        "Z": gt_z,
        "source": "tests/symbolic/test_utils.py:411-417",
    }
    write_full("ka_categorical_cpt", tc, worlds, masks=mask, extra=extra)
    with torch.no_grad():
        z = int_tc()
    low = plan_from_torch(int_tc, allow_external_params=False)
    out = {"plan": plan_array(low.plan), "y": z.numpy()}
    for i, p in enumerate(low.leaves):
        out[f"leaf_{i}"] = p.detach().numpy()
    out["meta"] = np.frombuffer(json.dumps({"kind": "constant", "Z": gt_z}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "ka_categorical_cpt_Z.npz"), **out)
    print("ka_categorical_cpt_Z: logZ", z.item(), "exp", z.exp().item())


def case_ka_gaussian():
    """Hand-computed bivariate Gaussian circuit, tests/symbolic/test_utils.py:420-503."""
    from tests.symbolic.test_utils import build_monotonic_bivariate_gaussian_hadamard_dense_pc

    sc, gt, gt_z = build_monotonic_bivariate_gaussian_hadamard_dense_pc(return_ground_truth=True)
    ctx = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True)
    tc = ctx.compile(sc)
    x = torch.tensor([[0.3, 1.2], [0.3, 1.2], [-1.0, 0.25], [2.0, -3.0]])
    mask = torch.tensor([[False, False], [False, True], [True, False], [True, True]])
    extra = {
        "evi": {"0": 3.744904862456293},
        "mar": {"1": 23.528960785605985},
        "Z": gt_z,
        "source": "tests/symbolic/test_utils.py:497-503",
    }
    write_full("ka_gaussian", tc, x, masks=mask, extra=extra)


def case_random_small():
    B = 24
    torch.manual_seed(42)
    x8 = torch.randint(0, 256, (B, 64))
    m8 = torch.rand(B, 64) < 0.3
    _, tc = compile_ref(image_circuit((1, 8, 8), "quad-tree-2", "cp", 4))
    write_full("qt8_cp_k4", tc, x8, masks=m8)
    _, tc = compile_ref(image_circuit((1, 8, 8), "quad-graph", "cp", 4))
    write_full("qg8_cp_k4", tc, x8)
    _, tc = compile_ref(image_circuit((1, 8, 8), "quad-tree-2", "tucker", 4))
    write_full("qt8_tucker_k4", tc, x8)
    _, tc = compile_ref(image_circuit((1, 8, 8), "quad-tree-2", "tucker", 3), optimize=False)
    write_full("qt8_kronecker_k3_unopt", tc, x8)
    _, tc = compile_ref(image_circuit((1, 8, 8), "quad-tree-4", "cp-t", 5))
    write_full("qt8x4_cpt_k5", tc, x8)
    _, tc = compile_ref(image_circuit((1, 6, 6), "poon-domingos", "cp", 3), optimize=False)
    write_full("pd6_cp_k3_unopt", tc, torch.randint(0, 256, (B, 36)))
    _, tc = compile_ref(image_circuit((1, 8, 8), "quad-graph", "cp", 4, use_mixing_weights=False))
    write_full("qg8_cp_k4_densemix", tc, x8)
    _, tc = compile_ref(image_circuit((1, 8, 8), "quad-tree-2", "cp", 6, input_layer="embedding"))
    # embedding weights must be positive for the log-space path: re-draw them uniform
    low = plan_from_torch(tc)
    with torch.no_grad():
        for p, s in zip(low.leaves, low.plan.leaves):
            if len(s.shape) == 3 and s.shape[-1] == 256:
                p.uniform_(0.05, 1.0)
    write_full("qt8_cp_k6_embedding", tc, x8)
    # Gaussian leaves over a random binary tree (tabular route)
    sc = data_modalities.tabular_data(
        "random-binary-tree", num_features=12, input_layers={"name": "gaussian", "args": {}},
        num_input_units=5, sum_product_layer="cp", num_sum_units=5,
    )
    _, tc = compile_ref(sc)
    xg = torch.randn(B, 12)
    write_full("rbt12_gaussian_k5", tc, xg, masks=torch.rand(B, 12) < 0.4)
    # 1-D Gaussian mixture, 8 components (BASELINE.json configs[0])
    g = GaussianLayer(Scope((0,)), 8)
    s = SumLayer(8, 1, 1, weight_factory=lambda shape: Parameter.from_unary(
        SoftmaxParameter(shape), TensorParameter(*shape, initializer=NormalInitializer())))
    sc = Circuit([g, s], in_layers={s: [g]}, outputs=[s])
    _, tc = compile_ref(sc)
    write_full("gmm1d_k8", tc, torch.randn(64, 1))
    # categorical layers parameterised by unnormalised logits (integrate -> logsumexp)
    ins = [CategoricalLayer(Scope((v,)), 3, num_categories=5,
                            logits_factory=lambda shape: Parameter.from_input(
                                TensorParameter(*shape, initializer=NormalInitializer())))
           for v in range(4)]
    h1, h2 = HadamardLayer(3, 2), HadamardLayer(3, 2)
    wf = lambda shape: Parameter.from_unary(
        SoftmaxParameter(shape), TensorParameter(*shape, initializer=NormalInitializer()))
    s1, s2 = SumLayer(3, 3, 1, weight_factory=wf), SumLayer(3, 3, 1, weight_factory=wf)
    h3 = HadamardLayer(3, 2)
    s3 = SumLayer(3, 1, 1, weight_factory=wf)
    sc = Circuit(ins + [h1, h2, s1, s2, h3, s3],
                 in_layers={h1: ins[:2], h2: ins[2:], s1: [h1], s2: [h2], h3: [s1, s2], s3: [h3]},
                 outputs=[s3])
    _, tc = compile_ref(sc)
    xl = torch.randint(0, 5, (B, 4))
    write_full("cat_logits_k3", tc, xl)
    # Marginals of unnormalised categoricals: integrating a variable yields logsumexp(logits)
    # (layers/input.py:414-421).  The reference returns that as (F, K) instead of (F, 1, K), so
    # its IntegrateQuery (queries.py:143, torch.where) only broadcasts correctly when every
    # Categorical layer has ONE fold -- i.e. for the un-folded circuit, which is what this fixture
    # is generated from (a folded circuit raises "size of tensor a must match ..." there).
    _, tcu = compile_ref(sc, fold=False)
    write_full("cat_logits_k3_unfolded", tcu, xl, masks=torch.rand(B, 4) < 0.4)


def case_bench():
    torch.manual_seed(42)
    x = torch.randint(0, 256, (8, 784))
    _, tc = compile_ref(image_circuit((1, 28, 28), "quad-tree-2", "cp", 32))
    write_seeded("qt28_cp_k32", tc, x, seed=1234, extra={"units": 32, "config": 1})
    _, tc = compile_ref(image_circuit((1, 28, 28), "quad-tree-2", "cp", 64))
    write_seeded("qt28_cp_k64", tc, x, seed=1234, extra={"units": 64, "config": "north-star"})
    _, tc = compile_ref(image_circuit((1, 28, 28), "quad-tree-2", "tucker", 64))
    write_seeded("qt28_tucker_k64", tc, x[:4], seed=1234, extra={"units": 64, "config": 2})


def case_pd32():
    """BASELINE.json configs[3] structure: PoonDomingos over 3x32x32 Categorical inputs, CP layers
    (fold + optimize), built at K = 4 (the fold / gather structure does not depend on K; tests and
    bench.py resize it with CircuitPlan.with_units)."""
    torch.manual_seed(42)
    x = torch.randint(0, 256, (4, 3 * 32 * 32))
    _, tc = compile_ref(image_circuit((3, 32, 32), "poon-domingos", "cp", 4))
    write_seeded("pd32_cp_k4", tc, x, seed=1234, extra={"units": 4, "config": 3})


def case_complex():
    """'complex-lse-sum' circuits (BASELINE.json configs[4] structure at small size; SURVEY §8 a7,
    a16): complex Embedding inputs and complex sum weights, real part and imaginary part uniform
    in [0, 1) as in notebooks/sum-of-squares-circuits.ipynb cell 12.  Fixture kind "complex": the
    CUDA runtime has no kernels for this semiring yet, the oracle is pinned here ahead of them."""
    B = 16
    torch.manual_seed(42)
    cplx = utils.Parameterization(dtype="complex", initialization="uniform")
    sc = data_modalities.tabular_data(
        "random-binary-tree", num_features=16,
        input_layers={"name": "embedding", "args": {
            "num_states": 16, "weight_factory": utils.parameterization_to_factory(cplx)}},
        num_input_units=4, sum_product_layer="cp-t", num_sum_units=4, sum_weight_param=cplx)
    ctx, tc = compile_ref(sc, semiring="complex-lse-sum")
    x = torch.randint(0, 16, (B, 16))
    write_full("rbt16_cpt_k4_complex", tc, x, kind="complex")
    # the conjugate circuit shares c's leaves through `conj` parameter nodes
    tcc = ctx.compile(SF.conjugate(sc))
    write_full("rbt16_cpt_k4_complex_conj", tcc, x, kind="complex")
    # Tucker, un-optimised CP (sum / hadamard / mixing layers) and a circuit with real-valued
    # Categorical inputs under complex sum weights (map_from(LSE) = cast, semiring.py:511-514)
    emb = {"name": "embedding", "args": {
        "num_states": 16, "weight_factory": utils.parameterization_to_factory(cplx)}}
    cat = {"name": "categorical", "args": {"num_categories": 16}}
    for name, layers, spl, K, D, optimize in [
            ("rbt8_tucker_k3_complex", emb, "tucker", 3, 8, True),
            ("rbt12_cp_k3_complex_unopt", emb, "cp", 3, 12, False),
            ("rbt8_cpt_k4_complex_categorical", cat, "cp-t", 4, 8, True)]:
        sc = data_modalities.tabular_data(
            "random-binary-tree", num_features=D, input_layers=layers, num_input_units=K,
            sum_product_layer=spl, num_sum_units=K, sum_weight_param=cplx)
        _, tc = compile_ref(sc, optimize=optimize, semiring="complex-lse-sum")
        write_full(name, tc, torch.randint(0, 16, (B, D)), kind="complex")


def case_sampling():
    """Fixtures of kind "sampling" (SamplingQuery, queries.py:187-275): small circuits over a few
    categorical variables with x = EVERY joint state, so that `y` is the reference's exact joint
    log-distribution -- what the sampler's empirical frequencies are compared with, exactly as the
    reference's own test does (tests/backend/torch/test_queries/test_sampling.py:18-53).  Also
    stored: what the reference's SamplingQuery itself returns for 20000 samples under seed 123
    (its empirical world frequencies; the CUDA sampler uses another random stream)."""
    import functools

    from cirkit.backend.torch.queries import SamplingQuery
    from cirkit.symbolic.parameters import mixing_weight_factory
    from cirkit.templates.region_graph import PoonDomingos, QuadGraph, QuadTree, RandomBinaryTree
    from cirkit.templates.utils import name_to_input_layer_factory, parameterization_to_factory

    soft = utils.Parameterization(activation="softmax", initialization="normal")

    def build(rg, spl, K, V, mixing=True):
        wf = parameterization_to_factory(soft)
        return rg.build_circuit(
            input_factory=name_to_input_layer_factory("categorical", num_categories=V, probs_factory=wf),
            sum_product=spl, sum_weight_factory=wf,
            nary_sum_weight_factory=functools.partial(mixing_weight_factory, param_factory=wf) if mixing else wf,
            num_input_units=K, num_sum_units=K, num_classes=1, factorize_multivariate=True)

    cases = [
        ("smp_qt16_cp_k3", QuadTree((1, 4, 4), num_patch_splits=2), "cp", 3, 2, True, True, True),
        ("smp_qg9_cp_k3_unopt", QuadGraph((1, 3, 3)), "cp", 3, 3, True, True, False),
        ("smp_qg9_cpt_k4", QuadGraph((1, 3, 3)), "cp-t", 4, 3, True, True, True),
        ("smp_pd16_cp_k2", PoonDomingos((1, 4, 4), delta=1), "cp", 2, 2, True, True, True),
        ("smp_pd9_cp_k3_concat", PoonDomingos((1, 3, 3), delta=1), "cp", 3, 3, False, True, True),
        ("smp_rbt8_tucker_k3", RandomBinaryTree(8), "tucker", 3, 3, True, True, True),
        ("smp_qt4_tucker_k3_unopt", QuadTree((1, 2, 2), num_patch_splits=2), "tucker", 3, 5, True, True, False),
    ]
    for name, rg, spl, K, V, mixing, fold, optimize in cases:
        if NAMES is not None and name not in NAMES:
            continue
        _, tc = compile_ref(build(rg, spl, K, V, mixing), fold=fold, optimize=optimize)
        D = tc.num_variables
        worlds = torch.tensor(list(itertools.product(range(V), repeat=D)), dtype=torch.int64)
        extra = {"num_categories": V}
        try:  # the reference has no sample() for Tucker layers (layers/inner.py:66-71)
            torch.manual_seed(123)
            smp, _ = SamplingQuery(tc)(num_samples=20000)
            idx = (smp * torch.tensor([V ** (D - 1 - i) for i in range(D)])).sum(dim=1)
            extra["ref_counts"] = torch.bincount(idx, minlength=V ** D).tolist()
        except (TypeError, NotImplementedError, RuntimeError) as exc:  # Kronecker.sample: shape bug -> 76 GB alloc
            extra["ref_sampling_error"] = type(exc).__name__
        write_full(name, tc, worlds, kind="sampling", extra=extra)


CASES = {
    "sampling": case_sampling,
    "complex": case_complex,
    "pd32": case_pd32,
    "ka_categorical": case_ka_categorical,
    "ka_gaussian": case_ka_gaussian,
    "random_small": case_random_small,
    "bench": case_bench,
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--names", nargs="*", default=None, help="write only these fixtures")
    args = ap.parse_args()
    if args.names:
        NAMES = set(args.names)
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(42)
    np.random.seed(42)
    for name, fn in CASES.items():
        if args.only is None or name in args.only:
            fn()
