"""Circuit plan: the static dataflow program a compiled (folded) circuit lowers to.

The reference executes a compiled circuit by looping over *address-book entries*
(`cirkit/backend/torch/graph/modules.py:303-335`): every entry names a folded layer, the
producer layers whose outputs are concatenated along the fold axis, and an `(F, H)` index
tensor selecting the rows each fold reads (`cirkit/backend/torch/circuits.py:30-71`,
`cirkit/backend/torch/graph/folding.py:202-243`).  A :class:`CircuitPlan` is that same
information as plain data (numpy arrays + small dataclasses), independent of both the
reference package and of the device runtime:

* one :class:`StepSpec` per folded layer, in topological order;
* gathers expressed as ``(source step, fold)`` pairs resolved from the reference's
  concatenation order (first-seen producer order, `folding.py:210-222`);
* one :class:`ParamSpec` per layer parameter: a leaf tensor plus the chain of
  re-parameterisation ops the reference evaluates every forward
  (`cirkit/backend/torch/parameters/parameter.py:180-188`).

A plan is what the runtime (`cirkit_b200.runtime`), the CPU oracle (`oracle/`) and the golden
fixtures (`tests/golden/`) all consume, and it can be stored as a single ``.npz`` file.
"""

from __future__ import annotations

import dataclasses
import io
import json
from dataclasses import dataclass, field
from typing import Any, Callable, Sequence

import numpy as np

# Step kinds (names follow the reference layer classes they restate).
# "external": an input layer of a kind without a kernel, evaluated by the host per call (adapter only)
INPUT_KINDS = ("categorical", "gaussian", "embedding", "constant", "external")
INNER_KINDS = ("sum", "cpt", "mixing", "hadamard", "kronecker", "tucker", "tensordot")
ALL_KINDS = INPUT_KINDS + INNER_KINDS

# Re-parameterisation ops a ParamSpec chain may hold (reference:
# `cirkit/backend/torch/parameters/nodes.py`): name -> attrs.
PARAM_OPS = (
    "softmax",  # TorchSoftmaxParameter nodes.py:764-772 (attrs: dim, relative to un-folded shape)
    "log_softmax",  # TorchLogSoftmaxParameter nodes.py:775-783
    "scaled_sigmoid",  # TorchScaledSigmoidParameter nodes.py:682-699 (attrs: vmin, vmax)
    "sigmoid",  # TorchSigmoidParameter nodes.py:677-679
    "exp",  # nodes.py:656-660
    "log",  # nodes.py:663-667
    "square",  # nodes.py:670-674
    "softplus",  # nodes.py:730-738
    "clamp",  # nodes.py:702-727 (attrs: vmin, vmax)
    "mixing",  # TorchMixingWeightParameter nodes.py:847-862
    # TorchMatMulParameter nodes.py:786-805 (SumCollapse): attrs {"rhs": {"leaf", "ops", "fold_idx"}},
    # the right operand being another leaf -> op chain of the same plan
    "matmul",
    "conj",  # TorchConjugateParameter nodes.py (complex circuits: conjugate(c) shares c's leaves)
)

SEMIRINGS = ("lse-sum", "complex-lse-sum")  # cirkit/backend/torch/semiring.py:326-408, :410-476


@dataclass
class LeafSpec:
    """A learnable leaf tensor (`TorchTensorParameter._ptensor`, nodes.py:188-201)."""

    shape: tuple[int, ...]  # including the fold axis
    init: str = "normal"  # how reset_parameters() fills it ("normal" = nn.init.normal_)
    requires_grad: bool = True
    name: str = ""  # state_dict key of the tensor in the reference module, if known
    dtype: str = "float"  # "float" | "complex" (complex64 leaves of 'complex-lse-sum' circuits)


@dataclass
class ParamSpec:
    """One layer parameter = leaf tensor -> optional fold slice -> chain of ops."""

    leaf: int  # index into CircuitPlan.leaves; -1 = external (evaluated by the host, adapter only)
    ops: list[tuple[str, dict[str, Any]]] = field(default_factory=list)
    shape: tuple[int, ...] = ()  # effective shape, including the fold axis
    fold_idx: np.ndarray | None = None  # TorchPointerParameter slice (nodes.py:223-279)


@dataclass
class StepSpec:
    """One folded layer of the circuit."""

    kind: str
    num_folds: int
    arity: int
    num_input_units: int
    num_output_units: int
    params: dict[str, ParamSpec] = field(default_factory=dict)
    # inner layers: for every fold f and input h, the producer step and the fold within it
    in_step: np.ndarray | None = None  # (F, H) int32
    in_fold: np.ndarray | None = None  # (F, H) int32
    # input layers: variable read by every fold (all input layers on the path are univariate)
    scope_idx: np.ndarray | None = None  # (F,) int32
    config: dict[str, Any] = field(default_factory=dict)

    @property
    def is_input(self) -> bool:
        return self.kind in INPUT_KINDS


@dataclass
class CircuitPlan:
    steps: list[StepSpec]
    leaves: list[LeafSpec]
    out_step: np.ndarray  # (O,) int32 producer step of every circuit output
    out_fold: np.ndarray  # (O,) int32 fold within it
    num_variables: int  # max(scope) + 1: the width D of the input matrix
    scope: tuple[int, ...]  # variables the circuit is defined on
    semiring: str = "lse-sum"
    meta: dict[str, Any] = field(default_factory=dict)

    # ---------------------------------------------------------------- derived sizes
    @property
    def num_outputs(self) -> int:
        return int(self.out_step.shape[0])

    @property
    def num_output_units(self) -> int:
        return self.steps[int(self.out_step[0])].num_output_units

    def activation_units(self) -> int:
        """A of SURVEY §8(d): activation units per sample, summed over steps."""
        return sum(s.num_folds * s.num_output_units for s in self.steps)

    def parameter_elements(self) -> int:
        """P of SURVEY §8(d): learnable scalars."""
        return sum(int(np.prod(l.shape)) for l in self.leaves)

    def algorithmic_bytes(self, batch: int, x_itemsize: int = 8) -> int:
        """Bytes one forward+backward pass must move (SURVEY §8(d) formula; complex circuits move
        8 bytes per unit and per complex parameter)."""
        unit = 8 if self.semiring == "complex-lse-sum" else 4
        pbytes = sum(int(np.prod(l.shape)) * (8 if l.dtype == "complex" else 4) for l in self.leaves)
        return (
            x_itemsize * batch * self.num_variables
            + 5 * unit * self.activation_units() * batch
            + 7 * pbytes
        )

    # ---------------------------------------------------------------- validation
    def validate(self) -> None:
        if self.semiring not in SEMIRINGS:
            raise ValueError(f"unknown semiring {self.semiring!r}")
        for lid, l in enumerate(self.leaves):
            if l.dtype not in ("float", "complex"):
                raise ValueError(f"leaf {lid}: unknown dtype {l.dtype!r}")
            if l.dtype == "complex" and self.semiring != "complex-lse-sum":
                raise ValueError(f"leaf {lid}: complex leaf in a {self.semiring!r} circuit")
        for sid, s in enumerate(self.steps):
            if s.kind not in ALL_KINDS:
                raise ValueError(f"step {sid}: unknown kind {s.kind!r}")
            if s.is_input:
                if s.kind not in ("constant", "external"):
                    if s.scope_idx is None or s.scope_idx.shape != (s.num_folds,):
                        raise ValueError(f"step {sid}: scope_idx must have shape (F,)")
            else:
                if s.in_step is None or s.in_fold is None:
                    raise ValueError(f"step {sid}: inner step without inputs")
                if s.in_step.shape != (s.num_folds, s.arity) or s.in_fold.shape != s.in_step.shape:
                    raise ValueError(f"step {sid}: gather index must have shape (F, H)")
                if np.any(s.in_step >= sid) or np.any(s.in_step < 0):
                    raise ValueError(f"step {sid}: gather from a later step")
                for p, f in zip(s.in_step.ravel(), s.in_fold.ravel()):
                    src = self.steps[int(p)]
                    if not 0 <= f < src.num_folds:
                        raise ValueError(f"step {sid}: fold {f} out of range for step {p}")
                    if src.num_output_units != s.num_input_units:
                        raise ValueError(
                            f"step {sid}: producer {p} has {src.num_output_units} units, "
                            f"expected {s.num_input_units}"
                        )
            for name, p in s.params.items():
                if p.leaf >= len(self.leaves):
                    raise ValueError(f"step {sid}: parameter {name} points past the leaf table")
                for op, attrs in p.ops:
                    if op not in PARAM_OPS:
                        raise ValueError(f"step {sid}: parameter op {op!r} is not supported")
                    if op == "conj" and self.semiring != "complex-lse-sum":
                        raise ValueError(f"step {sid}: 'conj' outside a complex circuit")
                    if op == "matmul" and not 0 <= attrs["rhs"]["leaf"] < len(self.leaves):
                        raise ValueError(f"step {sid}: matmul operand points past the leaf table")

    # ---------------------------------------------------------------- (de)serialisation
    def to_bytes(self) -> bytes:
        arrays: dict[str, np.ndarray] = {
            "out_step": self.out_step.astype(np.int32),
            "out_fold": self.out_fold.astype(np.int32),
        }
        steps_js = []
        for sid, s in enumerate(self.steps):
            js: dict[str, Any] = {
                "kind": s.kind,
                "num_folds": s.num_folds,
                "arity": s.arity,
                "num_input_units": s.num_input_units,
                "num_output_units": s.num_output_units,
                "config": s.config,
                "params": {},
            }
            if s.in_step is not None:
                arrays[f"s{sid}_in_step"] = s.in_step.astype(np.int32)
                arrays[f"s{sid}_in_fold"] = s.in_fold.astype(np.int32)
            if s.scope_idx is not None:
                arrays[f"s{sid}_scope"] = s.scope_idx.astype(np.int32)
            for name, p in s.params.items():
                if p.leaf < 0:
                    raise ValueError("a plan holding external parameters cannot be serialised")
                pj = {"leaf": p.leaf, "ops": p.ops, "shape": list(p.shape), "fold_idx": False}
                if p.fold_idx is not None:
                    arrays[f"s{sid}_p_{name}_fold_idx"] = np.asarray(p.fold_idx, dtype=np.int64)
                    pj["fold_idx"] = True
                js["params"][name] = pj
            steps_js.append(js)
        header = {
            "version": 1,
            "steps": steps_js,
            "leaves": [dataclasses.asdict(l) for l in self.leaves],
            "num_variables": self.num_variables,
            "scope": list(self.scope),
            "semiring": self.semiring,
            "meta": self.meta,
        }
        arrays["header"] = np.frombuffer(json.dumps(header).encode(), dtype=np.uint8)
        buf = io.BytesIO()
        np.savez_compressed(buf, **arrays)
        return buf.getvalue()

    def save(self, path: str) -> None:
        with open(path, "wb") as fh:
            fh.write(self.to_bytes())

    @classmethod
    def load(cls, path_or_bytes: str | bytes) -> "CircuitPlan":
        if isinstance(path_or_bytes, (bytes, bytearray)):
            z = np.load(io.BytesIO(path_or_bytes))
        else:
            z = np.load(path_or_bytes)
        header = json.loads(bytes(z["header"]).decode())
        steps = []
        for sid, js in enumerate(header["steps"]):
            params = {}
            for name, pj in js["params"].items():
                params[name] = ParamSpec(
                    leaf=pj["leaf"],
                    ops=[(o, dict(a)) for o, a in pj["ops"]],
                    shape=tuple(pj["shape"]),
                    fold_idx=z[f"s{sid}_p_{name}_fold_idx"] if pj["fold_idx"] else None,
                )
            steps.append(
                StepSpec(
                    kind=js["kind"],
                    num_folds=js["num_folds"],
                    arity=js["arity"],
                    num_input_units=js["num_input_units"],
                    num_output_units=js["num_output_units"],
                    params=params,
                    in_step=z[f"s{sid}_in_step"] if f"s{sid}_in_step" in z else None,
                    in_fold=z[f"s{sid}_in_fold"] if f"s{sid}_in_fold" in z else None,
                    scope_idx=z[f"s{sid}_scope"] if f"s{sid}_scope" in z else None,
                    config=js["config"],
                )
            )
        leaves = [
            LeafSpec(tuple(l["shape"]), l["init"], l["requires_grad"], l.get("name", ""),
                     l.get("dtype", "float"))
            for l in header["leaves"]
        ]
        plan = cls(
            steps=steps,
            leaves=leaves,
            out_step=z["out_step"],
            out_fold=z["out_fold"],
            num_variables=header["num_variables"],
            scope=tuple(header["scope"]),
            semiring=header["semiring"],
            meta=header.get("meta", {}),
        )
        plan.validate()
        return plan

    # ---------------------------------------------------------------- utilities
    def with_units(self, k: int) -> "CircuitPlan":
        """Return the same structure with every unit axis of size ``meta['units']`` resized to
        ``k`` (layers with a single output unit, e.g. the root sum, keep it).

        The fold/gather structure a region graph compiles to does not depend on the number of
        units, so the committed structure fixtures (built from the reference at a small K) are
        re-sized to the benchmark K here.  Parameter shapes are re-derived from the layer kind
        (`cirkit/backend/torch/layers/*`: Categorical/Embedding (F,K,V), Gaussian (F,K), sum
        (F,Ko,H*Ki), CP-T (F,Ko,Ki), Tucker (F,Ko,Ki^H), mixing (F,K,H)), not matched by value, so
        a fixture with K = 3 and an arity-3 layer resizes correctly.
        """
        k0 = self.meta.get("units")
        if k0 is None:
            raise ValueError("plan has no meta['units']; cannot resize")
        if k0 == k:
            return self

        def r(n: int) -> int:
            return k if n == k0 else n

        steps, leaf_shapes = [], {}

        def resize_leaf(leaf: int, ops, eff_shape) -> None:
            """New shape of a leaf from the op chain that leads from it to a layer parameter."""
            old = self.leaves[leaf].shape
            names = [o for o, _ in ops]
            if names and names[-1] == "matmul":
                # W1 @ W2 (SumCollapse): W1 is (F, Ko, K); W2 has its own chain
                new = (old[0], r(old[1]), r(old[2]))
                rhs = ops[-1][1]["rhs"]
                resize_leaf(rhs["leaf"], [(o, a) for o, a in rhs["ops"]], None)
            elif "mixing" in names:
                new = (old[0], r(old[1])) + tuple(old[2:])  # mixing weights are (F, K, H)
            elif eff_shape is not None:
                new = (old[0],) + tuple(eff_shape[1:])
            else:
                new = (old[0],) + tuple(r(d) for d in old[1:])
            if leaf_shapes.setdefault(leaf, new) != new:
                raise ValueError(f"leaf {leaf} is shared by layers of different shapes")

        for sid, s in enumerate(self.steps):
            ki = s.num_input_units if s.is_input else r(s.num_input_units)
            ko = r(s.num_output_units)
            params = {}
            for n, p in s.params.items():
                F = p.shape[0]
                if s.kind in ("categorical", "embedding"):
                    shape = (F, ko, p.shape[2])
                elif s.kind in ("gaussian", "constant"):
                    shape = (F, ko) + tuple(p.shape[2:])
                elif s.kind == "sum":
                    shape = (F, ko, s.arity * ki)
                elif s.kind == "cpt":
                    shape = (F, ko, ki)
                elif s.kind == "tucker":
                    shape = (F, ko, ki ** s.arity)
                elif s.kind == "mixing":
                    shape = (F, ko, s.arity)
                else:
                    shape = tuple(p.shape)
                if len(shape) != len(p.shape):
                    raise ValueError(f"step {sid} parameter {n!r}: cannot resize shape {p.shape}")
                params[n] = ParamSpec(p.leaf, [(o, dict(a)) for o, a in p.ops], shape, p.fold_idx)
                if p.leaf >= 0:
                    resize_leaf(p.leaf, p.ops, shape)
            steps.append(dataclasses.replace(s, num_input_units=ki, num_output_units=ko, params=params))
        leaves = [dataclasses.replace(l, shape=tuple(leaf_shapes.get(i, l.shape)))
                  for i, l in enumerate(self.leaves)]
        meta = dict(self.meta)
        meta["units"] = k
        plan = dataclasses.replace(self, steps=steps, leaves=leaves, meta=meta)
        plan.validate()
        return plan


# -------------------------------------------------------------------- memory layout
@dataclass
class PlanLayout:
    """Arena addressing derived from a plan (all offsets are *per sample*, in floats).

    Activations live in one arena: step ``s`` writes a contiguous ``(F_s, B, Ko_s)`` block at
    float offset ``B * out_off[s]``.  Because every offset scales linearly with the batch size
    the layout is built once per plan.  The reference instead keeps one tensor per layer and
    materialises ``cat(...)[idx]`` copies for every gather (`circuits.py:42-47`); here a gather
    is just the list of row offsets ``in_rows``.

    Backward is pull-based: step ``c`` stores the gradient w.r.t. its *inputs* in the grad arena
    (block ``gin_off[c]``, shape ``(F_c, Hg, B, Ki)`` with ``Hg = 1`` when all inputs of a fold
    receive the same gradient, i.e. Hadamard-style products), and every producer row sums the
    rows listed in its CSR consumer list.  This makes the accumulation over multiple consumers
    (`graph/modules.py:325-334` keeps outputs alive for exactly that reason) deterministic and
    atomics-free.
    """

    out_off: np.ndarray  # (S,) int64
    in_rows: list[np.ndarray | None]  # per step (F*H,) int64 arena offsets of the gathered rows
    gin_off: np.ndarray  # (S,) int64 (-1: step has no input gradient block)
    gin_h: np.ndarray  # (S,) int32
    cons_ptr: list[np.ndarray]  # per step (F+1,) int32
    cons_rows: list[np.ndarray]  # per step (nnz,) int64 grad-arena offsets
    out_rows: np.ndarray  # (O,) int64 arena offsets of the circuit outputs
    out_goff: int  # grad-arena offset of the (O, B, K) output-gradient block
    arena_units: int  # floats per sample in the activation arena
    garena_units: int  # floats per sample in the gradient arena


SHARED_GRAD_KINDS = ("cpt", "hadamard")  # every input of a fold receives the same gradient


def _align(n: int, a: int = 64) -> int:
    """Blocks start at multiples of 64 floats per sample (256 bytes times the batch size): the
    TMA-fed kernels address the arenas as matrices of 64-float rows, and every vector access of
    the others stays 16-byte aligned."""
    return (n + a - 1) // a * a


def build_layout(plan: CircuitPlan) -> PlanLayout:
    S = len(plan.steps)
    out_off = np.zeros(S, dtype=np.int64)
    off = 0
    for sid, s in enumerate(plan.steps):
        out_off[sid] = off
        off = _align(off + s.num_folds * s.num_output_units)
    arena_units = off

    gin_off = np.full(S, -1, dtype=np.int64)
    gin_h = np.ones(S, dtype=np.int32)
    goff = 0
    in_rows: list[np.ndarray | None] = []
    for sid, s in enumerate(plan.steps):
        if s.is_input:
            in_rows.append(None)
            continue
        src_k = s.num_input_units
        rows = out_off[s.in_step.astype(np.int64)] + s.in_fold.astype(np.int64) * src_k
        in_rows.append(np.ascontiguousarray(rows.reshape(-1)))
        gin_h[sid] = 1 if s.kind in SHARED_GRAD_KINDS else s.arity
        gin_off[sid] = goff
        goff = _align(goff + s.num_folds * int(gin_h[sid]) * s.num_input_units)
    K = plan.num_output_units
    out_goff = goff
    goff = _align(goff + plan.num_outputs * K)
    garena_units = goff

    # consumer lists: for each producer row, the grad-arena rows to sum
    cons: list[list[list[int]]] = [[[] for _ in range(s.num_folds)] for s in plan.steps]
    for cid, c in enumerate(plan.steps):
        if c.is_input:
            continue
        hg = int(gin_h[cid])
        for f in range(c.num_folds):
            for h in range(c.arity):
                p, pf = int(c.in_step[f, h]), int(c.in_fold[f, h])
                hh = 0 if hg == 1 else h
                cons[p][pf].append(int(gin_off[cid]) + (f * hg + hh) * c.num_input_units)
    for o in range(plan.num_outputs):
        cons[int(plan.out_step[o])][int(plan.out_fold[o])].append(out_goff + o * K)
    cons_ptr, cons_rows = [], []
    for sid, s in enumerate(plan.steps):
        ptr = np.zeros(s.num_folds + 1, dtype=np.int32)
        flat: list[int] = []
        for f in range(s.num_folds):
            flat.extend(cons[sid][f])
            ptr[f + 1] = len(flat)
        cons_ptr.append(ptr)
        cons_rows.append(np.asarray(flat, dtype=np.int64))
    out_rows = out_off[plan.out_step.astype(np.int64)] + plan.out_fold.astype(np.int64) * K
    return PlanLayout(
        out_off=out_off,
        in_rows=in_rows,
        gin_off=gin_off,
        gin_h=gin_h,
        cons_ptr=cons_ptr,
        cons_rows=cons_rows,
        out_rows=out_rows.astype(np.int64),
        out_goff=out_goff,
        arena_units=arena_units,
        garena_units=garena_units,
    )


def init_leaf_(t, spec: LeafSpec) -> None:
    """Fill a leaf the way the reference's ``reset_parameters`` does (nodes.py:188-201)."""
    import torch

    with torch.no_grad():
        if spec.init == "normal":
            torch.nn.init.normal_(t)
        elif spec.init == "uniform":
            torch.nn.init.uniform_(t)
        elif spec.init == "zeros":
            t.zero_()
        elif spec.init == "ones":
            t.fill_(1.0)
        else:
            raise ValueError(f"unknown initialiser {spec.init!r}")


def seeded_leaves(plan: CircuitPlan, seed: int) -> list:
    """Deterministic leaf values shared by the fixture generator, the tests and the bench:
    one CPU generator, float32 standard normals drawn leaf by leaf in plan order."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for l in plan.leaves:
        if l.dtype == "complex":
            # real and imaginary parts uniform in [0, 1), as notebooks/sum-of-squares-circuits.ipynb
            # initialises complex weights (cell 12)
            out.append(torch.view_as_complex(torch.rand((*l.shape, 2), generator=g, dtype=torch.float32)))
        else:
            out.append(torch.randn(l.shape, generator=g, dtype=torch.float32))
    return out
