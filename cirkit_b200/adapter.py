"""Bridge from a circuit compiled by the reference (`cirkit.pipeline.compile`) to a plan.

This module is the only place that touches the reference package, and it imports it lazily:
everything else in `cirkit_b200` runs without `cirkit` installed (e.g. from a stored plan).

* :func:`plan_from_torch` walks `TorchCircuit.address_book`
  (`cirkit/backend/torch/graph/modules.py:168-187`) and every layer's `config` / `params`
  (`layers/base.py:54-75`) and lowers them to a :class:`~cirkit_b200.plan.CircuitPlan`.
* :func:`accelerate` swaps the executor of an existing `TorchCircuit` in place: the object keeps
  its class hierarchy, layers, address book, `state_dict` keys and -- crucially -- the very same
  `nn.Parameter` leaves (`TorchTensorParameter._ptensor`, `parameters/nodes.py:193-201`), so
  optimisers, checkpoints and pointer-shared integrate/multiply circuits keep working; only
  `forward` / `evaluate` are re-routed to the CUDA runtime.
* :func:`register_backend` makes `PipelineContext(backend="b200", ...)` work by extending the
  reference's backend switch (`cirkit/pipeline.py:348-356`, `backend/compiler.py:11`).
"""

from __future__ import annotations

import functools
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch
from torch import nn

from .plan import CircuitPlan, LeafSpec, ParamSpec, StepSpec


class UnsupportedCircuitError(NotImplementedError):
    """The circuit holds a layer or parameterisation the CUDA runtime has no kernel for."""


@dataclass
class LoweredCircuit:
    plan: CircuitPlan
    leaves: list[nn.Parameter]  # the reference's own leaf tensors, in plan.leaves order
    externals: dict[tuple[int, str], Any] = field(default_factory=dict)  # (step, name) -> TorchParameter


# --------------------------------------------------------------------------- parameters
def _lower_param(p, leaf_ids: dict[int, int], leaves: list, leaf_specs: list[LeafSpec],
                 names: dict[int, str]) -> ParamSpec | None:
    """Recognise `leaf [-> pointer slice] -> unary op chain` parameter graphs.

    Returns None when the graph is anything else (kron/matmul/einsum nodes produced by
    multiply/integrate or by the SumCollapse rule) -- the caller then keeps it *external*:
    the host evaluates the reference's own `TorchParameter` with PyTorch and feeds the result
    to the kernels (such graphs run once per step on small tensors).
    """
    from cirkit.backend.torch.parameters import nodes as N

    op_names = {
        N.TorchSoftmaxParameter: "softmax",
        N.TorchLogSoftmaxParameter: "log_softmax",
        N.TorchScaledSigmoidParameter: "scaled_sigmoid",
        N.TorchSigmoidParameter: "sigmoid",
        N.TorchExpParameter: "exp",
        N.TorchLogParameter: "log",
        N.TorchSquareParameter: "square",
        N.TorchSoftplusParameter: "softplus",
        N.TorchClampParameter: "clamp",
        N.TorchMixingWeightParameter: "mixing",
        N.TorchConjugateParameter: "conj",
    }
    entries = list(p.address_book)
    mods = [e.module for e in entries[:-1]]
    if not mods:
        return None

    def identity(fi, n_folds: int) -> bool:
        if isinstance(fi, torch.Tensor):
            return fi.tolist() == list(range(n_folds))
        return fi == ()

    def leaf_index(t) -> int:
        key = id(t)
        if key not in leaf_ids:
            leaf_ids[key] = len(leaves)
            leaves.append(t)
            leaf_specs.append(LeafSpec(tuple(t.shape), "normal", bool(t.requires_grad), names.get(key, ""),
                                       "complex" if t.is_complex() else "float"))
        return leaf_ids[key]

    def chain(i: int):
        """(leaf tensor, fold_idx, ops) of node i, or None when the sub-graph is not a chain of
        unary ops over one leaf with at most `matmul` nodes joining two such chains."""
        m, e = mods[i], entries[i]
        if isinstance(m, (N.TorchTensorParameter, N.TorchPointerParameter)):
            fold_idx = None
            if isinstance(m, N.TorchPointerParameter):
                fold_idx = None if m._fold_idx is None else m._fold_idx.cpu().numpy().astype(np.int64)
                m = m.deref()
            if not isinstance(m, N.TorchTensorParameter) or m._ptensor is None:
                return None
            return m._ptensor, fold_idx, []
        if type(m) in op_names:
            # must consume exactly one earlier node, un-permuted
            if len(e.in_module_ids) != 1 or len(e.in_module_ids[0]) != 1 or len(e.in_fold_idx) != 1:
                return None
            j = e.in_module_ids[0][0]
            if not identity(e.in_fold_idx[0], mods[j].num_folds):
                return None
            sub = chain(j)
            if sub is None:
                return None
            attrs: dict[str, Any] = {}
            if hasattr(m, "dim"):
                attrs["dim"] = int(m.dim)
            if isinstance(m, (N.TorchScaledSigmoidParameter, N.TorchClampParameter)):
                attrs["vmin"] = None if m.vmin is None else float(m.vmin)
                attrs["vmax"] = None if m.vmax is None else float(m.vmax)
            return sub[0], sub[1], sub[2] + [(op_names[type(m)], attrs)]
        if isinstance(m, N.TorchMatMulParameter):
            # SumCollapse (cirkit/backend/torch/optimization/layers.py:30-47): W = W1 @ W2
            if len(e.in_module_ids) != 2 or any(len(ids) != 1 for ids in e.in_module_ids):
                return None
            (j1,), (j2,) = e.in_module_ids
            if not identity(e.in_fold_idx[0], mods[j1].num_folds) or not identity(e.in_fold_idx[1], mods[j2].num_folds):
                return None
            lhs, rhs = chain(j1), chain(j2)
            if lhs is None or rhs is None:
                return None
            rhs_spec = {"leaf": leaf_index(rhs[0]), "ops": [[o, a] for o, a in rhs[2]],
                        "fold_idx": None if rhs[1] is None else [int(v) for v in rhs[1]]}
            return lhs[0], lhs[1], lhs[2] + [("matmul", {"rhs": rhs_spec})]
        return None

    last = entries[-1]
    if last.in_module_ids != [[len(mods) - 1]]:
        return None
    if last.in_fold_idx[0].tolist() != list(range(mods[-1].num_folds)):
        return None
    low = chain(len(mods) - 1)
    if low is None:
        return None
    tensor, fold_idx, ops = low
    return ParamSpec(leaf_index(tensor), ops, (p.num_folds, *p.shape), fold_idx)


# --------------------------------------------------------------------------- layers
def plan_from_torch(tc, *, allow_external_params: bool = True,
                    semirings: tuple[str, ...] = ("lse-sum",)) -> LoweredCircuit:
    """Lower a compiled `TorchCircuit` to a plan (see module docstring).

    `semirings` lists the semirings the caller can execute.  The CUDA runtime implements
    'lse-sum' only, so that is the default and anything else raises UnsupportedCircuitError;
    the plan format itself also describes 'complex-lse-sum' circuits (complex64 leaves), which
    the test oracle evaluates -- fixtures for the complex kernels are generated that way."""
    from cirkit.backend.torch.layers.inner import (
        TorchHadamardLayer,
        TorchKroneckerLayer,
        TorchSumLayer,
    )
    from cirkit.backend.torch.layers.input import (
        TorchInputLayer,
        TorchCategoricalLayer,
        TorchConstantValueLayer,
        TorchEmbeddingLayer,
        TorchGaussianLayer,
    )
    from cirkit.backend.torch.layers.optimized import TorchCPTLayer, TorchTensorDotLayer, TorchTuckerLayer
    from cirkit.backend.torch.semiring import ComplexLSESumSemiring, LSESumSemiring

    semiring_names = {LSESumSemiring: "lse-sum", ComplexLSESumSemiring: "complex-lse-sum"}
    semiring = None

    names = {id(t): n for n, t in tc.named_parameters()}
    leaf_ids: dict[int, int] = {}
    leaves: list[nn.Parameter] = []
    leaf_specs: list[LeafSpec] = []
    externals: dict[tuple[int, str], Any] = {}
    steps: list[StepSpec] = []
    entries = list(tc.address_book)
    num_folds: list[int] = []

    def resolve(in_ids: list[int], idx, F: int, H: int) -> tuple[np.ndarray, np.ndarray]:
        sizes = [num_folds[i] for i in in_ids]
        total = sum(sizes)
        if isinstance(idx, torch.Tensor):
            flat = idx.cpu().numpy().astype(np.int64).reshape(-1)
        else:  # unsqueeze shortcuts, graph/folding.py:235-241
            flat = np.arange(total, dtype=np.int64)
        bounds = np.cumsum([0] + sizes)
        which = np.searchsorted(bounds, flat, side="right") - 1
        step = np.asarray(in_ids, dtype=np.int64)[which]
        fold = flat - bounds[which]
        return step.reshape(F, H).astype(np.int32), fold.reshape(F, H).astype(np.int32)

    for sid, e in enumerate(entries[:-1]):
        m = e.module
        if semiring_names.get(m.semiring) not in semirings:
            raise UnsupportedCircuitError(f"semiring {m.semiring.__name__} has no CUDA path")
        if semiring not in (None, m.semiring):
            raise UnsupportedCircuitError("layers of one circuit disagree on the semiring")
        semiring = m.semiring
        F = m.num_folds
        params: dict[str, ParamSpec] = {}
        for name, p in m.params.items():
            spec = _lower_param(p, leaf_ids, leaves, leaf_specs, names)
            if spec is None:
                if not allow_external_params:
                    raise UnsupportedCircuitError(f"step {sid}: parameter graph of {name!r}")
                spec = ParamSpec(-1, [], (p.num_folds, *p.shape), None)
                externals[(sid, name)] = p
            params[name] = spec
        kind: str
        config: dict[str, Any] = {}
        scope_idx = None
        in_step = in_fold = None
        if isinstance(m, TorchCategoricalLayer):
            kind = "categorical"
            config["num_categories"] = int(m.num_categories)
        elif isinstance(m, TorchEmbeddingLayer):
            kind = "embedding"
            config["num_states"] = int(m.num_states)
        elif isinstance(m, TorchGaussianLayer):
            kind = "gaussian"
        elif isinstance(m, TorchConstantValueLayer):
            kind = "constant"
            config["log_space"] = bool(m.log_space)
        elif isinstance(m, TorchCPTLayer):
            kind = "cpt"
        elif isinstance(m, TorchTuckerLayer):
            kind = "tucker"
        elif isinstance(m, TorchTensorDotLayer):  # layers/optimized.py:205-300
            kind = "tensordot"
            config["kq"] = int(m._num_batch_units)
        elif isinstance(m, TorchSumLayer):
            kind = "sum"
            w = params["weight"]
            if w.leaf >= 0 and w.ops and w.ops[-1][0] == "mixing":
                kind = "mixing"
                params["weight"] = ParamSpec(
                    w.leaf, w.ops[:-1], (F, m.num_output_units, m.arity), w.fold_idx
                )
        elif isinstance(m, TorchHadamardLayer):
            kind = "hadamard"
        elif isinstance(m, TorchKroneckerLayer):
            kind = "kronecker"
        elif isinstance(m, TorchInputLayer) and allow_external_params and semiring_names.get(m.semiring) == "lse-sum":
            # per-step fallback: an input layer kind without a kernel (Binomial,
            # Polynomial, Evidence, multivariate layers ...) is evaluated by the reference's own
            # module with PyTorch on every call; its (F, B, K) output enters the arena as a
            # differentiable external tensor, everything above it runs on the CUDA kernels
            kind = "external"
            params = {"output": ParamSpec(-1, [], (F, 0, int(m.num_output_units)), None)}
            externals[(sid, "output")] = m
        else:
            raise UnsupportedCircuitError(f"step {sid}: no CUDA kernel for {type(m).__name__}")
        if kind == "external":
            arity, k_in = 1, 0
        elif kind in ("categorical", "embedding", "gaussian"):
            if m.num_variables != 1:
                raise UnsupportedCircuitError(f"step {sid}: multivariate input layer")
            scope_idx = m.scope_idx.cpu().numpy().astype(np.int32).reshape(F)
            arity, k_in = 1, 1
        elif kind == "constant":
            arity, k_in = 1, 0
        else:
            arity, k_in = int(m.arity), int(m.num_input_units)
            if kind in ("tucker", "kronecker") and arity != 2:
                # the library rejects these at ckb_plan_create (plan.cu:check_step); say so here,
                # while accelerate() can still leave the circuit on the reference path
                raise UnsupportedCircuitError(
                    f"step {sid}: {kind} layers have CUDA kernels for arity 2 only (got {arity})")
            if kind == "mixing" and k_in != int(m.num_output_units):
                raise UnsupportedCircuitError(f"step {sid}: mixing layer with {k_in} != {m.num_output_units} units")
            in_step, in_fold = resolve(e.in_module_ids[0], e.in_fold_idx[0], F, arity)
        steps.append(
            StepSpec(kind, F, arity, k_in, int(m.num_output_units), params, in_step, in_fold,
                     scope_idx, config)
        )
        num_folds.append(F)

    last = entries[-1]
    out_step, out_fold = resolve(last.in_module_ids[0], last.in_fold_idx[0], -1, 1)
    scope = tuple(sorted(tc.scope))
    plan = CircuitPlan(
        steps=steps,
        leaves=leaf_specs,
        out_step=out_step.reshape(-1),
        out_fold=out_fold.reshape(-1),
        num_variables=(max(scope) + 1) if scope else 0,
        scope=scope,
        semiring=semiring_names.get(semiring, "lse-sum"),
    )
    plan.validate()
    return LoweredCircuit(plan, leaves, externals)


# --------------------------------------------------------------------------- executor swap
def _integrate_mask_of(module_fn):
    """The (B or 1, D) bool mask when `module_fn` is what the reference's own `IntegrateQuery`
    hands to `evaluate` -- `functools.partial(IntegrateQuery._layer_fn, integrate_vars_mask=m)`,
    cirkit/backend/torch/queries.py:101-107 -- else None."""
    from cirkit.backend.torch.queries import IntegrateQuery

    if (
        isinstance(module_fn, functools.partial)
        and module_fn.func is IntegrateQuery._layer_fn
        and not module_fn.args
        and set(module_fn.keywords) == {"integrate_vars_mask"}
    ):
        return module_fn.keywords["integrate_vars_mask"]
    return None


def _sampling_args_of(module_fn):
    """(num_samples, mixture_samples list) when `module_fn` is what the reference's own
    `SamplingQuery` hands to `evaluate` -- `functools.partial(query._layer_fn, num_samples=n,
    mixture_samples=lst)`, cirkit/backend/torch/queries.py:233-241 -- else None."""
    from cirkit.backend.torch.queries import SamplingQuery

    if (
        isinstance(module_fn, functools.partial)
        and getattr(module_fn.func, "__func__", None) is SamplingQuery._layer_fn
        and not module_fn.args
        and set(module_fn.keywords) == {"num_samples", "mixture_samples"}
    ):
        return module_fn.keywords["num_samples"], module_fn.keywords["mixture_samples"]
    return None


def accelerate(tc, *, strict: bool = False):
    """Re-route `tc(x)` to the CUDA runtime, in place; returns `tc`.

    The circuit object, its parameters and its `state_dict` are untouched.  When the circuit
    holds something the runtime cannot lower, `strict=False` leaves it on the reference path
    (and records why in `tc._b200_reason`); `strict=True` raises UnsupportedCircuitError.
    """
    from .runtime import PlanRuntime

    try:
        lowered = plan_from_torch(tc, semirings=("lse-sum", "complex-lse-sum"))
        runtime = PlanRuntime(lowered.plan)
    except NotImplementedError as exc:  # UnsupportedCircuitError, or a plan the runtime has no kernels for
        if strict:
            raise exc if isinstance(exc, UnsupportedCircuitError) else UnsupportedCircuitError(str(exc)) from exc
        tc._b200_reason = str(exc)
        return tc

    base = type(tc)

    def _tensors(self, x=None, mask=None):
        """Leaf tensors + what the host evaluates per call: parameter graphs the plan does not
        model (the reference's own TorchParameter) and the outputs of input layers without a
        kernel (the reference's own layer module, gathered as LayerAddressBook.lookup does,
        circuits.py:57-71; under an integration mask through IntegrateQuery._layer_fn)."""
        from cirkit.backend.torch.layers.input import TorchInputLayer
        from cirkit.backend.torch.queries import IntegrateQuery

        ext = {}
        for k, p in lowered.externals.items():
            if not isinstance(p, TorchInputLayer):
                ext[k] = p()
                continue
            if p.num_variables:
                if x is None:
                    raise ValueError(f"Expected some input 'x', as the circuit has scope '{self._scope}'")
                if x.ndim != 2:
                    raise ValueError(
                        "The input to the circuit should have shape (B, D), "
                        "where B is the batch size and D is the number of variables "
                        "the circuit is defined on")
                args = (x.to(p.scope_idx.device)[..., p.scope_idx].permute(1, 0, 2),)
            else:
                args = (1 if x is None else x.shape[0],)
            if mask is not None:
                m = mask if mask.ndim == 2 else mask.unsqueeze(0)
                y = IntegrateQuery._layer_fn(p, *args, integrate_vars_mask=m.to(args[0].device) if p.num_variables else m)
            else:
                y = p(*args)
            ext[k] = y
        return lowered.leaves, ext

    class B200Circuit(base):  # type: ignore[misc, valid-type]
        """`TorchCircuit` whose layer loop runs as sm_100a kernels (cirkit_b200)."""

        def forward(self, x=None):  # circuits.py:242-261
            if self._scope and x is None:
                raise ValueError(
                    f"Expected some input 'x', as the circuit has scope '{self._scope}'"
                )
            leaves, ext = _tensors(self, x)
            y = runtime.evaluate(x, leaves, ext)  # (B, O, K)
            if not self._scope:
                y = y.squeeze(dim=0)
            return y

        def evaluate(self, x=None, module_fn=None):  # graph/modules.py:303-335
            # `IntegrateQuery(circuit)(x, integrate_vars=...)` of the reference package arrives
            # here: its per-layer callback is recognised and becomes the runtime's masked
            # evaluation (queries.py:112-143).  Any other callback is arbitrary Python per layer
            # and runs on the reference's own executor, on the circuit's device.
            mask = _integrate_mask_of(module_fn)
            sampling = _sampling_args_of(module_fn) if x is None else None
            if sampling is not None and not lowered.externals:
                # `SamplingQuery(circuit)(num_samples)` of the reference package: ancestral
                # sampling on the device; the caller expects (O, K, N, D) and takes [0, 0]
                # (queries.py:242-246)
                samples, mixtures = runtime.sample(sampling[0], lowered.leaves)
                sampling[1].extend(mixtures)
                return samples.unsqueeze(0).unsqueeze(0)
            if module_fn is not None and mask is None:
                return super().evaluate(x, module_fn)
            leaves, ext = _tensors(self, x, mask)
            return runtime.evaluate(x, leaves, ext, integrate_mask=mask).transpose(0, 1)

        def integrate_query(self, x, mask):
            leaves, ext = _tensors(self, x, mask)
            return runtime.evaluate(x, leaves, ext, integrate_mask=mask)

        def sample_query(self, num_samples, *, seed=None):
            if lowered.externals:
                raise TypeError("Sampling needs every layer and parameter on the CUDA runtime")
            return runtime.sample(num_samples, lowered.leaves, seed=seed)

    B200Circuit.__name__ = f"B200{base.__name__}"
    tc.__class__ = B200Circuit
    tc._b200_runtime = runtime
    tc._b200_lowered = lowered
    return tc


def register_backend() -> None:
    """Teach the reference's pipeline the backend name ``"b200"``.

    `PipelineContext(backend="b200", semiring=..., fold=..., optimize=...)` then compiles with
    the reference's own torch front-end (rules, optimiser, folding: SURVEY §2 rows 11-15) and
    hands every compiled circuit to :func:`accelerate`.
    """
    import cirkit.backend.compiler as BC
    import cirkit.pipeline as P
    from cirkit.backend.torch.compiler import TorchCompiler

    if "b200" in BC.SUPPORTED_BACKENDS:
        return

    class B200Compiler(TorchCompiler):
        def compile_pipeline(self, sc):  # backend/torch/compiler.py:140-151
            return accelerate(super().compile_pipeline(sc))

    BC.SUPPORTED_BACKENDS.append("b200")
    prev = P.retrieve_compiler

    def retrieve_compiler(backend: str, **backend_kwargs):
        if backend == "b200":
            return B200Compiler(**backend_kwargs)
        return prev(backend, **backend_kwargs)

    P.retrieve_compiler = retrieve_compiler
