"""CPU restatement of the reference's folded-circuit evaluator.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows.  The arithmetic is issued as the same
sequence of stock PyTorch ops the reference issues (`torch.cat` + fancy index for gathers,
`amax/clamp/sub/exp/einsum/log/add` for the LSE semiring), so that (a) results agree with the
reference to the last bit on the same dtype, and (b) timing this module on the host cores is a
faithful stand-in for "the reference's own PyTorch-CPU path" where the reference package itself
is not installed (GPU box).  Gradients come from autograd, exactly as in the reference
(no custom backward on the real-valued path).

Parity status: pinned against the live reference in `tests/test_oracle_vs_reference.py` and
against `tests/golden/*.npz` -- for the 'lse-sum' semiring and, ahead of the CUDA kernels for it,
for 'complex-lse-sum' (semiring.py:410-476; fixtures of kind "complex").
"""

from __future__ import annotations

import functools
import math

import numpy as np
import torch
from torch import Tensor, nn

from cirkit_b200.plan import CircuitPlan, ParamSpec, StepSpec, init_leaf_


# --------------------------------------------------------------------------- semiring
def lse_apply_reduce(func, *xs: Tensor, dim: int = -1) -> Tensor:
    """`LSESumSemiring.apply_reduce`, cirkit/backend/torch/semiring.py:382-408 (keepdim=True).

    max over `dim` clamped to the finite range (so an all -inf row stays finite-safe), exp of the
    shifted inputs, the linear-space contraction `func`, log, and the shifts added back.
    """
    max_xs = [
        torch.clamp(
            torch.amax(xi, dim=dim, keepdim=True),
            min=torch.finfo(xi.dtype).min,
            max=torch.finfo(xi.dtype).max,
        )
        for xi in xs
    ]
    exp_xs = [torch.exp(xi - mi) for xi, mi in zip(xs, max_xs)]
    y = func(*exp_xs)
    return torch.log(y) + functools.reduce(torch.add, max_xs)


class _ComplexSafeLog(torch.autograd.Function):
    """`csafelog`, cirkit/backend/torch/utils.py:32-50: complex log whose backward replaces the
    NaN / inf of d log(z) = g / conj(z) at z = 0 by 0 / the largest finite values."""

    @staticmethod
    def forward(ctx, z: Tensor) -> Tensor:
        ctx.save_for_backward(z)
        return torch.log(z)

    @staticmethod
    def backward(ctx, g: Tensor) -> Tensor:
        (z,) = ctx.saved_tensors
        return torch.nan_to_num(g / z.conj())


def complex_cast(t: Tensor) -> Tensor:
    """`ComplexLSESumSemiring.cast`, semiring.py:414-421."""
    if t.is_complex():
        return t
    if t.is_floating_point():
        return t.to(t.dtype.to_complex())
    return t.to(torch.get_default_dtype().to_complex())


def complex_lse_apply_reduce(func, *xs: Tensor, dim: int = -1) -> Tensor:
    """`ComplexLSESumSemiring.apply_reduce`, semiring.py:440-476 (keepdim=True): as the real
    block, with the shift taken over the REAL parts, complex exp, and the safe complex log."""
    max_xs = [
        torch.clamp(
            torch.amax(xi.real, dim=dim, keepdim=True),
            min=torch.finfo(xi.real.dtype).min,
            max=torch.finfo(xi.real.dtype).max,
        )
        for xi in xs
    ]
    exp_xs = [torch.exp(xi - mi) for xi, mi in zip(xs, max_xs)]
    y = func(*exp_xs)
    return _ComplexSafeLog.apply(y) + functools.reduce(torch.add, max_xs)


# --------------------------------------------------------------------------- parameters
def apply_param_op(t: Tensor, op: str, attrs: dict) -> Tensor:
    """One re-parameterisation node, cirkit/backend/torch/parameters/nodes.py (dim is given
    relative to the un-folded shape and shifted by one at run time, nodes.py:751,772,783)."""
    if op == "softmax":
        return torch.softmax(t, dim=attrs.get("dim", t.ndim - 2) + 1)  # nodes.py:764-772
    if op == "log_softmax":
        return torch.log_softmax(t, dim=attrs.get("dim", t.ndim - 2) + 1)  # nodes.py:775-783
    if op == "scaled_sigmoid":
        vmin, vmax = attrs["vmin"], attrs["vmax"]
        return torch.sigmoid(t) * (vmax - vmin) + vmin  # nodes.py:682-699
    if op == "sigmoid":
        return torch.sigmoid(t)
    if op == "exp":
        return torch.exp(t)
    if op == "log":
        return torch.log(t)
    if op == "square":
        return torch.square(t)
    if op == "softplus":
        return torch.nn.functional.softplus(t)
    if op == "clamp":
        return torch.clamp(t, min=attrs.get("vmin"), max=attrs.get("vmax"))
    if op == "conj":
        return torch.conj(t)  # TorchConjugateParameter, nodes.py:742-746
    if op == "mixing":
        # TorchMixingWeightParameter.forward, nodes.py:857-862: (F, K, H) -> (F, K, H*K)
        d = torch.vmap(torch.vmap(torch.diag, in_dims=1))(t)
        return d.permute(0, 2, 1, 3).flatten(start_dim=2)
    raise ValueError(f"unknown parameter op {op!r}")


# --------------------------------------------------------------------------- the circuit
class OracleCircuit(nn.Module):
    """Evaluates a :class:`CircuitPlan` the way `TorchCircuit` evaluates its address book."""

    def __init__(self, plan: CircuitPlan, dtype: torch.dtype = torch.float32):
        super().__init__()
        if plan.semiring not in ("lse-sum", "complex-lse-sum"):
            raise NotImplementedError(
                "the oracle restates the 'lse-sum' and 'complex-lse-sum' semiring paths")
        plan.validate()
        self.plan = plan
        # semiring.py:382-408 (real) / :440-476 (complex: activations are complex logs)
        self.is_complex = plan.semiring == "complex-lse-sum"
        self._reduce = complex_lse_apply_reduce if self.is_complex else lse_apply_reduce
        self.leaves = nn.ParameterList(
            [nn.Parameter(torch.empty(l.shape, dtype=dtype.to_complex() if l.dtype == "complex" else dtype),
                          requires_grad=l.requires_grad)
             for l in plan.leaves]
        )
        self.reset_parameters()
        # Address-book view of the gathers, rebuilt the way the reference builds it
        # (graph/folding.py:202-243): unique producers in first-seen order, cumulative offsets.
        self._sources: list[list[int]] = []
        self._idx: list[Tensor | tuple | None] = []
        for s in plan.steps:
            if s.is_input:
                self._sources.append([])
                self._idx.append(None)
                continue
            pairs = list(zip(s.in_step.ravel().tolist(), s.in_fold.ravel().tolist()))
            sources = list(dict.fromkeys(p for p, _ in pairs))
            sizes = [plan.steps[p].num_folds for p in sources]
            cum = dict(zip(sources, np.cumsum([0] + sizes[:-1]).tolist()))
            flat = [cum[p] + f for p, f in pairs]
            total = sum(sizes)
            idx: Tensor | tuple
            if flat == list(range(total)) and s.num_folds == 1 and s.arity == total:
                idx = (None,)  # unsqueeze(0) shortcut, folding.py:235-238
            elif flat == list(range(total)) and s.num_folds == total and s.arity == 1:
                idx = (slice(None), None)  # unsqueeze(1) shortcut, folding.py:239-241
            else:
                idx = torch.tensor(flat, dtype=torch.int64).view(s.num_folds, s.arity)
            self._sources.append(sources)
            self._idx.append(idx)
        self._out_sources = list(dict.fromkeys(plan.out_step.tolist()))
        sizes = [plan.steps[p].num_folds for p in self._out_sources]
        cum = dict(zip(self._out_sources, np.cumsum([0] + sizes[:-1]).tolist()))
        self._out_idx = torch.tensor(
            [cum[int(p)] + int(f) for p, f in zip(plan.out_step, plan.out_fold)], dtype=torch.int64
        )

    def reset_parameters(self) -> None:
        for t, spec in zip(self.leaves, self.plan.leaves):
            # complex leaves: real and imaginary parts drawn independently
            init_leaf_(torch.view_as_real(t.data) if t.is_complex() else t.data, spec)

    def _from_lse(self, y: Tensor) -> Tensor:
        """`semiring.map_from(y, LSESumSemiring)`: identity / cast (semiring.py:511-514)."""
        return complex_cast(y) if self.is_complex else y

    def _from_linear(self, v: Tensor) -> Tensor:
        """`semiring.map_from(v, SumProductSemiring)`: log (semiring.py:497-499) / the safe
        complex log of the cast value (:506-508)."""
        return _ComplexSafeLog.apply(complex_cast(v)) if self.is_complex else torch.log(v)

    def _operand(self, w: Tensor) -> Tensor:
        """`SemiringImpl.einsum` casts the weight operands, semiring.py:181-192."""
        return complex_cast(w) if self.is_complex else w

    # ---------------------------------------------------------------- parameters
    def param(self, p: ParamSpec) -> Tensor:
        """`TorchParameter.forward`, parameters/parameter.py:180-188, for a leaf->ops chain."""
        if p.leaf < 0:
            # graph the plan does not model (kron/matmul nodes...): the caller evaluated the
            # reference's own TorchParameter and passes the result in
            return self._externals[id(p)]
        t: Tensor = self.leaves[p.leaf]
        if p.fold_idx is not None:
            t = t[torch.as_tensor(p.fold_idx, dtype=torch.int64)]  # nodes.py:277-279
        for op, attrs in p.ops:
            if op == "matmul":
                # TorchMatMulParameter.forward, parameters/nodes.py:802-805 (SumCollapse rule,
                # optimization/layers.py:30-47): (F, d1, d2) @ (F, d2, d3)
                t = torch.matmul(t, self._param_chain(attrs["rhs"]))
            else:
                t = apply_param_op(t, op, attrs)
        return t

    def _param_chain(self, spec: dict) -> Tensor:
        """The right operand of a matmul node: another leaf -> [fold slice] -> op chain."""
        t: Tensor = self.leaves[spec["leaf"]]
        if spec.get("fold_idx") is not None:
            t = t[torch.as_tensor(spec["fold_idx"], dtype=torch.int64)]
        for op, attrs in spec["ops"]:
            if op == "matmul":
                t = torch.matmul(t, self._param_chain(attrs["rhs"]))
            else:
                t = apply_param_op(t, op, attrs)
        return t

    # ---------------------------------------------------------------- layers
    def _input_layer(self, s: StepSpec, x: Tensor | None, batch: int) -> Tensor:
        F, K = s.num_folds, s.num_output_units
        if s.kind == "constant":
            # TorchConstantValueLayer.forward, layers/input.py:739-743
            v = self.param(s.params["value"])
            v = v.unsqueeze(1).expand(F, batch, K)
            return self._from_lse(v) if s.config.get("log_space", False) else self._from_linear(v)
        assert x is not None
        scope_idx = torch.as_tensor(s.scope_idx, dtype=torch.int64).view(F, 1)
        xs = x[..., scope_idx].permute(1, 0, 2)  # circuits.py:66 -> (F, B, 1)
        if s.kind == "categorical":
            # TorchCategoricalLayer.log_unnormalized_likelihood, layers/input.py:399-412
            if xs.is_floating_point():
                xs = xs.long()
            xs = xs.squeeze(2)
            if "probs" in s.params:
                logits = torch.log(self.param(s.params["probs"]))
            else:
                logits = self.param(s.params["logits"])
            idx_fold = torch.arange(F)
            return self._from_lse(logits[idx_fold[:, None], :, xs])
        if s.kind == "embedding":
            # TorchEmbeddingLayer.forward, layers/input.py:258-266 (+ log morphism semiring.py:499)
            if xs.is_floating_point():
                xs = xs.long()
            xs = xs.squeeze(2)
            w = self.param(s.params["weight"])
            idx_fold = torch.arange(F)
            return self._from_linear(w[idx_fold[:, None], :, xs])
        if s.kind == "gaussian":
            # TorchGaussianLayer.log_unnormalized_likelihood, layers/input.py:661-670
            mean = self.param(s.params["mean"]).unsqueeze(1)
            stddev = self.param(s.params["stddev"]).unsqueeze(1)
            lp = torch.distributions.Normal(loc=mean, scale=stddev).log_prob(xs)
            if "log_partition" in s.params:
                lp = lp + self.param(s.params["log_partition"]).unsqueeze(1)
            return self._from_lse(lp)
        raise ValueError(s.kind)

    def _integrate(self, s: StepSpec) -> Tensor:
        """`TorchInputLayer.integrate` for the layers on the path: layers/input.py:280-282,
        :414-421 (Categorical), :672-678 (Gaussian).  Shape (F, 1, K)."""
        F, K = s.num_folds, s.num_output_units
        ref = next(iter(s.params.values()))
        dtype = self.leaves[ref.leaf].dtype if ref.leaf >= 0 else self.leaves[0].dtype
        if s.kind == "categorical":
            if "probs" in s.params:
                return self._from_lse(torch.zeros(F, 1, K, dtype=dtype))
            return self._from_lse(torch.logsumexp(self.param(s.params["logits"]), dim=2).unsqueeze(1))
        if s.kind == "gaussian":
            if "log_partition" in s.params:
                return self._from_lse(self.param(s.params["log_partition"]).unsqueeze(1))
            return self._from_lse(torch.zeros(F, 1, K, dtype=dtype))
        raise TypeError(f"integration is not supported for {s.kind} layers")

    def _inner_layer(self, s: StepSpec, x: Tensor) -> Tensor:
        """x: (F, H, B, Ki) -> (F, B, Ko)."""
        if s.kind == "hadamard":
            return x.sum(dim=1)  # TorchHadamardLayer.forward, layers/inner.py:126-127
        if s.kind == "kronecker":
            # TorchKroneckerLayer.forward, layers/inner.py:178-187
            y0 = x[:, 0]
            for i in range(1, x.shape[1]):
                y0 = torch.flatten(y0.unsqueeze(-1) + x[:, i].unsqueeze(-2), start_dim=-2)
            return y0
        if s.kind in ("sum", "mixing"):
            # TorchSumLayer.forward, layers/inner.py:266-273; for 'mixing' the weight is the
            # block-diagonal expansion of (F, K, H) mixing weights (nodes.py:857-862)
            w = self.param(s.params["weight"])
            if s.kind == "mixing":
                w = apply_param_op(w, "mixing", {})
            w = self._operand(w)
            xf = x.permute(0, 2, 1, 3).flatten(start_dim=2)
            return self._reduce(lambda e: torch.einsum("fbi,foi->fbo", e, w), xf)
        if s.kind == "cpt":
            # TorchCPTLayer.forward, layers/optimized.py:171-178
            w = self._operand(self.param(s.params["weight"]))
            u = x.sum(dim=1)
            return self._reduce(lambda e: torch.einsum("fbi,foi->fbo", e, w), u)
        if s.kind == "tensordot":
            # TorchTensorDotLayer.forward, layers/optimized.py:287-300: (F, B, Kj*Kq) -> (F, B, Kq, Kj),
            # contraction over j with weight (F, Kk, Kj), flattened back to (F, B, Kq*Kk)
            w = self._operand(self.param(s.params["weight"]))
            kq = int(s.config["kq"])
            xs = x.squeeze(dim=1)
            xs = xs.view(xs.shape[0], xs.shape[1], w.shape[2], kq).permute(0, 1, 3, 2)
            y = self._reduce(lambda e: torch.einsum("fbqj,fkj->fbqk", e, w), xs)
            return y.reshape(y.shape[0], y.shape[1], s.num_output_units)
        if s.kind == "tucker":
            # TorchTuckerLayer.forward, layers/optimized.py:89-103 (einsum spec :62-66)
            H, Ki, Ko = s.arity, s.num_input_units, s.num_output_units
            w = self._operand(self.param(s.params["weight"])).view(-1, Ko, *([Ki] * H))
            spec = (
                tuple((0, 1, i + 2) for i in range(H))
                + ((0, H + 2, *tuple(i + 2 for i in range(H))),)
                + ((0, 1, H + 2),)
            )

            def f(*es: Tensor) -> Tensor:
                args = []
                for t, sub in zip((*es, w), spec[:-1]):
                    args += [t, list(sub)]
                return torch.einsum(*args, list(spec[-1]))

            return self._reduce(f, *x.unbind(dim=1))
        raise ValueError(s.kind)

    # ---------------------------------------------------------------- executor
    def forward(
        self,
        x: Tensor | None = None,
        integrate_mask: Tensor | None = None,
        externals: dict[tuple[int, str], Tensor] | None = None,
    ) -> Tensor:
        """`TorchCircuit.forward` -> `TorchDiAcyclicGraph.evaluate` -> `LayerAddressBook.lookup`
        (circuits.py:242-278, graph/modules.py:303-335, circuits.py:30-71).  With
        ``integrate_mask`` (bool, (B or 1, D)) it follows `IntegrateQuery._layer_fn`
        (queries.py:112-143)."""
        plan = self.plan
        self._externals = {
            id(plan.steps[sid].params[name]): t for (sid, name), t in (externals or {}).items()
        }
        if plan.scope and x is None:
            raise ValueError(f"Expected some input 'x', as the circuit has scope '{plan.scope}'")
        if x is not None and x.ndim != 2:
            raise ValueError(
                "The input to the circuit should have shape (B, D), "
                "where B is the batch size and D is the number of variables "
                "the circuit is defined on"
            )
        batch = 1 if x is None else x.shape[0]
        outs: list[Tensor] = []
        for sid, s in enumerate(plan.steps):
            if s.is_input:
                y = self._input_layer(s, x, batch)
                if integrate_mask is not None and s.kind != "constant":
                    scope_idx = torch.as_tensor(s.scope_idx, dtype=torch.int64).view(-1, 1)
                    m = integrate_mask[:, scope_idx].permute(1, 0, 2)  # (F, B|1, 1)
                    if bool(torch.any(m)):
                        y = torch.where(m, self._integrate(s), y)
            else:
                src = self._sources[sid]
                t = outs[src[0]] if len(src) == 1 else torch.cat([outs[p] for p in src], dim=0)
                y = self._inner_layer(s, t[self._idx[sid]])
            outs.append(y)
        self.last_outputs = outs  # per-step (F, B, K) tensors, for layer-by-layer parity reports
        src = self._out_sources
        t = outs[src[0]] if len(src) == 1 else torch.cat([outs[p] for p in src], dim=0)
        y = t[self._out_idx].transpose(0, 1)  # (O, B, K) -> (B, O, K)
        if not plan.scope:
            y = y.squeeze(dim=0)
        return y


def make_inputs(plan: CircuitPlan, batch: int, seed: int = 0) -> Tensor:
    """Synthetic evidence for a plan: uniform categories for discrete inputs, N(0,1) otherwise
    (SURVEY §8(d) "concrete synthetic inputs")."""
    g = torch.Generator().manual_seed(seed)
    kinds = {s.kind for s in plan.steps if s.is_input}
    D = plan.num_variables
    if kinds <= {"categorical", "embedding", "constant"}:
        V = min(
            s.config.get("num_categories", s.config.get("num_states", 2))
            for s in plan.steps
            if s.kind in ("categorical", "embedding")
        )
        return torch.randint(0, V, (batch, D), generator=g, dtype=torch.int64)
    x = torch.randn(batch, D, generator=g)
    for s in plan.steps:
        if s.kind in ("categorical", "embedding"):
            V = s.config.get("num_categories", s.config.get("num_states", 2))
            cols = torch.as_tensor(s.scope_idx, dtype=torch.int64)
            x[:, cols] = torch.randint(0, V, (batch, len(cols)), generator=g).to(x.dtype)
    return x
