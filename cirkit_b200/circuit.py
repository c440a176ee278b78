"""Stand-alone compiled-circuit module backed by the CUDA runtime.

`B200Circuit` mirrors the surface of the reference's `TorchCircuit`
(cirkit/backend/torch/circuits.py:122-278) for a circuit given as a :class:`CircuitPlan`:
`cc(x)` with `x: (B, D)` returns `(B, O, K)` log-values (or `(O, K)` when the scope is empty),
raises `ValueError` when `x` is missing or not 2-D, is differentiable w.r.t. `cc.parameters()`,
and exposes `scope`, `num_variables`, `layers`, `reset_parameters()`, `state_dict()`.
When the reference package is installed, `cirkit_b200.accelerate(tc)` is the drop-in route
instead (it keeps the reference object and its parameters); this class is what runs from a
stored plan, e.g. on a box without the reference.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from .plan import CircuitPlan, StepSpec, init_leaf_
from .runtime import PlanRuntime


@dataclass(frozen=True)
class LayerInfo:
    """Read-only view of one folded layer (the attributes `TorchLayer` exposes,
    cirkit/backend/torch/layers/base.py:13-119)."""

    kind: str
    num_folds: int
    arity: int
    num_input_units: int
    num_output_units: int
    config: dict

    @classmethod
    def from_step(cls, s: StepSpec) -> "LayerInfo":
        return cls(s.kind, s.num_folds, s.arity, s.num_input_units, s.num_output_units, dict(s.config))


class B200Circuit(nn.Module):
    def __init__(self, plan: CircuitPlan, *, seed: int | None = None, fuse_tables: bool = True,
                 fuse_table_inputs: bool = True):
        super().__init__()
        self.plan = plan
        self.runtime = PlanRuntime(plan, fuse_tables=fuse_tables, fuse_table_inputs=fuse_table_inputs)
        self.leaves = nn.ParameterList(
            [nn.Parameter(torch.empty(l.shape, dtype=torch.complex64 if l.dtype == "complex" else torch.float32),
                          requires_grad=l.requires_grad)
             for l in plan.leaves]
        )
        if seed is not None:
            from .plan import seeded_leaves

            with torch.no_grad():
                for p, v in zip(self.leaves, seeded_leaves(plan, seed)):
                    p.copy_(v)
        else:
            self.reset_parameters()

    @classmethod
    def from_torch(cls, tc, *, share_parameters: bool = True, fuse_tables: bool = True) -> "B200Circuit":
        """A stand-alone circuit from one compiled by the reference (`cirkit.pipeline.compile`):
        the address book and every layer's parameters are lowered to a plan
        (`adapter.plan_from_torch`), after which the reference object is no longer needed -- the
        plan can be saved and the circuit evaluated where `cirkit` is not installed.

        With `share_parameters` the module holds the reference circuit's OWN `nn.Parameter` leaves
        (`TorchTensorParameter._ptensor`, parameters/nodes.py:193-201), so both objects train the
        same storage; otherwise it gets copies.  Parameter graphs the plan cannot express
        (kron / einsum nodes of product circuits) need the reference at run time: use
        `cirkit_b200.accelerate(tc)` for those."""
        from .adapter import plan_from_torch

        low = plan_from_torch(tc, allow_external_params=False)
        self = cls(low.plan, fuse_tables=fuse_tables)
        if share_parameters:
            self.leaves = nn.ParameterList(low.leaves)
        else:
            with torch.no_grad():
                for p, v in zip(self.leaves, low.leaves):
                    p.copy_(v)
        return self

    # -- TorchCircuit surface -------------------------------------------------------------
    @property
    def scope(self) -> tuple[int, ...]:
        return self.plan.scope

    @property
    def num_variables(self) -> int:
        return len(self.plan.scope)

    @property
    def layers(self) -> list[LayerInfo]:
        return [LayerInfo.from_step(s) for s in self.plan.steps]

    @property
    def is_folded(self) -> bool:
        return True

    def reset_parameters(self) -> None:
        for t, spec in zip(self.leaves, self.plan.leaves):
            # complex leaves: real and imaginary parts drawn independently
            init_leaf_(torch.view_as_real(t.data) if t.is_complex() else t.data, spec)
        self.runtime.invalidate_parameter_cache()

    def forward(self, x: Tensor | None = None) -> Tensor:
        if self.plan.scope and x is None:
            raise ValueError(f"Expected some input 'x', as the circuit has scope '{self.plan.scope}'")
        y = self.runtime.evaluate(x, self._leaf_list())  # (B, O, K)
        if not self.plan.scope:
            y = y.squeeze(dim=0)
        return y

    def _leaf_list(self) -> list:
        # (iterating an nn.ParameterList costs ~3 us per entry: the list is rebuilt only when the
        # module's parameters were re-assigned)
        cached = self.__dict__.get("_leaves_cache")
        if cached is None or len(cached) != len(self.leaves) or any(
                a is not b for a, b in zip(cached, self.leaves._parameters.values())):
            cached = list(self.leaves)
            self.__dict__["_leaves_cache"] = cached
        return cached

    def integrate_query(self, x: Tensor, mask: Tensor) -> Tensor:
        return self.runtime.evaluate(x, list(self.leaves), integrate_mask=mask)

    def sample_query(self, num_samples: int, *, seed: int | None = None) -> tuple[Tensor, list[Tensor]]:
        """(samples (N, D), mixture_samples): see `PlanRuntime.sample`."""
        return self.runtime.sample(num_samples, list(self.leaves), seed=seed)
