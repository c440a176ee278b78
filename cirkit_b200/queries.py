"""Integration (marginal) and sampling queries, mirroring `cirkit.backend.torch.queries.IntegrateQuery`
(cirkit/backend/torch/queries.py:19-184): same constructor, call signature, mask formats and
error messages; the per-layer `torch.where(mask, layer.integrate(), output)` of `_layer_fn`
(:112-143) is folded into the input-layer kernels (the mask selects, per sample and variable,
between the layer's log-density and what integrating that variable yields)."""

from __future__ import annotations

from collections.abc import Sequence

import torch
from torch import Tensor


class IntegrateQuery:
    def __init__(self, circuit) -> None:
        props = getattr(circuit, "properties", None)
        if props is not None and (not props.smooth or not props.decomposable):
            raise ValueError(
                f"The circuit to integrate must be smooth and decomposable, but found {props}"
            )
        if not hasattr(circuit, "integrate_query"):
            raise TypeError("IntegrateQuery needs a circuit evaluated by cirkit_b200")
        self._circuit = circuit

    def __call__(self, x: Tensor, *, integrate_vars) -> Tensor:
        scope = tuple(self._circuit.scope)
        num_vars = max(scope) + 1
        if isinstance(integrate_vars, Tensor):
            if integrate_vars.dtype != torch.bool:
                raise ValueError(f"Expected dtype of tensor to be torch.bool, got {integrate_vars.dtype}")
            if integrate_vars.ndim == 1:
                integrate_vars = integrate_vars.unsqueeze(0)
            if integrate_vars.shape[1] != num_vars:
                raise ValueError(
                    f"Circuit scope has {num_vars} variables but integrate_vars "
                    f"was defined over {integrate_vars.shape[1]} != {num_vars} variables"
                )
            mask = integrate_vars
        else:
            mask = self.scopes_to_mask(self._circuit, integrate_vars)
        if mask.shape[0] not in (1, x.shape[0]):
            raise ValueError(
                "The number of scopes to integrate over must "
                "either match the batch size of x, or be 1 if you "
                "want to broadcast. Found #inputs = "
                f"{x.shape[0]} != {mask.shape[0]} = len(integrate_vars)"
            )
        return self._circuit.integrate_query(x, mask)  # (B, O, K)

    @staticmethod
    def scopes_to_mask(circuit, batch_integrate_vars) -> Tensor:
        scope = set(circuit.scope)
        if not isinstance(batch_integrate_vars, Sequence) or (
            batch_integrate_vars and isinstance(next(iter(batch_integrate_vars)), int)
        ):
            batch_integrate_vars = [batch_integrate_vars]
        num_rvs = max(scope) + 1
        mask = torch.zeros((len(batch_integrate_vars), num_rvs), dtype=torch.bool)
        for i, idxs in enumerate(batch_integrate_vars):
            idxs = list(idxs)
            invalid = [v for v in idxs if v not in scope]
            if invalid:
                raise ValueError(
                    "The variables to marginalize must be a subset of "
                    "the circuit scope. Invalid variables "
                    f"not in scope: {invalid} "
                )
            if idxs:
                mask[i, idxs] = True
        return mask


class SamplingQuery:
    """Mirror of `cirkit.backend.torch.queries.SamplingQuery` (queries.py:187-275): same
    constructor check, call signature and errors.  `query(num_samples)` returns
    `(samples, mixture_samples)` with samples of shape (num_samples, D) drawn from the joint
    distribution of the circuit by ancestral sampling on the device
    (`csrc/sampling_kernels.cu`); see `PlanRuntime.sample` for the form of `mixture_samples`."""

    def __init__(self, circuit) -> None:
        props = getattr(circuit, "properties", None)
        if props is not None and (not props.smooth or not props.decomposable):
            raise ValueError(
                f"The circuit to sample from must be smooth and decomposable, but found {props}"
            )
        if not hasattr(circuit, "sample_query"):
            raise TypeError("SamplingQuery needs a circuit evaluated by cirkit_b200")
        self._circuit = circuit

    def __call__(self, num_samples: int = 1, *, seed: int | None = None):
        if num_samples <= 0:
            raise ValueError("The number of samples must be a positive number")
        return self._circuit.sample_query(num_samples, seed=seed)
