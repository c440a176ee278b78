"""Shared helpers for the parity tests."""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

from cirkit_b200.plan import CircuitPlan, seeded_leaves

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    """One tests/golden/*.npz fixture (written by tests/golden/make_golden.py)."""

    def __init__(self, name: str):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        self.meta = json.loads(bytes(self.z["meta"]).decode())
        self.plan = CircuitPlan.load(bytes(self.z["plan"]))

    @property
    def kind(self) -> str:
        return self.meta["kind"]

    def leaves(self, dtype=torch.float64) -> list[torch.Tensor]:
        if self.kind == "seeded":
            return [t.to(dtype) for t in seeded_leaves(self.plan, self.meta["seed"])]
        vals = [torch.from_numpy(self.z[f"leaf_{i}"]) for i in range(len(self.plan.leaves))]
        return [v.to(dtype.to_complex() if v.is_complex() else dtype) for v in vals]

    def x(self) -> torch.Tensor | None:
        return torch.from_numpy(self.z["x"]) if "x" in self.z else None

    def y(self) -> torch.Tensor:
        return torch.from_numpy(self.z["y"])

    def grads(self) -> list[torch.Tensor]:
        return [torch.from_numpy(self.z[f"grad_{i}"]) for i in range(len(self.plan.leaves))]

    def grads_mask(self) -> list[torch.Tensor] | None:
        """Gradients of -mean(IntegrateQuery output) the reference's autograd produced."""
        if "grad_mask_0" not in self.z:
            return None
        return [torch.from_numpy(self.z[f"grad_mask_{i}"]) for i in range(len(self.plan.leaves))]

    def grad_full(self, i: int) -> torch.Tensor | None:
        """Whole float64 gradient of leaf i (seeded fixtures keep it for the small leaves)."""
        return torch.from_numpy(self.z[f"gfull_{i}"]) if f"gfull_{i}" in self.z else None

    def mask(self):
        if "mask" not in self.z:
            return None, None
        return torch.from_numpy(self.z["mask"]), torch.from_numpy(self.z["y_mask"])


def golden_names(kind: str | None = None) -> list[str]:
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    if kind is None:
        return names
    out = []
    for n in names:
        z = np.load(os.path.join(GOLDEN, f"{n}.npz"))
        if json.loads(bytes(z["meta"]).decode())["kind"] == kind:
            out.append(n)
    return out


GRAD_FLOOR = 2e-6


EPS32 = 1.1920929e-07


def grad_tolerance(g_ref: torch.Tensor, gout_l1: float = 1.0, ll_max: float = 0.0, w_max: float = 1.0) -> float:
    """Per-tensor bound for fp32 gradients against the float64 reference:
    max(2e-6, 1e-4 * max|g_ref64|).

    SURVEY §8(d) derives max(5e-7, 1e-4*max|g|) from the reference's own fp32-vs-fp64 gap at
    B=256.  The floor is an *absolute* error that does not shrink with the gradient: through a
    softmax re-parameterisation d(theta) = W * (dW - sum_j W_j dW_j) cancels O(1) entries of dW
    (each a sum over the batch of terms of size 1/B, so |dW| ~ 1 regardless of B) down to
    1e-3..1e-7 near the root, leaving ~4 ulp(1) = 5e-7 of fp32 rounding; the reference's fp32
    run shows the same (SURVEY §7 "gradient parity is ill-conditioned").  2e-6 = 4x that.
    `gout_l1` is the L1 norm of d(loss)/d(output) (1 for a mean log-likelihood): |dW| and with
    it the floor scale linearly with it.

    `ll_max` / `w_max` add the bound that follows from the fp32 storage of the activations
    themselves: a layer input y carries a rounding error of eps32*|y| (2.4e-4 at |y| = 2000), so
    e = exp(u - m) carries 4*eps32*|y| relative, dW the same times gout_l1, and
    d(theta) = W*(dW - sum W dW) at most 8*eps32*|y|*max(W)*gout_l1.  Near the root of a
    randomly initialised circuit all units of a layer compute (almost) the same value, the true
    gradient is ~1e-8 and this bound is what any fp32 evaluation can promise.
    """
    ulp_bound = 8.0 * EPS32 * ll_max * w_max
    return max(GRAD_FLOOR, ulp_bound, 1e-4 * float(g_ref.abs().max())) * max(gout_l1, 1.0)
