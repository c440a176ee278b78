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
        return [torch.from_numpy(self.z[f"leaf_{i}"]).to(dtype) for i in range(len(self.plan.leaves))]

    def x(self) -> torch.Tensor | None:
        return torch.from_numpy(self.z["x"]) if "x" in self.z else None

    def y(self) -> torch.Tensor:
        return torch.from_numpy(self.z["y"])

    def grads(self) -> list[torch.Tensor]:
        return [torch.from_numpy(self.z[f"grad_{i}"]) for i in range(len(self.plan.leaves))]

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


def grad_tolerance(g_ref: torch.Tensor) -> float:
    """per-tensor bound max(5e-7, 1e-4 * max|g_ref64|) for fp32 gradients."""
    return max(5e-7, 1e-4 * float(g_ref.abs().max()))
