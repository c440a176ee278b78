"""CPU checkers for SamplingQuery.  TEST INFRASTRUCTURE ONLY (imported by tests/ alone).

Two independent restatements:

* `reference_sample` -- the reference's own BOTTOM-UP algorithm (every unit of every layer draws
  num_samples values; sum units gather), issued as the same torch calls in the same order as
  `SamplingQuery.__call__` / `_layer_fn` / `_pad_samples` (cirkit/backend/torch/queries.py:219-275)
  and the layers' `sample` methods (layers/input.py:423-434, :680-685; layers/inner.py:129-133,
  :189-197, :275-300; layers/optimized.py:180-202), so that under the same `torch.manual_seed` it
  returns the reference's samples bit for bit (pinned in tests/test_oracle_vs_reference.py).
  Memory is O(F K N D): small circuits only.

* `ancestral_sample` -- a numpy restatement of the TOP-DOWN sampler of
  `cirkit_b200/csrc/sampling_kernels.cu` with the same Philox4x32-10 stream and the same
  inverse-CDF rule, to check the kernel sample by sample.

Both are compared with the exact joint distribution (enumeration through the pinned forward
oracle) in tests/test_gpu_sampling.py, as the reference's own test does
(tests/backend/torch/test_queries/test_sampling.py:18-53).
"""

from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from cirkit_b200.plan import CircuitPlan

from .reference_eval import OracleCircuit


# --------------------------------------------------------------------------- bottom-up (reference)
def _check_weight(w: Tensor, exc) -> None:
    # layers/inner.py:277-281, layers/optimized.py:182-188
    if torch.any(w < 0.0):
        raise exc("Sampling only works with positive weights")
    if not torch.allclose(torch.sum(w, dim=-1), torch.ones(1, dtype=w.dtype)):
        raise exc("Sampling only works with a normalized parametrization")


def _gather_mixture(x: Tensor, weight: Tensor) -> tuple[Tensor, Tensor]:
    """x (F, Kred, N, D), weight (F, Ko, Kred) -> ((F, Ko, N, D), (F, Ko, N));
    layers/inner.py:283-300 == layers/optimized.py:190-202."""
    n, d = x.shape[2], x.shape[3]
    dist = torch.distributions.Categorical(probs=weight)
    mixing = dist.sample((n,)).permute(1, 2, 0)  # (N, F, Ko) -> (F, Ko, N)
    idx = mixing.unsqueeze(-1).expand(-1, -1, -1, d)
    return torch.gather(x, dim=1, index=idx), mixing


def reference_sample(oc: OracleCircuit, num_samples: int) -> tuple[Tensor, list[Tensor]]:
    """(samples (N, D), mixture_samples) exactly as `SamplingQuery(tc)(num_samples)` returns them,
    drawing from torch's global generator in the reference's order (one layer after the other in
    the plan's = the reference's topological order)."""
    if num_samples <= 0:
        raise ValueError("The number of samples must be a positive number")
    plan: CircuitPlan = oc.plan
    D = len(plan.scope)
    outs: list[Tensor] = []
    mixtures: list[Tensor] = []
    for s in plan.steps:
        if s.is_input:
            if s.kind == "categorical":
                # layers/input.py:423-434
                name = "probs" if "probs" in s.params else "logits"
                p = oc.param(s.params[name])
                logits = torch.log(p) if name == "probs" else p
                smp = torch.distributions.Categorical(logits=logits).sample((num_samples,)).permute(1, 2, 0)
            elif s.kind == "gaussian":
                # layers/input.py:680-685
                dist = torch.distributions.Normal(loc=oc.param(s.params["mean"]), scale=oc.param(s.params["stddev"]))
                smp = dist.sample((num_samples,)).permute(1, 2, 0)
            else:
                raise TypeError(f"Sampling is not supported for layers of type {s.kind}")
            # queries.py:258-275 (_pad_samples)
            padded = torch.zeros((*smp.shape, D), dtype=smp.dtype)
            fold_idx = torch.arange(smp.shape[0])
            padded[fold_idx, :, :, torch.as_tensor(s.scope_idx, dtype=torch.int64)] = smp
            mixtures.append(padded)
            outs.append(padded)
            continue
        # LayerAddressBook.lookup, circuits.py:57-71: (F, H, K, N, D)
        x = torch.stack([
            torch.stack([outs[int(s.in_step[f, h])][int(s.in_fold[f, h])] for h in range(s.arity)])
            for f in range(s.num_folds)
        ])
        if s.kind == "hadamard":
            y = torch.sum(x, dim=1)  # layers/inner.py:129-133
        elif s.kind == "kronecker":
            y = x[:, 0]  # layers/inner.py:189-197
            for i in range(1, x.shape[1]):
                y = torch.flatten(y.unsqueeze(2) + x[:, i].unsqueeze(1), start_dim=1, end_dim=2)
        elif s.kind == "sum":
            w = oc.param(s.params["weight"])
            _check_weight(w, TypeError)
            y, mix = _gather_mixture(x.flatten(1, 2), w)
            mixtures.append(mix)
        elif s.kind == "mixing":
            # a sum layer whose (F, K, H) weights the reference expands to the block-diagonal
            # (F, K, H*K) matrix (parameters/nodes.py:847-862) before TorchSumLayer.sample
            w3 = oc.param(s.params["weight"])
            w = torch.vmap(torch.vmap(torch.diag, in_dims=1))(w3).permute(0, 2, 1, 3).flatten(start_dim=2)
            _check_weight(w, TypeError)
            y, mix = _gather_mixture(x.flatten(1, 2), w)
            mixtures.append(mix)
        elif s.kind == "cpt":
            w = oc.param(s.params["weight"])
            _check_weight(w, ValueError)
            y, mix = _gather_mixture(torch.sum(x, dim=1), w)
            mixtures.append(mix)
        else:
            raise TypeError(f"Sampling is not supported for layers of type {s.kind}")
        outs.append(y)
    root = outs[int(plan.out_step[0])][int(plan.out_fold[0])]  # (K, N, D)
    return root[0], mixtures  # queries.py:242-246: samples[:, 0, 0]


# --------------------------------------------------------------------------- top-down (kernel)
def philox4x32_10(c: np.ndarray, k0: int, k1: int) -> np.ndarray:
    """c: (..., 4) uint32 counters -> (..., 4) uint32 random words (Salmon et al., SC'11)."""
    c = c.astype(np.uint64)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m32
        n1 = p1 & m32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m32
        n3 = p0 & m32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _u01(r: np.ndarray) -> np.ndarray:
    return (r >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def _draw(cdf_rows: np.ndarray, u: np.ndarray) -> np.ndarray:
    """First j with cdf[j] > u * cdf[-1], never an entry of probability zero (device `draw`)."""
    t = (u * cdf_rows[:, -1]).astype(np.float32)
    gt = cdf_rows > t[:, None]
    j = np.where(gt.any(axis=1), gt.argmax(axis=1), cdf_rows.shape[1] - 1)
    for r in np.nonzero(~gt.any(axis=1))[0]:  # t rounded up to the total: step off a flat tail
        while j[r] > 0 and cdf_rows[r, j[r]] == cdf_rows[r, j[r] - 1]:
            j[r] -= 1
    return j.astype(np.int64)


def effective_cdfs(oc: OracleCircuit) -> dict[int, np.ndarray]:
    """fp32 row-wise CDFs per step, summed left to right like the device kernel."""
    out = {}
    with torch.no_grad():
        for sid, s in enumerate(oc.plan.steps):
            if s.kind == "categorical":
                name = "probs" if "probs" in s.params else "logits"
                p = oc.param(s.params[name]).float()
                logp = torch.log(p) if name == "probs" else p
                out[sid] = np.cumsum(torch.exp(logp).numpy().astype(np.float32), axis=-1, dtype=np.float32)
            elif s.kind in ("sum", "cpt", "mixing", "tucker"):
                w = oc.param(s.params["weight"]).float().numpy().astype(np.float32)
                out[sid] = np.cumsum(w, axis=-1, dtype=np.float32)
    return out


def ancestral_sample(oc: OracleCircuit, num_samples: int, seed: int, cdfs: dict | None = None,
                     sample_base: int = 0) -> tuple[np.ndarray, dict[int, np.ndarray]]:
    """Top-down sampler, same rules and random stream as csrc/sampling_kernels.cu.  Returns
    (x (N, D), {sum step id: (F, N) mixture draws, -1 off the path})."""
    plan = oc.plan
    N = num_samples
    cdfs = effective_cdfs(oc) if cdfs is None else cdfs
    row0 = np.concatenate([[0], np.cumsum([s.num_folds for s in plan.steps])])
    sel = [np.full((s.num_folds, N), -1, dtype=np.int64) for s in plan.steps]
    mixes: dict[int, np.ndarray] = {}
    sel[int(plan.out_step[0])][int(plan.out_fold[0])] = 0
    is_float = any(s.kind == "gaussian" for s in plan.steps)
    x = np.zeros((N, plan.num_variables), dtype=np.float32 if is_float else np.int64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    n_idx = np.arange(N, dtype=np.uint64) + np.uint64(sample_base)
    with torch.no_grad():
        for sid in range(len(plan.steps) - 1, -1, -1):
            s = plan.steps[sid]
            F, H, Ki = s.num_folds, s.arity, s.num_input_units
            cur = sel[sid]
            ff, nn_ = np.nonzero(cur >= 0)
            o = cur[ff, nn_]
            ctr = np.zeros((len(ff), 4), dtype=np.uint32)
            ctr[:, 0] = (n_idx[nn_] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            ctr[:, 1] = (n_idx[nn_] >> np.uint64(32)).astype(np.uint32)
            ctr[:, 2] = (row0[sid] + ff).astype(np.uint32)
            rnd = philox4x32_10(ctr, k0, k1)
            u = _u01(rnd[:, 0])

            def put(h, units, rows=None):
                """the paths `rows` (default: all active ones) go through `units` of input h"""
                rows = np.arange(len(ff)) if rows is None else rows
                st, fo = s.in_step[ff[rows], h], s.in_fold[ff[rows], h]
                for t in np.unique(st):
                    m = st == t
                    sel[int(t)][fo[m], nn_[rows][m]] = units[m]

            if s.kind in ("sum", "cpt", "tucker"):
                j = _draw(cdfs[sid][ff, o], u)
                mixes.setdefault(sid, np.full((F, N), -1, dtype=np.int64))[ff, nn_] = j
                if s.kind == "sum" and H > 1:  # concatenated inputs: component j = (h, unit)
                    hh = j // Ki
                    for h in range(H):
                        m = np.nonzero(hh == h)[0]
                        if len(m):
                            put(h, (j - h * Ki)[m], rows=m)
                elif s.kind == "tucker":
                    put(0, j // Ki)
                    put(1, j % Ki)
                else:
                    for h in range(H):
                        put(h, j)
            elif s.kind == "mixing":
                hh = _draw(cdfs[sid][ff, o], u)
                mixes.setdefault(sid, np.full((F, N), -1, dtype=np.int64))[ff, nn_] = hh
                for h in range(H):
                    m = np.nonzero(hh == h)[0]
                    if len(m):
                        put(h, o[m], rows=m)
            elif s.kind == "hadamard":
                for h in range(H):
                    put(h, o)
            elif s.kind == "kronecker":
                put(0, o // Ki)
                put(1, o % Ki)
            elif s.kind == "categorical":
                v = _draw(cdfs[sid][ff, o], u)
                x[nn_, np.asarray(s.scope_idx)[ff]] = v
            elif s.kind == "gaussian":
                mean = oc.param(s.params["mean"]).float().numpy().reshape(F, -1)
                std = oc.param(s.params["stddev"]).float().numpy().reshape(F, -1)
                u1 = np.float32(1.0) - u
                u2 = _u01(rnd[:, 1])
                z = np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(2.0 * np.pi) * u2)
                x[nn_, np.asarray(s.scope_idx)[ff]] = (mean[ff, o] + std[ff, o] * z).astype(np.float32)
            else:
                raise TypeError(f"Sampling is not supported for layers of type {s.kind}")
    return x, mixes
