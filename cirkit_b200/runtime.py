"""Device runtime: binds a :class:`CircuitPlan` to the CUDA library and to autograd.

One `PlanRuntime` per circuit; per device it owns the index tables, the C plan handles, the
effective-parameter buffers and a scratch workspace.  A forward pass is ONE call into the library
(`ckb_plan_forward`: parameter ops + every folded layer), a backward pass is one more
(`ckb_plan_backward`), wrapped in a single `torch.autograd.Function` whose differentiable inputs
are the parameter tensors.  PyTorch is used for memory, streams and autograd plumbing only.

What this replaces in the reference: `TorchDiAcyclicGraph.evaluate`
(cirkit/backend/torch/graph/modules.py:303-335), `LayerAddressBook.lookup`
(circuits.py:30-71), every `TorchLayer.forward` on the path and the autograd graph they record.

Execution plans.  The same circuit is lowered to two step lists:
  * "plain": one kernel group per folded layer, exactly the reference's layer sequence;
  * "fused": an input table layer (Categorical / Embedding) that is consumed fold-by-fold by an
    arity-1 sum layer is merged with it into one TABLE_DENSE step: the sum layer is applied to
    the V rows of the table once per step instead of to every sample, and samples only gather
    rows of the resulting table.  Used when the batch is at least V/2 and no integration mask is
    given; values are the same as the plain plan's up to fp32 rounding.
"""

from __future__ import annotations

import ctypes as C
import dataclasses
import os
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib as L
from .plan import CircuitPlan, ParamSpec, StepSpec, build_layout

_DTYPES = {
    torch.uint8: L.U8,
    torch.int16: L.I16,
    torch.int32: L.I32,
    torch.int64: L.I64,
    torch.float32: L.F32,
    torch.float64: L.F64,
}
STEP_TABLE_DENSE = 8


def apply_param_op(t: Tensor, op: str, attrs: dict) -> Tensor:
    """Host (PyTorch) evaluation of one re-parameterisation node, used only for the part of a
    parameter chain that has no fused kernel (small tensors, once per step)."""
    nd = t.ndim - 1
    if op == "softmax":
        return torch.softmax(t, dim=attrs.get("dim", nd - 1) + 1)
    if op == "log_softmax":
        return torch.log_softmax(t, dim=attrs.get("dim", nd - 1) + 1)
    if op == "scaled_sigmoid":
        return torch.sigmoid(t) * (attrs["vmax"] - attrs["vmin"]) + attrs["vmin"]
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
    if op == "mixing":
        d = torch.vmap(torch.vmap(torch.diag, in_dims=1))(t)
        return d.permute(0, 2, 1, 3).flatten(start_dim=2)
    if op == "conj":
        return torch.conj(t).resolve_conj()
    if op == "clog":  # semiring.map_from(value, SumProduct) of a complex constant (semiring.py:506-508)
        return torch.log(t.to(torch.complex64) if not t.is_complex() else t)
    raise ValueError(f"unknown parameter op {op!r}")


def eval_param_chain(leaves: Sequence[Tensor], spec: dict) -> Tensor:
    """Host (PyTorch, differentiable) evaluation of a `leaf [-> fold slice] -> ops` chain given as
    the plain dict a `matmul` op carries for its right operand."""
    t = leaves[spec["leaf"]]
    if spec.get("fold_idx") is not None:
        t = t.index_select(0, torch.as_tensor(spec["fold_idx"], dtype=torch.int64, device=t.device))
    for op, attrs in spec["ops"]:
        if op == "matmul":
            t = torch.matmul(t, eval_param_chain(leaves, attrs["rhs"]))
        else:
            t = apply_param_op(t, op, attrs)
    return t


@dataclass
class _Binding:
    """One layer parameter: source tensor -> host prefix ops -> optional fused op -> slot."""

    sid: int
    name: str
    spec: ParamSpec
    prefix: list
    native: tuple | None  # (pop kind, rows, cols, aux, a, b)
    src_shape: tuple
    eff_shape: tuple
    src_slot: int = -1
    dst_slot: int = -1
    is_complex: bool = False  # the tensor holds complex numbers ((re, im) float pairs on the device)


def _is_last_dim(op: tuple, name: str, shape: tuple) -> bool:
    nd = len(shape) - 1
    return op[0] == name and op[1].get("dim", nd - 1) in (nd - 1, -1)


def _bind_complex(sid: int, step: StepSpec, name: str, spec: ParamSpec, leaf_is_complex: bool) -> _Binding:
    """Parameters of a 'complex-lse-sum' plan.  Complex tensors stay in the reference's layout
    (Embedding weights (F, K, V), sum weights (F, Ko, Kred)); the only fused op is the conjugate
    (TorchConjugateParameter, nodes.py:742-746).  A Categorical layer keeps its REAL parameters and
    the real log-softmax-transpose op: its log-probabilities are cast to complex by the kernel."""
    ops = list(spec.ops)
    shape = tuple(spec.shape)
    if step.kind == "categorical" and not leaf_is_complex:
        return _bind(sid, step, name, spec)
    if step.kind not in ("embedding", "sum", "cpt", "tucker", "tensordot", "constant"):
        raise NotImplementedError(f"step {sid}: no complex kernel for {step.kind!r} parameters")
    native, prefix = None, ops
    if ops and ops[-1][0] == "conj":
        prefix, native = ops[:-1], (L.POP_CONJ, int(np.prod(shape)), 1, 0, 0.0, 0.0)
    if step.kind == "constant" and not step.config.get("log_space", False):
        prefix = prefix + [("clog", {})]  # batch-free (F, K): the host maps it to log space
    return _Binding(sid, name, spec, prefix, native, shape, shape, is_complex=True)


def _bind(sid: int, step: StepSpec, name: str, spec: ParamSpec) -> _Binding:
    ops = list(spec.ops)
    shape = tuple(spec.shape)
    F = shape[0]
    native = None
    prefix = ops
    eff_shape = shape
    kind = step.kind
    if kind == "categorical":
        K, V = shape[1], shape[2]
        eff_shape = (F, V, K)
        if name == "probs":
            if ops and _is_last_dim(ops[-1], "softmax", shape):
                prefix, native = ops[:-1], (L.POP_LOG_SOFTMAX_T, F, V, K, 0.0, 0.0)
            else:
                native = (L.POP_LOG_T, F, V, K, 0.0, 0.0)
        else:
            if ops and _is_last_dim(ops[-1], "log_softmax", shape):
                prefix, native = ops[:-1], (L.POP_LOG_SOFTMAX_T, F, V, K, 0.0, 0.0)
            else:
                native = (L.POP_COPY_T, F, V, K, 0.0, 0.0)
    elif kind == "embedding":
        K, V = shape[1], shape[2]
        eff_shape = (F, V, K)
        native = (L.POP_LOG_T, F, V, K, 0.0, 0.0)
    elif kind == "gaussian":
        if name == "stddev" and ops and ops[-1][0] == "scaled_sigmoid":
            a = ops[-1][1]
            prefix = ops[:-1]
            native = (L.POP_SCALED_SIGMOID, int(np.prod(shape)), 1, 0, float(a["vmin"]), float(a["vmax"]))
    elif kind == "constant":
        if not step.config.get("log_space", False):
            native = (L.POP_LOG, int(np.prod(shape)), 1, 0, 0.0, 0.0)
    elif kind in ("sum", "cpt", "mixing", "tucker"):
        if ops and _is_last_dim(ops[-1], "softmax", shape):
            prefix = ops[:-1]
            native = (L.POP_SOFTMAX, int(np.prod(shape[:-1])), shape[-1], 0, 0.0, 0.0)
    return _Binding(sid, name, spec, prefix, native, shape, eff_shape)


_KIND = {
    "categorical": L.STEP_TABLE,
    "embedding": L.STEP_TABLE,
    "gaussian": L.STEP_GAUSSIAN,
    "constant": L.STEP_CONSTANT,
    "external": L.STEP_EXTERNAL,
    "sum": L.STEP_DENSE,
    "cpt": L.STEP_DENSE,
    "mixing": L.STEP_MIXING,
    "hadamard": L.STEP_HADAMARD,
    "kronecker": L.STEP_KRONECKER,
    "tucker": L.STEP_TUCKER,
    "tensordot": L.STEP_TENSORDOT,
}
_PARAM_ORDER = {
    "categorical": ("probs|logits",),
    "embedding": ("weight",),
    "gaussian": ("mean", "stddev", "log_partition"),
    "constant": ("value",),
    "external": ("output",),
    "sum": ("weight",),
    "cpt": ("weight",),
    "mixing": ("weight",),
    "tucker": ("weight",),
    "tensordot": ("weight",),
    "hadamard": (),
    "kronecker": (),
}


@dataclass
class ExecStep:
    """One entry of an execution plan (what a ckb_step_desc_t is built from)."""

    kind: int  # ckb_step_kind
    label: str
    sids: tuple  # plan steps it covers
    out_sid: int  # plan step whose arena block / consumer lists it uses
    slots: list = field(default_factory=list)
    scratch: tuple | None = None  # (slot, shape) of a runtime-owned buffer (TABLE_DENSE: T2)
    folds: tuple | None = None  # (f0, f1): the step covers this fold range only (staged backward)
    flags: int = 0  # extra ckb_step_desc_t flags (CKB_STEP_NO_GATHER / CKB_STEP_TABLE_INPUT)
    # CKB_STEP_TABLE_INPUT: (table step id, absorbed sum step id, T2 slot) of the fused input pair
    table_input: tuple | None = None


@dataclass
class GradStage:
    """One stage of a staged backward pass (`PlanRuntime.enable_gradient_stages`): the steps and
    parameter ops after which a group of parameter gradients is final and may be all-reduced."""

    steps: tuple  # [begin, end) in the "fused_sync" execution plan
    ops: tuple  # [begin, end) in that plan's parameter-op list
    pieces: list  # (binding index, fold begin, fold end) of the gradients that become final


def _rows64(lay, sid: int, gathered: bool = True) -> bool:
    """All arena / gradient-arena row offsets of step `sid` are multiples of 64 floats
    (CKB_STEP_ROWS64: the TMA-fed kernels address the arenas as matrices of 64-float rows)."""
    for arr in ((lay.in_rows[sid] if gathered else None), lay.cons_rows[sid]):
        if arr is not None and arr.size and np.any(arr % 64):
            return False
    return lay.out_off[sid] % 64 == 0 and (lay.gin_off[sid] < 0 or lay.gin_off[sid] % 64 == 0)


def find_table_dense_pairs(plan: CircuitPlan) -> dict[int, int]:
    """input step -> sum step, for every table layer whose only consumer is an arity-1 sum layer
    reading it fold by fold (e.g. the Categorical -> Sum pair every region-graph circuit starts
    with, `cirkit/templates/region_graph/graph.py:344-588`)."""
    consumers: dict[int, set] = {i: set() for i in range(len(plan.steps))}
    for cid, c in enumerate(plan.steps):
        if not c.is_input:
            for p in np.unique(c.in_step):
                consumers[int(p)].add(cid)
    outs = set(int(s) for s in plan.out_step)
    pairs = {}
    for sid, s in enumerate(plan.steps):
        if s.kind not in ("categorical", "embedding") or sid in outs or len(consumers[sid]) != 1:
            continue
        (cid,) = consumers[sid]
        c = plan.steps[cid]
        if c.kind not in ("sum", "cpt") or c.arity != 1 or c.num_folds != s.num_folds:
            continue
        if not np.array_equal(c.in_fold.reshape(-1), np.arange(s.num_folds)):
            continue
        if "weight" not in c.params or max(c.num_input_units, c.num_output_units) > 128:
            continue
        pairs[sid] = cid
    return pairs


class _DeviceState:
    """Everything that lives on one GPU for one plan."""

    def __init__(self, rt: "PlanRuntime", device: torch.device):
        self.device = device
        self.rt = rt
        self.lib = L.load()
        self.keep: list[Tensor] = []  # index tables referenced by the C plans
        lay, plan = rt.layout, rt.plan

        def up(a, dtype) -> int:
            if a is None or a.size == 0:
                t = torch.zeros(1, dtype=dtype, device=device)
            else:
                t = torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)
            self.keep.append(t)
            return t.data_ptr()

        self.tables = []
        for sid, s in enumerate(plan.steps):
            self.tables.append({
                "in_rows": up(lay.in_rows[sid], torch.int64) if lay.in_rows[sid] is not None else None,
                "scope_var": up(s.scope_idx, torch.int32) if s.scope_idx is not None else None,
                "cons_ptr": up(lay.cons_ptr[sid], torch.int32),
                "cons_rows": up(lay.cons_rows[sid], torch.int64),
            })
        self.handles: dict[str, C.c_void_p] = {}
        # effective parameters and their gradients (runtime-owned, persistent)
        self.eff: dict[int, Tensor] = {}
        self.eff_grad: dict[int, Tensor] = {}
        for b in rt.bindings:
            if b.native is not None:
                shape = (*b.eff_shape, 2) if b.is_complex else b.eff_shape
                self.eff[b.dst_slot] = torch.empty(shape, dtype=torch.float32, device=device)
        for es in rt.exec_plans["fused"]:
            if es.scratch is not None:
                slot, shape = es.scratch
                self.eff[slot] = torch.empty(shape, dtype=torch.float32, device=device)
        self.int_buf = {
            sid: torch.zeros(plan.steps[sid].num_folds, plan.steps[sid].num_output_units,
                             dtype=torch.float32, device=device)
            for sid in rt.int_buf_sids
        }
        self.ws: Tensor | None = None
        # effective-parameter cache: identity + version of every parameter tensor the
        # buffers in `eff` (and the logsumexp buffers of the masked plan) were last computed from
        self.param_key: tuple | None = None
        self.lse_key: tuple | None = None
        self.tensor_table: tuple | None = None  # (parameter pointers, n_slots, ctypes table): see _prepare_call
        self.grad_layouts: dict = {}  # cached gradient-table layouts: see _grad_table
        # SamplingQuery: CDF buffers per step (valid for sample_key) and the selection-arena tables
        self.cdf: dict[int, Tensor] = {}
        self.sample_key: tuple | None = None
        self._sampling: dict | None = None

    def sampling_tables(self) -> dict:
        """Selection arena of the sampler: one row per (step, fold); per inner step the rows of
        its inputs as a device (F*H) int32 table."""
        if self._sampling is None:
            plan = self.rt.plan
            row0 = np.concatenate([[0], np.cumsum([s.num_folds for s in plan.steps])]).astype(np.int64)
            in_rows = []
            for s in plan.steps:
                if s.is_input:
                    in_rows.append(None)
                    continue
                r = row0[s.in_step.astype(np.int64)] + s.in_fold.astype(np.int64)  # (F, H)
                t = torch.from_numpy(np.ascontiguousarray(r.astype(np.int32))).to(self.device)
                self.keep.append(t)
                in_rows.append(t.data_ptr())
            self._sampling = {"row0": row0, "rows": int(row0[-1]), "in_rows": in_rows}
        return self._sampling

    def table_folds(self, sid: int) -> int:
        """Device (F*H) int64 table of the folds step `sid` reads (CKB_STEP_TABLE_INPUT)."""
        key = ("table_folds", sid)
        if key not in self.tables[sid]:
            a = np.ascontiguousarray(self.rt.plan.steps[sid].in_fold.astype(np.int64))
            t = torch.from_numpy(a).to(self.device)
            self.keep.append(t)
            self.tables[sid][key] = t.data_ptr()
        return self.tables[sid][key]

    def grad_buffer(self, slot: int) -> Tensor:
        g = self.eff_grad.get(slot)
        if g is None:
            like = self.eff.get(slot)
            if like is None:  # an integrate-value buffer
                like = next(b for sid, b in self.int_buf.items() if self.rt.int_slots[sid] == slot)
            g = self.eff_grad[slot] = torch.empty_like(like)
        return g

    def handle(self, which: str) -> C.c_void_p:
        h = self.handles.get(which)
        if h is not None:
            return h
        rt, lay, plan = self.rt, self.rt.layout, self.rt.plan
        steps = rt.exec_plans[which]
        descs = (L.StepDesc * len(steps))()
        for i, es in enumerate(steps):
            s = plan.steps[es.out_sid]
            first = plan.steps[es.sids[0]]
            t_out, t_first = self.tables[es.out_sid], self.tables[es.sids[0]]
            d = descs[i]
            d.kind = es.kind
            f0, f1 = es.folds if es.folds is not None else (0, s.num_folds)
            d.num_folds, d.arity = f1 - f0, s.arity
            d.k_in, d.k_out = max(s.num_input_units, 0), s.num_output_units
            d.flags = (L.DENSE_CONCAT if (s.kind == "sum" and s.arity > 1) else 0) | es.flags
            if _rows64(lay, es.out_sid, gathered=es.table_input is None):
                d.flags |= L.STEP_ROWS64
            if rt.is_complex:
                d.flags |= L.STEP_COMPLEX
                if s.kind == "categorical":
                    d.flags |= L.STEP_REAL_TABLE
            d.num_states = int(first.config.get("num_categories", first.config.get("num_states", 0)))
            if s.kind == "tensordot":
                d.num_states = int(s.config["kq"])  # vectors interleaved in a sample row
            d.gin_h = int(lay.gin_h[es.out_sid])
            d.out_off = int(lay.out_off[es.out_sid]) + f0 * s.num_output_units
            d.gin_off = int(lay.gin_off[es.out_sid])
            d.in_rows = t_out["in_rows"] if es.kind != STEP_TABLE_DENSE else None
            # a fold range of an input step: the per-fold tables start at f0 (int32 entries; the
            # CSR values index cons_rows absolutely)
            d.scope_var = t_first["scope_var"] + 4 * f0 if t_first["scope_var"] else None
            d.cons_ptr = t_out["cons_ptr"] + 4 * f0
            d.cons_rows = t_out["cons_rows"]
            for j in range(4):
                d.slot[j] = es.slots[j] if j < len(es.slots) else -1
            d.int_slot = rt.int_slots.get(es.sids[0], -1) if es.kind != STEP_TABLE_DENSE else -1
            cp = lay.cons_ptr[es.out_sid]
            d.max_consumers = int(np.max(np.diff(cp))) if len(cp) > 1 else 0
            if es.table_input is not None:
                # the layer gathers rows of the pair's T2 table: in_rows = table FOLDS, u lives in
                # the arena block of the absorbed sum layer (never materialised in this plan)
                tsid, cid, t2 = es.table_input
                d.slot[1] = t2
                d.scope_var = self.tables[tsid]["scope_var"]
                d.in_rows = self.table_folds(es.out_sid)
                d.num_states = int(plan.steps[tsid].config.get("num_categories", plan.steps[tsid].config.get("num_states", 0)))
                d.aux_off = int(lay.out_off[cid])
        # the logsumexp ops come first: the backward pass runs the ops in reverse, and theirs ADDS
        # to the logits gradient the table op has written by then
        op_list = [(L.POP_LSE_ROWS, src, dst, rows, cols, 0, 0.0, 0.0)
                   for src, dst, rows, cols in (rt.lse_ops if which == "masked" else [])]
        if which in rt.sync_ops:
            op_list += rt.sync_ops[which]
        else:
            op_list += [(kind, b.src_slot, b.dst_slot, rows, cols, aux, a, bb)
                        for b, (kind, rows, cols, aux, a, bb) in rt.native_ops]
        ops = (L.ParamOp * max(1, len(op_list)))()
        for i, (kind, src, dst, rows, cols, aux, a, bb) in enumerate(op_list):
            ops[i].kind, ops[i].src, ops[i].dst = kind, src, dst
            ops[i].rows, ops[i].cols, ops[i].aux, ops[i].a, ops[i].b = rows, cols, aux, a, bb
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(
                self.lib.ckb_plan_create(descs, len(steps), ops, len(op_list), rt.n_slots, C.byref(h)),
                "ckb_plan_create",
            )
        self.handles[which] = h
        return h

    def workspace(self, which: str, batch: int) -> Tensor:
        need = int(self.lib.ckb_plan_workspace_bytes(self.handle(which), batch))
        if which + "_sync" in self.rt.exec_plans:
            # a fold chunk is a smaller launch and may be cut into more split-K slabs
            need = max(need, int(self.lib.ckb_plan_workspace_bytes(self.handle(which + "_sync"), batch)))
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self.ws

    def __del__(self):  # pragma: no cover
        try:
            for h in self.handles.values():
                self.lib.ckb_plan_destroy(h)
        except Exception:
            pass


class PlanRuntime:
    def __init__(self, plan: CircuitPlan, *, fuse_tables: bool = True, fuse_table_inputs: bool = True):
        plan.validate()
        self.is_complex = plan.semiring == "complex-lse-sum"
        if not self.is_complex and any(s.kind == "tensordot" for s in plan.steps):
            raise NotImplementedError("tensordot layers have kernels for the 'complex-lse-sum' semiring only")
        if self.is_complex:
            fuse_tables = False
            for sid, s in enumerate(plan.steps):
                if s.kind in ("mixing", "kronecker", "gaussian") or (s.kind == "sum" and s.arity > 1):
                    raise NotImplementedError(
                        f"step {sid}: no 'complex-lse-sum' kernel for {s.kind!r} layers"
                        + (" over concatenated inputs" if s.kind == "sum" else ""))
                if s.kind == "tucker" and s.arity != 2:
                    raise NotImplementedError(f"step {sid}: complex tucker layers of arity {s.arity}")
        self.plan = plan
        self.layout = build_layout(plan)
        self.bindings: list[_Binding] = []
        self.step_slots: list[list[int]] = []
        self.int_slots: dict[int, int] = {}
        self.int_buf_sids: list[int] = []  # steps whose integrate values live in a runtime buffer
        self.lse_ops: list[tuple] = []  # (src slot, dst slot, rows, cols) of CKB_POP_LSE_ROWS ops
        n = 0
        for sid, s in enumerate(plan.steps):
            slots = []
            for pname in _PARAM_ORDER[s.kind]:
                name = next((c for c in pname.split("|") if c in s.params), None)
                if name is None:
                    slots.append(-1)
                    continue
                if self.is_complex:
                    spec = s.params[name]
                    leaf_c = spec.leaf >= 0 and plan.leaves[spec.leaf].dtype == "complex"
                    b = _bind_complex(sid, s, name, spec, leaf_c or spec.leaf < 0)
                else:
                    b = _bind(sid, s, name, s.params[name])
                b.src_slot = n
                n += 1
                if b.native is not None:
                    b.dst_slot = n
                    n += 1
                else:
                    b.dst_slot = b.src_slot
                self.bindings.append(b)
                slots.append(b.dst_slot)
            self.step_slots.append(slots)
            if s.kind == "categorical" and "logits" in s.params and b.native[0] == L.POP_COPY_T:
                # unnormalised logits: integrating the variable yields logsumexp(logits)
                # (layers/input.py:414-421), a parameter op of the masked plan (normalised
                # logits integrate to log 1 = 0, the kernels' default)
                self.int_slots[sid] = n
                self.int_buf_sids.append(sid)
                K, V = b.src_shape[1], b.src_shape[2]
                self.lse_ops.append((b.src_slot, n, s.num_folds * K, V))
                n += 1
            if s.kind == "gaussian" and "log_partition" in s.params:
                self.int_slots[sid] = slots[2]
        # execution plans
        plain = [
            ExecStep(_KIND[s.kind], f"{sid}:{s.kind}", (sid,), sid, list(self.step_slots[sid]))
            for sid, s in enumerate(plan.steps)
        ]
        fused: list[ExecStep] = []
        pairs = find_table_dense_pairs(plan) if fuse_tables else {}
        # the fused step back-propagates into runtime-owned table buffers: both parameters must
        # come out of a fused parameter op (true for every softmax/log-softmax parameterisation)
        def owned(sid: int) -> bool:
            return all(b.native is not None for b in self.bindings if b.sid == sid)

        self.table_pairs = {a: c for a, c in pairs.items() if owned(a)}
        absorbed = set(self.table_pairs.values())
        self.table_states = 0
        for sid, s in enumerate(plan.steps):
            if sid in absorbed:
                continue
            if sid in self.table_pairs:
                cid = self.table_pairs[sid]
                c = plan.steps[cid]
                V = int(s.config.get("num_categories", s.config.get("num_states", 0)))
                self.table_states = max(self.table_states, V)
                t2 = n
                n += 1
                fused.append(ExecStep(STEP_TABLE_DENSE, f"{sid}+{cid}:table_dense", (sid, cid), cid,
                                      [self.step_slots[sid][0], self.step_slots[cid][0], t2],
                                      scratch=(t2, (c.num_folds, V, c.num_output_units))))
            else:
                fused.append(plain[sid])
        if fuse_table_inputs:
            self._fuse_table_inputs(fused)
        # "masked": the plain step list plus the logsumexp ops of unnormalised categoricals
        self.exec_plans = {"plain": plain, "fused": fused, "masked": plain}
        self.n_slots = n
        self.native_ops = [(b, b.native) for b in self.bindings if b.native is not None]
        self.reads_evidence = any(s.kind in ("categorical", "embedding", "gaussian") for s in plan.steps)
        self.needs_batch = any(s.kind == "external" for s in plan.steps)
        self._states: dict[torch.device, _DeviceState] = {}
        self.last_launches = 0
        self.cache_parameters = True  # skip the parameter ops of no_grad calls on unchanged parameters
        self.check_evidence = False  # validate integer evidence against the number of states (costs a sync)
        # replay repeated forward / backward calls as CUDA graphs (CKB_USE_GRAPHS); CKB_GRAPHS=0 disables.
        # External steps hand the library per-call PyTorch tensors: their plans stay eager.
        # "auto": only where the host is the bottleneck -- activation arenas up to 256 MB (measured:
        # K = 32, B = 512: 0.70 -> 0.66 ms per step; the K = 64, B = 2048 step is GPU-bound and 1 % slower replayed)
        env = os.environ.get("CKB_GRAPHS", "auto")
        self.use_graphs = False if (env == "0" or self.needs_batch) else (True if env == "1" else "auto")
        self.keep_arena = False
        self.last_arena: Tensor | None = None
        self.last_flat_grad: Tensor | None = None  # flat buffer behind the last backward's gradients
        self.last_flat_target: Tensor | None = None  # where the kernels wrote them (the same, or the hook's buffer)
        # staged backward (data-parallel overlap, see enable_gradient_stages)
        self.aliases: dict[int, tuple] = {}  # slot -> (base slot, float offset) into the same buffer
        self.sync_ops: dict[str, list] = {}
        self.grad_stages: dict[str, list] = {}  # execution plan -> [GradStage] in backward order
        self.grad_sync = None  # callable(list of gradient tensors) with a .finish() method
        self.last_synced_bytes = 0

    # ------------------------------------------------------------------ staged backward
    def enable_gradient_stages(self, chunks: int = 4, bucket_bytes: int = 8 << 20,
                               chunk_steps: bool = False) -> bool:
        """Prepare a backward pass that finishes the parameter gradients GROUP BY GROUP, so that a
        data-parallel wrapper can all-reduce one group while the next is still being computed
        ("bottom layers hold most of the bytes and finish last, so bucket
        top-down").  The inner layers go first, in backward order, cut into stages of at least
        `bucket_bytes` of gradients (each stage = some steps + the parameter ops of their
        weights); the large input tables (Categorical / Embedding, alone or fused with their
        dense sum: the bulk of the parameters, and last in the backward pass) follow in `chunks`
        fold ranges, each with the parameter ops of its slice -- by default only the PARAMETER OP
        of a large table is cut into fold ranges (its gradient leaves in `chunks` pieces while the
        next range is transformed); `chunk_steps` also cuts the layer's own backward kernels, which
        starts the first collective earlier but runs them as smaller, less efficient launches
        (measured on 2 GPUs: not worth it at the north-star shape).  The stages live in extra
        execution plans ("fused_sync", "plain_sync") used by the backward pass only while
        `grad_sync` is set; values are bit-equal to the unstaged pass (same kernels, same order
        within every fold).  Returns False (nothing changes) for complex plans and plans with
        per-sample PyTorch inputs."""
        if self.grad_stages:
            return True
        if self.is_complex or self.needs_batch:
            return False
        bidx = {id(b): i for i, b in enumerate(self.bindings)}
        by_sid: dict[int, list] = {}
        for b in self.bindings:
            by_sid.setdefault(b.sid, []).append(b)

        def gbytes(es: ExecStep) -> int:
            return sum(4 * int(np.prod(b.src_shape)) for sid in es.sids for b in by_sid.get(sid, []))

        def op_of(b, src=None, dst=None, rows=None):
            kind, r, cols, aux, a, bb = b.native
            return (kind, b.src_slot if src is None else src, b.dst_slot if dst is None else dst,
                    r if rows is None else rows, cols, aux, a, bb)

        n = self.n_slots

        def alias(base: int, off: int) -> int:
            nonlocal n
            if off == 0:
                return base
            self.aliases[n] = (base, off)
            n += 1
            return n - 1

        for which in ("fused", "plain"):
            base = self.exec_plans[which]
            if which == "fused" and not self.table_pairs:
                continue

            def chunkable(es: ExecStep) -> bool:
                if es.kind not in (STEP_TABLE_DENSE, L.STEP_TABLE) or gbytes(es) < bucket_bytes:
                    return False
                bs = [b for sid in es.sids for b in by_sid.get(sid, [])]
                return self.plan.steps[es.out_sid].num_folds >= chunks and all(
                    b.native is not None and b.spec.fold_idx is None for b in bs)

            big = [es for es in base if chunkable(es)] if chunks > 1 else []
            rest = [es for es in base if not any(es is t for t in big)]
            steps: list[ExecStep] = []
            ops: list[tuple] = []
            tail_stages = []  # stages of the big input steps, in backward order
            if big and chunk_steps:
                for c in range(chunks):
                    s0, o0, pieces = len(steps), len(ops), []
                    for es in big:
                        F = self.plan.steps[es.out_sid].num_folds
                        f0, f1 = c * F // chunks, (c + 1) * F // chunks
                        slots = []
                        for slot in es.slots:
                            owner = [b for sid in es.sids for b in by_sid.get(sid, []) if b.dst_slot == slot]
                            if owner:
                                (b,) = owner
                                src = alias(b.src_slot, f0 * int(np.prod(b.src_shape[1:])))
                                dst = alias(b.dst_slot, f0 * int(np.prod(b.eff_shape[1:])))
                                ops.append(op_of(b, src, dst, b.native[1] // F * (f1 - f0)))
                                pieces.append((bidx[id(b)], f0, f1))
                                slots.append(dst)
                            elif es.scratch is not None and slot == es.scratch[0]:
                                slots.append(alias(slot, f0 * int(np.prod(es.scratch[1][1:]))))
                            else:
                                slots.append(slot)
                        steps.append(ExecStep(es.kind, f"{es.label}[{f0}:{f1}]", es.sids, es.out_sid, slots,
                                              scratch=es.scratch, folds=(f0, f1), flags=es.flags))
                    tail_stages.insert(0, GradStage((s0, len(steps)), (o0, len(ops)), pieces))
            elif big:
                # whole steps; the parameter op of every large tensor in `chunks` fold ranges
                steps = list(big)
                o0, pieces, split = 0, [], []
                for es in big:
                    for sid in es.sids:
                        for b in by_sid.get(sid, []):
                            if 4 * int(np.prod(b.src_shape)) >= bucket_bytes and b.src_shape[0] >= chunks:
                                split.append(b)
                                continue
                            if b.native is not None:
                                ops.append(op_of(b))
                            pieces.append((bidx[id(b)], 0, b.src_shape[0]))
                tail_stages.append(GradStage((0, len(steps)), (o0, len(ops)), pieces))
                for c in range(chunks):
                    o0, pieces = len(ops), []
                    for b in split:
                        F = b.src_shape[0]
                        f0, f1 = c * F // chunks, (c + 1) * F // chunks
                        src = alias(b.src_slot, f0 * int(np.prod(b.src_shape[1:])))
                        dst = alias(b.dst_slot, f0 * int(np.prod(b.eff_shape[1:])))
                        ops.append(op_of(b, src, dst, b.native[1] // F * (f1 - f0)))
                        pieces.append((bidx[id(b)], f0, f1))
                    tail_stages.append(GradStage((0, 0), (o0, len(ops)), pieces))
            # the other steps keep their order; stages are cut walking them backwards
            n_big = len(steps)
            steps += rest
            stages, hi, acc = [], len(steps), 0
            for i in range(len(steps) - 1, n_big - 1, -1):
                acc += gbytes(steps[i])
                if acc >= bucket_bytes or i == n_big:
                    o0, pieces = len(ops), []
                    for es in steps[i:hi]:
                        for sid in es.sids:
                            for b in by_sid.get(sid, []):
                                if b.native is not None:
                                    ops.append(op_of(b))
                                pieces.append((bidx[id(b)], 0, b.src_shape[0]))
                    if pieces or not stages:
                        stages.append(GradStage((i, hi), (o0, len(ops)), pieces))
                    else:  # steps without parameters: extend the previous stage downwards
                        last = stages[-1]
                        stages[-1] = GradStage((i, last.steps[1]), last.ops, last.pieces)
                    hi, acc = i, 0
            self.exec_plans[which + "_sync"] = steps
            self.grad_stages[which] = stages + tail_stages
            self.sync_ops[which + "_sync"] = ops
        self.n_slots = n
        return True

    def _fuse_table_inputs(self, fused: list) -> None:
        """A CP-T layer (Hadamard + dense) all of whose inputs are rows of ONE fused table pair, and
        which is that pair's only reader, gathers the table rows itself (CKB_STEP_TABLE_INPUT): the
        pair's (F', B, K) output block is never written -- the sum of the H gathered rows goes into
        it instead, (F, B, K), and both passes of the layer read that."""
        plan, lay = self.plan, self.layout
        readers: dict[int, set] = {}
        for cid, c in enumerate(plan.steps):
            if not c.is_input:
                for p in np.unique(c.in_step):
                    readers.setdefault(int(p), set()).add(cid)
        outs = set(int(s) for s in plan.out_step)
        by_out = {es.out_sid: es for es in fused}
        for td in [es for es in fused if es.kind == STEP_TABLE_DENSE]:
            tsid, cid = td.sids
            if cid in outs or len(readers.get(cid, ())) != 1:
                continue
            (rid,) = readers[cid]
            r, es = plan.steps[rid], by_out.get(rid)
            if (es is None or r.kind != "cpt" or not 1 < r.arity <= 4 or r.num_input_units % 4
                    or r.num_folds * r.arity > 2 * plan.steps[cid].num_folds
                    or lay.out_off[cid] % 64 or lay.gin_h[rid] != 1):
                continue
            td.flags |= L.STEP_NO_GATHER
            # (the plain plan shares its ExecStep objects with the fused one: replace, not mutate)
            fused[fused.index(es)] = dataclasses.replace(
                es, flags=es.flags | L.STEP_TABLE_INPUT, table_input=(tsid, cid, td.scratch[0]))

    def graphs_for(self, batch: int) -> bool:
        if self.use_graphs == "auto":
            return (2 if self.is_complex else 1) * batch * self.layout.arena_units * 4 <= (256 << 20)
        return bool(self.use_graphs)

    def invalidate_parameter_cache(self) -> None:
        """Forget the cached effective parameters (call after changing parameter storage behind
        autograd's back, e.g. through `.data`, which does not bump the version counter)."""
        for st in self._states.values():
            st.param_key = st.lse_key = None

    def step_output(self, sid: int, batch: int) -> Tensor:
        """(F, B, K) activations of plan step `sid` from the last forward pass (keep_arena=True)."""
        s = self.plan.steps[sid]
        off = batch * int(self.layout.out_off[sid])
        n = s.num_folds * batch * s.num_output_units
        return self.last_arena[off : off + n].view(s.num_folds, batch, s.num_output_units)

    def choose_plan(self, batch: int, masked: bool) -> str:
        if masked:
            return "masked" if self.lse_ops else "plain"
        if self.table_pairs and 2 * batch >= self.table_states:
            return "fused"
        return "plain"

    # ------------------------------------------------------------------ device state
    def state(self, device: torch.device) -> _DeviceState:
        if device.type != "cuda":
            raise RuntimeError(
                "cirkit_b200 evaluates circuits with CUDA kernels only: parameters must live on a "
                f"CUDA device (found {device}).  There is no CPU fallback."
            )
        device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        st = self._states.get(device)
        if st is None:
            st = self._states[device] = _DeviceState(self, device)
        return st

    # ------------------------------------------------------------------ parameters
    def parameter_tensors(self, leaves: Sequence[Tensor], externals: dict | None) -> list[Tensor]:
        out = []
        for b in self.bindings:
            if b.spec.leaf < 0:
                if externals is None or (b.sid, b.name) not in externals:
                    raise ValueError(f"missing external parameter for step {b.sid} {b.name!r}")
                t = externals[(b.sid, b.name)]
            else:
                t = leaves[b.spec.leaf]
                if b.spec.fold_idx is not None:
                    idx = torch.as_tensor(b.spec.fold_idx, dtype=torch.int64, device=t.device)
                    t = t.index_select(0, idx)
            for op, attrs in b.prefix:
                if op == "matmul":  # W1 @ W2, the right operand is another leaf -> op chain
                    t = torch.matmul(t, eval_param_chain(leaves, attrs["rhs"]))
                else:
                    t = apply_param_op(t, op, attrs)
            want = torch.complex64 if b.is_complex else torch.float32
            if t.dtype != want:
                t = t.to(want)
            if b.name == "output" and self.plan.steps[b.sid].kind == "external":
                pass  # (F, B, K): the batch axis is only known per call
            elif tuple(t.shape) != b.src_shape:
                raise ValueError(
                    f"step {b.sid} parameter {b.name!r}: expected shape {b.src_shape}, got {tuple(t.shape)}"
                )
            out.append(t.resolve_conj().contiguous() if t.is_complex() else t.contiguous())
        return out

    # ------------------------------------------------------------------ sampling
    _SAMPLE_KINDS = ("categorical", "gaussian", "sum", "cpt", "mixing", "hadamard", "kronecker", "tucker")

    def sample(self, num_samples: int, leaves: Sequence[Tensor], externals: dict | None = None, *,
               seed: int | None = None, return_mixtures: bool = True,
               chunk: int = 1 << 16) -> tuple[Tensor, list[Tensor]]:
        """SamplingQuery (cirkit/backend/torch/queries.py:187-275) by ancestral sampling on the
        device (csrc/sampling_kernels.cu).  Returns (samples (N, D), mixture_samples): the latter
        holds, per sum-type layer in plan order, an (F, N) int32 tensor with the mixture component
        sample n drew in fold f, or -1 when the sample's path does not visit that fold (the
        reference returns the draws of EVERY unit, (F, Ko, N); only the ones on the path take
        part in the sample).  `seed` defaults to a draw from torch's global generator, so
        `torch.manual_seed` makes the query reproducible."""
        if num_samples <= 0:
            raise ValueError("The number of samples must be a positive number")
        if self.is_complex:
            raise TypeError("Sampling needs a monotonic circuit (semiring 'lse-sum')")
        plan = self.plan
        for sid, s in enumerate(plan.steps):
            if s.kind not in self._SAMPLE_KINDS:
                raise TypeError(f"Sampling is not supported for layers of type {s.kind} (step {sid})")
            if s.kind in ("kronecker", "tucker") and s.arity != 2:
                raise TypeError(f"Sampling {s.kind} layers of arity {s.arity} is not supported")
        if seed is None:
            seed = int(torch.randint(0, 2**62, (1,)).item())
        with torch.no_grad():
            P = self.parameter_tensors(leaves, externals)
            st = self.state(P[0].device)
            lib, dev = st.lib, st.device
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                tensors = (C.c_void_p * self.n_slots)()
                for b, p in zip(self.bindings, P):
                    tensors[b.src_slot] = p.data_ptr()
                for slot, buf in st.eff.items():
                    tensors[slot] = buf.data_ptr()
                key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in P)
                if st.param_key != key:
                    L.check(lib.ckb_plan_param_ops(st.handle("plain"), 0, len(self.native_ops), 0,
                                                   tensors, None, stream), "ckb_plan_param_ops")
                    st.param_key = key
                    st.sample_key = None
                tab = st.sampling_tables()
                if st.sample_key != key:
                    self._sampling_cdfs(st, tensors, P, stream)
                    st.sample_key = key
                descs = (L.SampleStep * len(plan.steps))()
                by_sid = {}
                for b, p in zip(self.bindings, P):
                    by_sid.setdefault(b.sid, {})[b.name] = (b, p)
                for sid, s in enumerate(plan.steps):
                    d = descs[sid]
                    d.kind = _KIND[s.kind]
                    d.num_folds, d.arity = s.num_folds, s.arity
                    d.k_in, d.k_out = max(s.num_input_units, 0), s.num_output_units
                    d.num_states = int(s.config.get("num_categories", 0))
                    d.flags = L.DENSE_CONCAT if (s.kind == "sum" and s.arity > 1) else 0
                    d.sel_row = int(tab["row0"][sid])
                    d.in_sel_rows = tab["in_rows"][sid]
                    d.scope_var = st.tables[sid]["scope_var"]
                    if sid in st.cdf:
                        d.cdf = st.cdf[sid].data_ptr()
                    if s.kind == "gaussian":
                        d.p0 = tensors[by_sid[sid]["mean"][0].dst_slot]
                        d.p1 = tensors[by_sid[sid]["stddev"][0].dst_slot]
                R, D = int(tab["rows"]), plan.num_variables
                is_float = any(s.kind == "gaussian" for s in plan.steps)
                x = torch.zeros((num_samples, D), dtype=torch.float32 if is_float else torch.int64, device=dev)
                sum_sids = [sid for sid, s in enumerate(plan.steps) if s.kind in ("sum", "cpt", "mixing", "tucker")]
                mixes: list[list[Tensor]] = [[] for _ in sum_sids]
                root = int(tab["row0"][int(plan.out_step[0])]) + int(plan.out_fold[0])
                n_launch = 0
                for base in range(0, num_samples, chunk):
                    n = min(chunk, num_samples - base)
                    sel = torch.empty((R, n), dtype=torch.int32, device=dev)
                    mix = torch.empty((R, n), dtype=torch.int32, device=dev) if return_mixtures else None
                    L.check(lib.ckb_plan_sample(descs, len(plan.steps), n, base, seed, R, root, 0,
                                                sel.data_ptr(), mix.data_ptr() if mix is not None else None,
                                                x[base : base + n].data_ptr(), D, 1 if is_float else 0, stream),
                            "ckb_plan_sample")
                    n_launch += len(plan.steps) + 1
                    if mix is not None:
                        for i, sid in enumerate(sum_sids):
                            r0 = int(tab["row0"][sid])
                            mixes[i].append(mix[r0 : r0 + plan.steps[sid].num_folds])
                self.last_launches = n_launch
        mixture = [torch.cat(m, dim=1) if len(m) > 1 else m[0] for m in mixes] if return_mixtures else []
        return x, mixture

    def _sampling_cdfs(self, st: "_DeviceState", tensors, P, stream) -> None:
        """Row-wise CDFs of every mixture / category distribution (one small kernel per layer,
        batch-free) and the reference's admissibility checks on sum weights
        (layers/inner.py:277-281, layers/optimized.py:182-188)."""
        lib, dev, plan = st.lib, st.device, self.plan
        st.cdf = {}
        for b, p in zip(self.bindings, P):
            s = plan.steps[b.sid]
            src = st.eff[b.dst_slot] if b.native is not None else p
            if s.kind == "categorical":
                F, K, V = b.src_shape
                if b.native is None or b.eff_shape != (F, V, K):
                    raise TypeError(f"step {b.sid}: categorical table without a device layout")
                out = torch.empty((F, K, V), dtype=torch.float32, device=dev)
                L.check(lib.ckb_sample_cdf_rows(src.data_ptr(), out.data_ptr(), F * K, V, 1, K, stream),
                        "ckb_sample_cdf_rows")
            elif s.kind in ("sum", "cpt", "mixing", "tucker") and b.name == "weight":
                w = src.contiguous()
                cols = int(w.shape[-1])
                rows = w.numel() // cols
                # the reference raises TypeError in TorchSumLayer.sample, ValueError in TorchCPTLayer.sample
                exc = TypeError if s.kind in ("sum", "mixing") else ValueError
                if bool((w < 0.0).any()):
                    raise exc("Sampling only works with positive weights")
                out = torch.empty_like(w)
                L.check(lib.ckb_sample_cdf_rows(w.data_ptr(), out.data_ptr(), rows, cols, 0, 0, stream),
                        "ckb_sample_cdf_rows")
                total = out.reshape(rows, cols)[:, -1]
                if not torch.allclose(total, torch.ones(1, device=dev)):
                    raise exc("Sampling only works with a normalized parametrization")
            else:
                continue
            st.cdf[b.sid] = out

    # ------------------------------------------------------------------ evaluation
    def evaluate(
        self,
        x: Tensor | None,
        leaves: Sequence[Tensor],
        externals: dict | None = None,
        integrate_mask: Tensor | None = None,
    ) -> Tensor:
        """Returns the circuit output, shape (B, O, K) (B = 1 when the circuit reads no evidence)."""
        if x is not None and x.ndim != 2:
            raise ValueError(
                "The input to the circuit should have shape (B, D), "
                "where B is the batch size and D is the number of variables "
                "the circuit is defined on"
            )
        if x is None and self.reads_evidence:
            raise ValueError(
                f"Expected some input 'x', as the circuit has scope '{self.plan.scope}'"
            )
        if x is not None and x.shape[1] < self.plan.num_variables:
            raise IndexError(
                f"the circuit reads variable {self.plan.num_variables - 1} but the input has "
                f"{x.shape[1]} columns"
            )
        if self.check_evidence and x is not None and not x.is_floating_point():
            # The table kernels clamp a state to [0, V) instead of faulting; the reference's fancy
            # index (layers/input.py:399-412) raises on such evidence.  Opt-in: the check costs a
            # pass over x and a device synchronisation per call.
            for sid, s in enumerate(self.plan.steps):
                V = int(s.config.get("num_categories", s.config.get("num_states", 0)))
                if s.kind in ("categorical", "embedding") and V > 0:
                    cols = torch.as_tensor(s.scope_idx, dtype=torch.int64, device=x.device)
                    xs = x.index_select(1, cols)
                    if bool(((xs < 0) | (xs >= V)).any()):
                        raise IndexError(f"evidence out of range for step {sid}: states must lie in [0, {V})")
        if integrate_mask is not None and self.is_complex:
            raise NotImplementedError("integration masks are not implemented for 'complex-lse-sum' plans")
        if integrate_mask is not None:
            # TorchInputLayer.integrate raises for layers that cannot be integrated
            # (layers/input.py:82-92: Embedding), and IntegrateQuery._layer_fn calls it only when
            # one of the layer's variables is actually masked (queries.py:136-139)
            m = integrate_mask if integrate_mask.ndim == 2 else integrate_mask.unsqueeze(0)
            for sid, s in enumerate(self.plan.steps):
                if s.kind == "embedding" and m.shape[1] > int(s.scope_idx.max()):
                    cols = torch.as_tensor(s.scope_idx, dtype=torch.int64, device=m.device)
                    if bool(m[:, cols].any()):
                        raise TypeError(
                            f"Integration is not supported for layers of type embedding (step {sid})"
                        )
        P = self.parameter_tensors(leaves, externals)
        if not P:
            raise ValueError("circuit without parameters")
        st = self.state(P[0].device)
        # (inside autograd.Function.forward grad mode is always off: sample it here)
        cache_ok = self.cache_parameters and not torch.is_grad_enabled()
        return _PlanFn.apply(self, st, x, integrate_mask, cache_ok, *P)


@dataclass
class _Call:
    """Device-side arguments shared by the forward and the backward library calls."""

    which: str
    B: int
    xT: Tensor | None
    x_is_float: int
    maskT: Tensor | None
    mask_rows: int
    tensors: object
    n_steps: int


def _prepare_call(rt: PlanRuntime, st: _DeviceState, x, mask, P, stream) -> _Call:
    lib, dev, plan = st.lib, st.device, rt.plan
    xT, x_is_float, B = None, 0, 1
    if x is not None:
        B = int(x.shape[0])
        if B == 0:
            raise ValueError("empty batch")
        if x.dtype not in _DTYPES:
            x = x.to(torch.float32 if x.is_floating_point() else torch.int64)
        xd = x.detach()
        if xd.device != dev:
            xd = xd.to(dev, non_blocking=True)
        if xd.stride(1) != 1:
            xd = xd.contiguous()
        x_is_float = 1 if xd.is_floating_point() else 0
        D = plan.num_variables
        xT = torch.empty((max(D, 1), B), dtype=torch.float32 if x_is_float else torch.int32, device=dev)
        if rt.reads_evidence:
            L.check(
                lib.ckb_transpose_input(xd.data_ptr(), _DTYPES[xd.dtype], B, D, xd.stride(0),
                                        xT.data_ptr(), stream),
                "ckb_transpose_input",
            )
    maskT, mask_rows = None, 0
    if mask is not None:
        m = mask.to(device=dev, dtype=torch.uint8)
        if m.ndim == 1:
            m = m.unsqueeze(0)
        if m.shape[0] not in (1, B) or m.shape[1] < plan.num_variables:
            raise ValueError(f"integration mask of shape {tuple(mask.shape)} does not match x")
        m = m[:, : plan.num_variables].contiguous()
        mask_rows = int(m.shape[0])
        maskT = torch.empty((plan.num_variables, mask_rows), dtype=torch.uint8, device=dev)
        L.check(
            lib.ckb_transpose_mask(m.data_ptr(), mask_rows, plan.num_variables, maskT.data_ptr(), stream),
            "ckb_transpose_mask",
        )
    which = rt.choose_plan(B, mask is not None)
    # the tensor table only changes when a parameter tensor moves (the effective-parameter and
    # integrate buffers are persistent): rebuilt then, reused otherwise
    ptrs = tuple(p.data_ptr() for p in P)
    cached = st.tensor_table
    if cached is not None and cached[0] == ptrs and cached[1] == rt.n_slots:
        tensors = cached[2]
    else:
        tensors = (C.c_void_p * rt.n_slots)()
        for b, ptr in zip(rt.bindings, ptrs):
            tensors[b.src_slot] = ptr  # complex64 storage = interleaved (re, im) floats
        for slot, buf in st.eff.items():
            tensors[slot] = buf.data_ptr()
        for sid, slot in rt.int_slots.items():
            if sid in st.int_buf:
                tensors[slot] = st.int_buf[sid].data_ptr()
        for slot, (base, off) in rt.aliases.items():
            if tensors[base]:
                tensors[slot] = tensors[base] + 4 * off
        st.tensor_table = (ptrs, rt.n_slots, tensors)
    return _Call(which, B, xT, x_is_float, maskT, mask_rows, tensors, len(rt.exec_plans[which]))


def _grad_table(rt: PlanRuntime, st: _DeviceState, call: _Call, P, need) -> tuple[object, list]:
    if not rt.is_complex:
        return _grad_table_cached(rt, st, call, P, need)
    return _grad_table_build(rt, st, call, P, need)[:2]


def _grad_table_cached(rt: PlanRuntime, st: _DeviceState, call: _Call, P, need) -> tuple[object, list]:
    """`_grad_table_build` with everything that does not depend on the call cached per (plan,
    wanted gradients, shapes): the table itself (its entries for runtime-owned buffers never change),
    the offsets into the flat buffer and the slots that follow it.  Per call: one allocation, one
    split, one view per tensor, a few integer stores."""
    key = (call.which, tuple(need), tuple(tuple(p.shape) for p in P), rt.n_slots)
    lay = st.grad_layouts.get(key)
    if lay is None:
        grads, outs, flat_slots = _grad_table_build(rt, st, call, P, need)
        st.grad_layouts[key] = {
            "grads": grads, "sizes": [o.numel() if o is not None else 0 for o in outs],
            "offs": list(rt.last_flat_offsets), "total": 0 if rt.last_flat_grad is None else rt.last_flat_grad.numel(),
            "flat_slots": flat_slots, "shapes": [tuple(p.shape) for p in P],
        }
        return grads, outs
    total = lay["total"]
    flat = torch.empty(total, dtype=torch.float32, device=st.device) if total else None
    target = flat
    alloc = getattr(rt.grad_sync, "alloc", None)
    if alloc is not None and flat is not None and not rt.needs_batch:
        owned = alloc(total, st.device)
        if owned is not None:
            target = owned
    rt.last_flat_grad, rt.last_flat_target = flat, target
    rt.last_flat_offsets = lay["offs"]
    grads = lay["grads"]
    outs: list[Tensor | None] = []
    if flat is not None:
        base = target.data_ptr()
        for slot, off in lay["flat_slots"]:
            grads[slot] = base + 4 * off
        for off, n, shape in zip(lay["offs"], lay["sizes"], lay["shapes"]):
            outs.append(None if off < 0 else flat.narrow(0, off, n).view(shape))
    else:
        outs = [None] * len(P)
    return grads, outs


def _grad_table_build(rt: PlanRuntime, st: _DeviceState, call: _Call, P, need) -> tuple[object, list, list]:
    grads = (C.c_void_p * rt.n_slots)()
    outs: list[Tensor | None] = []
    # all requested parameter gradients are views of ONE flat buffer (16-byte aligned pieces), so
    # that a data-parallel wrapper can sum them over the ranks with a single all-reduce
    nfl = [p.numel() * (2 if p.is_complex() else 1) for p in P]  # floats per tensor
    sizes = [(-(-n // 4) * 4) if nd else 0 for n, nd in zip(nfl, need)]
    flat = torch.empty(sum(sizes), dtype=torch.float32, device=st.device) if sum(sizes) else None
    # A gradient-sync hook may own the memory the kernels write into (a symmetric-memory buffer for
    # the in-switch all-reduce): the tensors handed to autograd stay views of the fresh `flat`,
    # which the hook fills with the reduced values at the end of the backward pass.
    target = flat
    alloc = getattr(rt.grad_sync, "alloc", None)
    if alloc is not None and flat is not None and not rt.needs_batch:
        owned = alloc(flat.numel(), st.device)
        if owned is not None:
            target = owned
    rt.last_flat_grad, rt.last_flat_target = flat, target
    rt.last_flat_offsets = offs = []  # float offset of every binding's gradient in `flat` (-1: none)
    off = 0
    for b, p, nd, sz, n in zip(rt.bindings, P, need, sizes, nfl):
        offs.append(off if nd else -1)
        if not nd:
            outs.append(None)
            continue
        g = flat[off : off + n]
        g = torch.view_as_complex(g.view(*p.shape, 2)) if p.is_complex() else g.view(p.shape)
        outs.append(g)
        grads[b.src_slot] = target.data_ptr() + 4 * off
        off += sz
        if b.native is not None:
            grads[b.dst_slot] = st.grad_buffer(b.dst_slot).data_ptr()
    if call.which == "masked":
        # d(loss)/d(logsumexp(logits)) of the integrated variables: written by the table backward,
        # consumed by the CKB_POP_LSE_ROWS backward
        for src, dst, _, _ in rt.lse_ops:
            if grads[src]:
                grads[dst] = st.grad_buffer(dst).data_ptr()
    for es in rt.exec_plans[call.which]:
        if es.scratch is not None:
            grads[es.scratch[0]] = st.grad_buffer(es.scratch[0]).data_ptr()
            # the fused step back-propagates through the table: its gradient buffer must exist
            t_slot = es.slots[0]
            if not grads[t_slot]:
                grads[t_slot] = st.grad_buffer(t_slot).data_ptr()
    for slot, (base, off) in rt.aliases.items():
        if grads[base]:
            grads[slot] = grads[base] + 4 * off
    # the entries that point into the flat buffer (they move with every call): (slot, float offset)
    flat_slots = [(b.src_slot, o) for b, o in zip(rt.bindings, offs) if o >= 0]
    src_off = dict(flat_slots)
    flat_slots += [(slot, src_off[base] + off) for slot, (base, off) in rt.aliases.items() if base in src_off]
    return grads, outs, flat_slots


def _stage_pieces(rt: "PlanRuntime", stage: GradStage, flat: Tensor, offs: list) -> list[Tensor]:
    """The slices of the flat gradient buffer that are final after `stage`, adjacent ones merged."""
    spans = []
    for bi, f0, f1 in stage.pieces:
        b = rt.bindings[bi]
        row = int(np.prod(b.src_shape[1:])) * (2 if b.is_complex else 1)
        F = b.src_shape[0]
        lo, hi = offs[bi] + f0 * row, offs[bi] + f1 * row
        if f1 == F:
            hi = offs[bi] + -(-F * row // 4) * 4  # the padding up to the next piece travels along
        spans.append((lo, hi))
    spans.sort()
    merged = []
    for lo, hi in spans:
        if merged and merged[-1][1] == lo:
            merged[-1][1] = hi
        else:
            merged.append([lo, hi])
    return [flat[lo:hi] for lo, hi in merged]


class _PlanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rt: PlanRuntime, st: _DeviceState, x, mask, cache_ok, *P):
        lib, dev = st.lib, st.device
        lay, plan = rt.layout, rt.plan
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            call = _prepare_call(rt, st, x, mask, P, stream)
            B = call.B
            cpx = 2 if rt.is_complex else 1  # complex activations: (re, im) float pairs
            arena = torch.empty(cpx * B * lay.arena_units, dtype=torch.float32, device=dev)
            ws = st.workspace(call.which, B)
            # The reference re-materialises every effective weight on every call
            # (TorchParameter.forward, parameters/parameter.py:180-188).  In inference loops (no
            # gradient wanted: eval, IntegrateQuery sweeps) the parameter ops are skipped while the
            # parameter tensors are the same objects at the same version.
            key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in P)
            masked = call.which == "masked"
            fresh = cache_ok and st.param_key == key and (not masked or st.lse_key == key)
            L.check(
                lib.ckb_plan_forward(
                    st.handle(call.which), 0, call.n_steps, B,
                    call.xT.data_ptr() if call.xT is not None else None, call.x_is_float,
                    call.maskT.data_ptr() if call.maskT is not None else None, call.mask_rows,
                    call.tensors, arena.data_ptr(), ws.data_ptr(), ws.numel(),
                    (0 if fresh else L.RUN_PARAM_OPS) | (L.USE_GRAPHS if rt.graphs_for(B) else 0), stream),
                "ckb_plan_forward",
            )
            st.param_key = key
            if masked:
                st.lse_key = key
            rt.last_launches = int(lib.ckb_plan_last_launches(st.handle(call.which))) + (
                1 if call.xT is not None else 0)
            K = plan.num_output_units
            rows = lay.out_rows
            if rt.is_complex:
                ac = torch.view_as_complex(arena.view(-1, 2))
                out = torch.stack([ac[B * int(r) : B * int(r) + B * K].view(B, K) for r in rows], dim=1)
            elif len(rows) == 1:
                r = int(rows[0])
                out = arena[B * r : B * r + B * K].view(B, 1, K).clone()
            else:
                out = torch.stack([arena[B * int(r) : B * int(r) + B * K].view(B, K) for r in rows], dim=1)
        ctx.rt, ctx.st, ctx.call = rt, st, call
        ctx.arena = arena
        if rt.keep_arena:  # debugging aid: per-step activations of the last forward pass
            rt.last_arena = arena
        ctx.P = P
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        rt, st, call = ctx.rt, ctx.st, ctx.call
        lib, dev = st.lib, st.device
        lay, plan = rt.layout, rt.plan
        B = call.B
        need = ctx.needs_input_grad[5:]
        rt.last_synced_bytes = 0
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            cpx = 2 if rt.is_complex else 1
            garena = torch.empty(cpx * B * lay.garena_units, dtype=torch.float32, device=dev)
            O, K = plan.num_outputs, plan.num_output_units
            if rt.is_complex:
                gc = torch.view_as_complex(garena.view(-1, 2))
                gc[B * lay.out_goff : B * lay.out_goff + O * B * K].view(O, B, K).copy_(
                    gout.to(torch.complex64).transpose(0, 1))
            else:
                go = garena[B * lay.out_goff : B * lay.out_goff + O * B * K].view(O, B, K)
                go.copy_(gout.to(torch.float32).transpose(0, 1))
            grads, outs = _grad_table(rt, st, call, ctx.P, need)
            ws = st.workspace(call.which, B)
            xp = call.xT.data_ptr() if call.xT is not None else None
            mp = call.maskT.data_ptr() if call.maskT is not None else None
            sync = rt.grad_sync if rt.last_flat_grad is not None else None
            if sync is not None and call.which in rt.grad_stages and all(need):
                # staged: a group of gradients is final after each stage; its all-reduce (issued by
                # `sync` on the communication stream) overlaps the stages that follow
                h = st.handle(call.which + "_sync")
                flat, offs, n = rt.last_flat_grad, rt.last_flat_offsets, 0
                flat = rt.last_flat_target
                for stage in rt.grad_stages[call.which]:
                    if stage.steps[1] > stage.steps[0]:
                        L.check(
                            lib.ckb_plan_backward(h, stage.steps[0], stage.steps[1], B, xp, call.x_is_float,
                                                  mp, call.mask_rows, call.tensors, grads,
                                                  ctx.arena.data_ptr(), garena.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), 0, stream), "ckb_plan_backward")
                        n += int(lib.ckb_plan_last_launches(h))
                    if stage.ops[1] > stage.ops[0]:
                        L.check(lib.ckb_plan_param_ops(h, stage.ops[0], stage.ops[1], 1, call.tensors,
                                                       grads, stream), "ckb_plan_param_ops")
                        n += int(lib.ckb_plan_last_launches(h))
                    sync(_stage_pieces(rt, stage, flat, offs))
                rt.last_launches = n
            else:
                L.check(
                    lib.ckb_plan_backward(
                        st.handle(call.which), 0, call.n_steps, B, xp, call.x_is_float,
                        mp, call.mask_rows,
                        call.tensors, grads, ctx.arena.data_ptr(), garena.data_ptr(),
                        ws.data_ptr(), ws.numel(), L.RUN_PARAM_OPS | (L.USE_GRAPHS if rt.graphs_for(B) else 0),
                        stream),
                    "ckb_plan_backward",
                )
                rt.last_launches = int(lib.ckb_plan_last_launches(st.handle(call.which)))
                if sync is not None and not rt.needs_batch:
                    sync([rt.last_flat_target])
            if sync is not None and not rt.needs_batch:
                # the reduced gradients are consumed on this stream (autograd accumulates them
                # into .grad right after this function returns)
                if rt.last_flat_target is not rt.last_flat_grad:
                    rt.last_synced_bytes = sync.finish(rt.last_flat_target, rt.last_flat_grad)
                else:
                    rt.last_synced_bytes = sync.finish()
        ctx.arena = None
        return (None, None, None, None, None, *outs)


# ----------------------------------------------------------------------------- per-step timing
def _exec_step_flops(rt: PlanRuntime, es: ExecStep, batch: int) -> tuple[int, int]:
    """Algorithmic floating-point operations (2 per multiply-add of the contraction the reference
    layer states) of the forward and backward launches of one step; the backward of
    a sum-product contraction is two contractions of the same size (input and weight gradients).
    The 3xTF32 kernels execute three tensor-core products per algorithmic product."""
    s = rt.plan.steps[es.out_sid]
    if s.is_input:
        return 0, 0
    F, H, Ki, Ko = s.num_folds, s.arity, s.num_input_units, s.num_output_units
    if s.kind == "tucker":
        red = Ki ** H
    elif s.kind == "sum":
        red = H * Ki
    elif s.kind == "cpt":
        red = Ki
    else:
        return 0, 0
    units = batch
    if es.kind == STEP_TABLE_DENSE:
        first = rt.plan.steps[es.sids[0]]
        units = int(first.config.get("num_categories", first.config.get("num_states", 0)))
    fwd = 2 * F * units * Ko * red
    return fwd, 2 * fwd


def _exec_step_bytes(rt: PlanRuntime, es: ExecStep, batch: int) -> tuple[int, int]:
    """Algorithmic HBM bytes of the forward and of the backward launches of one execution step:
    every tensor the step must read or write counted once, fp32 (see DESIGN.md "Kernels")."""
    plan = rt.plan
    s = plan.steps[es.out_sid]
    F, H, Ki, Ko, B = s.num_folds, s.arity, s.num_input_units, s.num_output_units, batch
    out = F * B * Ko * 4
    p = sum(int(np.prod(q.shape)) for sid in es.sids for q in plan.steps[sid].params.values()) * 4
    if es.kind == STEP_TABLE_DENSE:
        first = plan.steps[es.sids[0]]
        V = int(first.config.get("num_categories", first.config.get("num_states", 0)))
        t2 = F * V * Ko * 4
        x = B * F * 4
        return p + 2 * t2 + x + out, out + x + 3 * t2 + 2 * p
    if s.is_input:
        x = B * F * 4
        return out + x + p, out + x + p  # fwd: table + x -> y ; bwd: g + x -> dT
    inp = F * H * B * Ki * 4
    hg = 1 if s.kind in ("cpt", "hadamard") else H
    gin = F * hg * B * Ki * 4
    fwd = inp + p + out
    bwd = (inp + out if s.kind not in ("hadamard", "kronecker") else 0) + out + gin + 2 * p
    return fwd, bwd


def profile_steps(rt: PlanRuntime, x: Tensor, leaves: Sequence[Tensor], iters: int = 3) -> list[dict]:
    """Times every execution step's forward and backward launch group in isolation with CUDA
    events on the launching stream (bench.py's roofline leg).  Returns one dict per step."""
    P = rt.parameter_tensors(leaves, None)
    st = rt.state(P[0].device)
    lib, dev, plan, lay = st.lib, st.device, rt.plan, rt.layout
    with torch.cuda.device(dev), torch.no_grad():
        stream = torch.cuda.current_stream(dev).cuda_stream
        call = _prepare_call(rt, st, x, None, P, stream)
        B, S = call.B, call.n_steps
        steps = rt.exec_plans[call.which]
        handle = st.handle(call.which)
        grads, keep = _grad_table(rt, st, call, P, [True] * len(P))
        cpx = 2 if rt.is_complex else 1
        arena = torch.empty(cpx * B * lay.arena_units, dtype=torch.float32, device=dev)
        garena = torch.zeros(cpx * B * lay.garena_units, dtype=torch.float32, device=dev)
        O, K = plan.num_outputs, plan.num_output_units
        garena[cpx * B * lay.out_goff : cpx * (B * lay.out_goff + O * B * K) : cpx] = -1.0 / B
        ws = st.workspace(call.which, B)
        xp = call.xT.data_ptr() if call.xT is not None else None

        def fwd(s0, s1, flags):
            L.check(lib.ckb_plan_forward(handle, s0, s1, B, xp, call.x_is_float, None, 0,
                                         call.tensors, arena.data_ptr(), ws.data_ptr(), ws.numel(),
                                         flags, stream), "ckb_plan_forward")
            return int(lib.ckb_plan_last_launches(handle))

        def bwd(s0, s1, flags):
            L.check(lib.ckb_plan_backward(handle, s0, s1, B, xp, call.x_is_float, None, 0,
                                          call.tensors, grads, arena.data_ptr(), garena.data_ptr(),
                                          ws.data_ptr(), ws.numel(), flags, stream), "ckb_plan_backward")
            return int(lib.ckb_plan_last_launches(handle))

        def timed(fn, *a):
            ts, n = [], 0
            for _ in range(iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = fn(*a)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.mean(ts)), n

        fwd(0, S, L.RUN_PARAM_OPS)  # warm everything once
        bwd(0, S, L.RUN_PARAM_OPS)
        res = []
        t, n = timed(fwd, 0, 0, L.RUN_PARAM_OPS)
        pbytes = sum(int(np.prod(l.shape)) * (8 if l.dtype == "complex" else 4) for l in plan.leaves)
        res.append({"step": "param_ops", "kind": "param_ops", "fwd_ms": t, "fwd_launches": n,
                    "fwd_bytes": 2 * pbytes, "bwd_bytes": 3 * pbytes})
        for i, es in enumerate(steps):
            t, n = timed(fwd, i, i + 1, 0)
            fb, bb = (cpx * v for v in _exec_step_bytes(rt, es, B))
            ff, bf = _exec_step_flops(rt, es, B)
            res.append({"step": es.label, "kind": es.label.split(":")[1], "F": plan.steps[es.out_sid].num_folds,
                        "fwd_ms": t, "fwd_launches": n, "fwd_bytes": fb, "bwd_bytes": bb,
                        "fwd_flops": ff, "bwd_flops": bf})
            if es.table_input is not None:
                # The SURVEY formula charges the layer its H input rows; since the gather fusion the
                # launches move the gathered block u instead (forward: table slices + x in, u out and in
                # again, y out; backward: u, y, g in, du out).  Reported next to the formula's bytes.
                s_ = plan.steps[es.out_sid]
                blk = s_.num_folds * B * s_.num_input_units * 4
                res[-1]["fwd_bytes_moved"] = fb - s_.arity * blk + 2 * blk
                res[-1]["bwd_bytes_moved"] = bb - (s_.arity - 1) * blk
        for i in reversed(range(S)):
            t, n = timed(bwd, i, i + 1, 0)
            res[i + 1]["bwd_ms"] = t
            res[i + 1]["bwd_launches"] = n
        t, n = timed(bwd, 0, 0, L.RUN_PARAM_OPS)
        res[0]["bwd_ms"] = t
        res[0]["bwd_launches"] = n
    return res
