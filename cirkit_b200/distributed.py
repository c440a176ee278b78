"""Batch-sharded evaluation of one circuit on several GPUs.

Every operation of a circuit is independent across samples, so the multi-GPU scheme is plain
replication: each rank (one process per GPU, `torch.distributed`) holds the same plan and the same
parameters, evaluates a contiguous block of the rows of `x`, and the only exchange on the data
path is ONE all-gather of the root log-densities (4 bytes per sample and output).  For training,
the leaf gradients of the replicas are summed with one all-reduce per parameter tensor so that
every rank ends up with the gradient of the global-batch loss (single-GPU parity).

The reference has no multi-device support (it evaluates wherever its tensors live,
cirkit/backend/torch/circuits.py:242-278); this module is what a user of it would otherwise write
around `DistributedDataParallel`.  It is backend-agnostic (nccl on GPUs, gloo in the CPU tests) and
wraps any module mapping x:(B, D) -> (B, O, K).
"""

from __future__ import annotations

from typing import Iterable

import torch
import torch.distributed as dist
from torch import Tensor, nn


def shard_rows(num_rows: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous row block [begin, end) of `rank`: the first `num_rows % world_size` ranks get
    one extra row, so blocks differ by at most one row and concatenate in rank order."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    base, extra = divmod(num_rows, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _world(group) -> tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


class _GatheredRows:
    """Handle of an all-gather in flight: `wait()` returns the global tensor."""

    def __init__(self, work, finish):
        self._work, self._finish = work, finish

    def wait(self) -> Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._finish()


def all_gather_rows(local: Tensor, num_rows: int, group=None) -> Tensor:
    """Concatenate the row blocks of all ranks (the blocks of `shard_rows(num_rows, ...)`) into
    the global (num_rows, ...) tensor, on every rank.  One collective; blocks that are one row
    short are padded for the exchange and trimmed afterwards."""
    return all_gather_rows_async(local, num_rows, group).wait()


def all_gather_rows_async(local: Tensor, num_rows: int, group=None) -> _GatheredRows:
    """`all_gather_rows` issued asynchronously (on the process group's communication stream for
    NCCL): work enqueued afterwards, e.g. the backward pass, does not wait for the exchange."""
    world, rank = _world(group)
    if world == 1:
        if local.shape[0] != num_rows:
            raise ValueError(f"expected {num_rows} rows, got {local.shape[0]}")
        return _GatheredRows(None, lambda: local)
    begin, end = shard_rows(num_rows, world, rank)
    if local.shape[0] != end - begin:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, its block has {end - begin}")
    rows_max = -(-num_rows // world)
    tail = local.shape[1:]
    send = local.detach().contiguous()
    if send.shape[0] < rows_max:  # uneven split: pad to the common block size
        send = torch.cat([send, send.new_zeros((rows_max - send.shape[0], *tail))])
    recv = send.new_empty((world * rows_max, *tail))
    work = dist.all_gather_into_tensor(recv, send, group=group, async_op=True)

    def finish() -> Tensor:
        if num_rows == world * rows_max:
            return recv
        blocks = []
        for r in range(world):
            b, e = shard_rows(num_rows, world, r)
            blocks.append(recv[r * rows_max : r * rows_max + (e - b)])
        return torch.cat(blocks)

    return _GatheredRows(work, finish)


def _flat_gradient(params: list, flat: Tensor | None) -> Tensor | None:
    """`flat` if every gradient is a piece of that one buffer and together they fill it (the CUDA
    runtime returns its parameter gradients as views of one flat tensor, `runtime.last_flat_grad`):
    one collective instead of one per parameter tensor."""
    if flat is None or flat.numel() == 0:
        return None
    lo = flat.data_ptr()
    hi = lo + flat.numel() * flat.element_size()
    covered, n = 0, 0
    for p in params:
        if not p.requires_grad:
            continue
        g = p.grad
        if g is None or g.dtype != flat.dtype or g.device != flat.device or not g.is_contiguous():
            return None
        if not (lo <= g.data_ptr() and g.data_ptr() + g.numel() * g.element_size() <= hi):
            return None
        covered += g.numel()
        n += 1
    if covered + 4 * n < flat.numel():
        return None  # the buffer holds more than these parameters' gradients
    return flat


def all_reduce_gradients(params: Iterable[Tensor], *, average: bool = False, group=None,
                         flat: Tensor | None = None) -> int:
    """Sum (or average) `.grad` of every parameter over the ranks, in place.  Parameters without
    a gradient on this rank contribute zeros, so all ranks issue the same collectives.  Returns
    the number of bytes reduced."""
    world, _ = _world(group)
    total = 0
    params = list(params)
    flat = _flat_gradient(params, flat)
    if flat is not None:
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat.div_(world)
        return flat.numel() * flat.element_size()
    for p in params:
        if not p.requires_grad:
            continue
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        total += p.grad.numel() * p.grad.element_size()
        if world > 1:
            # complex gradients travel as (re, im) float pairs (NCCL has no complex types)
            g = torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            if average:
                p.grad.div_(world)
    return total


class OverlappedGradientReducer:
    """Sums gradient pieces over the ranks while the backward pass is still running.

    The CUDA runtime calls it once per stage of its staged backward pass
    (`PlanRuntime.enable_gradient_stages`) with the slices of the flat gradient buffer that have
    just become final.  Each slice goes out as an asynchronous all-reduce: the process group's
    communication stream waits for the kernels enqueued so far on the compute stream and runs
    next to the stages that follow; `finish()` (end of the backward pass) makes the compute stream
    wait for all of them.  Only the last group's collective is exposed."""

    def __init__(self, group=None, average: bool = False) -> None:
        self.group = group
        self.average = average
        self._works: list = []
        self._pieces: list[Tensor] = []
        self.bytes = 0

    def __call__(self, pieces: Iterable[Tensor]) -> None:
        world, _ = _world(self.group)
        for t in pieces:
            if t.numel() == 0:
                continue
            self.bytes += t.numel() * t.element_size()
            if world > 1:
                self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                self._pieces.append(t)

    def finish(self, target: Tensor | None = None, flat: Tensor | None = None) -> int:
        world, _ = _world(self.group)
        for w in self._works:
            w.wait()
        if self.average and world > 1:
            for t in self._pieces:
                t.div_(world)
        self._works.clear()
        self._pieces.clear()
        n, self.bytes = self.bytes, 0
        return n


class NvlsGradientReducer:
    """Sums the replicas' gradients in the NVSwitch (`csrc/nvls_allreduce.cu`) instead of calling
    NCCL.  The CUDA runtime writes its flat gradient buffer straight into a symmetric-memory
    buffer this object owns (`alloc`, allocated once and exchanged between the ranks through
    `torch.distributed._symmetric_memory`: plumbing); at the end of the backward pass `finish`
    orders the ranks with a barrier on the stream, launches the in-place two-shot kernel (rank r
    reduces its 1/N share with `multimem.ld_reduce` and broadcasts it with `multimem.st`), a second
    barrier, and copies the result into the tensors autograd hands out -- those must not alias a
    buffer that the next backward pass overwrites.  Every replica ends up with the same bits."""

    def __init__(self, group=None, average: bool = False, num_ctas: int = 0, fused: bool = True) -> None:
        self.group = group
        self.average = average
        self.num_ctas = num_ctas
        self.fused = fused  # barriers and copy-out inside the reduction kernel (ckb_nvls_allreduce_fused)
        self._epoch = 0
        self._counter: Tensor | None = None
        self._buf: Tensor | None = None
        self._hdl = None
        self.bytes = 0

    @staticmethod
    def available(device: torch.device) -> bool:
        try:
            from torch._C._distributed_c10d import _SymmetricMemory

            idx = device.index if device.index is not None else torch.cuda.current_device()
            return bool(_SymmetricMemory.has_multicast_support(torch._C._autograd.DeviceType.CUDA, idx))
        except Exception:  # pragma: no cover - depends on the build
            return False

    def alloc(self, numel: int, device: torch.device) -> Tensor | None:
        world, _ = _world(self.group)
        if world == 1:
            return None
        n4 = -(-numel // 4) * 4
        if self._buf is None or self._buf.numel() < n4 or self._buf.device != device:
            import torch.distributed._symmetric_memory as symm

            buf = symm.empty(n4, dtype=torch.float32, device=device)
            hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            if not hdl.multicast_ptr:
                raise RuntimeError("NvlsGradientReducer: this group has no NVLink multicast (NVLS) support")
            if self.fused and (symm.get_signal_pad_size() < 1536 or hdl.world_size > 64):
                self.fused = False  # the fused kernel keeps its flags at words 256..383 of the signal pads
            self._buf, self._hdl = buf, hdl
            self._counter = torch.zeros(4, dtype=torch.int32, device=device)
            self._epoch = 0
            hdl.barrier(channel=0)  # (also: nobody's first fused call can run ahead of a peer's rendezvous)
        return self._buf[:numel]

    def __call__(self, pieces: Iterable[Tensor]) -> None:
        self.bytes += sum(t.numel() * t.element_size() for t in pieces)

    def finish(self, target: Tensor | None = None, flat: Tensor | None = None) -> int:
        n, self.bytes = self.bytes, 0
        if target is None or flat is None:
            return n  # world 1: nothing to exchange
        from . import _lib as L

        hdl, buf = self._hdl, self._buf
        world, rank = hdl.world_size, hdl.rank
        stream = torch.cuda.current_stream(buf.device).cuda_stream
        n4 = -(-target.numel() // 4) * 4
        mc = hdl.multicast_ptr + (buf.data_ptr() - hdl.buffer_ptrs[rank])
        lib = L.load()
        if self.fused and not self.average and flat.numel() == n4:
            # one kernel: wait for the peers, reduce + broadcast, wait again, copy out
            self._epoch += 1
            L.check(lib.ckb_nvls_allreduce_fused(mc, buf.data_ptr(), flat.data_ptr(), n4, rank, world,
                                                 hdl.signal_pad_ptrs_dev, self._epoch, self._counter.data_ptr(),
                                                 self.num_ctas, stream), "ckb_nvls_allreduce_fused")
            return n
        hdl.barrier(channel=0)  # every replica's backward pass has written its gradients
        L.check(lib.ckb_nvls_allreduce(mc, n4, rank, world, self.num_ctas, stream), "ckb_nvls_allreduce")
        hdl.barrier(channel=1)  # every share has been broadcast
        if self.average:
            torch.mul(target, 1.0 / world, out=flat)
        else:
            flat.copy_(target)
        return n


def _runtime_of(circuit):
    return getattr(circuit, "runtime", None) or getattr(circuit, "_b200_runtime", None)


class BatchShardedCircuit(nn.Module):
    """A replica of `circuit` that evaluates this rank's rows of a global batch.

    forward(x_local)            -> local root log-densities (B_local, O, K), differentiable
    log_likelihoods(x_local, B) -> all-gathered (B, O, K) on every rank (no gradient)
    loss(x_local, B)            -> -sum(ll_local) / B: summing the gradients of this over the
                                   ranks (sync_gradients) gives the gradient of the mean negative
                                   log-likelihood of the global batch
    """

    def __init__(self, circuit: nn.Module, group=None) -> None:
        super().__init__()
        self.circuit = circuit
        self.group = group

    @property
    def world_size(self) -> int:
        return _world(self.group)[0]

    @property
    def rank(self) -> int:
        return _world(self.group)[1]

    def local_rows(self, num_rows: int) -> tuple[int, int]:
        return shard_rows(num_rows, self.world_size, self.rank)

    def shard(self, x_global: Tensor) -> Tensor:
        begin, end = self.local_rows(x_global.shape[0])
        return x_global[begin:end]

    def forward(self, x_local: Tensor) -> Tensor:
        return self.circuit(x_local)

    def log_likelihoods(self, x_local: Tensor, num_rows: int) -> Tensor:
        with torch.no_grad():
            ll = self.circuit(x_local)
        return all_gather_rows(ll, num_rows, self.group)

    def loss(self, x_local: Tensor, num_rows: int) -> Tensor:
        return -self.circuit(x_local).sum() / num_rows

    def nvls_gradient_sync(self, *, average: bool = False, num_ctas: int = 0, fused: bool = True) -> bool:
        """Sum the parameter gradients in the NVSwitch at the end of the backward pass (see
        `NvlsGradientReducer`).  Returns False -- and changes nothing -- without a CUDA runtime,
        without NVLink multicast support, or for plans with per-sample PyTorch inputs."""
        rt = _runtime_of(self.circuit)
        if rt is None or not hasattr(rt, "grad_sync") or rt.needs_batch or rt.is_complex:
            return False
        dev = next(self.circuit.parameters()).device
        ok = dev.type == "cuda" and NvlsGradientReducer.available(dev)
        red = None
        if ok:
            # Dry run on a small buffer: the symmetric-memory rendezvous and the multicast mapping are
            # the parts that can fail on a given box (driver, fabric manager, container limits).
            try:
                red = NvlsGradientReducer(self.group, average, num_ctas, fused)
                n = 4096
                buf = red.alloc(n, dev)
                if buf is not None:
                    buf.fill_(1.0)
                    out = torch.empty(n, dtype=torch.float32, device=dev)
                    red.finish(buf, out)
                    ok = bool((out == (1.0 if average else float(self.world_size))).all())
            except Exception:  # pragma: no cover - depends on the box
                ok = False
        # every rank must take the same path (the collectives differ)
        if self.world_size > 1 and dev.type == "cuda":
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            ok = bool(flag.item())
        if not ok:
            return False
        rt.grad_sync = red
        return True

    def overlap_gradient_sync(self, chunks: int = 4, *, average: bool = False,
                              bucket_bytes: int = 8 << 20, chunk_steps: bool = False) -> bool:
        """Sum the parameter gradients over the ranks INSIDE the backward pass, stage by stage
        (see `OverlappedGradientReducer`); `sync_gradients()` then only reports the bytes.  Returns
        False when the circuit has no CUDA runtime or its plan has per-sample inputs from PyTorch
        layers (external steps), in which case `sync_gradients()` keeps doing the collective."""
        rt = _runtime_of(self.circuit)
        if rt is None or not hasattr(rt, "enable_gradient_stages") or rt.needs_batch or rt.is_complex:
            return False
        rt.enable_gradient_stages(chunks, bucket_bytes, chunk_steps)
        rt.grad_sync = OverlappedGradientReducer(self.group, average)
        return True

    def sync_gradients(self, *, average: bool = False) -> int:
        rt = _runtime_of(self.circuit)
        if rt is not None and getattr(rt, "grad_sync", None) is not None and rt.last_synced_bytes:
            if average != rt.grad_sync.average:
                raise ValueError("overlap_gradient_sync was set up with a different `average`")
            return rt.last_synced_bytes  # already reduced inside the backward pass
        flat = getattr(rt, "last_flat_grad", None)
        return all_reduce_gradients(self.circuit.parameters(), average=average, group=self.group, flat=flat)

    def broadcast_parameters(self, src: int = 0) -> None:
        """Make every replica start from rank `src`'s parameters."""
        if self.world_size > 1:
            for p in self.circuit.parameters():
                dist.broadcast(p.data, src=src, group=self.group)
            rt = _runtime_of(self.circuit)
            if rt is not None and hasattr(rt, "invalidate_parameter_cache"):
                rt.invalidate_parameter_cache()  # `.data` writes do not bump the version counter
