#!/usr/bin/env python
"""Benchmark of the hot path: forward + backward log-likelihood of a compiled circuit.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on host cores

A *step* is `ll = circuit(x); (-ll.mean()).backward()` on one batch of synthetic evidence
(`notebooks/learning-a-circuit.ipynb` loop body without the optimiser).  Workload:
QuadTree 28x28, Categorical-256 inputs, CP layers, K=64 (BASELINE.json metric), per-GPU batch 2048
by default.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "samples/sec (fwd+bwd log-lik) QuadTree 28x28 K=64"  # BASELINE.json metric
REF_VERSION = "0.2.1"  # libcirkit, pyproject.toml of the reference checkout
UNIT = "samples/s"
WORKLOADS = {
    "qt28_cp_k64": "QuadTree(quad-tree-2) 28x28, Categorical-256 inputs, CP (Dense+Hadamard), K=64",
    "qt28_cp_k32": "QuadTree(quad-tree-2) 28x28, Categorical-256 inputs, CP (Dense+Hadamard), K=32",
    "qt28_tucker_k64": "QuadTree(quad-tree-2) 28x28, Categorical-256 inputs, Tucker, K=64",
    "pd32_cp_k128": "PoonDomingos 32x32x3, Categorical-256 inputs, CP (fold+optimize), K=128",
    "rbt64_sos_k64": "Sum-of-squares circuit: RandomBinaryTree 64 vars, complex Embedding-256 inputs, CP-T, K=64, "
                     "complex-lse-sum; log p(x) = 2 Re c(x) - Re Z, Z = integrate(c conj(c))",
}
# workloads built by resizing a structure fixture: name -> (fixture, its units, units to run at)
RESIZED = {"pd32_cp_k128": ("pd32_cp_k4", 4, 128)}


def metric_name(workload):
    """BASELINE.json's metric for the default workload; the other workloads (parity / coverage
    configurations of BASELINE.json) are labelled by what they are."""
    if workload == "qt28_cp_k64":
        return METRIC
    return "samples/sec (fwd+bwd log-lik) " + {"qt28_cp_k32": "QuadTree 28x28 K=32 (configs[1])",
                                               "qt28_tucker_k64": "QuadTree 28x28 Tucker K=64 (configs[2])",
                                               "pd32_cp_k128": "PoonDomingos 32x32x3 K=128 (configs[3])",
                                               "rbt64_sos_k64": "squared circuit RBT-64 K=64 complex (configs[4])"}[workload]


class _PlanOnly:
    def __init__(self, plan):
        self.plan = plan


def load_plan(name):
    import dataclasses

    from helpers import Golden

    if name == "rbt64_sos_k64":  # the plan of c(x), lowered from the reference's compiled circuit
        from cirkit_b200.adapter import plan_from_torch

        return _PlanOnly(plan_from_torch(build_squared("torch").c, semirings=("complex-lse-sum",)).plan)

    if name in RESIZED:
        fixture, k0, k = RESIZED[name]
        g = Golden(fixture)
        g.plan = dataclasses.replace(g.plan, meta={"units": k0}).with_units(k)
        return g
    return Golden(name)


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """Dense bf16 tensor throughput the driver measured (sustained: the kernel runs inside a long
    step); kind::tf32 runs at half of it."""
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained; tf32 = half)"
    return 1400.0, "fallback (B200_PROFILING.md sustained bf16; tf32 = half)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8: "hw_slowdown",
            0x40: "hw_thermal_slowdown",
            0x20: "sw_thermal_slowdown",
            0x4: "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=1)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def make_config(args, world, plan):
    """The `config` object of the JSON line -- the same for this repo's arm and the reference arm."""
    B = args.batch
    return {
        "workload": WORKLOADS[args.workload], "batch_per_gpu": B, "global_batch": world * B,
        "x_dtype": "int64", "parallelism": f"dp{world} (batch-sharded replicas)",
        "collectives": "all_gather(root ll)" + ("" if args.no_grad_allreduce or world == 1 else " + all_reduce(param grads, sum)"),
        "l2": f"no flush: per-step working set {plan.algorithmic_bytes(B) / 5 / 1e9:.2f} GB of activations >> 126 MB L2; 4 rotating input batches",
        "leaves": "seeded N(0,1), seed 1234",
        **({"tc_flags": args.tc_flags} if args.tc_flags is not None else {}),
    }


# ------------------------------------------------------------------------------- reference arm
REF_DIR = os.path.join(REPO, "baseline", "_ref")  # pip --target install of the unmodified reference
REF_SPECS = {  # workload -> image_data(shape, region graph, sum-product layer, units)
    "qt28_cp_k64": ((1, 28, 28), "quad-tree-2", "cp", 64),
    "qt28_cp_k32": ((1, 28, 28), "quad-tree-2", "cp", 32),
    "qt28_tucker_k64": ((1, 28, 28), "quad-tree-2", "tucker", 64),
    "pd32_cp_k128": ((3, 32, 32), "poon-domingos", "cp", 128),
}


class SquaredCircuit(torch.nn.Module):
    """log p(x) = 2 Re c(x) - Re Z of a squared circuit (notebooks/sum-of-squares-circuits.ipynb
    cell 32): `c` evaluates the complex log-scores, `z` the batch-free log-partition function of
    c conj(c); both are compiled by ONE PipelineContext and share their parameters."""

    def __init__(self, c, z):
        super().__init__()
        self.c, self.z = c, z

    def forward(self, x):
        return 2.0 * self.c(x).real - self.z().real


def build_squared(backend, units=64):
    """BASELINE.json configs[4] through the reference's own front-end (needs baseline/_ref):
    tabular_data('random-binary-tree', 64 features, complex Embedding(256) inputs, 'cp-t', K units),
    semiring 'complex-lse-sum', Z = integrate(multiply(c, conjugate(c)))."""
    if not os.path.isdir(os.path.join(REF_DIR, "cirkit")):
        raise RuntimeError(f"workload rbt64_sos_k64 is built by the reference's front-end: {REF_DIR} is missing")
    if REF_DIR not in sys.path:
        sys.path.insert(1, REF_DIR)
    import cirkit.symbolic.functional as SF
    from cirkit.pipeline import PipelineContext
    from cirkit.templates import data_modalities, utils

    if backend == "b200":
        import cirkit_b200

        cirkit_b200.register_backend()
    cplx = utils.Parameterization(dtype="complex", initialization="uniform")
    sc = data_modalities.tabular_data(
        "random-binary-tree", num_features=64,
        input_layers={"name": "embedding", "args": {
            "num_states": 256, "weight_factory": utils.parameterization_to_factory(cplx)}},
        num_input_units=units, sum_product_layer="cp-t", num_sum_units=units, sum_weight_param=cplx)
    zsc = SF.integrate(SF.multiply(sc, SF.conjugate(sc)))
    torch.manual_seed(1234)
    ctx = PipelineContext(backend=backend, semiring="complex-lse-sum", fold=True, optimize=True)
    return SquaredCircuit(ctx.compile(sc), ctx.compile(zsc))


def build_reference(workload, plan):
    """The UNMODIFIED reference (april-tools/cirkit, installed under baseline/_ref) compiled through
    its own public API -- `data_modalities.image_data` -> `PipelineContext(backend="torch")` -- for
    `workload`, holding the same seeded parameters as this repo's arm.  Returns (circuit, kind):
    kind "reference", or (OracleCircuit, "port") when the install is missing / cannot be imported."""
    from cirkit_b200.plan import seeded_leaves

    if workload == "rbt64_sos_k64":
        return build_squared("torch"), "reference"
    try:
        if not os.path.isdir(os.path.join(REF_DIR, "cirkit")):
            raise ImportError(f"{REF_DIR} not present")
        if REF_DIR not in sys.path:
            sys.path.insert(1, REF_DIR)
        from cirkit.pipeline import PipelineContext
        from cirkit.templates import data_modalities, utils

        from cirkit_b200.adapter import plan_from_torch

        shape, rg, spl, K = REF_SPECS[workload]
        sc = data_modalities.image_data(
            shape, region_graph=rg, input_layer="categorical", num_input_units=K,
            sum_product_layer=spl, num_sum_units=K,
            sum_weight_param=utils.Parameterization(activation="softmax", initialization="normal"))
        tc = PipelineContext(backend="torch", semiring="lse-sum", fold=True, optimize=True).compile(sc)
        low = plan_from_torch(tc, allow_external_params=False)  # only to name the leaves in plan order
        if [tuple(l.shape) for l in low.plan.leaves] != [tuple(l.shape) for l in plan.leaves]:
            raise RuntimeError("reference circuit and fixture plan disagree on the leaf shapes")
        with torch.no_grad():
            for p, v in zip(low.leaves, seeded_leaves(plan, 1234)):
                p.copy_(v)
        return tc, "reference"
    except Exception as exc:  # pragma: no cover - depends on the box
        sys.stderr.write(f"bench: reference install unusable ({exc!r}); timing the oracle port\n")
        from oracle import OracleCircuit

        oc = OracleCircuit(plan, dtype=torch.float32)
        with torch.no_grad():
            for p, v in zip(oc.leaves, seeded_leaves(plan, 1234)):
                p.copy_(v)
        return oc, "port"


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores: the unmodified
    reference package from baseline/_ref through its public API (fallback: the oracle port, same
    PyTorch op sequence, when the install is absent).  Same workload and batch as this repo's arm."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = load_plan(args.workload)
    oc, kind = build_reference(args.workload, g.plan)
    B = args.cpu_batch
    gen = torch.Generator().manual_seed(0)
    xs = [torch.randint(0, 256, (B, g.plan.num_variables), generator=gen) for _ in range(2)]

    def step(i):
        oc.zero_grad(set_to_none=True)
        ll = oc(xs[i % len(xs)])
        (-ll.mean()).backward()
        return ll

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    line = {
        "impl": "reference",
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c64" if args.workload == "rbt64_sos_k64" else "f32",
        "data": "synthetic",
        "config": make_config(args, world, g.plan),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps of batch {B} on rank 0's host cores "
                                   f"({'cirkit ' + REF_VERSION + ' from baseline/_ref, backend=torch' if kind == 'reference' else 'oracle port, torch CPU ops'}, "
                                   f"fp32, {cores} threads, device cpu)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(g, workload, batch, steps=3, warmup=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oc, kind = build_reference(workload, g.plan)
    x = torch.randint(0, 256, (batch, g.plan.num_variables), generator=torch.Generator().manual_seed(0))
    ts = []
    for i in range(warmup + steps):
        oc.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        ll = oc(x)
        (-ll.mean()).backward()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return {"value": batch / float(np.median(ts)), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"median of {steps} steps of batch {batch} (same circuit and parameters, fp32, "
                      f"{'unmodified reference from baseline/_ref' if kind == 'reference' else 'oracle port'}, "
                      f"{cores} threads)"}


# ------------------------------------------------------------------------------- CUDA arm
def run_b200(args):
    import torch.distributed as dist

    from cirkit_b200 import B200Circuit
    from cirkit_b200.runtime import profile_steps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.tc_flags is not None:
        from cirkit_b200 import _lib

        _lib.check(_lib.load().ckb_set_option(1, args.tc_flags), "ckb_set_option")
    if args.workload == "rbt64_sos_k64":
        cc = build_squared("b200").to(dev)
        assert type(cc.c).__name__ == "B200TorchCircuit", getattr(cc.c, "_b200_reason", "")
        runtime, plan = cc.c._b200_runtime, cc.c._b200_lowered.plan
        z_runtime = getattr(cc.z, "_b200_runtime", None)  # the partition-function circuit (batch-free)
        leaves, prof_leaves = list(cc.parameters()), cc.c._b200_lowered.leaves
        g = _PlanOnly(plan)
    else:
        g = load_plan(args.workload)
        plan = g.plan
        cc = B200Circuit(plan, seed=1234).to(dev)
        runtime, z_runtime = cc.runtime, None
        leaves = prof_leaves = list(cc.leaves)
    B, D = args.batch, plan.num_variables
    gen = torch.Generator().manual_seed(1000 + rank)
    n_batches = 4
    host_x = [torch.randint(0, 256, (B, D), generator=gen, dtype=torch.int64).pin_memory()
              for _ in range(n_batches)]
    dev_x = [h.to(dev) for h in host_x]
    flat_grads = None
    launches = 0

    from cirkit_b200.distributed import BatchShardedCircuit, all_gather_rows_async

    sharded = BatchShardedCircuit(cc)  # this rank's replica: rows [rank*B, (rank+1)*B) of the job
    allreduce = "none" if (world == 1 or args.no_grad_allreduce) else "nccl"
    if allreduce != "none" and z_runtime is None:
        # auto: the in-switch kernel on 8 GPUs (155 us for the 77 MB against NCCL's 312: 1.477 vs 1.511 ms per
        # step); on 2 and 4 GPUs a rank owns a half / a quarter of the buffer and NCCL's staged all-reduces
        # behind the backward pass are faster (N = 4: 1.425 vs 1.473 ms)
        want_nvls = args.allreduce == "nvls" or (args.allreduce == "auto" and world >= 8)
        if want_nvls and sharded.nvls_gradient_sync(num_ctas=args.nvls_ctas, fused=not args.nvls_unfused):
            allreduce = "nvls"
        elif args.allreduce == "nvls":
            raise SystemExit("bench: --allreduce nvls needs NVLink multicast support")
        elif args.grad_chunks > 0:
            sharded.overlap_gradient_sync(args.grad_chunks, chunk_steps=args.chunk_steps)
            allreduce = "nccl-staged"
    args.allreduce_used = allreduce
    if args.stage_only:  # developer A/B: the staged backward pass without any collective
        from cirkit_b200.distributed import OverlappedGradientReducer

        class _NoReduce(OverlappedGradientReducer):
            def __call__(self, pieces):
                self.bytes += sum(t.numel() * 4 for t in pieces)

        runtime.enable_gradient_stages(args.grad_chunks, chunk_steps=args.chunk_steps)
        runtime.grad_sync = _NoReduce()

    def step(x):
        nonlocal launches
        for p in leaves:
            p.grad = None
        ll = cc(x)
        n = runtime.last_launches + (z_runtime.last_launches if z_runtime else 0)
        # the one collective of the data path: all-gather of the root log-densities (runs on the
        # communication stream next to the backward pass) ...
        gathered = all_gather_rows_async(ll, world * B) if world > 1 else None
        loss = -ll.sum() / (world * B)  # this rank's share of the global-batch mean NLL
        loss.backward()
        launches = n + runtime.last_launches + (z_runtime.last_launches if z_runtime else 0)
        if world > 1:
            gathered.wait()
            # ... and, for single-GPU gradient parity, the sum of the replicas' leaf gradients
            # (already reduced stage by stage inside backward() when the overlap is on)
            if not args.no_grad_allreduce:
                sharded.sync_gradients()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        step(dev_x[i % n_batches])
    # ---- device-resident inputs: `value`
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for i in range(args.steps):
        step(dev_x[i % n_batches])
    host_issue_ms = 1e3 * (time.perf_counter() - t_host) / args.steps  # CPU time to enqueue a step
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end: pinned host evidence in, loss out, every step.  The H2D copy of step i+1 runs
    # on a copy stream while step i computes (two device staging buffers, as a prefetching data
    # loader does); every step's input crosses PCIe inside the timed region, K copies for K steps.
    stages = [torch.empty((B, D), dtype=torch.int64, device=dev) for _ in range(2)]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    def e2e_run(n_steps):
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def fetch(i):
            k = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(free[k])  # step i-2 has consumed this buffer
                stages[k].copy_(host_x[i % n_batches], non_blocking=True)
                ready[k].record(copy_stream)

        copy_stream.wait_stream(main)
        fetch(0)
        for i in range(n_steps):
            if i + 1 < n_steps:
                fetch(i + 1)
            main.wait_event(ready[i % 2])
            loss = step(stages[i % 2])
            free[i % 2].record(main)
            loss_host.copy_(loss.detach(), non_blocking=True)

    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(ms2.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0, live, CUDA events around each step's launches)
    prof = profile_steps(runtime, dev_x[0], prof_leaves, iters=5)
    peak, peak_src = peaks()
    # the dominant KERNEL: the step/direction with the largest time per launch (a step that is
    # two kernels, like the fused table+dense one, is not a single roofline point)
    best = None
    for r in prof:
        for d in ("fwd", "bwd"):
            t = r.get(f"{d}_ms")
            if t is None:
                continue
            per_launch = t / max(r.get(f"{d}_launches", 1), 1)
            if best is None or per_launch > best[3]:
                best = (t, r, d, per_launch)
    t, r, d, _ = best
    nbytes = r[f"{d}_bytes"]
    achieved = nbytes / (t * 1e-3) / 1e9
    step_ms_sum = sum(q.get("fwd_ms", 0) + q.get("bwd_ms", 0) for q in prof)
    whole = {
        "algorithmic_bytes": plan.algorithmic_bytes(B),
        "achieved_gbs": plan.algorithmic_bytes(B) / (ms / args.steps * 1e-3) / 1e9,
        "frac": plan.algorithmic_bytes(B) / (ms / args.steps * 1e-3) / 1e9 / peak,
    }
    common = {
        "traffic": ncu_traffic(f"{args.workload} B={B} {r['kind']} {d} F={r.get('F')}"),
        "kernel": f"step {r['step']} {r['kind']} {d} (F={r.get('F')}, {r.get(d + '_launches', 1)} launch)",
        "kernel_ms": t, "kernel_share_of_step": t / step_ms_sum,
        "algorithmic_bytes": nbytes, "whole_step": whole,
    }
    if f"{d}_bytes_moved" in r:
        # this launch reads a fused (pre-summed) input block: fewer bytes than the formula charges it
        common["bytes_moved"] = r[f"{d}_bytes_moved"]
        common["achieved_on_bytes_moved_gbs"] = r[f"{d}_bytes_moved"] / (t * 1e-3) / 1e9
    if r["kind"] == "tucker":
        # the Tucker contraction is tensor-pipe bound: 2*Ko*Ki^2 flop per (fold, sample) forward,
        # twice that backward; the kernels run it as 3xTF32 (three tf32 MMAs per product)
        tf_peak, tf_src = tensor_peak()
        flops = r[f"{d}_flops"]
        tfs = flops / (t * 1e-3) / 1e12
        total_flops = sum(q.get("fwd_flops", 0) + q.get("bwd_flops", 0) for q in prof)
        roofline = {
            "bound": "tensor", "achieved": tfs, "peak": tf_peak, "unit": "TFLOP/s", "frac": tfs / tf_peak,
            "peak_source": tf_src, "algorithmic_flops": flops,
            "tensor_core_flops_executed": 3 * flops, "executed_frac_of_tf32_peak": 3 * tfs / (tf_peak / 2),
            "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src},
            **common,
        }
        roofline["whole_step"]["algorithmic_flops"] = total_flops
        roofline["whole_step"]["achieved_tflops"] = total_flops / (ms / args.steps * 1e-3) / 1e12
    else:
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "peak_source": peak_src, **common}
    if args.profile_out:
        with open(args.profile_out, "w") as fh:
            json.dump(prof, fh, indent=1)
    base = cpu_baseline(g, args.workload, args.cpu_batch) if world == 1 and not args.no_cpu_baseline else None
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c64" if args.workload == "rbt64_sos_k64" else "f32",
        "data": "synthetic",
        "config": make_config(args, world, plan),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * D * 8,
                "d2h_bytes_per_step": 4, "ms_per_step": float(ms2.item()) / args.steps,
                "pipeline": "H2D of step i+1 on a copy stream overlaps step i (2 staging buffers)"},
        "gpu_launches": launches * args.steps,
        "host_issue_ms_per_step": host_issue_ms,  # CPU time to enqueue one step (must stay below ms_per_step)
        # how the gradient sum ran: "nvls" = in-switch multimem kernel (csrc/nvls_allreduce.cu) at the end of
        # backward(), "nccl-staged" = NCCL all-reduces issued per stage of the backward pass, "nccl", "none"
        "allreduce": args.allreduce_used,
        "roofline": roofline,
    }
    if base is not None:
        line["cpu_baseline"] = base
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel `key` names, from
    the committed `ncu --set full` captures (profiles/*_traffic.json, written by
    scripts/ncu_traffic.py); None when that kernel/shape has not been captured."""
    import glob

    for path in sorted(glob.glob(os.path.join(REPO, "profiles", "*_traffic.json")), reverse=True):
        with open(path) as fh:
            table = json.load(fh)
        if key in table:
            return table[key]["dram_bytes"]
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="qt28_cp_k64", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None,
                    help="per-GPU batch (default 2048; 512 for pd32_cp_k128 = configs[3]: 4096 over 8 GPUs)")
    ap.add_argument("--cpu-batch", type=int, default=None,
                    help="batch of the CPU arm (default: the per-GPU batch, i.e. the same config; "
                         "256 for the Tucker workload and 32 for pd32_cp_k128, whose full batch takes minutes per step)")
    ap.add_argument("--no-grad-allreduce", action="store_true")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nvls", "nccl"],
                    help="N > 1: how the replicas' gradients are summed (auto: the in-switch NVLS kernel when the "
                         "GPUs support NVLink multicast, else NCCL)")
    ap.add_argument("--nvls-ctas", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--nvls-unfused", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--grad-chunks", type=int, default=1,
                    help="N > 1: fold chunks of the input table in the staged backward pass whose "
                         "gradient all-reduces overlap the remaining backward work (0: one all-reduce after backward)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--chunk-steps", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--profile-out", default=None)
    ap.add_argument("--tc-flags", type=int, default=None,
                    help="developer switch: value for CKB_OPT_TC_FAST_MATH (3 = default kernels; "
                         "3|512 also routes K=128 layers to the experimental tcgen05 kernels)")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {"pd32_cp_k128": 512, "rbt64_sos_k64": 1024}.get(args.workload, 2048)
    if args.cpu_batch is None:
        args.cpu_batch = {"pd32_cp_k128": 32, "qt28_tucker_k64": 256}.get(args.workload, args.batch)
    if args.impl == "reference":
        run_reference(args)  # bounded sample: --cpu-batch samples per step
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
