"""Layer-by-layer report of the Tucker tcgen05 kernels against the float64 oracle (GPU box).
usage: python tests/tools/debug_tucker.py [batch ...]"""
import dataclasses
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from helpers import Golden  # noqa: E402

from cirkit_b200 import B200Circuit  # noqa: E402
from cirkit_b200.plan import seeded_leaves  # noqa: E402
from oracle import OracleCircuit  # noqa: E402
from oracle.reference_eval import make_inputs  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [300]
    g = Golden("qt8_tucker_k4")
    plan = dataclasses.replace(g.plan, meta={"units": 4}).with_units(64)
    dev = torch.device("cuda:0")
    cc = B200Circuit(plan, seed=1).to(dev)
    cc.runtime.keep_arena = True
    oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for q, v in zip(oc.leaves, seeded_leaves(plan, 1)):
            q.copy_(v)
    for B in batches:
        x = make_inputs(plan, B, seed=B)
        for p in cc.leaves:
            p.grad = None
        for q in oc.leaves:
            q.grad = None
        y = cc(x.to(dev))
        torch.cuda.synchronize()
        yo = oc(x)
        print(f"B={B}: root err {(y.detach().double().cpu() - yo.detach()).abs().max().item():.3e} "
              f"(|ll| {yo.abs().max().item():.1f})")
        for sid, s in enumerate(plan.steps):
            got = cc.runtime.step_output(sid, B).double().cpu()
            ref = oc.last_outputs[sid].detach()
            err = (got - ref).abs()
            print(f"  step {sid} {s.kind:12s} F={s.num_folds:3d} Ko={s.num_output_units:3d} "
                  f"max err {err.max().item():.3e} mean err {(got - ref).mean().item():+.3e} "
                  f"finite {bool(torch.isfinite(got).all())}")
        w = torch.randn(B, 1, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
        (y * w.to(dev, torch.float32)).sum().backward()
        torch.cuda.synchronize()
        (yo * w).sum().backward()
        for i, (p, q) in enumerate(zip(cc.leaves, oc.leaves)):
            gr = q.grad
            got = p.grad.double().cpu()
            err = (got - gr).abs().max().item()
            print(f"  leaf {i} {tuple(p.shape)} grad err {err:.3e} (max|g| {gr.abs().max().item():.3e}, "
                  f"rel {err / max(gr.abs().max().item(), 1e-30):.2e}) finite {bool(torch.isfinite(got).all())}")


if __name__ == "__main__":
    main()
