"""Functional + timing check of a K=128 circuit (the unit count of BASELINE.json configs[3]) on the
QuadTree 28x28 CP structure: the FP32 SIMT route, since the tcgen05 kernels exist for K=64 only."""
import dataclasses, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from helpers import Golden
from cirkit_b200 import B200Circuit
dev = torch.device("cuda:0")
g = Golden("qt28_cp_k64")
plan = dataclasses.replace(g.plan, meta={"units": 64}).with_units(128)
cc = B200Circuit(plan, seed=3).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.randint(0, 256, (B, 784), generator=torch.Generator().manual_seed(0)).to(dev)
for _ in range(3):
    for p in cc.leaves: p.grad = None
    y = cc(x); (-y.mean()).backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n):
    for p in cc.leaves: p.grad = None
    y = cc(x); (-y.mean()).backward()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"K=128 B={B}: {ms:.2f} ms/step = {B / ms * 1e3:.0f} samples/s, ll mean {y.mean().item():.3f}, finite {bool(torch.isfinite(y).all())}, "
      f"grads finite {all(bool(torch.isfinite(p.grad).all()) for p in cc.leaves)}")
