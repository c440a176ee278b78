"""Where does the HOST time of a training step go?  cProfile over the step loop of bench.py's
workload (developer tool; run on the GPU box)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import load_plan  # noqa: E402

from cirkit_b200 import B200Circuit  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "qt28_cp_k64"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
g = load_plan(wl)
dev = torch.device("cuda:0")
cc = B200Circuit(g.plan, seed=1234).to(dev)
x = torch.randint(0, 256, (B, g.plan.num_variables), device=dev)
leaves = list(cc.leaves)


def step():
    for p in leaves:
        p.grad = None
    ll = cc(x)
    loss = -ll.sum() / B
    loss.backward()


for _ in range(10):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{wl} B={B}: host issue {1e3 * (t1 - t0) / 200:.3f} ms/step, total {1e3 * (t2 - t0) / 200:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
