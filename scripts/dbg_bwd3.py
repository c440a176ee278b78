"""bwd3 (TMA-fed) vs the register-fed kernels on the north-star circuit: per-leaf gradient differences,
run-to-run determinism."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib
dev = torch.device("cuda:0")
g = Golden("qt28_cp_k64")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cc = B200Circuit(g.plan, seed=1234).to(dev)
x = torch.randint(0, 256, (B, 784), generator=torch.Generator().manual_seed(0)).to(dev)
lib = _lib.load()
def run(flags):
    lib.ckb_set_option(1, flags)
    for p in cc.leaves: p.grad = None
    y = cc(x); (-y.mean()).backward(); torch.cuda.synchronize()
    return [p.grad.clone() for p in cc.leaves]
ref = run(3 | 512 | 2048 | 1024)   # round-1 kernel
for trial in range(3):
    got = run(3 | 512)
    print("trial", trial, " ".join(f"{i}:{(a-b).abs().max().item():.2e}/{b.abs().max().item():.1e}" for i, (a, b) in enumerate(zip(got, ref))))
got2 = run(3 | 512 | 2048)  # bwd2
print("bwd2   ", " ".join(f"{i}:{(a-b).abs().max().item():.2e}" for i, (a, b) in enumerate(zip(got2, ref))))
# where do the differences sit for the worst leaf?
worst = max(range(len(ref)), key=lambda i: ((got[i]-ref[i]).abs().max() / (ref[i].abs().max() + 1e-30)).item())
d = (got[worst] - ref[worst]).abs()
print("worst leaf", worst, tuple(ref[worst].shape), "folds with diff > 1e-6:", (d.flatten(1).max(1).values > 1e-6).nonzero().flatten().tolist()[:40])
