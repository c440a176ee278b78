"""Stress: many fwd+bwd passes with the TMA-fed backward, each checked against the register-fed kernel."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib
dev = torch.device("cuda:0")
g = Golden("qt28_cp_k64")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cc = B200Circuit(g.plan, seed=1234).to(dev)
gen = torch.Generator().manual_seed(0)
xs = [torch.randint(0, 256, (B, 784), generator=gen).to(dev) for _ in range(4)]
lib = _lib.load()
def run(flags, x):
    lib.ckb_set_option(1, flags)
    for p in cc.leaves: p.grad = None
    y = cc(x); (-y.sum() / B).backward()
    return [p.grad.clone() for p in cc.leaves]
refs = [run(3 | 512 | 2048, x) for x in xs]
torch.cuda.synchronize()
bad = 0
for it in range(N):
    got = run(3 | 512, xs[it % 4])
    if it % 8 == 7: torch.cuda.synchronize()
    errs = [((a - b).abs().max() / (b.abs().max() + 1e-30)).item() for a, b in zip(got, refs[it % 4])]
    if max(errs) > 1e-3:
        bad += 1
        print("iter", it, "MISMATCH", " ".join(f"{i}:{e:.1e}" for i, e in enumerate(errs)))
torch.cuda.synchronize()
print("done", N, "iterations,", bad, "mismatches")
