"""Stress with allocator perturbation: arenas land at different addresses from iteration to iteration."""
import os, sys, random, ctypes
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib
dev = torch.device("cuda:0")
g = Golden("qt28_cp_k64")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
perturb = 0
extra = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ref_flags = int(sys.argv[4]) if len(sys.argv) > 4 else (3 | 512 | 2048)
cc = B200Circuit(g.plan, seed=1234).to(dev)
gen = torch.Generator().manual_seed(0)
xs = [torch.randint(0, 256, (B, 784), generator=gen).to(dev) for _ in range(4)]
lib = _lib.load()
def run(flags, x):
    lib.ckb_set_option(1, flags)
    for p in cc.leaves: p.grad = None
    y = cc(x); (-y.sum() / B).backward()
    return [p.grad.clone() for p in cc.leaves]
refs = [run(ref_flags, x) for x in xs]
torch.cuda.synchronize()
bad = 0
keep = []
random.seed(1)
for it in range(N):
    if perturb:
        keep.append(torch.empty(random.randint(1, 64) * 1024 * 1024, dtype=torch.uint8, device=dev))
        if len(keep) > 6: keep.pop(random.randrange(len(keep)))
        if it % 5 == 0: torch.cuda.empty_cache()
    got = run(3 | 512 | extra, xs[it % 4])
    errs = [((a - b).abs().max() / (b.abs().max() + 1e-30)).item() for a, b in zip(got[:5], refs[it % 4][:5])]
    if max(errs) > 1e-4:
        bad += 1
        for li in range(5):
            d = (got[li] - refs[it % 4][li]).abs().flatten(1).max(1).values
            scale = refs[it % 4][li].abs().max()
            badf = (d > 1e-4 * scale).nonzero().flatten().tolist()
            print(f"   leaf {li}: {len(badf)} of {d.numel()} folds differ:", badf[:24])
        print("iter", it, "MISMATCH", " ".join(f"{i}:{e:.1e}" for i, e in enumerate(errs)))
print("done", N, "iterations,", bad, "mismatches (leaves 0-4), flags", extra)
import ctypes
buf = (ctypes.c_ulonglong * 8)()
lib.ckb_set_option(1, 3 | 512 | 4096)
lib.ckb_debug_read(buf, 64)
print("ring mismatches x/y/g:", buf[0], buf[1], buf[2], "tiles checked:", buf[7])
