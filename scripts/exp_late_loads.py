"""A/B: dense_tc_bwd with the next tile's loads issued before (default) or after the proxy fence."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib
from cirkit_b200.runtime import profile_steps
dev = torch.device("cuda:0")
g = Golden("qt28_cp_k64")
cc = B200Circuit(g.plan, seed=1234).to(dev)
x = torch.randint(0, 256, (2048, 784), generator=torch.Generator().manual_seed(0)).to(dev)
lib = _lib.load()
for flags in (3, 3 | 16, 3, 3 | 16):
    lib.ckb_set_option(1, flags)
    prof = profile_steps(cc.runtime, x, list(cc.leaves), iters=10)
    tot_b = sum(r["bwd_ms"] for r in prof if r["kind"] == "cpt")
    print(f"flags {flags}: cpt bwd total {tot_b:.3f} ms; " + " ".join(f"F{r['F']}={r['bwd_ms']:.3f}" for r in prof if r["kind"] == "cpt" and r["F"] >= 49))
