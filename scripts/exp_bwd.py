import os, sys, json
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
for flags, label in ((3, "baseline"), (3 | 32, "no du stores"), (3 | 64, "no loads"), (3 | 8, "no math"), (3 | 8 | 32, "no math, no stores"), (3 | 32 | 64, "no stores/loads")):
    lib.ckb_set_option(1, flags)
    prof = profile_steps(cc.runtime, x, list(cc.leaves), iters=3)
    print(f"{label:24s}", " ".join(f"{r['step'].split(':')[0]}:{r['bwd_ms']*1000:.0f}" for r in prof[1:6]), " | fwd", " ".join(f"{r['fwd_ms']*1000:.0f}" for r in prof[1:6]))
