"""Layer-by-layer parity of a fixture: CUDA activations of every step vs the float64 oracle."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from helpers import Golden
from cirkit_b200 import B200Circuit
from oracle import OracleCircuit

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "pd32_cp_k4"
g = Golden(name)
x = g.x()
oc = OracleCircuit(g.plan, dtype=torch.float64)
with torch.no_grad():
    for p, v in zip(oc.leaves, g.leaves(torch.float64)):
        p.copy_(v)
    oc(x)
ref = oc.last_outputs
cc = B200Circuit(g.plan, fuse_tables=False)
with torch.no_grad():
    for p, v in zip(cc.leaves, g.leaves(torch.float32)):
        p.copy_(v)
cc = cc.to(dev)
cc.runtime.keep_arena = True
with torch.no_grad():
    cc(x.to(dev))
for sid, s in enumerate(g.plan.steps):
    y = cc.runtime.step_output(sid, x.shape[0]).double().cpu()
    r = ref[sid]
    err = (y - r).abs()
    print(f" step {sid:2d} {s.kind:12s} F={s.num_folds:4d} H={s.arity:2d} Ki={s.num_input_units:3d} Ko={s.num_output_units:3d} "
          f"|y|max {r.abs().max().item():9.2f}  max err {err.max().item():.3e}  ulps {err.max().item() / (r.abs().max().item() * 1.19e-7 + 1e-30):7.1f}")
