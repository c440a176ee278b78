"""Layer-by-layer parity: CUDA activations of every step vs the float64 oracle (plain plan)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib
from oracle import OracleCircuit

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "qt28_cp_k64"
g = Golden(name)
x = g.x()
oc = OracleCircuit(g.plan, dtype=torch.float64)
with torch.no_grad():
    for p, v in zip(oc.leaves, g.leaves(torch.float64)):
        p.copy_(v)
    oc(x)
ref = oc.last_outputs
for tc, fm in ((1, 0), (1, 1), (1, 2), (1, 3), (0, 0)):
    _lib.load().ckb_set_option(_lib.OPT_TENSOR_CORES, tc)
    _lib.load().ckb_set_option(1, fm)
    cc = B200Circuit(g.plan, fuse_tables=False)
    with torch.no_grad():
        for p, v in zip(cc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    cc = cc.to(dev)
    cc.runtime.keep_arena = True
    with torch.no_grad():
        cc(x.to(dev))
    print(f"== {name} tensor_cores={tc} fast_math={fm}")
    for sid, s in enumerate(g.plan.steps):
        if sid not in (1, 2, 6, 11): continue
        y = cc.runtime.step_output(sid, x.shape[0]).double().cpu()
        r = ref[sid]
        err = (y - r).abs()
        spread_ref = (r.max(dim=-1).values - r.min(dim=-1).values).abs().max().item()
        spread = (y.max(dim=-1).values - y.min(dim=-1).values).abs().max().item()
        print(f" step {sid:2d} {s.kind:12s} F={s.num_folds:4d} |y|max {r.abs().max().item():9.2f}  max err {err.max().item():.3e}"
              f"  mean signed err {(y - r).mean().item():+.3e}  unit spread ref {spread_ref:.3e} got {spread:.3e}")
_lib.load().ckb_set_option(_lib.OPT_TENSOR_CORES, 1)
