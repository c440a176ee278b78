"""Print forward / gradient errors of the CUDA path against the float64 reference fixtures."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib

dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["qt28_cp_k64", "qt28_cp_k32"]:
    g = Golden(name)
    for tc in (1, 0):
        _lib.load().ckb_set_option(_lib.OPT_TENSOR_CORES, tc)
        cc = B200Circuit(g.plan)
        with torch.no_grad():
            for p, v in zip(cc.leaves, g.leaves(torch.float32)):
                p.copy_(v)
        cc = cc.to(dev)
        y = cc(g.x().to(dev))
        (-y.mean()).backward()
        ferr = (y.detach().double().cpu() - g.y()).abs().max().item()
        print(f"{name} tc={tc}: forward max err {ferr:.3e} (|ll| {g.y().abs().max().item():.1f})")
        for i, p in enumerate(cc.leaves):
            flat = p.grad.double().cpu().reshape(-1)
            gsum, gabs, gmax = g.z[f"gsum_{i}"]
            probe = torch.from_numpy(g.z[f"gval_{i}"]); idx = torch.from_numpy(g.z[f"gidx_{i}"])
            e = (flat[idx] - probe).abs().max().item()
            print(f"   leaf {i:2d} shape {tuple(p.shape)}: probe err {e:.3e}  max|g| {gmax:.3e}  tol {max(2e-6, 1e-4*gmax):.1e}  abs-sum rel {abs(flat.abs().sum().item()-gabs)/gabs:.2e}")
_lib.load().ckb_set_option(_lib.OPT_TENSOR_CORES, 1)
