"""clock64 timeline of the first k-blocks of the Tucker forward kernel (library built with
CKB_NVCC_EXTRA=-DCKB_TIMELINE)."""
import ctypes, dataclasses, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib
from oracle.reference_eval import make_inputs
g = Golden("qt8_tucker_k4")
plan = dataclasses.replace(g.plan, meta={"units": 4}).with_units(64)
dev = torch.device("cuda:0")
cc = B200Circuit(plan, seed=1).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = make_inputs(plan, B, seed=1).to(dev)
lib = _lib.load()
with torch.no_grad():
    for _ in range(2):
        cc(x)
    torch.cuda.synchronize()
    lib.ckb_set_option(1, 3 | 256)
    cc(x)
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
lib.ckb_debug_read(buf, 512 * 8)
a = np.array(buf[:], dtype=np.int64)
t0 = a[16 + 3]
names = ["mma:wait", "mma:ready", "mma:issued", "A:start", "A:free", "A:written", "W:free", "epi:ready", "mma:1st", "mma:3rd", "mma:12th", "W:arrived", "W8:free", "W8:stored", "W8:arrived"]
order = [11, 12, 13, 14, 0, 1, 8, 9, 10, 2, 3, 4, 5, 6, 7]
for kb in range(8, 20):
    base = 16 + kb * 16
    print(f"kb {kb:2d} " + " ".join(f"{names[i]}=+{a[base+i]-t0}" for i in order if a[base + i] > 0))
