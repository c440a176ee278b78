"""clock64 timeline of stages 8.. of the Tucker backward-dX kernel (library built with
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
for _ in range(2):
    (-cc(x).mean()).backward()
torch.cuda.synchronize()
lib.ckb_set_option(1, 3 | 256)
(-cc(x).mean()).backward()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
lib.ckb_debug_read(buf, 1024 * 8)
a = np.array(buf[:], dtype=np.int64)
names = ["iss:start", "iss:tempty", "iss:wfull", "iss:mma_done", "iss:committed", "w0:epi_wait", "w0:T_ready",
         "w0:arrived", "w0:stored", "w8:epi_wait", "w8:T_ready", "w8:arrived", "w8:stored"]
t0 = a[512 + 5]
for i in range(8, 20):
    base = 512 + (i - 8) * 16
    print(f"i {i:2d} " + " ".join(f"{n}=+{a[base+k]-t0}" for k, n in enumerate(names) if a[base + k] > 0))

print("fwd kernel (last launch, CTA 0): start 0, setup +%d, prologue done +%d, last chunk ready +%d, all warps done +%d, end +%d" % tuple(a[k] - a[0] for k in (1, 2, 3, 4, 5)))
print("dX kernel (last launch, CTA 0): start 0, setup +%d, prologue done +%d, r in TMEM +%d, loop done +%d, end +%d" % tuple(a[k] - a[8] for k in (9, 10, 11, 12, 13)))

