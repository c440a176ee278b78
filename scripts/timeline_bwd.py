import os, sys, ctypes
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch, numpy as np
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib
dev = torch.device("cuda:0")
g = Golden("qt28_cp_k64")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cc = B200Circuit(g.plan, seed=1234, fuse_tables=False).to(dev)
x = torch.randint(0, 256, (B, 784), generator=torch.Generator().manual_seed(0)).to(dev)
lib = _lib.load()
for _ in range(2):
    (-cc(x).mean()).backward()
lib.ckb_set_option(1, 3 | 128)
(-cc(x).mean()).backward()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
lib.ckb_debug_read(buf, 512 * 8)
a = np.array(buf[:], dtype=np.int64)
t0 = a[0]
print("setup done +%d, loop end +%d, exit sync +%d  (clocks; last bwd TC launch = step 1, F=784)" % (a[1]-t0, a[2]-t0, a[3]-t0))
names = ["mma:ready", "mma:issued", "tr:loads_issued", "tr:buffers_free", "tr:written", "epi:T_ready", "epi:stored"]
for it in range(4):
    base = 16 + it * 8
    print("tile", it, " ".join(f"{n}=+{a[base+i]-t0}" for i, n in enumerate(names) if a[base+i] > 0))
