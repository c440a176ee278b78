"""Phase timeline of dense_tc_bwd2_kernel (needs a build with CKB_NVCC_EXTRA=-DCKB_TIMELINE):
clock64() stamps of warps 0 and 9 of the middle CTA of the F = 392 launch."""
import os, sys, ctypes
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch, numpy as np
from helpers import Golden
from cirkit_b200 import B200Circuit, _lib
from cirkit_b200.runtime import profile_steps, _prepare_call, _grad_table
dev = torch.device("cuda:0")
g = Golden("qt28_cp_k64")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
step = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # exec step index (fused plan): 1 = F=392 cpt
cc = B200Circuit(g.plan, seed=1234).to(dev)
x = torch.randint(0, 256, (B, 784), generator=torch.Generator().manual_seed(0)).to(dev)
lib = _lib.load()
for _ in range(2):
    (-cc(x).mean()).backward()
rt = cc.runtime
P = rt.parameter_tensors(list(cc.leaves), None)
st = rt.state(P[0].device)
import ctypes as C
with torch.no_grad():
    stream = torch.cuda.current_stream(dev).cuda_stream
    call = _prepare_call(rt, st, x, None, P, stream)
    lay = rt.layout
    grads, keep = _grad_table(rt, st, call, P, [True] * len(P))
    arena = torch.empty(B * lay.arena_units, dtype=torch.float32, device=dev)
    garena = torch.zeros(B * lay.garena_units, dtype=torch.float32, device=dev)
    garena[B * lay.out_goff : B * lay.out_goff + B] = -1.0 / B
    ws = st.workspace(call.which, B)
    h = st.handle(call.which)
    S = call.n_steps
    lib.ckb_plan_forward(h, 0, S, B, call.xT.data_ptr(), 0, None, 0, call.tensors, arena.data_ptr(), ws.data_ptr(), ws.numel(), 1, stream)
    lib.ckb_plan_backward(h, 0, S, B, call.xT.data_ptr(), 0, None, 0, call.tensors, grads, arena.data_ptr(), garena.data_ptr(), ws.data_ptr(), ws.numel(), 1, stream)
    torch.cuda.synchronize()
    lib.ckb_set_option(1, 3 | 512 | 128 | (int(sys.argv[3]) if len(sys.argv) > 3 else 0))
    lib.ckb_plan_backward(h, step, step + 1, B, call.xT.data_ptr(), 0, None, 0, call.tensors, grads, arena.data_ptr(), garena.data_ptr(), ws.data_ptr(), ws.numel(), 0, stream)
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
lib.ckb_debug_read(buf, 512 * 8)
a = np.array(buf[:], dtype=np.int64)
t0 = a[0]
print("loop end +%d, exit +%d" % (a[1] - t0, a[2] - t0))
names = ["top", "loads_landed", "pair_bar", "math", "ab_empty", "arrived", "next_loads", "all_arrived", "gemm1_issued", "d1_full", "tmem_ld", "du_stored"]
for w, base in ((0, 16), (9, 272)):
    print("warp", w)
    for it in range(2, 8):
        row = a[base + it * 16: base + it * 16 + 12]
        if row[0] == 0:
            break
        prev = row[0]
        out = [f"tile {it:2d} top=+{row[0]-t0:7d}"]
        for i in range(1, 12):
            if row[i] > 0:
                out.append(f"{names[i]}+{row[i]-prev}")
                prev = row[i]
        print("  " + " ".join(out))
