"""Timeline of one tile of the tcgen05 MLP kernel (needs build/libicon_dbg.so built with -DICON_TC_TRACE)."""
import sys, os, ctypes
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import icon_b200._C as C
dbg = ctypes.CDLL(os.path.join(root, "build", "libicon_dbg.so"))
for name, (res, args) in C._SIGS.items():
    fn = getattr(dbg, name); fn.restype = res; fn.argtypes = args
C.lib = dbg
import icon_b200.ops as ops
ops.lib = dbg
import torch
from icon_b200 import synthetic as S
dev = torch.device("cuda:0")
pk = ops.pack_mlp(S.mlp_state_dict(13, seed=3), 13, device=dev)
x = torch.randn(1, 13, 1 << 22, device=dev)
for _ in range(3): ops.mlp_only(x, pk)
torch.cuda.synchronize()
out = (ctypes.c_longlong * 256)()
dbg.icon_debug_tc_trace(out)
t = list(out); t0 = t[0]
rel = lambda i: (t[i] - t0) if t[i] else None
print("MMA: tile start 0, x0 ready", rel(1), " acc2free ok", rel(2))
for j in range(8):
    print(f"  L1({j}): begin {rel(10+4*j)}  a0_full {rel(11+4*j)}  w_full {rel(12+4*j)}")
print("  acc1 commit issued", rel(50))
for c in range(4): print(f"  L2 chunk {c}: act1 ready {rel(52+2*c)} w_full {rel(53+2*c)}")
print("  acc2 commit issued", rel(62))
print("worker(row0,h0):")
for j in range(8): print(f"  chunk {j}: acc0_full wake {rel(70+2*j)}  a0_full arrive {rel(71+2*j)}")
print("  acc1_full wake", rel(90), " act1 chunks done", [rel(91+i) for i in range(4)])
print("gather warp: loop top", rel(100), " G(next) done", rel(101), " acc2_full wake", rel(102), " acc2free arrive", rel(103))
