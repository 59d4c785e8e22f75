import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icon_b200._C as C
dbg = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "libicon_dbg.so"))
# rebind the ops to the debug library
for name, (res, args) in C._SIGS.items():
    fn = getattr(dbg, name); fn.restype = res; fn.argtypes = args
C.lib = dbg
import icon_b200.ops as ops
ops.lib = dbg
import torch
from icon_b200 import synthetic as S
dev = torch.device("cuda:0")
v, f = S.body_mesh(); cm, vi = S.body_attributes(v)
body = ops.SmplBody(torch.from_numpy(v)[None].to(dev), torch.from_numpy(f)[None].to(dev), torch.from_numpy(cm)[None].to(dev), torch.from_numpy(vi)[None].to(dev))
for res in (64, 128, 256):
    pts = S.lattice_points(res).permute(0, 2, 1).contiguous().to(dev)
    out = (ctypes.c_ulonglong * 8)()
    dbg.icon_debug_sdf_stats(out, 1)
    ops.sdf_only(pts, torch.eye(4)[None], body)
    torch.cuda.synchronize()
    dbg.icon_debug_sdf_stats(out, 1)
    w = max(out[0], 1)
    print(f"res {res}: warps {out[0]} overflow {out[1]} ({out[1]/w:.3f}) avg leaves {out[2]/w:.1f} avg faces {out[3]/w:.1f}")
