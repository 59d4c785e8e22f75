"""Candidate statistics of k_sdf_warp (overflowing warps, leaves and faces per warp).

Needs a debug build of the library with the counters compiled in:
    cd icon_b200/csrc && for f in *.cu; do nvcc -c $f -o /tmp/dbg_${f%.cu}.o -gencode arch=compute_100a,code=sm_100a \
        -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DICON_SDF_STATS $( [ $f = sdf.cu -o $f = smpl.cu ] && echo -fmad=false ); done
    nvcc -shared -o build/libicon_dbg.so /tmp/dbg_*.o -gencode arch=compute_100a,code=sm_100a -lcudart
"""
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
