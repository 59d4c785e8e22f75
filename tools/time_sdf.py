"""SDF block alone on the dense lattice for every points-per-warp setting (icon_set_sdf_policy)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icon_b200 import ops, synthetic as S

dev = torch.device("cuda:0")
v, f = S.body_mesh(); cm, vi = S.body_attributes(v)
body = ops.SmplBody(torch.from_numpy(v)[None].to(dev), torch.from_numpy(f)[None].to(dev),
                    torch.from_numpy(cm)[None].to(dev), torch.from_numpy(vi)[None].to(dev))
eye = torch.eye(4)[None]
for res in (128, 256):
    pts = S.lattice_points(res).permute(0, 2, 1).contiguous().to(dev)
    for ppw in (8, 16, 32):
        ops.set_sdf_policy(ppw)
        for _ in range(2):
            ops.sdf_only(pts, eye, body)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.sdf_only(pts, eye, body)
        e1.record(); torch.cuda.synchronize()
        print(f"res {res} ppw {ppw}: {e0.elapsed_time(e1) / 5:.3f} ms (bin+sort+sdf)")
ops.set_sdf_policy(0)
