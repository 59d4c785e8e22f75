"""get_visibility at the reference's 4096^2 on the synthetic body (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icon_b200 import ops, synthetic as S
dev = torch.device("cuda:0")
v, f = S.body_mesh()
xyz = torch.from_numpy(v).to(dev)
xyz = (torch.cat([xyz[:, :2], -xyz[:, 2:3]], 1) + 1) / 2
f = torch.from_numpy(f).to(dev)
for _ in range(3):
    vis = ops.visibility(xyz, f, 4096)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    vis = ops.visibility(xyz, f, 4096)
e1.record(); torch.cuda.synchronize()
print(f"get_visibility 4096^2, V={len(v)} F={len(f)}: {e0.elapsed_time(e1) / 10:.3f} ms, visible {vis.mean().item():.3f}")
