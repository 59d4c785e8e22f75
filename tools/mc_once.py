"""One marching-cubes extraction of an analytic ellipsoid at R = 513 and 257 (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icon_b200 import ops
dev = torch.device("cuda:0")
for R in (257, 513):
    a = torch.linspace(-1, 1, R, device=dev)
    z, y, x = torch.meshgrid(a, a, a, indexing="ij")
    occ = (0.5 + 2.0 * (0.8 - ((x / 0.45) ** 2 + (y / 0.8) ** 2 + (z / 0.3) ** 2).sqrt())).contiguous()
    del x, y, z
    for _ in range(2):
        v, f = ops.marching_cubes(occ, 0.5)
    torch.cuda.synchronize()
    print(R, v.shape, f.shape)
