"""filter() and NormalNet.forward with and without CUDA-graph replay (CUDA events, steady state)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icon_b200 import config, net, graphs
dev = torch.device("cuda:0")
netG = net.HGPIFuNet(config.preset("icon-filter")).to(dev).eval()
g = torch.Generator().manual_seed(0)
batch = {"image": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "T_normal_F": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "T_normal_B": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "smpl_verts": torch.zeros(1, 4, 3).to(dev), "smpl_faces": torch.zeros(1, 2, 3).long().to(dev),
         "smpl_vis": torch.zeros(1, 4, 1).to(dev), "smpl_cmap": torch.zeros(1, 4, 3).to(dev)}


def timed(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    for flag in (False, True):
        graphs.enable(flag)
        print(f"cuda graphs {flag}: filter {timed(lambda: netG.filter(batch)):.2f} ms, "
              f"NormalNet {timed(lambda: netG.normal_filter(batch)):.2f} ms")
