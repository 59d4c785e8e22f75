import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icon_b200 import config, net
dev = torch.device("cuda:0")
cfg = config.preset("icon-filter")
netG = net.HGPIFuNet(cfg).to(dev).eval()
g = torch.Generator().manual_seed(0)
batch = {"image": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "T_normal_F": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "T_normal_B": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "smpl_verts": torch.zeros(1, 4, 3).to(dev), "smpl_faces": torch.zeros(1, 2, 3).long().to(dev),
         "smpl_vis": torch.zeros(1, 4, 1).to(dev), "smpl_cmap": torch.zeros(1, 4, 3).to(dev)}
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        netG.filter(batch)
torch.cuda.synchronize()
