"""Two eager (no CUDA graphs) filter() calls at 512 x 512 for an ncu launch list: the second call is the steady state."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icon_b200 import config, net, graphs, synthetic as S
dev = torch.device("cuda:0")
graphs.enable(False)
netG = net.HGPIFuNet(config.preset(sys.argv[1] if len(sys.argv) > 1 else "icon-filter")).to(dev).eval()
batch = {k: v.to(dev) for k, v in S.encoder_inputs_512(seed=5).items()}
batch.update({"smpl_verts": torch.zeros(1, 4, 3).to(dev), "smpl_faces": torch.zeros(1, 2, 3).long().to(dev),
              "smpl_vis": torch.zeros(1, 4, 1).to(dev), "smpl_cmap": torch.zeros(1, 4, 3).to(dev)})
with torch.no_grad():
    for _ in range(2):
        netG.filter(batch)
        torch.cuda.synchronize()
        print("MARK filter done", flush=True)
