"""Eager (no CUDA graphs) filter() calls at 512 x 512 for an ncu launch list.  Three calls: the first packs weights,
the last is the steady state (use `ncu -s <launches of the first two>` or tools/launch_summary.py's skip fraction)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icon_b200 import _C, config, net, graphs, synthetic as S
dev = torch.device("cuda:0")
graphs.enable(False)
netG = net.HGPIFuNet(config.preset(sys.argv[1] if len(sys.argv) > 1 else "icon-filter")).to(dev).eval()
batch = {k: v.to(dev) for k, v in S.encoder_inputs_512(seed=5).items()}
batch.update({"smpl_verts": torch.zeros(1, 4, 3).to(dev), "smpl_faces": torch.zeros(1, 2, 3).long().to(dev),
              "smpl_vis": torch.zeros(1, 4, 1).to(dev), "smpl_cmap": torch.zeros(1, 4, 3).to(dev)})
with torch.no_grad():
    for i in range(3):
        l0 = _C.launch_count()
        netG.filter(batch)
        torch.cuda.synchronize()
        print(f"MARK filter {i} done: {_C.launch_count() - l0} icon launches", flush=True)
