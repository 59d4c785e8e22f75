"""Run the dense 256^3 query several times and report where (tile / row / CTA) the outputs differ, if anywhere."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from icon_b200 import net, synthetic as S

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["icon-filter-256"]
cfg, netG, _ = bench.build_model(dev, wl)
im = bench.DeviceImage(bench.build_image(wl, 0), dev)
im.bind(netG)
pts = S.lattice_points(256).to(dev)
outs = [net.query_func(cfg, netG, [im.feat], pts).clone() for _ in range(6)]
torch.cuda.synchronize()
ref = outs[0].flatten()
for k, o in enumerate(outs[1:], 1):
    o = o.flatten()
    bad = torch.nonzero(o != ref).flatten()
    print(f"run {k}: {bad.numel()} differing outputs")
    if bad.numel():
        b = bad[:4000].cpu()
        tile = b // 128
        print("  max |diff|", (o - ref).abs().max().item())
        print("  rows      ", sorted(set((b % 128).tolist()))[:40], "...")
        print("  tiles     ", sorted(set(tile.tolist()))[:20], "... (", len(set(tile.tolist())), "distinct )")
        print("  cta       ", sorted(set((tile % 148).tolist()))[:40])
        print("  tile//148 ", sorted(set((tile // 148).tolist()))[:40])
        i = int(b[0])
        print("  first: idx", i, "ref", ref[i].item(), "got", o[i].item())
