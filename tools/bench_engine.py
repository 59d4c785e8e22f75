"""Secondary measurements on the GPU (not the bench.py headline): reconstruction engine with an analytic
query function, marching cubes, and the full per-image pipeline (filter + engine + export) with the network.
Prints a JSON dict; CUDA-event timed, 3 warm-ups."""
import json, os, sys, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import torch
from icon_b200 import config, net, ops, synthetic as S
from icon_b200.engine import Seg3dLossless

dev = torch.device("cuda:0")
HBM = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def field(points, **kw):
    p = points[0]
    s = torch.tensor([0.45, 0.8, 0.3], device=p.device)
    r = (p / s).norm(dim=1)
    return (0.5 + 2.0 * (0.8 - r)).view(1, 1, -1)


res_out = {}
for mc in (256, 512):
    res = [2 ** k + 1 for k in range(5, mc.bit_length())]
    eng = Seg3dLossless(query_func=field, b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]], resolutions=res,
                        align_corners=True, balance_value=0.5, faster=True).to(dev)
    ms, occ = timed(lambda: eng())
    R = res[-1]
    ms_mc, (v, f) = timed(lambda: ops.marching_cubes(occ, 0.5))
    ms_exp, _ = timed(lambda: eng.export_mesh(occ), n=3, warm=1)
    G = R + 1 if (R - 1) <= 256 else R - 1
    mc_bytes = G ** 3 * 4 + v.numel() * v.element_size() + f.numel() * 8    # grid read once + outputs
    up_bytes = ((R + 1) // 2) ** 3 * 4 + R ** 3 * 4      # last-level upsample
    ms_up, _ = timed(lambda: ops.grid_upsample(torch.zeros(((R + 1) // 2,) * 3, device=dev), None, 0.5, want_mask=False))
    res_out[f"mcube_res_{mc}"] = {
        "levels": res, "engine_ms_analytic_query": ms, "query_points_per_call": eng.last_query_counts,
        "marching_cubes_ms_device": ms_mc, "export_mesh_ms_incl_d2h": ms_exp, "verts": int(v.shape[0]), "faces": int(f.shape[0]),
        "mc_algorithmic_GBs": mc_bytes / (ms_mc * 1e-3) / 1e9, "mc_frac_of_measured_hbm": mc_bytes / (ms_mc * 1e-3) / 1e9 / HBM,
        "last_upsample_ms": ms_up, "last_upsample_GBs": up_bytes / (ms_up * 1e-3) / 1e9,
        "last_upsample_frac_of_measured_hbm": up_bytes / (ms_up * 1e-3) / 1e9 / HBM}

# full per-image pipeline with the network (icon-filter, 512^2 inputs, mcube_res 256)
cfg = config.preset("icon-filter")
netG = net.HGPIFuNet(cfg).to(dev).eval()
sd = S.mlp_state_dict(c0=13, seed=1)
sd["filters.3.bias"] = sd["filters.3.bias"] + 0.5
netG.if_regressor.load_state_dict(sd)
v, f = S.body_mesh(seed=0)
cm, vi = S.body_attributes(v, seed=0)
g = torch.Generator().manual_seed(0)
batch = {"image": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "T_normal_F": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "T_normal_B": (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev),
         "smpl_verts": torch.from_numpy(v)[None].to(dev), "smpl_faces": torch.from_numpy(f)[None].to(dev),
         "smpl_cmap": torch.from_numpy(cm)[None].to(dev), "smpl_vis": torch.from_numpy(vi)[None].to(dev)}
eng = Seg3dLossless(query_func=net.query_func, b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]],
                    resolutions=[33, 65, 129, 257], align_corners=True, balance_value=0.5, faster=True).to(dev)
with torch.no_grad():
    ms_f, feats = timed(lambda: netG.filter(batch), n=3, warm=1)
    ms_e, occ = timed(lambda: eng(opt=cfg, netG=netG, features=feats, proj_matrix=None), n=3, warm=1)
    entry = {"filter_ms (NormalNet + 2 x HGFilter)": ms_f, "engine_ms_with_network": ms_e,
             "query_points_per_call": eng.last_query_counts}
    if occ is not None:
        ms_x, (vv, ff) = timed(lambda: eng.export_mesh(occ), n=3, warm=1)
        entry.update({"export_mesh_ms": ms_x, "verts": int(vv.shape[0]), "faces": int(ff.shape[0])})
res_out["per_image_icon_filter_256"] = entry
print(json.dumps(res_out, indent=1))
