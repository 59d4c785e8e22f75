"""Host/device time breakdown of one reconstruction (engine with the network), development helper."""
import os, sys, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import torch
from icon_b200 import config, net, ops, synthetic as S
from icon_b200.engine import Seg3dLossless
dev = torch.device("cuda:0")
cfg = config.preset("icon-filter")
netG = net.HGPIFuNet(cfg).to(dev).eval()
sd = S.mlp_state_dict(c0=13, seed=1); sd["filters.3.bias"] = sd["filters.3.bias"] + 0.5
netG.if_regressor.load_state_dict(sd)
v, f = S.body_mesh(seed=0); cm, vi = S.body_attributes(v, seed=0)
netG.smpl_feat_dict = {"smpl_verts": torch.from_numpy(v)[None].to(dev), "smpl_faces": torch.from_numpy(f)[None].to(dev),
                       "smpl_cmap": torch.from_numpy(cm)[None].to(dev), "smpl_vis": torch.from_numpy(vi)[None].to(dev)}
feats = [S.feature_map(12, 128, seed=0).to(dev)]
T = {}
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return r
import ctypes
from icon_b200 import _C
_C.lib.icon_profile_enable(1)
stages = []
_q = ops.query
def q_prof(*a, **kw):
    r = _q(*a, **kw)
    buf = (ctypes.c_float * 4)()
    _C.check(_C.lib.icon_profile_last_query(buf), "prof")
    stages.append((a[1].shape[2], [round(x, 3) for x in buf]))
    return r
ops.query = q_prof
orig = {k: getattr(ops, k) for k in ("grid_init_points", "grid_count_above", "grid_upsample", "grid_dilate", "grid_compact", "grid_scatter", "query")}
for k, fn in orig.items():
    setattr(ops, k, (lambda k, fn: lambda *a, **kw: timed(k, lambda: fn(*a, **kw)))(k, fn))
eng = Seg3dLossless(query_func=net.query_func, b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]],
                    resolutions=[33, 65, 129, 257], align_corners=True, balance_value=0.5, faster=True).to(dev)
import sys as _s
force = int(_s.argv[1]) if len(_s.argv) > 1 else 0
_C.check(_C.lib.icon_set_sdf_policy(force, -1, -1), "policy")
print("forced ppw:", force)
with torch.no_grad():
    for it in range(3):
        T.clear(); stages.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        occ = eng(opt=cfg, netG=netG, features=feats, proj_matrix=None)
        torch.cuda.synchronize(); tot = (time.perf_counter() - t0) * 1e3
print("total (with per-op syncs) %.2f ms; points %s" % (tot, eng.last_query_counts))
for k, v_ in sorted(T.items(), key=lambda kv: -kv[1]):
    print("  %-18s %.2f ms" % (k, v_))
print("per query call: N, [bin+sort, sdf, outlier rank, gather+mlp] ms")
for st in stages: print("  ", st)
