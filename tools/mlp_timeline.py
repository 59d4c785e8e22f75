"""SM-clock timeline of ONE 128-point tile of k_query_mlp_tc (one CTA): when the MMA issuer, a worker warp and a gather
warp pass each hand-off.  Uses the diagnostics build (python -m icon_b200.build --timeline -> libicon_b200_tl.so).
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ICON_B200_LIB"] = os.path.join(ROOT, "icon_b200", "libicon_b200_tl.so")
import torch
import bench
from icon_b200 import _C, synthetic as S

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["icon-filter-256"]
cfg, netG, _ = bench.build_model(dev, wl)
im = bench.DeviceImage(bench.build_image(wl, 0), dev)
im.bind(netG)
pts = S.lattice_points(256).to(dev)
from icon_b200 import net
for _ in range(3):
    net.query_func(cfg, netG, [im.feat], pts)
torch.cuda.synchronize()
lib = ctypes.CDLL(_C.LIB_PATH)
buf = (ctypes.c_longlong * 256)()
assert lib.icon_debug_mlp_timeline(buf) == 0
t = list(buf)
t0 = t[0]
names = {0: "mma tile start", 1: "mma ACC2E seen (layer 1 may start)", 19: "mma ACC1 committed", 25: "mma ACC2 committed",
         20: "mma W2(0,1) seen", 22: "mma W2(2,3) seen", 100: "wrk ACC1 seen",
         128: "epi iteration start", 131: "epi next tile's x0 published", 132: "epi ACC2 seen", 133: "epi ACC2 read (ACC2E arrived)",
         134: "epi tile stored", 140: "epi warp 13 ACC2 seen", 141: "epi warp 13 ACC2E arrived", 142: "epi warp 17 ACC2 seen",
         143: "epi warp 17 ACC2E arrived"}
for j in range(8):
    names[32 + j] = f"mma A0F({j}) seen"; names[3 + 2 * j] = f"mma W({j}) seen"; names[4 + 2 * j] = f"mma L1({j}) issued"
    names[64 + 4 * j] = f"wrk ACC0F({j}) seen"; names[65 + 4 * j] = f"wrk ld({j}) done"; names[66 + 4 * j] = f"wrk math({j}) done"
    names[67 + 4 * j] = f"wrk A0F({j}) arrived"
for c in range(4):
    names[40 + c] = f"mma ACT1({c}) seen"; names[101 + c] = f"wrk ACT1({c}) arrived"
ev = sorted((t[k] - t0, names[k]) for k in names if t[k])
prev = 0
for dt, n in ev:
    print(f"{dt:8d}  (+{dt - prev:6d})  {n}")
    prev = dt
