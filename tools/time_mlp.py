"""Quick timing + error report of the fused MLP kernels on the GPU (development helper)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icon_b200 import ops, synthetic as S
from oracle import query as OQ
dev = torch.device("cuda:0")
sd = S.mlp_state_dict(13, seed=3)
pk = ops.pack_mlp(sd, 13, device=dev)
g = torch.Generator().manual_seed(1)
x = torch.randn(1, 13, 200000, generator=g) * 1.5
ref = OQ.mlp_forward(sd, x, dtype=torch.float64).float()
for impl in ("fp32", "tcgen05"):
    ops.set_mlp_impl(impl)
    y = ops.mlp_only(x.to(dev), pk).cpu()
    e = (y - ref).abs()
    print(f"{impl:8s} max|err| {e.max().item():.3e}  p99.9 {e.flatten().kthvalue(int(0.999*e.numel())).values.item():.3e}  |ref| max {ref.abs().max().item():.2f}")
N = 1 << 24
xb = torch.randn(1, 13, N, device=dev)
for impl in ("fp32", "tcgen05"):
    ops.set_mlp_impl(impl)
    for _ in range(2): ops.mlp_only(xb, pk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ops.mlp_only(xb, pk)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{impl:8s} N=2^24: {ms:.2f} ms  -> {N/ms/1e3:.1f} M pts/s, {N*344602/ms/1e9:.1f} TFLOP/s algorithmic")
