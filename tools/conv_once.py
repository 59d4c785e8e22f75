"""A few launches of the dominant encoder kernels at their real shapes, for `ncu --set full` captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from icon_b200 import nhwc as T
dev = torch.device("cuda:0")
with torch.no_grad():
    m = nn.Conv2d(1024, 1024, 3, padding=0).to(dev)                 # ResnetBlock conv (FBNet.py:268-319)
    raw = T.raw_from_nchw(torch.randn(1, 1024, 32, 32, device=dev))
    op, _ = T.act(raw, halo=1)
    m2 = nn.Conv2d(256, 128, 3, padding=1, bias=False).to(dev)      # ConvBlock conv1 at 128 x 128 (net_util.py:258-280)
    raw2 = T.raw_from_nchw(torch.randn(1, 256, 128, 128, device=dev))
    op2, _ = T.act(raw2)
    for _ in range(4):
        r = T.conv(op, m)
        T.act(r, T.finalize(r), relu=True, halo=1)
        T.conv(op2, m2)
    torch.cuda.synchronize()
print("done")
