"""Per-layer timing of the NHWC encoder kernels at the shapes of the 512 x 512 encoders (CUDA events, median of 20)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from icon_b200 import nhwc as T

dev = torch.device("cuda:0")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
with torch.no_grad():
    for name, cin, cout, k, kind, h in [
            ("res 1024->1024 3x3 reflect @32", 1024, 1024, 3, "reflect", 32),
            ("down 64->128 s2 @512", 64, 128, 3, "s2", 512), ("down 128->256 s2 @256", 128, 256, 3, "s2", 256),
            ("down 256->512 s2 @128", 256, 512, 3, "s2", 128), ("down 512->1024 s2 @64", 512, 1024, 3, "s2", 64),
            ("up 1024->512 T @32", 1024, 512, 3, "T", 32), ("up 512->256 T @64", 512, 256, 3, "T", 64),
            ("up 256->128 T @128", 256, 128, 3, "T", 128), ("up 128->64 T @256", 128, 64, 3, "T", 256),
            ("hg 256->128 3x3 @128", 256, 128, 3, "zero", 128), ("hg 128->64 3x3 @128", 128, 64, 3, "zero", 128),
            ("hg 64->64 3x3 @128", 64, 64, 3, "zero", 128), ("hg 64->64 3x3 @256", 64, 64, 3, "zero", 256),
            ("hg 32->32 3x3 @256", 32, 32, 3, "zero", 256), ("hg 256->256 1x1 @128", 256, 256, 1, "zero", 128),
            ("hg 256->128 3x3 @64", 256, 128, 3, "zero", 64), ("hg 256->128 3x3 @32", 256, 128, 3, "zero", 32)]:
        x = torch.randn(1, cin, h, h, device=dev)
        raw = T.raw_from_nchw(x)
        if kind == "T":
            m = nn.ConvTranspose2d(cin, cout, 3, stride=2, padding=1, output_padding=1).to(dev)
            op, _ = T.act(raw)
            fn = lambda: T.conv_transpose(op, m)
            flop = 2.0 * cin * 9 * cout * h * h
        else:
            m = nn.Conv2d(cin, cout, k, stride=2 if kind == "s2" else 1,
                          padding=0 if kind == "reflect" else (k // 2)).to(dev)
            op, _ = T.act(raw, halo=1 if kind == "reflect" else 0, s2d=(kind == "s2"))
            fn = lambda: T.conv(op, m)
            oh = h // 2 if kind == "s2" else h
            flop = 2.0 * cin * k * k * cout * oh * oh
        ms = timed(fn)
        ms_act = timed(lambda: T.act(raw, T.finalize(raw), relu=True, halo=1 if kind == "reflect" else 0, s2d=(kind == "s2")))
        rows.append({"layer": name, "conv_ms": ms, "tflops": flop / ms / 1e9, "executed_tflops": 3 * flop / ms / 1e9,
                     "finalize_act_ms": ms_act})
        print(json.dumps(rows[-1]), flush=True)
    for cin, s, refl, h in [(6, 1, True, 512), (3, 2, False, 512)]:
        m = nn.Conv2d(cin, 64, 7, stride=s, padding=0 if refl else 3).to(dev)
        x = torch.randn(1, cin, h, h, device=dev)
        ms = timed(lambda: T.stem_conv7(x, m, reflect=refl))
        print(json.dumps({"layer": f"stem {cin}->64 7x7 s{s} @{h}", "conv_ms": ms,
                          "tflops": 2.0 * cin * 49 * 64 * (h // s) ** 2 / ms / 1e9}), flush=True)
    m = nn.Conv2d(64, 3, 7).to(dev)
    oph, _ = T.act(T.raw_from_nchw(torch.randn(1, 64, 512, 512, device=dev)))
    print(json.dumps({"layer": "head 64->3 7x7 @512", "conv_ms": timed(lambda: T.conv7_head(oph, m, True))}), flush=True)
