"""Tensor-level wrappers of the encoder kernels (csrc/conv.cu) used by icon_b200/encoders.py.

Each function mirrors the torch call it replaces in the reference (cited in encoders.py) and
enqueues hand-written kernels on the current stream; NCHW fp32, no autograd (inference path)."""
import torch

from . import _C
from ._C import check, lib
from .ops import _need_cuda, _p, _stream


def _c(t):
    return t.detach().float().contiguous()


def conv2d(x, conv, reflect=0, tanh=False, relu=False, residual=None):
    """nn.Conv2d forward; `reflect=p` folds a preceding nn.ReflectionPad2d(p) into the gather."""
    _need_cuda(x, conv.weight)
    x = _c(x)
    w = _c(conv.weight)
    N, Cin, H, W = x.shape
    Cout, _, KH, KW = w.shape
    stride = conv.stride[0]
    pad = reflect if reflect else conv.padding[0]
    if conv.dilation[0] != 1 or conv.groups != 1 or conv.stride[0] != conv.stride[1]:
        raise NotImplementedError("conv2d kernel: dilation 1, groups 1, square stride (all convs of the path)")
    OH = (H + 2 * pad - KH) // stride + 1
    OW = (W + 2 * pad - KW) // stride + 1
    y = torch.empty(N, Cout, OH, OW, dtype=torch.float32, device=x.device)
    b = _c(conv.bias) if conv.bias is not None else None
    r = _c(residual) if residual is not None else None
    check(lib.icon_conv2d(_p(x), _p(w), _p(b), _p(r), _p(y), N, Cin, H, W, Cout, KH, KW, stride, pad, 0,
                          1 if reflect else 0, 0, 2 if tanh else (1 if relu else 0), _stream()), "icon_conv2d")
    return y


def conv_transpose2d(x, conv):
    """nn.ConvTranspose2d forward (FBNet.py:245-252: k3, s2, p1, output_padding 1)."""
    _need_cuda(x, conv.weight)
    x = _c(x)
    w = _c(conv.weight)                       # [Cin, Cout, KH, KW]
    N, Cin, H, W = x.shape
    _, Cout, KH, KW = w.shape
    s, p, op = conv.stride[0], conv.padding[0], conv.output_padding[0]
    OH = (H - 1) * s - 2 * p + KH + op
    OW = (W - 1) * s - 2 * p + KW + op
    y = torch.empty(N, Cout, OH, OW, dtype=torch.float32, device=x.device)
    b = _c(conv.bias) if conv.bias is not None else None
    check(lib.icon_conv2d(_p(x), _p(w), _p(b), None, _p(y), N, Cin, H, W, Cout, KH, KW, s, p, op, 0, 1, 0, _stream()),
          "icon_conv2d(transposed)")
    return y


def group_norm(x, gn, relu=False):
    """nn.GroupNorm forward (+ fused ReLU)."""
    x = _c(x)
    N, C, H, W = x.shape
    y = torch.empty_like(x)
    check(lib.icon_group_norm(_p(x), _p(_c(gn.weight)), _p(_c(gn.bias)), None, _p(y), N, C, H * W, gn.num_groups,
                              float(gn.eps), 1 if relu else 0, _stream()), "icon_group_norm")
    return y


def instance_norm(x, relu=False, residual=None, eps=1e-5):
    """nn.InstanceNorm2d(affine=False) forward (+ fused ReLU or residual add: x + IN(y))."""
    x = _c(x)
    N, C, H, W = x.shape
    y = torch.empty_like(x)
    r = _c(residual) if residual is not None else None
    check(lib.icon_group_norm(_p(x), None, None, _p(r), _p(y), N, C, H * W, C, float(eps), 1 if relu else 0, _stream()),
          "icon_group_norm(instance)")
    return y


def avg_pool2(x):
    x = _c(x)
    N, C, H, W = x.shape
    y = torch.empty(N, C, H // 2, W // 2, dtype=torch.float32, device=x.device)
    check(lib.icon_avg_pool2(_p(x), _p(y), N * C, H, W, _stream()), "icon_avg_pool2")
    return y


def bicubic_up2_add(x, add):
    x, add = _c(x), _c(add)
    N, C, H, W = x.shape
    y = torch.empty(N, C, 2 * H, 2 * W, dtype=torch.float32, device=x.device)
    if tuple(add.shape) != tuple(y.shape):
        raise _C.IconError("bicubic_up2_add: shape mismatch")
    check(lib.icon_bicubic_up2_add(_p(x), _p(add), _p(y), N * C, H, W, _stream()), "icon_bicubic_up2_add")
    return y


def cat_add(parts, residual):
    a, b, c = [_c(t) for t in parts]
    r = _c(residual)
    N, _, H, W = a.shape
    y = torch.empty_like(r)
    check(lib.icon_cat3_add(_p(a), _p(b), _p(c), _p(r), _p(y), N, a.shape[1], b.shape[1], c.shape[1], H * W, _stream()),
          "icon_cat3_add")
    return y


def add3(a, b, c):
    a, b, c = _c(a), _c(b), _c(c)
    y = torch.empty_like(a)
    check(lib.icon_add3(_p(a), _p(b), _p(c), _p(y), a.numel(), _stream()), "icon_add3")
    return y


def normalize_mask(nml, image):
    nml, image = _c(nml), _c(image)
    N, C, H, W = nml.shape
    if C != 3:
        raise _C.IconError("normalize_mask: 3-channel normal map expected")
    y = torch.empty_like(nml)
    check(lib.icon_normalize_mask(_p(nml), _p(image), _p(y), N, image.shape[1], H * W, _stream()), "icon_normalize_mask")
    return y


def conv3d_bn(x, conv, bn, relu=False, residual=None):
    raise NotImplementedError("VolumeEncoder conv3d kernels (PaMIR, lib/net/VE.py) are not built yet: "
                              "pass a pre-encoded in_tensor_dict['vol_feat'] (DESIGN.md section 7)")
