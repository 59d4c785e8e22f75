"""Host-side logic that needs no GPU: C-ABI surface, weight packing, state_dict parity, configs."""
import ctypes
import json
import os
import re

import pytest
import torch

from icon_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from icon_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "icon_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(icon_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/icon_b200.h but not exported"
    from icon_b200 import _C
    assert set(_C.EXPORTS) == declared          # the ctypes table covers the whole header
    assert _C.lib.icon_version() == 1
    # size queries are host-only and must work without a GPU
    assert _C.lib.icon_smpl_workspace_bytes(6890, 13776) > 13776 * 64
    assert _C.lib.icon_query_workspace_bytes(1 << 20, 13776, 0) > (1 << 20) * 32
    assert _C.lib.icon_mc_workspace_bytes(257, 1) >= 258 ** 3 * 6        # voff i32 + flags + case per voxel
    assert _C.lib.icon_voxelize_workspace_bytes(128) >= 128 ** 3
    assert _C.lib.icon_visibility_workspace_bytes(4096) >= 4096 * 4096 * 8


def test_ops_refuse_cpu_tensors(built_lib):
    from icon_b200 import _C, ops
    sd = S.mlp_state_dict(13, seed=1)
    packed = ops.pack_mlp(sd, 13)
    with pytest.raises(_C.IconError):
        ops.mlp_only(torch.zeros(1, 13, 8), packed)      # no CPU fallback


@pytest.mark.parametrize("c0", [13, 10])
def test_pack_mlp_folding_matches_oracle(built_lib, c0):
    """Unpack the BN-folded k-major block and evaluate it with plain matmuls on the CPU."""
    from icon_b200 import ops
    from oracle import query as OQ
    sd = S.mlp_state_dict(c0, seed=4)
    pk = ops.pack_mlp(sd, c0)
    assert pk.tc.numel() == ops.MLP_TC_BYTES and pk.tc.dtype == torch.uint8
    p = pk.f32.double()
    o = 0

    def take(n):
        nonlocal o
        r = p[o:o + n]
        o += n
        return r

    W0t, b0 = take(16 * 512).view(16, 512), take(512)
    W1t, b1 = take(512 * 256).view(512, 256), take(256)
    W2t, b2 = take(272 * 128).view(272, 128), take(128)
    W3, b3 = take(144), take(1)
    assert o == ops.MLP_PACKED_FLOATS
    x = torch.randn(1, c0, 500, generator=torch.Generator().manual_seed(0)).double()
    x16 = torch.zeros(16, 500, dtype=torch.float64)
    x16[:c0] = x[0]
    lre = torch.nn.functional.leaky_relu
    h0 = lre(W0t.t() @ x16 + b0[:, None], 0.01)
    h1 = lre(W1t.t() @ h0 + b1[:, None], 0.01)
    h2 = lre(W2t.t() @ torch.cat([h1, x16]) + b2[:, None], 0.01)
    y = W3 @ torch.cat([h2, x16]) + b3
    ref = OQ.mlp_forward(sd, x, dtype=torch.float64)[0, 0]
    assert (y - ref).abs().max() < 1e-5


def _unswizzle128(buf, rows):
    """Inverse of ops._img_sw128: K-major SWIZZLE_128B tile bytes -> [rows, 64] fp16."""
    import numpy as np
    t = np.frombuffer(buf, dtype=np.float16).reshape(rows, 8, 8)
    out = np.zeros_like(t)
    r = np.arange(rows)[:, None]
    c = np.arange(8)[None, :]
    out[r, c] = t[r, c ^ (r % 8)]
    return out.reshape(rows, 64).astype(np.float64)


def _unpack_nosw(buf, rows):
    """Inverse of ops._img_nosw: [k-core][row group][8 rows][8 elems] -> [rows, 16] fp16."""
    import numpy as np
    t = np.frombuffer(buf, dtype=np.float16).reshape(2, rows // 8, 8, 8)
    return np.ascontiguousarray(t.transpose(1, 2, 0, 3)).reshape(rows, 16).astype(np.float64)


@pytest.mark.parametrize("c0", [13, 10])
def test_tensor_core_blob_evaluated_like_the_kernel_matches_oracle(built_lib, c0):
    """Decode the UMMA tiles of the tcgen05 blob (hi + lo) and run the network the way k_query_mlp_tc does: x0 column
    15 is the constant 1, b0 and b2 are NOT added (they sit in row 15 of W0 and of the x0 tail of W2), b1 / b3 are."""
    import numpy as np
    from icon_b200 import ops
    from oracle import query as OQ
    sd = S.mlp_state_dict(c0, seed=9)
    blob = ops.pack_mlp(sd, c0).tc.numpy().tobytes()
    o = 0

    def take(n):
        nonlocal o
        r = blob[o:o + n]
        o += n
        return r

    W0 = _unpack_nosw(take(16384), 512) + _unpack_nosw(take(16384), 512)                       # [512, 16]
    W1 = np.concatenate([_unswizzle128(take(32768), 256) + _unswizzle128(take(32768), 256) for _ in range(8)], 1)   # [256, 512]
    W2 = np.concatenate([_unswizzle128(take(16384), 128) + _unswizzle128(take(16384), 128) for _ in range(4)], 1)   # [128, 256]
    W2t = _unpack_nosw(take(4096), 128) + _unpack_nosw(take(4096), 128)                        # [128, 16]
    f32 = np.frombuffer(take((512 + 256 + 128 + 144 + 4) * 4), dtype=np.float32).astype(np.float64)
    assert o == ops.MLP_TC_BYTES
    b1, w3, b3 = f32[512:768], f32[896:1040], f32[1040]
    x = torch.randn(1, c0, 300, generator=torch.Generator().manual_seed(1)).double()
    x16 = np.zeros((16, 300)); x16[:c0] = x[0].numpy(); x16[15] = 1.0
    lre = lambda v: np.maximum(v, 0.01 * v)
    h0 = lre(W0 @ x16)
    h1 = lre(W1 @ h0 + b1[:, None])
    h2 = lre(W2 @ h1 + W2t @ x16)
    xs = x16.copy(); xs[15] = 0.0                                # layer 3 reads the fp32 feature copy: no constant row
    y = w3[:128] @ h2 + w3[128:] @ xs + b3
    ref = OQ.mlp_forward(sd, x, dtype=torch.float64)[0, 0].numpy()
    assert np.abs(y - ref).max() < 2e-5                          # fp16 hi + lo weights: 22 significant bits


def test_pack_mlp_refuses_16_input_channels(built_lib):
    """x0 column 15 is taken by the constant that carries the folded biases."""
    from icon_b200 import ops
    with pytest.raises(NotImplementedError):
        ops.pack_mlp(S.mlp_state_dict(16, seed=1), 16)


def test_state_dict_keys_match_reference(golden_dir):
    from icon_b200 import config, net
    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))

    def shapes(m):
        return {k: list(v.shape) for k, v in m.state_dict().items()}

    g = net.HGPIFuNet(config.preset("icon-filter"))
    assert shapes(g.F_filter) == keys["HGFilter(opt,2,3)"]
    assert shapes(g.normal_filter.netF) == keys["define_G(6,3,64,global,4,9,1,3,instance)"]
    assert shapes(g.normal_filter.netB) == keys["define_G(6,3,64,global,4,9,1,3,instance)"]
    assert shapes(g.if_regressor) == keys["MLP([13,512,256,128,1])"]
    p = net.HGPIFuNet(config.preset("pamir"))
    assert shapes(p.ve) == keys["VolumeEncoder(3,7,2)"]
    assert shapes(p.F_filter) == keys["HGFilter(opt,2,9)"]
    top = set(k.split(".")[0] for k in g.state_dict())
    assert top == {"if_regressor", "F_filter", "normal_filter"}


@pytest.mark.parametrize("name,c0", [("icon-filter", 13), ("icon-nofilter", 10), ("pamir", 13), ("pifu", 13)])
def test_presets_give_reference_mlp_width(name, c0):
    from icon_b200 import config, net
    g = net.HGPIFuNet(config.preset(name))
    assert g.if_regressor.filter_channels == [c0, 512, 256, 128, 1]
    assert g.sdf_clip == pytest.approx(0.05)
    assert g.if_regressor.last_op is None          # test_mode: no sigmoid (HGPIFuNet.py:133)


def test_engine_constructor_contract():
    """Seg3dLossless ctor as apps/ICON.py:78-90 calls it (the import-path side is tests/test_overlay_cpu.py)."""
    from icon_b200.engine import Seg3dLossless
    eng = Seg3dLossless(query_func=None, b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]],
                        resolutions=[33, 65, 129, 257], align_corners=True, faster=True)
    assert set(dict(eng.named_buffers())) >= {"b_min", "b_max", "resolutions"}
    with pytest.raises(AssertionError):
        Seg3dLossless(None, [[-1.0, 1, -1]], [[1.0, -1, 1]], resolutions=[32, 64])
    with pytest.raises(NotImplementedError):
        Seg3dLossless(None, [[-1.0, 1, -1]], [[1.0, -1, 1]], resolutions=[33, 65], faster=False)()


def test_oracle_marching_cubes_is_watertight_and_matches_analytic_sphere():
    import numpy as np
    from oracle import mcubes as OM
    R = 41
    a = np.linspace(-1, 1, R)
    z, y, x = np.meshgrid(a, a, a, indexing="ij")
    occ = (0.5 + (0.6 - np.sqrt(x * x + y * y + z * z))).astype(np.float32)
    v, f = OM.export_mesh(occ, 0.5)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    _, c = np.unique(np.sort(e, 1), axis=0, return_counts=True)
    assert (c == 2).all()
    # vertices (grid-index units of the ORIGINAL grid) lie on the radius-0.6 sphere
    p = v / ((R - 1) / 2.0) - 1.0
    assert np.abs(np.linalg.norm(p, axis=1) - 0.6).max() < 2e-3
    # Euler characteristic of a sphere
    V, E, F = len(v), len(np.unique(np.sort(e, 1), axis=0)), len(f)
    assert V - E + F == 2
    # consistent orientation: signed volume is positive or negative for ALL faces summed, and matches 4/3 pi r^3
    vol = np.einsum("ij,ij->i", p[f[:, 0]], np.cross(p[f[:, 1]], p[f[:, 2]])).sum() / 6.0
    assert abs(abs(vol) - 4.0 / 3.0 * np.pi * 0.6 ** 3) < 2e-2


def test_display_refuses_cpu_tensors():
    """R14 (training preview) is a kernel now (icon_display): like every op it has no CPU path."""
    from icon_b200 import _C
    from icon_b200.engine import Seg3dLossless
    eng = Seg3dLossless(None, [[-1.0, 1, -1]], [[1.0, -1, 1]], resolutions=[17, 33], align_corners=True, faster=True)
    with pytest.raises(_C.IconError):
        eng.display(torch.zeros(33, 33, 33))


def test_source_cache_is_keyed_on_tensor_identity_and_version():
    """A (data_ptr, _version) key can go stale when a freed tensor's address is reused; the caches of the prepared
    body / pamir volume hold their sources and compare identity + version instead."""
    import torch
    from icon_b200.net import _SourceCache
    c = _SourceCache()
    a, b = torch.zeros(4), torch.zeros(3)
    assert c.get([a, b]) is None
    c.put([a, b], "value")
    assert c.get([a, b]) == "value"
    assert c.get([a.clone(), b]) is None          # equal content, other object
    a.add_(1)                                     # in-place update bumps the version
    assert c.get([a, b]) is None
    c.put([a, b], "new")
    assert c.get([a, b]) == "new"
    c.clear()
    assert c.get([a, b]) is None
