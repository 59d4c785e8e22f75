"""PaMIR semantic voxelisation: oracle known-answer tests on CPU, CUDA vs oracle on the GPU.
The reference holds no test for voxelize_cuda and its source is absent -> parity unpinned (oracle/voxelize.py)."""
import numpy as np
import pytest
import torch

from icon_b200 import synthetic as S


@pytest.fixture(scope="module")
def built_lib():
    from icon_b200 import build
    return build.build()


def _unit_tet(scale=0.4):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32) * scale - 0.1
    return v, np.array([[0, 1, 2, 3]], np.int32)


def test_oracle_single_tet_matches_analytic_halfspaces():
    from oracle import voxelize as OV
    v, t = _unit_tet()
    res = 32
    occ = OV.occupancy(v, t, res)
    c = (np.arange(res, dtype=np.float64) + 0.5) / res - 0.5
    Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
    x, y, z = (X + 0.1) / 0.4, (Y + 0.1) / 0.4, (Z + 0.1) / 0.4
    inside = (x >= 0) & (y >= 0) & (z >= 0) & (x + y + z <= 1)
    margin = np.minimum.reduce([np.abs(x), np.abs(y), np.abs(z), np.abs(1 - x - y - z)]) > 1e-4
    assert np.array_equal(occ.astype(bool)[margin], inside[margin])
    assert occ.sum() > 50
    # orientation of the tetrahedron must not matter, degenerate / padded rows are ignored
    occ2 = OV.occupancy(v, np.array([[0, 2, 1, 3], [0, 0, 0, 0]], np.int32), res)
    assert np.array_equal(occ, occ2)


def test_oracle_semantic_volume_properties():
    from oracle import voxelize as OV
    verts, n_surf, tets, code = S.tet_body(rings=10, segs=12)
    vol, occ = OV.semantic_volume(verts, n_surf, code, tets, 32, 0.05)
    assert vol.shape == (3, 32, 32, 32) and occ.sum() > 200
    assert np.all(vol[:, occ == 0] == 0)
    on = vol[:, occ == 1]
    assert on.min() >= 0 and on.max() <= 1.0            # convex combination of codes in [0,1], damped by the 1e-3
    # a constant code comes back as that constant times sum_w / (1e-3 + sum_w)
    const = np.full_like(code, 0.5)
    vol_c, _ = OV.semantic_volume(verts, n_surf, const, tets, 32, 0.05)
    assert np.all(vol_c[:, occ == 1] <= 0.5 + 1e-6)
    # the code field follows position: x-code grows with x
    xs = np.nonzero(occ)[2]
    assert np.corrcoef(xs, vol[0][occ == 1])[0, 1] > 0.8


def test_voxelization_module_refuses_cpu(built_lib):
    from icon_b200.voxelize import Voxelization
    verts, n_surf, tets, code = S.tet_body(rings=6, segs=8)
    m = Voxelization(code, code[:1], np.zeros((1, 3), np.int32), tets, 16, 0.05, 7, 1, "cpu")
    with pytest.raises(TypeError):
        m(torch.from_numpy(verts)[None])


@pytest.mark.gpu
@pytest.mark.parametrize("res,rings,segs", [(32, 10, 12), (128, 82, 84)])
def test_voxelize_vs_oracle(res, rings, segs):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from icon_b200 import ops
    from oracle import voxelize as OV
    dev = torch.device("cuda:0")
    verts, n_surf, tets, code = S.tet_body(rings=rings, segs=segs)
    # padded like TestDataset.compute_voxel_verts: zero rows at the end of both tables
    tets_p = np.concatenate([tets, np.zeros((7, 4), np.int32)], 0)
    ref, occ = OV.semantic_volume(verts, n_surf, code, tets_p, res, 0.05)
    out = ops.voxelize(torch.from_numpy(verts).to(dev), n_surf, torch.from_numpy(code).to(dev),
                       torch.from_numpy(tets_p).to(dev), res, 0.05)[0].cpu().numpy()
    got_occ = (np.abs(out).sum(0) > 0)
    assert np.array_equal(got_occ, occ.astype(bool)), f"occupancy differs on {(got_occ != occ.astype(bool)).sum()} voxels"
    assert np.abs(out - ref).max() <= 2e-5           # fp32 sums over 6890 vertices vs the oracle's fp64


@pytest.mark.gpu
def test_pamir_filter_query_with_voxelisation():
    """filter() with voxel_verts -> Voxelization -> VolumeEncoder -> query equals the same query fed the
    oracle's semantic volume through the 'vol' entry."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from icon_b200 import config, net
    from oracle import voxelize as OV
    dev = torch.device("cuda:0")
    cfg = config.preset("pamir")
    model = net.HGPIFuNet(cfg).to(dev).eval()
    model.load_state_dict(S.seeded_like(model.state_dict(), 5))
    verts, n_surf, tets, code = S.tet_body()
    model.set_smpl_constants(code, code[:1], np.zeros((1, 3), np.int32), tets)
    pad_v, pad_f = 8000 - len(verts), 25100 - len(tets)
    vv = torch.from_numpy(np.pad(verts, ((0, pad_v), (0, 0))))[None].to(dev)
    vf = torch.from_numpy(np.pad(tets, ((0, pad_f), (0, 0))))[None].long().to(dev)
    g = torch.Generator().manual_seed(3)
    base = {"image": torch.rand(1, 3, 128, 128, generator=g).to(dev) * 2 - 1,
            "normal_F": torch.rand(1, 3, 128, 128, generator=g).to(dev) * 2 - 1,
            "normal_B": torch.rand(1, 3, 128, 128, generator=g).to(dev) * 2 - 1}
    pts = (torch.rand(1, 3, 5000, generator=g) * 1.0 - 0.5).to(dev)
    calib = torch.eye(4)[None].to(dev)
    with torch.no_grad():
        f1 = model.filter({**base, "voxel_verts": vv, "voxel_faces": vf,
                           "pad_v_num": torch.tensor([pad_v]).to(dev), "pad_f_num": torch.tensor([pad_f]).to(dev)})
        p1 = model.query(f1, pts, calib, regressor=model.if_regressor)[0]
        # the same query fed (a) the CUDA volume through the 'vol' entry: identical wiring -> identical bits;
        # (b) the oracle's fp64 volume: equal up to the 2e-5 volume tolerance amplified by VolumeEncoder + MLP
        vol_cuda = model.voxelization(vv[:, :len(verts)])
        p2 = model.query(model.filter({**base, "vol": vol_cuda}), pts, calib, regressor=model.if_regressor)[0]
        ref_vol, _ = OV.semantic_volume(verts, n_surf, code, tets, 128, 0.05)
        p3 = model.query(model.filter({**base, "vol": torch.from_numpy(ref_vol)[None].to(dev)}), pts, calib,
                         regressor=model.if_regressor)[0]
    assert p1.shape == (1, 1, 5000)
    assert torch.isfinite(p1).all()
    assert torch.equal(p1, p2)
    assert (p1 - p3).abs().max().item() <= 2e-3 * max(1.0, p3.abs().max().item())
