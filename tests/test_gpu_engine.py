"""Engine (Seg3dLossless) and marching-cubes parity on the GPU against the oracle and against
the fixtures produced by the reference's own Seg3dLossless (tests/golden/engine.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from icon_b200 import synthetic as S  # noqa: E402


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _field(points):
    """Same analytic occupancy as tests/golden/make_golden.py, evaluated in float64 on the CPU and
    rounded once, so that CPU and GPU callers see identical values."""
    p = points[0].detach().cpu().double()
    s = torch.tensor([0.45, 0.8, 0.3], dtype=torch.float64)
    r = (p / s).norm(dim=1)
    bump = 0.15 * torch.sin(9.0 * p[:, 0]) * torch.cos(7.0 * p[:, 1])
    return (0.5 + 2.0 * (0.8 - r) + bump).float().view(1, 1, -1)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_engine_query_sets_match_reference_engine(tag, golden_dir):
    dev = _cuda()
    from icon_b200.engine import Seg3dLossless
    g = np.load(os.path.join(golden_dir, "engine.npz"))
    res = [int(r) for r in g[f"{tag}_res"]]
    calls = []

    def qf(points, **kw):
        calls.append(points.detach().cpu().clone())
        return _field(points).to(points.device)

    eng = Seg3dLossless(query_func=qf, b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]], resolutions=res,
                        align_corners=True, balance_value=0.5, faster=True).to(dev)
    occ = eng()
    assert len(calls) == int(g[f"{tag}_ncalls"])
    for i, p in enumerate(calls):
        assert np.array_equal(p.numpy(), g[f"{tag}_pts{i}"]), (tag, i)     # same points, same ORDER
    # the golden grid was produced with a float32 field; ours is float64-rounded: tolerance only
    assert np.abs(occ.cpu().numpy() - g[f"{tag}_occ"]).max() <= 5e-6


def test_engine_grid_identical_to_oracle_engine():
    dev = _cuda()
    from icon_b200.engine import Seg3dLossless
    from oracle.engine import Seg3dOracle
    res = [17, 33, 65, 129]
    eng = Seg3dLossless(query_func=lambda points, **kw: _field(points).to(points.device),
                        b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]], resolutions=res,
                        align_corners=True, balance_value=0.5, faster=True).to(dev)
    occ = eng().cpu()
    ora = Seg3dOracle([[-1.0, 1.0, -1.0]], [[1.0, -1.0, 1.0]], res)
    ref = ora.forward(_field)
    assert [int(c.shape[0]) for c in ora.log] == eng.last_query_counts
    assert torch.equal(occ, ref)                      # bit-identical grid (same field values in)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_display_matches_reference_preview(tag, golden_dir):
    """Seg3dLossless.display (4-view first-hit normal preview) as ONE kernel vs the uint8 image the reference's own
    find_vertices / render_normal / display produced for its own volume (tests/golden/engine.npz)."""
    dev = _cuda()
    from icon_b200.engine import Seg3dLossless
    g = np.load(os.path.join(golden_dir, "engine.npz"))
    res = [int(r) for r in g[f"{tag}_res"]]
    eng = Seg3dLossless(query_func=None, b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]], resolutions=res,
                        align_corners=True, balance_value=0.5, faster=True).to(dev)
    img = eng.display(torch.from_numpy(g[f"{tag}_occ"]).to(dev))
    ref = g[f"{tag}_display"]
    assert img.dtype == np.uint8 and img.shape == ref.shape == (res[-1], 4 * res[-1], 3)
    diff = np.abs(img.astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1                                  # float rounding of the normalisation at a truncation boundary
    assert (diff > 0).mean() < 0.01 and (ref != 255).mean() > 0.05


def test_engine_returns_none_on_empty_volume():
    dev = _cuda()
    from icon_b200.engine import Seg3dLossless
    eng = Seg3dLossless(query_func=lambda points, **kw: torch.zeros(1, 1, points.shape[1], device=points.device),
                        b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]], resolutions=[9, 17, 33],
                        align_corners=True, balance_value=0.5, faster=True).to(dev)
    assert eng() is None


@pytest.mark.parametrize("R", [33, 65])
def test_marching_cubes_index_exact_vs_oracle(R):
    dev = _cuda()
    from icon_b200 import ops
    from oracle import mcubes as OM
    a = torch.linspace(-1, 1, R, dtype=torch.float64)
    z, y, x = torch.meshgrid(a, a, a, indexing="ij")
    pts = torch.stack([x, y, z], -1).reshape(1, -1, 3)
    occ = _field(pts).reshape(R, R, R)
    g = torch.Generator().manual_seed(R)
    occ = occ + 0.05 * torch.randn(R, R, R, generator=g)               # ambiguous cases too
    v, f = ops.marching_cubes(occ.to(dev), 0.5)
    rv, rf = OM.export_mesh(occ.numpy(), 0.5)
    assert v.dtype == torch.float32 and f.dtype == torch.int64
    assert np.array_equal(f.cpu().numpy(), rf)
    assert np.array_equal(v.cpu().numpy(), rv)
    # watertight: every undirected edge is shared by exactly two triangles
    e = np.concatenate([rf[:, [0, 1]], rf[:, [1, 2]], rf[:, [2, 0]]])
    _, c = np.unique(np.sort(e, 1), axis=0, return_counts=True)
    assert (c == 2).all()


def test_marching_cubes_plain_branch_above_256():
    """final.shape[0] > 256 -> the PyMCubes branch of export_mesh (float64 verts, no padding)."""
    dev = _cuda()
    from icon_b200 import ops
    from oracle import mcubes as OM
    R = 259                                     # final = 258^3 > 256
    a = torch.linspace(-1, 1, R, dtype=torch.float64)
    z, y, x = torch.meshgrid(a, a, a, indexing="ij")
    occ = (0.5 + 2.0 * (0.6 - torch.sqrt((x / 0.5) ** 2 + (y / 0.8) ** 2 + (z / 0.3) ** 2))).float()
    v, f = ops.marching_cubes(occ.to(dev), 0.5)
    rv, rf = OM.export_mesh(occ.numpy(), 0.5)
    assert v.dtype == torch.float64
    assert np.array_equal(f.cpu().numpy(), rf)
    assert np.array_equal(v.cpu().numpy(), rv)


def test_marching_cubes_lexicographic_order_contract():
    """order='lex' = triangle soup + unique(dim=0) numbering (the order SURVEY 8c attributes to kaolin): same surface,
    vertices sorted by coordinate row with coincident ones merged; index-exact vs the oracle's restatement.  The
    field has a node exactly at the iso value so that the merge really collapses vertices."""
    dev = _cuda()
    from icon_b200 import ops
    from oracle import mcubes as OM
    R = 41
    a = torch.linspace(-1, 1, R)
    z, y, x = torch.meshgrid(a, a, a, indexing="ij")
    occ = (0.5 + (0.6 - torch.sqrt(x * x + y * y + z * z))).float()
    occ[20, 20, 32] = 0.5                        # a grid node ON the level set: up to 3 edge vertices coincide there
    v, f = ops.marching_cubes(occ.to(dev), 0.5, order="lex")
    rv, rf = OM.export_mesh(occ.numpy(), 0.5, order="lex")
    assert np.array_equal(f.cpu().numpy(), rf)
    assert np.array_equal(v.cpu().numpy(), rv)
    ve, fe = ops.marching_cubes(occ.to(dev), 0.5)
    assert len(v) <= len(ve) and f.shape == fe.shape
    assert torch.equal(v[f], ve[fe])             # identical triangles, vertex for vertex


def test_end_to_end_engine_with_network_vs_oracle():
    """filter-less icon config: engine + fused query + marching cubes vs the CPU oracle chain."""
    dev = _cuda()
    from icon_b200 import config, net
    from icon_b200.engine import Seg3dLossless
    from oracle import query as OQ
    from oracle import mcubes as OM
    from oracle.engine import Seg3dOracle
    cfg = config.preset("icon-nofilter")
    netG = net.HGPIFuNet(cfg).to(dev).eval()
    sd = S.mlp_state_dict(c0=10, seed=21)
    # bias the last layer so that the 0.5 level set crosses the volume
    sd["filters.3.bias"] = sd["filters.3.bias"] + 0.5
    netG.if_regressor.load_state_dict(sd)
    v, f = S.body_mesh(rings=20, segs=24, seed=3)
    cm, vi = S.body_attributes(v, seed=3)
    verts, faces = torch.from_numpy(v)[None], torch.from_numpy(f)[None]
    cmap, vis = torch.from_numpy(cm)[None], torch.from_numpy(vi)[None]
    gen = torch.Generator().manual_seed(4)
    nF = torch.randn(1, 3, 64, 64, generator=gen)
    nB = torch.randn(1, 3, 64, 64, generator=gen)
    batch = {"normal_F": nF.to(dev), "normal_B": nB.to(dev), "smpl_verts": verts.to(dev),
             "smpl_faces": faces.to(dev), "smpl_cmap": cmap.to(dev), "smpl_vis": vis.to(dev)}
    res = [9, 17, 33]
    eng = Seg3dLossless(query_func=net.query_func, b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]],
                        resolutions=res, align_corners=True, balance_value=0.5, faster=True).to(dev)
    with torch.no_grad():
        features = netG.filter(batch)
        occ = eng(opt=cfg, netG=netG, features=features, proj_matrix=None)
    assert occ is not None
    smpl = {"smpl_verts": verts, "smpl_faces": faces, "smpl_cmap": cmap, "smpl_vis": vis}
    feat_cpu = torch.cat([nF, nB], 1)
    ora = Seg3dOracle([[-1.0, 1.0, -1.0]], [[1.0, -1.0, 1.0]], res)
    ref = ora.forward(lambda p: OQ.query_func(sd, [feat_cpu], p, prior="icon", smpl=smpl))
    assert ref is not None
    assert [int(c.shape[0]) for c in ora.log] == eng.last_query_counts
    assert (occ.cpu() - ref).abs().max() <= 1e-4
    verts_o, faces_o = eng.export_mesh(occ)
    assert verts_o.device.type == "cpu" and faces_o.dtype == torch.int64
    rv, rf = OM.export_mesh(occ.cpu().numpy(), 0.5)
    assert np.array_equal(faces_o.numpy(), rf)
    assert np.array_equal(verts_o.numpy(), rv)
