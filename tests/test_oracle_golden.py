"""The oracle (oracle/) pinned against fixtures produced by the reference's own modules
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import torch

from icon_b200 import synthetic as S
from oracle import engine as OE
from oracle import query as OQ


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_mlp_matches_reference_module(golden_dir):
    g = _load(golden_dir, "mlp_index.npz")
    for c0 in (13, 10):
        sd = S.mlp_state_dict(c0=c0, seed=3)
        y = OQ.mlp_forward(sd, torch.from_numpy(g[f"mlp{c0}_x"]))
        assert np.abs(y.numpy() - g[f"mlp{c0}_y"]).max() <= 1e-6


def test_index_and_orthogonal_match_reference(golden_dir):
    g = _load(golden_dir, "mlp_index.npz")
    o2 = OQ.index(torch.from_numpy(g["index2d_feat"]), torch.from_numpy(g["index2d_uv"]))
    assert np.array_equal(o2.numpy(), g["index2d_out"])
    o3 = OQ.index(torch.from_numpy(g["index3d_feat"]), torch.from_numpy(g["index3d_uv"]))
    assert np.array_equal(o3.numpy(), g["index3d_out"])
    oo = OQ.orthogonal(torch.from_numpy(g["ortho_pts"]), torch.from_numpy(g["ortho_calib"]))
    assert np.array_equal(oo.numpy(), g["ortho_out"])


def _field(points):
    s = torch.tensor([0.45, 0.8, 0.3])
    r = (points[0] / s).norm(dim=1)
    bump = 0.15 * torch.sin(9.0 * points[0, :, 0]) * torch.cos(7.0 * points[0, :, 1])
    return (0.5 + 2.0 * (0.8 - r) + bump).view(1, 1, -1)


def test_engine_matches_reference_engine(golden_dir):
    g = _load(golden_dir, "engine.npz")
    for tag in ("a", "b"):
        res = [int(r) for r in g[f"{tag}_res"]]
        eng = OE.Seg3dOracle([[-1.0, 1.0, -1.0]], [[1.0, -1.0, 1.0]], res)
        calls = []

        def q(p):
            calls.append(p.clone())
            return _field(p)

        occ = eng.forward(q)
        assert len(calls) == int(g[f"{tag}_ncalls"])
        for i, p in enumerate(calls):
            assert np.array_equal(p.numpy(), g[f"{tag}_pts{i}"]), (tag, i)
        # the analytic field's sin/norm are not bit-reproducible run to run (vectorisation);
        # the engine itself adds no arithmetic beyond trilinear interpolation
        assert np.abs(occ.numpy() - g[f"{tag}_occ"]).max() <= 2e-6


def test_sdf_c_oracle_agrees_with_torch_transcription():
    v, f = S.body_mesh(rings=10, segs=12, seed=1)
    cm, vi = S.body_attributes(v)
    verts, faces = torch.from_numpy(v)[None], torch.from_numpy(f)[None]
    cmap, vis = torch.from_numpy(cm)[None], torch.from_numpy(vi)[None]
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(1, 4000, 3, generator=g) * 2 - 1
    a = OQ.cal_sdf_batch_c(verts, faces, cmap, vis, pts, return_face=True)
    b = OQ.cal_sdf_batch_torch(verts, faces, cmap, vis, pts)
    # distance and sign agree everywhere; the winning face agrees except on float-level ties
    assert (a[0] - b[0]).abs().max() < 1e-6
    same = a[4] == b[4]
    assert same.float().mean() > 0.95
    assert (a[1][0][same] - b[1][0][same]).abs().max() < 1e-4
    assert (a[2][0][same] - b[2][0][same]).abs().max() < 1e-4
    assert (a[3][0][same] == b[3][0][same]).float().mean() > 0.999


def test_sdf_known_answer_sphere():
    """Analytic known-answer: unit-ish sphere, distance / sign / normal direction."""
    v, f = S.body_mesh(rings=40, segs=48, seed=0, extent=(0.6, 0.6, 0.6))
    # make it an exact sphere of radius 0.6
    v = (v / np.linalg.norm(v, axis=1, keepdims=True) * 0.6).astype(np.float32)
    cm, vi = S.body_attributes(v)
    g = torch.Generator().manual_seed(2)
    pts = torch.rand(1, 3000, 3, generator=g) * 2 - 1
    sdf, norm, _, _ = OQ.cal_sdf_batch_c(torch.from_numpy(v)[None], torch.from_numpy(f)[None],
                                         torch.from_numpy(cm)[None], torch.from_numpy(vi)[None], pts)
    r = pts[0].norm(dim=1)
    expect = (0.6 - r) / np.sqrt(3.0)             # inside positive, divided by sqrt(3)
    assert (sdf[0, :, 0] - expect).abs().max() < 2e-3       # faceting error of the polyhedron
    assert ((sdf[0, :, 0] > 0) == (r < 0.6)).float().mean() > 0.995
    # normal = interpolated outward vertex normal * (-1, 1, -1)
    n = norm[0] * torch.tensor([-1.0, 1.0, -1.0])
    cos = (n * pts[0] / r[:, None]).sum(1) / n.norm(dim=1)
    assert cos.min() > 0.98


def test_outlier_cmap_rule_is_order_dependent():
    """k-th outlier's channel c receives s[(3k+c) mod K]."""
    sdf = torch.tensor([[[0.2], [0.01], [-0.3], [0.4], [-0.02], [-0.5]]])
    cmap = torch.arange(18, dtype=torch.float32).view(1, 6, 3) / 100.0
    out = sdf.abs().ge(0.05)
    sdf2 = sdf.clone()
    sdf2[out] = torch.sign(sdf2[out])
    c2 = cmap.clone()
    c2[out.repeat(1, 1, 3)] = sdf2[out].repeat(1, 1, 3)
    s = sdf2[out]                # [1,-1,1,-1]
    K = len(s)
    k = 0
    for i in range(6):
        if out[0, i, 0]:
            for c in range(3):
                assert c2[0, i, c] == s[(3 * k + c) % K]
            k += 1
        else:
            assert torch.equal(c2[0, i], cmap[0, i])


def test_sdf_oracle_tie_rule_lowest_face_index():
    """kaolin's brute-force scan keeps the first strict minimum, i.e. the LOWEST face index on exact ties
    (SURVEY 8c).  Outside a tetrahedron most points are nearest to an edge or a vertex, where the incident faces
    tie exactly; per-face distances from single-face meshes say which faces tie."""
    v = np.array([[0.0, 0.0, 0.5], [0.5, 0.0, -0.25], [-0.25, 0.4, -0.25], [-0.25, -0.4, -0.25]], np.float32)
    tet = np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 3, 2]], np.int64)
    g = torch.Generator().manual_seed(4)
    pts = torch.rand(1, 2000, 3, generator=g) * 2 - 1
    cm = np.zeros_like(v); vi = np.ones((4, 1), np.float32)

    def run(faces):
        out = OQ.cal_sdf_batch_c(torch.from_numpy(v)[None], torch.from_numpy(faces)[None], torch.from_numpy(cm)[None],
                                 torch.from_numpy(vi)[None], pts, return_face=True)
        return out[0][0, :, 0].abs().numpy(), out[4].numpy()

    D = np.stack([run(tet[k:k + 1])[0] for k in range(4)])            # [4, N] |sdf| to each face alone
    dist0, face0 = run(tet)
    assert np.array_equal(dist0, D.min(0))
    assert np.array_equal(D[face0, np.arange(D.shape[1])], D.min(0))   # the winner attains the minimum
    near_ties = (D == D.min(0, keepdims=True)).sum(0) > 1             # equal after sqrt: edge / vertex regions
    assert near_ties.mean() > 0.2
    for k in range(4):
        # a bit-identical twin of face k placed FIRST takes over every point face k used to win (lower index) ...
        _, face1 = run(np.concatenate([tet[k:k + 1], tet], 0))
        won = face0 == k
        assert won.any() and (face1[won] == 0).all()
        # ... and may take a point from another face only where that face tied with k
        moved = (~won) & (face1 != face0 + 1)
        assert (face1[moved] == 0).all() and (D[k, moved] == D.min(0)[moved]).all()
        # placed LAST it never wins: only a strictly smaller distance replaces the current best
        _, face2 = run(np.concatenate([tet, tet[k:k + 1]], 0))
        assert np.array_equal(face2, face0)


def test_oracle_outlier_context_equals_full_call():
    """oracle.query(outlier_context=...) on a subset == the full call restricted to the subset: the restatement
    of HGPIFuNet.py:303-304 used by the dense-256^3 GPU parity test."""
    from icon_b200 import synthetic as S
    from oracle import query as OQ
    sd = S.mlp_state_dict(c0=13, seed=0)
    v, f = S.body_mesh(rings=20, segs=24, seed=0)
    cm, vi = S.body_attributes(v, seed=0)
    smpl = {"smpl_verts": torch.from_numpy(v)[None], "smpl_faces": torch.from_numpy(f)[None],
            "smpl_cmap": torch.from_numpy(cm)[None], "smpl_vis": torch.from_numpy(vi)[None]}
    feat = S.feature_map(12, 32, seed=0)
    pts = S.lattice_points(24)
    full = OQ.query_func(sd, [feat], pts, prior="icon", smpl=smpl)
    sdf = OQ.cal_sdf_batch_c(smpl["smpl_verts"], smpl["smpl_faces"], smpl["smpl_cmap"], smpl["smpl_vis"], pts)[0][0, :, 0]
    out = sdf.abs() >= 0.05
    assert 0.5 < out.float().mean() < 1.0
    signs, rank = torch.sign(sdf[out]), torch.cumsum(out.long(), 0) - 1
    idx = torch.arange(0, pts.shape[1], 7)
    sub = OQ.query_func(sd, [feat], pts[:, idx].contiguous(), prior="icon", smpl=smpl,
                        outlier_context=(signs, rank[idx]))
    assert (full[:, :, idx] - sub).abs().max() <= 2e-6


def test_c_oracle_equals_torch_transcription_on_real_scan_body(golden_dir):
    """The two restatements of cal_sdf_batch (oracle/sdf_oracle.c and the line-by-line torch transcription) agree
    on the decimated THuman2 scan (non-manifold, slivers, zero-area faces) at adversarial points: same nearest
    face, same sign, sdf / attributes to float rounding."""
    import os
    import numpy as np
    from icon_b200 import synthetic as S
    from oracle import query as OQ
    g = np.load(os.path.join(golden_dir, "scan_body.npz"))
    v, f = g["verts"], g["faces"].astype(np.int64)
    cm, vi = S.body_attributes(v, seed=3)
    verts, faces = torch.from_numpy(v)[None], torch.from_numpy(f)[None]
    cmap, vis = torch.from_numpy(cm)[None], torch.from_numpy(vi)[None]
    pts = S.adversarial_points(v, f, n_each=40, seed=7)
    sdf, norm, cmo, vo, face = OQ.cal_sdf_batch_c(verts, faces, cmap, vis, pts, return_face=True)
    sdf_t, norm_t, cm_t, vis_t, face_t = OQ.cal_sdf_batch_torch(verts, faces, cmap, vis, pts)
    same = face == face_t.int()
    # the vectorised torch walk evaluates every region formula for every pair, so float-level ties may resolve to a
    # different equidistant face; what must agree everywhere: distance and sign
    assert same.float().mean() > 0.9
    off = sdf_t.abs() > 1e-5          # ray origins ON the surface: t > 0 is decided by the last bit, either answer is "right"
    assert off.float().mean() > 0.3
    assert torch.equal(torch.sign(sdf)[off], torch.sign(sdf_t)[off])
    assert (sdf - sdf_t).abs().max() <= 1e-6
    assert (cmo[0][same] - cm_t[0][same]).abs().max() <= 1e-4
    assert torch.equal(vo[0][same], vis_t[0][same])


def test_stock_torch_encoder_restatement_matches_reference_goldens(golden_dir):
    """tools/torch_encoders.py (stock torch ops on icon_b200's parameter containers = the cuDNN baseline of bench.py)
    against the outputs of the reference's own HGFilter / GlobalGenerator (tests/golden/encoders.npz, 64 x 64)."""
    import os
    import numpy as np
    from icon_b200 import config, synthetic as S
    from icon_b200.encoders import GlobalGenerator, HGFilter
    from tools import torch_encoders as OE
    g = np.load(os.path.join(golden_dir, "encoders.npz"))
    hg = HGFilter(config.preset("icon-filter").net, 2, 3)
    hg.load_state_dict(S.seeded_like(hg.state_dict(), 21)); hg.eval()
    gg = GlobalGenerator(6, 3, 64, 4, 9)
    gg.load_state_dict(S.seeded_like(gg.state_dict(), 22)); gg.eval()
    with torch.no_grad():
        y = OE.hgfilter(hg, torch.from_numpy(g["hg_x"]))[-1]
        y6 = OE.global_generator(gg, torch.from_numpy(g["gg_x"]))
    assert np.abs(y.numpy() - g["hg_y"]).max() <= 1e-5
    assert np.abs(y6.numpy() - g["gg_y"]).max() <= 1e-5


def test_oracle_clean_mesh_on_two_tetrahedra():
    """Known answer: two disjoint closed tetrahedra (4 and 4 vertices -> first maximum = the first), plus a fin
    attached along a 3-face (non-manifold) edge, which trimesh's face adjacency ignores."""
    import numpy as np
    from oracle import mesh as OMesh
    tet = np.array([[0, 1, 2], [0, 3, 1], [1, 3, 2], [0, 2, 3]])
    verts = np.random.RandomState(0).rand(9, 3)
    faces = np.concatenate([tet + 4, tet, [[0, 1, 8]]])          # tetra B first in face order, then A, then the fin
    v, f = OMesh.clean_mesh(verts, faces)
    assert OMesh.face_components(faces)[0] == 3                  # the fin is its own component: edge (0,1) has 3 faces
    # both tetrahedra have 4 vertices; A keeps edge (0,1) out of its adjacency but stays connected through the rest
    assert len(v) == 4 and len(f) == 4 and f.dtype == np.int32 and v.dtype == np.float32
    assert np.array_equal(v, verts[4:8].astype(np.float32))      # first maximum in component (= face) order: B
