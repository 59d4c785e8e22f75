"""get_visibility (producer of smpl_vis): oracle known-answer tests on the CPU, CUDA z-buffer
against the oracle on the GPU.  pytorch3d is absent -> parity unpinned (oracle/visibility.py header)."""
import numpy as np
import pytest
import torch

from icon_b200 import synthetic as S


def _two_quads():
    """A near square (z = 0.3 after the screen transform) hiding the centre of a far one (z = 0.7)."""
    def quad(h, z, base):
        v = np.array([[-h, -h, z], [h, -h, z], [h, h, z], [-h, h, z]], np.float32)
        f = np.array([[0, 1, 2], [0, 2, 3]], np.int64) + base
        return v, f
    v0, f0 = quad(0.2, 0.6, 0)      # get_visibility negates z: screen z = (1 - 0.6) / 2 = 0.2  (near)
    v1, f1 = quad(0.1, -0.2, 4)     # screen z = 0.6 (far), fully behind the near quad
    v2, f2 = quad(0.05, 0.9, 8)     # screen z = 0.05 (nearest) but shifted away below
    v2[:, 0] += 0.6
    return np.concatenate([v0, v1, v2]), np.concatenate([f0, f1, f2])


def _oriented(v, f):
    """flip faces so that the oracle's front-face convention (signed screen area >= 0) holds for all of them"""
    xyz = (np.concatenate([v[:, :2], -v[:, 2:3]], 1) + 1) / 2
    a, b, c = xyz[f[:, 0]], xyz[f[:, 1]], xyz[f[:, 2]]
    area = (a[:, 0] - b[:, 0]) * (c[:, 1] - b[:, 1]) - (a[:, 1] - b[:, 1]) * (c[:, 0] - b[:, 0])
    f = f.copy()
    f[area < 0] = f[area < 0][:, [0, 2, 1]]
    return f


def test_oracle_occlusion_backface_and_last_face_rule():
    from oracle import visibility as OV
    v, f = _two_quads()
    f = _oriented(v, f)
    vis = OV.get_visibility(v[:, :2], v[:, 2:3], f, image_size=128)[:, 0]
    assert vis[:4].tolist() == [1, 1, 1, 1]          # near quad
    assert vis[4:8].tolist() == [0, 0, 0, 0]         # hidden behind it
    assert vis[8:].tolist() == [1, 1, 1, 1]          # separate quad
    # back faces are culled: flipping the near quad uncovers the far one
    f2 = f.copy(); f2[:2] = f2[:2][:, [0, 2, 1]]
    vis2 = OV.get_visibility(v[:, :2], v[:, 2:3], f2, image_size=128)[:, 0]
    assert vis2[:4].tolist() == [0, 0, 0, 0] and vis2[4:8].tolist() == [1, 1, 1, 1]
    # faces[-1] rule: hide the last quad behind the near one -> still marked through the background index -1
    v3 = v.copy(); v3[8:, 0] -= 0.6; v3[8:, 2] = -0.5
    vis3 = OV.get_visibility(v3[:, :2], v3[:, 2:3], f, image_size=128)[:, 0]
    assert set(np.nonzero(vis3[8:])[0].tolist()) == set((f[-1] - 8).tolist())


def test_oracle_closed_body_front_half_is_visible():
    from oracle import visibility as OV
    verts, faces = S.body_mesh(rings=20, segs=24)
    faces = _oriented(verts, faces)
    vis = OV.get_visibility(verts[:, :2], verts[:, 2:3], faces, image_size=256)[:, 0]
    assert 0.25 < vis.mean() < 0.8
    # all faces oriented to the front here, so only occlusion hides vertices: the nearest ones must be visible
    zs = -verts[:, 2]
    assert vis[zs < np.quantile(zs, 0.2)].mean() > 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("S_img,rings,segs", [(256, 20, 24), (4096, 82, 84)])
def test_visibility_vs_oracle(S_img, rings, segs):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from icon_b200.visibility import get_visibility
    from oracle import visibility as OV
    verts, faces = S.body_mesh(rings=rings, segs=segs)         # outward-oriented closed mesh: half the faces are culled
    ref = OV.get_visibility(verts[:, :2], verts[:, 2:3], faces, image_size=S_img)
    out = get_visibility(torch.from_numpy(verts[:, :2]), torch.from_numpy(verts[:, 2:3]), torch.from_numpy(faces),
                         image_size=S_img)
    assert out.shape == (len(verts), 1) and out.device.type == "cpu"
    assert np.array_equal(out.numpy(), ref), f"{(out.numpy() != ref).sum()} vertices differ"
    assert 0.2 < ref.mean() < 0.8


@pytest.mark.gpu
def test_visibility_quads_and_cpu_inputs():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from icon_b200.visibility import get_visibility
    from oracle import visibility as OV
    v, f = _two_quads()
    f = _oriented(v, f)
    out = get_visibility(torch.from_numpy(v[:, :2]), torch.from_numpy(v[:, 2:3]), torch.from_numpy(f), image_size=128)
    assert np.array_equal(out.numpy(), OV.get_visibility(v[:, :2], v[:, 2:3], f, image_size=128))
