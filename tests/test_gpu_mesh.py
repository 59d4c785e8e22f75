"""clean_mesh on the device (csrc/clean.cu) against the oracle's restatement of trimesh's split (oracle/mesh.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _blobs(R, centres, radii):
    a = torch.linspace(-1, 1, R)
    z, y, x = torch.meshgrid(a, a, a, indexing="ij")
    occ = torch.zeros(R, R, R)
    for (cx, cy, cz), r in zip(centres, radii):
        occ = torch.maximum(occ, 0.5 + 2.0 * (r - torch.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2)))
    return occ.float()


@pytest.mark.parametrize("R", [65, 259])
def test_clean_mesh_keeps_the_largest_component(R):
    dev = _cuda()
    from icon_b200 import mesh, ops
    from oracle import mesh as OMesh
    occ = _blobs(R, [(-0.45, -0.4, 0.0), (0.35, 0.3, 0.1), (0.0, -0.7, -0.6), (0.7, -0.7, 0.7)], [0.3, 0.42, 0.15, 0.08])
    v, f = ops.marching_cubes(occ.to(dev), 0.5)
    cv, cf = mesh.clean_mesh_device(v, f)
    rv, rf = OMesh.clean_mesh(v.cpu().numpy(), f.cpu().numpy())
    assert OMesh.face_components(f.cpu().numpy())[0] == 4
    assert cv.dtype == torch.float32 and cf.dtype == torch.int32
    assert np.array_equal(cf.cpu().numpy(), rf) and np.array_equal(cv.cpu().numpy(), rv)
    assert 0 < len(cv) < len(v)
    # the kept component is closed and it is the big sphere
    e = np.concatenate([rf[:, [0, 1]], rf[:, [1, 2]], rf[:, [2, 0]]])
    _, c = np.unique(np.sort(e, 1), axis=0, return_counts=True)
    assert (c == 2).all()


def test_clean_mesh_drop_in_signature_and_device_shortcut():
    dev = _cuda()
    from icon_b200 import mesh
    from icon_b200.engine import Seg3dLossless
    from oracle import mesh as OMesh

    def field(points, **kw):
        p = points[0]
        a = 0.5 + 2.0 * (0.45 - (p - torch.tensor([0.2, 0.1, 0.0], device=p.device)).norm(dim=1))
        b = 0.5 + 2.0 * (0.2 - (p + torch.tensor([0.6, 0.6, 0.5], device=p.device)).norm(dim=1))
        return torch.maximum(a, b).view(1, 1, -1)

    eng = Seg3dLossless(query_func=field, b_min=[[-1.0, 1.0, -1.0]], b_max=[[1.0, -1.0, 1.0]], resolutions=[17, 33, 65],
                        align_corners=True, balance_value=0.5, faster=True).to(dev)
    verts, faces = eng.export_mesh(eng())
    assert verts.device.type == "cpu" and hasattr(verts, "_icon_device_mesh")
    cv, cf = mesh.clean_mesh(verts, faces)                         # reference call: clean_mesh(verts_pr, faces_pr)
    assert cv.device.type == "cpu" and cv.dtype == torch.float32 and cf.dtype == torch.int32
    rv, rf = OMesh.clean_mesh(verts.numpy(), faces.numpy())
    assert np.array_equal(cf.numpy(), rf) and np.array_equal(cv.numpy(), rv)
    cv2, cf2 = mesh.clean_mesh(verts.clone(), faces.clone())       # plain CPU tensors (no tag): upload path
    assert torch.equal(cv2, cv) and torch.equal(cf2, cf)
