"""Mesh post-processing right after export_mesh, on the device.

  clean_mesh   lib/dataset/mesh_util.py:778-791 (trimesh split + component with the most vertices), called at
               apps/ICON.py:614-615, 755-756.  Same signature: (verts, faces) -> (verts float32, faces int32) on
               verts.device.  Seg3dLossless.export_mesh tags the CPU tensors it returns with the device copies they
               came from, so the usual `clean_mesh(*reconEngine.export_mesh(sdf))` sequence does no H2D upload.
"""
import torch

from . import _C
from ._C import check, lib
from .ops import _p, _stream


def clean_mesh_device(verts, faces):
    """CUDA (verts [nv,3] f32|f64, faces [nf,3] i64) -> CUDA (verts f32, faces i32): largest component."""
    if not verts.is_cuda:
        raise _C.IconError("clean_mesh_device needs CUDA tensors: there is no CPU path")
    v = verts.detach().contiguous()
    f = faces.detach().to(torch.int64).contiguous()
    nv, nf = v.shape[0], f.shape[0]
    if nv == 0 or nf == 0:
        return v.float(), f.int()
    if v.dtype not in (torch.float32, torch.float64):
        v = v.float()
    nbytes = lib.icon_clean_mesh_workspace_bytes(nv, nf)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=v.device)
    counts = torch.zeros(2, dtype=torch.int64, device=v.device)
    check(lib.icon_clean_mesh_count(_p(f), nv, nf, _p(ws), nbytes, _p(counts), _stream()), "icon_clean_mesh_count")
    kv, kf = (int(c) for c in counts.tolist())
    out_v = torch.empty(kv, 3, dtype=torch.float32, device=v.device)
    out_f = torch.empty(kf, 3, dtype=torch.int32, device=v.device)
    check(lib.icon_clean_mesh_emit(_p(v), 1 if v.dtype == torch.float64 else 0, _p(f), nv, nf, _p(ws), _p(out_v),
                                   _p(out_f), _stream()), "icon_clean_mesh_emit")
    return out_v, out_f


def clean_mesh(verts, faces):
    """Drop-in for lib.dataset.mesh_util.clean_mesh: results live on verts.device (the reference's contract)."""
    device = verts.device
    dev_copy = getattr(verts, "_icon_device_mesh", None)
    if dev_copy is not None and dev_copy[0].shape == verts.shape and dev_copy[1].shape == faces.shape:
        v, f = dev_copy                                   # what export_mesh downloaded: still on the GPU
    elif verts.is_cuda:
        v, f = verts, faces
    else:
        if not torch.cuda.is_available():
            raise _C.IconError("clean_mesh: no CUDA device (icon_b200 has no CPU path)")
        v, f = verts.cuda(), faces.cuda()
    out_v, out_f = clean_mesh_device(v, f)
    return out_v.to(device), out_f.to(device)
