"""Vertex visibility of the body mesh: mirror of `lib/dataset/mesh_util.py::get_visibility` (reference :280-316),
the producer of the query path's `smpl_vis` input (TestDataset.compute_vis_cmap, TestDataset.py:134-148).

The reference builds a pytorch3d `Meshes`, rasterises at 4096^2 and runs two `torch.unique`; here the screen
transform is one torch expression and the rest is `icon_visibility` (csrc/visibility.cu): a 64-bit
(depth, face) z-buffer filled with atomicMin by one warp per face, then one pass marking the owners' vertices.
"""
import torch

from . import ops


def get_visibility(xy, z, faces, image_size=2 ** 12):
    """xy [N,2], z [N,1], faces [F,3] -> vis_mask float32 [N,1] on the CPU (as the reference returns it).
    Inputs on the CPU are moved to the current CUDA device; there is no CPU implementation."""
    xy, z = torch.as_tensor(xy), torch.as_tensor(z)
    dev = xy.device if xy.is_cuda else torch.device("cuda", torch.cuda.current_device())
    xyz = torch.cat((xy.to(dev).float(), -z.to(dev).float().reshape(-1, 1)), dim=1)
    xyz = (xyz + 1.0) / 2.0
    vis = ops.visibility(xyz, torch.as_tensor(faces).to(dev).long(), image_size)
    return vis.reshape(-1, 1).cpu()
