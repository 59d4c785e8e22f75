"""PaMIR semantic voxelisation module: mirror of `lib/net/voxelize.py::Voxelization` (reference :66-137).

Same constructor, `update_param(batch_size, smpl_tetra)` and `forward(smpl_vertices)` ->
`[B, 3, res, res, res]` (b, c, d, h, w).  The reference tiles the constant tables with numpy, uploads
them and calls `voxelize_cuda.forward_semantic_voxelization` (source absent, see csrc/voxelize.cu) for
every query call; here the tables are uploaded once per `update_param` and the volume comes from
`icon_voxelize` in channel-first layout directly (no bzyxc -> bcdhw permute).  The face centre /
normal / face-code inputs of the reference's autograd Function are accepted but unused, as in its
forward (voxelize.py:57-59 passes only vertices, vertex codes and tetrahedra to the kernel).
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops


def read_smpl_constants(folder):
    """lib/dataset/mesh_util.py:240-263: vertex codes = rest vertices normalised to [0,1] per axis."""
    v = np.loadtxt(os.path.join(folder, "vertices.txt"))
    lo, hi = v.min(0, keepdims=True), v.max(0, keepdims=True)
    smpl_vertex_code = np.float32((v - lo) / (hi - lo))
    smpl_faces = np.loadtxt(os.path.join(folder, "faces.txt"), dtype=np.int32) - 1
    smpl_face_code = (smpl_vertex_code[smpl_faces[:, 0]] + smpl_vertex_code[smpl_faces[:, 1]] +
                      smpl_vertex_code[smpl_faces[:, 2]]) / 3.0
    smpl_tetras = np.loadtxt(os.path.join(folder, "tetrahedrons.txt"), dtype=np.int32) - 1
    return smpl_vertex_code, smpl_face_code, smpl_faces, smpl_tetras


class Voxelization(nn.Module):
    def __init__(self, smpl_vertex_code, smpl_face_code, smpl_face_indices, smpl_tetraderon_indices,
                 volume_res, sigma, smooth_kernel_size, batch_size, device):
        super().__init__()
        smpl_face_indices = np.asarray(smpl_face_indices)
        smpl_tetraderon_indices = np.asarray(smpl_tetraderon_indices)
        assert smpl_face_indices.ndim == 2 and smpl_face_indices.shape[1] == 3
        assert smpl_tetraderon_indices.ndim == 2 and smpl_tetraderon_indices.shape[1] == 4
        self.volume_res = volume_res
        self.sigma = sigma
        self.smooth_kernel_size = smooth_kernel_size
        self.batch_size = batch_size
        self.device = device
        self.smpl_vertex_code = np.asarray(smpl_vertex_code, dtype=np.float32)
        self.smpl_face_code = np.asarray(smpl_face_code, dtype=np.float32)
        self.smpl_face_indices = smpl_face_indices
        self.smpl_tetraderon_indices = smpl_tetraderon_indices
        self._codes = None
        self._tets = None

    def update_param(self, batch_size, smpl_tetra):
        """voxelize.py:89-118.  `smpl_tetra`: [NT,4] numpy array or tensor (device tensors are used in place)."""
        self.batch_size = batch_size
        dev = torch.device(self.device)
        if torch.is_tensor(smpl_tetra):
            tets = smpl_tetra.to(device=dev, dtype=torch.int32)
            self.smpl_tetraderon_indices = smpl_tetra
        else:
            self.smpl_tetraderon_indices = np.asarray(smpl_tetra)
            tets = torch.from_numpy(np.ascontiguousarray(self.smpl_tetraderon_indices, dtype=np.int32)).to(dev)
        self._tets = tets.contiguous()
        if self._codes is None or self._codes.device != dev:
            self._codes = torch.from_numpy(self.smpl_vertex_code).to(dev).contiguous()

    def forward(self, smpl_vertices):
        """smpl_vertices [B, NV, 3] CUDA fp32 (surface vertices first) -> [B, 3, res, res, res]."""
        assert smpl_vertices.size(0) == self.batch_size
        self.check_input(smpl_vertices)
        if self._tets is None:
            self.update_param(self.batch_size, self.smpl_tetraderon_indices)
        n_surf = self._codes.shape[0]
        vols = [ops.voxelize(smpl_vertices[b], n_surf, self._codes, self._tets, self.volume_res, self.sigma)
                for b in range(smpl_vertices.size(0))]
        return vols[0] if len(vols) == 1 else torch.cat(vols, 0)

    def check_input(self, x):
        if not x.is_cuda:
            raise TypeError("Voxelization module supports only cuda tensors")
        if x.dtype != torch.float32:
            raise TypeError("Voxelization module supports only float32 tensors")
