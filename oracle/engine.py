"""oracle/engine.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU torch restatement of the reference's coarse-to-fine reconstruction engine in the only
live mode (`faster=True`, `align_corners=True`):

  lib/common/seg3d_lossless.py:37-123   Seg3dLossless.__init__
  lib/common/seg3d_lossless.py:125-144  batch_eval
  lib/common/seg3d_lossless.py:152-265  _forward_faster
  lib/common/seg3d_utils.py:122-136     create_grid3D
  lib/common/seg3d_utils.py:169-181     SmoothConv3D

Pinned against the reference class itself imported live (tests/golden/make_golden.py):
identical per-level query coordinate sets and identical final grid on an analytic field.
"""
import torch
import torch.nn.functional as F


def create_grid3D(lo, hi, steps):
    ar = torch.linspace(lo, hi, steps).long()
    gridD, gridH, gridW = torch.meshgrid([ar, ar, ar], indexing="ij")
    coords = torch.stack([gridW, gridH, gridD])
    return coords.view(3, -1).t()


def _smooth(mask_f, k):
    w = torch.ones((1, 1, k, k, k), dtype=torch.float32) / (k ** 3)
    return F.conv3d(mask_f, w, padding=(k - 1) // 2)


class Seg3dOracle:
    def __init__(self, b_min, b_max, resolutions, balance_value=0.5):
        self.b_min = torch.tensor(b_min).float().unsqueeze(1)
        self.b_max = torch.tensor(b_max).float().unsqueeze(1)
        self.resolutions = torch.tensor([(r, r, r) for r in resolutions])
        self.balance_value = balance_value
        self.init_coords = create_grid3D(0, int(self.resolutions[-1][0]) - 1,
                                         int(self.resolutions[0][0])).unsqueeze(0)
        self.log = []            # per level: coords evaluated (torch tensors [n,3])

    def batch_eval(self, coords, query_fn):
        coords2D = coords.float() / (self.resolutions[-1] - 1)
        coords2D = coords2D * (self.b_max - self.b_min) + self.b_min
        occ = query_fn(coords2D)
        assert occ.dim() == 3
        return occ

    def forward(self, query_fn):
        self.log = []
        for resolution in self.resolutions:
            W, H, D = [int(v) for v in resolution]
            stride = (self.resolutions[-1] - 1) / (resolution - 1)
            if torch.equal(resolution, self.resolutions[0]):
                coords = self.init_coords.clone()
                occupancys = self.batch_eval(coords, query_fn).view(1, 1, D, H, W)
                self.log.append(coords[0].clone())
                if (occupancys > 0.5).sum() == 0:
                    return None
                coords_accum = coords / stride
            elif torch.equal(resolution, self.resolutions[-1]):
                occupancys = F.interpolate(occupancys.float(), size=(D, H, W), mode="trilinear",
                                           align_corners=True)
            else:
                coords_accum *= 2
                valid = F.interpolate((occupancys > self.balance_value).float(), size=(D, H, W),
                                      mode="trilinear", align_corners=True)
                occupancys = F.interpolate(occupancys.float(), size=(D, H, W), mode="trilinear",
                                           align_corners=True)
                is_boundary = (valid > 0.0) & (valid < 1.0)
                if torch.equal(resolution, self.resolutions[1]):
                    k = 9
                elif torch.equal(resolution, self.resolutions[2]):
                    k = 7
                else:
                    k = 3
                is_boundary = (_smooth(is_boundary.float(), k) > 0)[0, 0]
                coords_accum = coords_accum.long()
                is_boundary[coords_accum[0, :, 2], coords_accum[0, :, 1], coords_accum[0, :, 0]] = False
                point_coords = is_boundary.permute(2, 1, 0).nonzero(as_tuple=False).unsqueeze(0)
                point_indices = (point_coords[:, :, 2] * H * W + point_coords[:, :, 1] * W +
                                 point_coords[:, :, 0])
                coords = point_coords * stride
                if coords.size(1) == 0:
                    continue
                occupancys_topk = self.batch_eval(coords, query_fn)
                self.log.append(coords[0].clone())
                point_indices = point_indices.unsqueeze(1)
                occupancys = occupancys.reshape(1, 1, D * H * W).scatter_(
                    2, point_indices, occupancys_topk).view(1, 1, D, H, W)
                voxels = coords / stride
                coords_accum = torch.cat([voxels, coords_accum], dim=1).unique(dim=1)
        return occupancys[0, 0]
