"""CPU restatement of the PaMIR semantic voxelisation -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference calls `voxelize_cuda.forward_semantic_voxelization(smpl_vertices,
smpl_vertex_code, smpl_tetrahedrons, occ_volume, semantic_volume, weight_sum_volume, sigma)`
(lib/net/voxelize.py:57-59); `voxelize_cuda` 0.0.0 (requirements.txt:34) is an un-vendored CUDA
extension whose source is not under /root/reference and which cannot be installed here, and the
reference holds no test or golden vector for it.  What is restated is what the call site and the
PaMIR paper determine:

* `occ_volume` [B,res,res,res] starts at 0, `semantic_volume` [B,res,res,res,3] at 0 and
  `weight_sum_volume` at 1e-3 (voxelize.py:40-51); index order (z, y, x), returned permuted to
  (b, c, d, h, w) (voxelize.py:137);
* the volume spans [-0.5, 0.5]^3 (TestDataset.py:174 scales the tetra-SMPL vertices by 0.5);
  voxel i along an axis is centred at (i + 0.5) / res - 0.5;
* a voxel is occupied iff its centre lies in (or on the boundary of) some tetrahedron;
* an occupied voxel holds sum_v w_v code_v / (1e-3 + sum_v w_v) over the SURFACE vertices,
  w_v = exp(-|p - v|^2 / (2 sigma^2)); unoccupied voxels hold 0.

The occupancy arithmetic is float32 with one rounding per operation, in the order written here;
icon_b200/csrc/voxelize.cu follows it operation for operation (compiled without FMA contraction),
so occupancy is compared bit for bit; the Gaussian sums are accumulated in float64 here and
compared to ~1e-5.
"""
import numpy as np

F = np.float32


def _det3(u, v, w):
    c0 = v[..., 1] * w[..., 2] - v[..., 2] * w[..., 1]
    c1 = v[..., 0] * w[..., 2] - v[..., 2] * w[..., 0]
    c2 = v[..., 0] * w[..., 1] - v[..., 1] * w[..., 0]
    return (u[..., 0] * c0 - u[..., 1] * c1) + u[..., 2] * c2


def _orient(a, b, c, p):
    return _det3(b - a, c - a, p - a)


def occupancy(verts, tets, res):
    """verts [NV,3] f32, tets [NT,4] int -> uint8 [res,res,res] indexed (z, y, x)."""
    verts = np.asarray(verts, dtype=F)
    occ = np.zeros((res, res, res), dtype=np.uint8)
    inv = F(1.0) / F(res)
    half = F(0.5)
    for t in np.asarray(tets):
        v = verts[np.clip(t, 0, len(verts) - 1)]
        vol = _orient(v[0], v[1], v[2], v[3])
        if vol == 0:
            continue
        lo, hi = v.min(0), v.max(0)
        i0 = np.maximum(0, np.floor((lo + half) * F(res) - half).astype(np.int64))
        i1 = np.minimum(res - 1, np.ceil((hi + half) * F(res) - half).astype(np.int64))
        if np.any(i1 < i0):
            continue
        ax = [np.arange(i0[a], i1[a] + 1) for a in range(3)]
        X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
        p = np.stack([(X.astype(F) + half) * inv - half, (Y.astype(F) + half) * inv - half,
                      (Z.astype(F) + half) * inv - half], -1).astype(F)
        d0 = _orient(v[0], v[1], v[2], p)
        d1 = _orient(v[0], v[3], v[1], p)
        d2 = _orient(v[0], v[2], v[3], p)
        d3 = _orient(v[1], v[3], v[2], p)
        if vol > 0:
            inside = (d0 >= 0) & (d1 >= 0) & (d2 >= 0) & (d3 >= 0)
        else:
            inside = (d0 <= 0) & (d1 <= 0) & (d2 <= 0) & (d3 <= 0)
        occ[Z[inside], Y[inside], X[inside]] = 1
    return occ


def semantic_volume(verts, n_surface, codes, tets, res, sigma):
    """-> float32 [3,res,res,res] (c, z, y, x)."""
    verts = np.asarray(verts, dtype=F)
    occ = occupancy(verts, tets, res)
    out = np.zeros((3, res, res, res), dtype=F)
    zz, yy, xx = np.nonzero(occ)
    if len(zz) == 0:
        return out, occ
    inv = F(1.0) / F(res)
    half = F(0.5)
    p = np.stack([(xx.astype(F) + half) * inv - half, (yy.astype(F) + half) * inv - half,
                  (zz.astype(F) + half) * inv - half], -1).astype(np.float64)
    sv = verts[:n_surface].astype(np.float64)
    cd = np.asarray(codes, dtype=np.float64)
    k = -1.0 / (2.0 * float(sigma) ** 2)
    step = 1024
    for s in range(0, len(p), step):
        q = p[s:s + step]
        d2 = (q[:, None, 0] - sv[None, :, 0]) ** 2
        d2 += (q[:, None, 1] - sv[None, :, 1]) ** 2
        d2 += (q[:, None, 2] - sv[None, :, 2]) ** 2
        w = np.exp(d2 * k)
        val = (w @ cd) / (1e-3 + w.sum(1, keepdims=True))
        out[:, zz[s:s + step], yy[s:s + step], xx[s:s + step]] = val.T.astype(F)
    return out, occ
