"""CPU restatement of get_visibility (lib/dataset/mesh_util.py:280-316) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference calls pytorch3d.renderer.mesh.rasterize_meshes (requirements.txt:33,
git HEAD, not installable here, source not under /root/reference) with the settings of
lib/common/render_utils.py:178-186 (blur 0, 1 face per pixel, perspective_correct, cull_backfaces) at
image_size 2**12, then `vis[unique(faces[unique(pix_to_face)])] = 1`.  Restated here:

* screen vertices xyz = (cat(xy, -z) + 1) / 2 (mesh_util.py:291-292);
* pytorch3d's pixel grid: pixel (yi, xi) is centred at NDC (1 - (2 xi + 1)/S, 1 - (2 yi + 1)/S);
* per face: skip if max z < 0, signed area e(v0, v1, v2) < 0 (back face) or |area| <= 1e-8, with
  e(p, a, b) = (p.x - a.x)(b.y - a.y) - (p.y - a.y)(b.x - a.x);
* coverage: barycentrics e(p, v1, v2), e(p, v2, v0), e(p, v0, v1) over (area + 1e-8) all > 0;
* depth: perspective-corrected weights (w0 z1 z2, z0 w1 z2, z0 z1 w2) / (sum + 1e-8), pz = sum w_i z_i,
  pz < 0 dropped; nearest pz wins, ties to the lowest face index;
* `unique(pix_to_face)` contains -1 whenever a pixel is empty and `faces[-1]` is the LAST face: its
  vertices are marked visible as well (bug-compatible).

All arithmetic is float32, one rounding per operation, in the order written; csrc/visibility.cu follows
it operation for operation (no FMA contraction), so pix_to_face and the mask are compared bit for bit.
"""
import numpy as np

F32 = np.float32


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def rasterize(xyz, faces, S):
    """-> (pix_to_face int64 [S,S] with -1 for background, zbuf float32 [S,S])."""
    xyz = np.asarray(xyz, dtype=F32)
    faces = np.asarray(faces, dtype=np.int64)
    p2f = np.full((S, S), -1, dtype=np.int64)
    zb = np.full((S, S), np.inf, dtype=F32)
    Sf = F32(S)
    one = F32(1.0)
    eps = F32(1e-8)
    for f, (a, b, c) in enumerate(faces):
        v0, v1, v2 = xyz[a], xyz[b], xyz[c]
        zmax = max(v0[2], v1[2], v2[2])
        area = _edge(v0[0], v0[1], v1[0], v1[1], v2[0], v2[1])
        if zmax < 0 or area < 0 or (-eps <= area <= eps):
            continue
        xmin, xmax = min(v0[0], v1[0], v2[0]), max(v0[0], v1[0], v2[0])
        ymin, ymax = min(v0[1], v1[1], v2[1]), max(v0[1], v1[1], v2[1])
        xi0 = max(int(np.floor(((1.0 - float(xmax)) * S - 1.0) * 0.5)) - 2, 0)
        xi1 = min(int(np.ceil(((1.0 - float(xmin)) * S - 1.0) * 0.5)) + 2, S - 1)
        yi0 = max(int(np.floor(((1.0 - float(ymax)) * S - 1.0) * 0.5)) - 2, 0)
        yi1 = min(int(np.ceil(((1.0 - float(ymin)) * S - 1.0) * 0.5)) + 2, S - 1)
        if xi1 < xi0 or yi1 < yi0:
            continue
        xs = np.arange(xi0, xi1 + 1)
        ys = np.arange(yi0, yi1 + 1)
        px = (one - (2 * xs + 1).astype(F32) / Sf)[None, :]
        py = (one - (2 * ys + 1).astype(F32) / Sf)[:, None]
        px, py = np.broadcast_arrays(px, py)
        inb = (px >= xmin) & (px <= xmax) & (py >= ymin) & (py <= ymax)
        den = area + eps
        w0 = _edge(px, py, v1[0], v1[1], v2[0], v2[1]) / den
        w1 = _edge(px, py, v2[0], v2[1], v0[0], v0[1]) / den
        w2 = _edge(px, py, v0[0], v0[1], v1[0], v1[1]) / den
        cov = inb & (w0 > 0) & (w1 > 0) & (w2 > 0)
        if not cov.any():
            continue
        z0, z1, z2 = v0[2], v1[2], v2[2]
        t0 = w0 * z1 * z2
        t1 = z0 * w1 * z2
        t2 = z0 * z1 * w2
        ds = (t0 + t1 + t2) + eps
        with np.errstate(divide="ignore", invalid="ignore"):
            pz = (t0 / ds) * z0 + (t1 / ds) * z1 + (t2 / ds) * z2
        cov &= pz >= 0
        sub_z = zb[yi0:yi1 + 1, xi0:xi1 + 1]
        sub_f = p2f[yi0:yi1 + 1, xi0:xi1 + 1]
        win = cov & (pz < sub_z)
        sub_z[win] = pz[win]
        sub_f[win] = f
    return p2f, zb


def get_visibility(xy, z, faces, image_size=4096):
    """mesh_util.py:280-316 -> float32 [N,1]."""
    xy = np.asarray(xy, dtype=F32)
    z = np.asarray(z, dtype=F32).reshape(-1, 1)
    xyz = (np.concatenate([xy, -z], 1) + F32(1.0)) / F32(2.0)
    faces = np.asarray(faces, dtype=np.int64)
    p2f, _ = rasterize(xyz, faces, image_size)
    ids = np.unique(faces[np.unique(p2f)])          # -1 -> last face, as in the reference
    vis = np.zeros((len(xy), 1), dtype=F32)
    vis[ids] = 1.0
    return vis
