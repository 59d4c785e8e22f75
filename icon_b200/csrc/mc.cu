// Marching cubes with shared (edge-owned) vertices.
//
// Replaces Seg3dLossless.export_mesh (lib/common/seg3d_lossless.py:583-604): the
// occupancys[1:,1:,1:] crop, kaolin voxelgrids_to_trianglemeshes (zero pad, iso 0.5) or
// PyMCubes marching_cubes, and the [:, [2,1,0]] / [:, [0,2,1]] permutations -- everything
// stays on the device; only the finished vertex / face arrays travel to the host.
//
// Indexing contract (oracle/mcubes.py restates it; DESIGN.md "marching cubes"):
//   * working grid g[i][j][k] (i = z, j = y, k = x of the occupancy frame); padded branch:
//     G = R+1, g = occ for 1 <= i,j,k <= R-1 and 0 on the two outer shells (crop + zero pad);
//     plain branch: G = R-1, g[i][j][k] = occ[i+1][j+1][k+1];
//   * a vertex lives on a grid edge, owned by the edge's lower voxel; id order = ascending
//     3*linear(owner)+axis (axis 0 = i, 1 = j, 2 = k); position = owner + t along the axis,
//     t = (iso - f_lo) / (f_hi - f_lo); output column order (k, j, i) = (x, y, z);
//   * triangles: cells in linear order, table order inside a cell, corners (0, 2, 1).
// Three passes, all HBM-bound: classify (read 4 B, write 10 B per voxel), two scans,
// emit (read 10 B per voxel + 12 B per vertex, 24 B per face written).
#include "common.cuh"

namespace icon {

#include "mc_tables.inc"   // __constant__ c_num_verts[256], c_tri_table[256][16], c_edge_owner[12][4]

struct McGrid {
    const float *occ;
    int R, G, shift;   // g[i][j][k] = inside ? occ[i+shift][j+shift][k+shift] : 0
    int lo, hi;        // inside when lo <= i,j,k <= hi
    __device__ __forceinline__ float at(int i, int j, int k) const {
        bool in = i >= lo && i <= hi && j >= lo && j <= hi && k >= lo && k <= hi;
        return in ? occ[((size_t)(i + shift) * R + (j + shift)) * R + (k + shift)] : 0.f;
    }
};

static McGrid make_grid(const float *occ, int R, int padded) {
    McGrid g;
    g.occ = occ; g.R = R;
    if (padded) { g.G = R + 1; g.shift = 0; g.lo = 1; g.hi = R - 1; }
    else        { g.G = R - 1; g.shift = 1; g.lo = 0; g.hi = R - 2; }
    return g;
}

__global__ void k_mc_classify(McGrid g, float iso, int32_t *__restrict__ vcount, int32_t *__restrict__ tcount,
                              uint8_t *__restrict__ vflags, uint8_t *__restrict__ vcase) {
    const int G = g.G;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, i = blockIdx.z;
    if (k >= G) return;
    const size_t v = ((size_t)i * G + j) * G + k;
    const bool i1 = i + 1 < G, j1 = j + 1 < G, k1 = k + 1 < G;
    // corners in table numbering: (di,dj,dk) = CORNERS[c]
    float f[8];
    f[0] = g.at(i, j, k);
    f[1] = i1 ? g.at(i + 1, j, k) : 0.f;
    f[3] = j1 ? g.at(i, j + 1, k) : 0.f;
    f[4] = k1 ? g.at(i, j, k + 1) : 0.f;
    const bool b0 = f[0] < iso;
    unsigned fl = 0;
    if (i1 && ((f[1] < iso) != b0)) fl |= 1;
    if (j1 && ((f[3] < iso) != b0)) fl |= 2;
    if (k1 && ((f[4] < iso) != b0)) fl |= 4;
    int cs = 0, nt = 0;
    if (i1 && j1 && k1) {
        f[2] = g.at(i + 1, j + 1, k);
        f[5] = g.at(i + 1, j, k + 1);
        f[6] = g.at(i + 1, j + 1, k + 1);
        f[7] = g.at(i, j + 1, k + 1);
#pragma unroll
        for (int c = 0; c < 8; ++c) cs |= (f[c] < iso) ? (1 << c) : 0;
        nt = c_num_verts[cs] / 3;
    }
    vcount[v] = __popc(fl);
    tcount[v] = nt;
    vflags[v] = (uint8_t)fl;
    vcase[v] = (uint8_t)cs;
}

template <typename VT>
__global__ void k_mc_emit(McGrid g, float iso, const int32_t *__restrict__ voff, const int32_t *__restrict__ toff,
                          const uint8_t *__restrict__ vflags, const uint8_t *__restrict__ vcase,
                          VT *__restrict__ verts, int64_t *__restrict__ faces) {
    const int G = g.G;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, i = blockIdx.z;
    if (k >= G) return;
    const size_t v = ((size_t)i * G + j) * G + k;
    const unsigned fl = vflags[v];
    if (fl) {
        const VT f0 = (VT)g.at(i, j, k);
        int64_t id = voff[v];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!(fl & (1u << a))) continue;
            const VT f1 = (VT)g.at(i + (a == 0), j + (a == 1), k + (a == 2));
            const VT t = ((VT)iso - f0) / (f1 - f0);
            VT pi = (VT)i, pj = (VT)j, pk = (VT)k;
            if (a == 0) pi += t; else if (a == 1) pj += t; else pk += t;
            verts[3 * id + 0] = pk;   // x
            verts[3 * id + 1] = pj;   // y
            verts[3 * id + 2] = pi;   // z
            ++id;
        }
    }
    const int cs = vcase[v];
    const int nv = c_num_verts[cs];
    if (nv) {
        int64_t t0 = toff[v];
        for (int t = 0; t < nv; t += 3) {
            int64_t ids[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int e = c_tri_table[cs][t + c];
                const int oi = i + c_edge_owner[e][0], oj = j + c_edge_owner[e][1], ok = k + c_edge_owner[e][2];
                const int ax = c_edge_owner[e][3];
                const size_t ov = ((size_t)oi * G + oj) * G + ok;
                ids[c] = (int64_t)voff[ov] + __popc((unsigned)vflags[ov] & ((1u << ax) - 1u));
            }
            faces[3 * t0 + 0] = ids[0];
            faces[3 * t0 + 1] = ids[2];
            faces[3 * t0 + 2] = ids[1];
            ++t0;
        }
    }
}

struct McWs {
    int32_t *voff, *toff;
    uint8_t *vflags, *vcase;
    void *scan_ws;
};
static McWs carve_mc(Carver &c, int G) {
    const size_t n = (size_t)G * G * G;
    McWs w;
    w.voff = c.take<int32_t>(n);
    w.toff = c.take<int32_t>(n);
    w.vflags = c.take<uint8_t>(n);
    w.vcase = c.take<uint8_t>(n);
    w.scan_ws = c.take<char>(scan_ws_bytes((int64_t)n));
    return w;
}

}  // namespace icon

using namespace icon;

extern "C" size_t icon_mc_workspace_bytes(int R, int padded) {
    Carver c(nullptr);
    carve_mc(c, padded ? R + 1 : R - 1);
    return c.total();
}

extern "C" int icon_mc_count(const float *occ, int R, float iso, int padded, void *ws, size_t ws_bytes,
                             int64_t *d_counts, icon_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ICON_CHECK_ARG(occ && ws && d_counts && R >= 3 && R <= 1290, "icon_mc_count: bad argument (R=%d)", R);
    if (ws_bytes < icon_mc_workspace_bytes(R, padded)) {
        set_error("icon_mc_count: workspace %zu < %zu", ws_bytes, icon_mc_workspace_bytes(R, padded));
        return ICON_ENOSPC;
    }
    McGrid g = make_grid(occ, R, padded);
    Carver c(ws);
    McWs w = carve_mc(c, g.G);
    dim3 grid((g.G + 127) / 128, g.G, g.G);
    k_mc_classify<<<grid, 128, 0, stream>>>(g, iso, w.voff, w.toff, w.vflags, w.vcase);
    ICON_LAUNCHED();
    const int64_t n = (int64_t)g.G * g.G * g.G;
    int rc = scan_exclusive_i32(w.voff, w.voff, n, d_counts, w.scan_ws, stream);
    if (rc) return rc;
    return scan_exclusive_i32(w.toff, w.toff, n, d_counts + 1, w.scan_ws, stream);
}

extern "C" int icon_mc_emit(const float *occ, int R, float iso, int padded, const void *ws, void *verts,
                            int64_t *faces, int64_t n_verts, int64_t n_tris, icon_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_verts == 0 && n_tris == 0) return ICON_OK;
    ICON_CHECK_ARG(occ && ws && verts && faces, "icon_mc_emit: null pointer");
    McGrid g = make_grid(occ, R, padded);
    Carver c((void *)ws);
    McWs w = carve_mc(c, g.G);
    dim3 grid((g.G + 127) / 128, g.G, g.G);
    if (padded)
        k_mc_emit<float><<<grid, 128, 0, stream>>>(g, iso, w.voff, w.toff, w.vflags, w.vcase, (float *)verts, faces);
    else
        k_mc_emit<double><<<grid, 128, 0, stream>>>(g, iso, w.voff, w.toff, w.vflags, w.vcase, (double *)verts, faces);
    ICON_LAUNCHED();
    return ICON_OK;
}
