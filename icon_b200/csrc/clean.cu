// clean_mesh on the device: keep the connected component with the most vertices.
//
// Replaces lib/dataset/mesh_util.py:778-791 (trimesh: Trimesh(verts, faces).split(only_watertight=False), component
// with the most vertices), the step apps/ICON.py:755-756 runs right after export_mesh -- the marching-cubes output
// never has to leave the GPU just to be split on the CPU.
//
// trimesh connects FACES that share an edge used by exactly two faces.  On a marching-cubes surface with edge-owned
// vertices (mc.cu) every mesh edge is such an edge and the faces around a vertex form one fan, so that relation has
// the same components as "vertices joined by a triangle edge" -- which is what this file computes, with a lock-free
// union-find over the vertex ids:
//   k_cc_union     for every face: unite(v0, v1), unite(v0, v2)       (roots only ever decrease: atomicMin hooks)
//   k_cc_flatten   parent[v] = root(v) (= smallest vertex id of the component); count[root] += 1
//   k_cc_best      arg max of count (ties: smallest root id = the component that appears first)
//   flags + scans  kept vertices / faces in ascending original order (what trimesh's submesh re-indexing yields)
//   k_cc_emit      compacted float32 vertices and int32 faces (the reference returns .float() / .int())
#include "common.cuh"

namespace icon {

__device__ __forceinline__ int cc_find(const int *parent_, int v) {
    const volatile int *parent = parent_;                // other threads hook roots concurrently: always re-read
    int p = parent[v];
    while (p != v) { v = p; p = parent[v]; }
    return v;
}

__device__ __forceinline__ void cc_unite(int *parent, int a, int b) {
    while (true) {
        a = cc_find(parent, a);
        b = cc_find(parent, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&parent[b], a);        // hook the larger root under the smaller one
        if (old == b) return;
        b = old;                                         // somebody hooked b first: continue from where it points
    }
}

__global__ void k_cc_init(int *parent, int *count, int nv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv) { parent[i] = i; count[i] = 0; }
}

__global__ void k_cc_union(const int64_t *__restrict__ faces, int nf, int *parent) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int a = (int)faces[3 * (size_t)f], b = (int)faces[3 * (size_t)f + 1], c = (int)faces[3 * (size_t)f + 2];
    cc_unite(parent, a, b);
    cc_unite(parent, a, c);
}

__global__ void k_cc_flatten(int *parent, int *count, int nv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const int r = cc_find(parent, i);
    parent[i] = r;                                       // other threads may still walk through i: r is on their path too
    atomicAdd(&count[r], 1);
}

__global__ void k_cc_best(const int *__restrict__ count, int nv, unsigned long long *best) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0;
    if (i < nv && count[i] > 0) key = ((unsigned long long)(unsigned)count[i] << 32) | (unsigned)(0x7fffffff - i);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0 && key) atomicMax(best, key);
}

__global__ void k_cc_flag(const int *__restrict__ parent, const int64_t *__restrict__ faces, int nv, int nf,
                          const unsigned long long *__restrict__ best, int32_t *vflag, int32_t *fflag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int root = 0x7fffffff - (int)(unsigned)(*best & 0xffffffffull);
    if (i < nv) vflag[i] = (cc_find(parent, i) == root) ? 1 : 0;
    if (i < nf) fflag[i] = (cc_find(parent, (int)faces[3 * (size_t)i]) == root) ? 1 : 0;
}

template <typename VT>
__global__ void k_cc_emit(const VT *__restrict__ verts, const int64_t *__restrict__ faces, int nv, int nf,
                          const int *__restrict__ parent, const unsigned long long *__restrict__ best,
                          const int32_t *__restrict__ vpos, const int32_t *__restrict__ fpos, float *__restrict__ out_v,
                          int32_t *__restrict__ out_f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int root = 0x7fffffff - (int)(unsigned)(*best & 0xffffffffull);
    if (i < nv && cc_find(parent, i) == root) {
        const int d = vpos[i];
        out_v[3 * (size_t)d] = (float)verts[3 * (size_t)i];
        out_v[3 * (size_t)d + 1] = (float)verts[3 * (size_t)i + 1];
        out_v[3 * (size_t)d + 2] = (float)verts[3 * (size_t)i + 2];
    }
    if (i < nf && cc_find(parent, (int)faces[3 * (size_t)i]) == root) {
        const int d = fpos[i];
        out_f[3 * (size_t)d] = vpos[(int)faces[3 * (size_t)i]];
        out_f[3 * (size_t)d + 1] = vpos[(int)faces[3 * (size_t)i + 1]];
        out_f[3 * (size_t)d + 2] = vpos[(int)faces[3 * (size_t)i + 2]];
    }
}

struct CleanWs {
    int *parent, *count;
    int32_t *vpos, *fpos;
    unsigned long long *best;
    void *scan;
};
static size_t clean_carve(void *ws, int64_t nv, int64_t nf, CleanWs *o) {
    Carver c(ws);
    CleanWs w;
    w.parent = c.take<int>(nv);
    w.count = c.take<int>(nv);
    w.vpos = c.take<int32_t>(nv);
    w.fpos = c.take<int32_t>(nf);
    w.best = c.take<unsigned long long>(1);
    w.scan = c.take<char>(scan_ws_bytes(nv > nf ? nv : nf));
    if (o) *o = w;
    return c.total();
}

}  // namespace icon

using namespace icon;

extern "C" size_t icon_clean_mesh_workspace_bytes(int64_t nv, int64_t nf) { return clean_carve(nullptr, nv, nf, nullptr); }

extern "C" int icon_clean_mesh_count(const int64_t *faces, int64_t nv, int64_t nf, void *ws, size_t ws_bytes,
                                     int64_t *d_counts /* [2]: vertices, faces kept */, icon_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ICON_CHECK_ARG(faces && ws && d_counts && nv > 0 && nf > 0 && nv < 0x7fffffff && nf < 0x7fffffff,
                   "icon_clean_mesh_count: bad argument");
    if (ws_bytes < icon_clean_mesh_workspace_bytes(nv, nf)) { set_error("icon_clean_mesh_count: workspace too small"); return ICON_ENOSPC; }
    CleanWs w;
    clean_carve(ws, nv, nf, &w);
    const unsigned gv = (unsigned)((nv + 255) / 256), gf = (unsigned)((nf + 255) / 256), gm = gv > gf ? gv : gf;
    k_cc_init<<<gv, 256, 0, stream>>>(w.parent, w.count, (int)nv);
    ICON_LAUNCHED();
    ICON_CUDA(cudaMemsetAsync(w.best, 0, sizeof(unsigned long long), stream));
    k_cc_union<<<gf, 256, 0, stream>>>(faces, (int)nf, w.parent);
    ICON_LAUNCHED();
    k_cc_flatten<<<gv, 256, 0, stream>>>(w.parent, w.count, (int)nv);
    ICON_LAUNCHED();
    k_cc_best<<<gv, 256, 0, stream>>>(w.count, (int)nv, w.best);
    ICON_LAUNCHED();
    k_cc_flag<<<gm, 256, 0, stream>>>(w.parent, faces, (int)nv, (int)nf, w.best, w.vpos, w.fpos);
    ICON_LAUNCHED();
    int rc = scan_exclusive_i32(w.vpos, w.vpos, nv, d_counts, w.scan, stream);
    if (rc) return rc;
    return scan_exclusive_i32(w.fpos, w.fpos, nf, d_counts + 1, w.scan, stream);
}

extern "C" int icon_clean_mesh_emit(const void *verts, int verts_f64, const int64_t *faces, int64_t nv, int64_t nf,
                                    const void *ws, float *out_verts, int32_t *out_faces, icon_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ICON_CHECK_ARG(verts && faces && ws && out_verts && out_faces && nv > 0 && nf > 0, "icon_clean_mesh_emit: bad argument");
    CleanWs w;
    clean_carve(const_cast<void *>(ws), nv, nf, &w);
    const unsigned gm = (unsigned)(((nv > nf ? nv : nf) + 255) / 256);
    if (verts_f64)
        k_cc_emit<double><<<gm, 256, 0, stream>>>((const double *)verts, faces, (int)nv, (int)nf, w.parent, w.best, w.vpos, w.fpos,
                                                  out_verts, out_faces);
    else
        k_cc_emit<float><<<gm, 256, 0, stream>>>((const float *)verts, faces, (int)nv, (int)nf, w.parent, w.best, w.vpos, w.fpos,
                                                 out_verts, out_faces);
    ICON_LAUNCHED();
    return ICON_OK;
}
