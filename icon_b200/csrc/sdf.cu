// SMPL-body SDF block of the occupancy query, brick-culled.  Compiled with -fmad=false.
//
// Replaces, per query point (reference: cal_sdf_batch, lib/dataset/mesh_util.py:357-396):
//   kaolin point_to_mesh_distance  -> nearest face (lowest index on ties) + squared distance
//   kaolin check_sign              -> +x ray parity
//   barycentric_coordinates_of_projection + the four gathers -> cmap / normal / vis
//
// Design (DESIGN.md "SDF bricks"): the cube [-1,1]^3 is cut into 32^3 bricks of edge 1/16.
// Points are counting-sorted by brick; one CTA owns one non-empty brick and
//   phase 1  culls the F faces against the brick centre c: with d_c = min_f d(c,f) and brick
//            radius r, only faces with d(c,f) <= d_c + 2r can be nearest to ANY point of the
//            brick (triangle inequality) -> candidate list in shared memory; a second list
//            keeps the faces whose yz-box meets the brick's (+x ray candidates);
//   phase 2  every thread takes a point and scans the candidate list with a bounding-sphere
//            reject in front of the exact point-triangle distance, then the ray list.
// Results are identical to the brute-force scan over all faces (icon_sdf_bruteforce and the CPU
// oracle): culling is conservative and the per-face arithmetic is the same code (geom.cuh).
#include <float.h>

#include "common.cuh"
#include "geom.cuh"

namespace icon {

constexpr int NB = 32;                 // bricks per axis
constexpr int NBRICK = NB * NB * NB;   // + 1 overflow brick for points outside [-1,1]^3
constexpr float BRICK_H = 2.0f / NB;
constexpr int SDF_T = 256;
constexpr int CAND1_MAX = 6144;        // faces surviving the sphere cull (with exact distance)
constexpr int CAND_MAX = 6144;         // final nearest-face candidates
constexpr int RCAND_MAX = 3072;        // +x ray candidates

struct Calib {
    float r[9];
    float t[3];
};

__device__ __forceinline__ V3 load_point(const float *__restrict__ pts, int64_t sc, int64_t sn,
                                         int64_t i, const Calib &cb) {
    float px = pts[i * sn], py = pts[sc + i * sn], pz = pts[2 * sc + i * sn];
    V3 o;
    o.x = fmaf(cb.r[2], pz, fmaf(cb.r[1], py, cb.r[0] * px)) + cb.t[0];
    o.y = fmaf(cb.r[5], pz, fmaf(cb.r[4], py, cb.r[3] * px)) + cb.t[1];
    o.z = fmaf(cb.r[8], pz, fmaf(cb.r[7], py, cb.r[6] * px)) + cb.t[2];
    return o;
}

__device__ __forceinline__ int brick_of(V3 p) {
    bool inside = p.x >= -1.f && p.x <= 1.f && p.y >= -1.f && p.y <= 1.f && p.z >= -1.f && p.z <= 1.f;
    if (!inside) return NBRICK;
    int bx = min(NB - 1, (int)((p.x + 1.f) * (NB * 0.5f)));
    int by = min(NB - 1, (int)((p.y + 1.f) * (NB * 0.5f)));
    int bz = min(NB - 1, (int)((p.z + 1.f) * (NB * 0.5f)));
    return (bz * NB + by) * NB + bx;
}

// xyz4[i] = (x, y, z, in_cube); bid[i]; count[brick]++ (warp-aggregated)
__global__ void k_points_bin(const float *__restrict__ pts, int64_t sc, int64_t sn, int64_t N,
                             Calib cb, float4 *__restrict__ xyz4, uint16_t *__restrict__ bid,
                             int32_t *__restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < N;
    int b = -1;
    if (live) {
        V3 p = load_point(pts, sc, sn, i, cb);
        // HGPIFuNet.py:270-275: in_cube = all(-1 < xyz < 1), strict
        float in_cube = (p.x > -1.f && p.x < 1.f && p.y > -1.f && p.y < 1.f && p.z > -1.f && p.z < 1.f)
                            ? 1.f : 0.f;
        xyz4[i] = make_float4(p.x, p.y, p.z, in_cube);
        b = brick_of(p);
        bid[i] = (uint16_t)b;
    }
    unsigned act = __ballot_sync(0xffffffffu, live);
    if (live) {
        unsigned peers = __match_any_sync(act, b);
        int leader = __ffs(peers) - 1;
        if ((threadIdx.x & 31) == leader) atomicAdd(&count[b], __popc(peers));
    }
}

__global__ void k_points_scatter(const uint16_t *__restrict__ bid, int64_t N,
                                 const int32_t *__restrict__ offset, int32_t *__restrict__ cursor,
                                 int32_t *__restrict__ perm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < N;
    int b = live ? (int)bid[i] : -1;
    unsigned act = __ballot_sync(0xffffffffu, live);
    if (live) {
        unsigned peers = __match_any_sync(act, b);
        int lane = threadIdx.x & 31;
        int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cursor[b], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        int rank = __popc(peers & ((1u << lane) - 1u));
        perm[offset[b] + base + rank] = (int32_t)i;
    }
}

// cmap / normal / vis of the winning face + final sdf; mirrors the tail of oracle_cal_sdf().
__device__ __forceinline__ void emit_record(V3 p, int bi, float best, int hits, const MeshView &m,
                                            float *__restrict__ rec, int32_t *__restrict__ face,
                                            int64_t idx) {
    Tri t = load_tri(m.tri + 3 * (size_t)bi);
    V3 n = cross3(t.ab, t.ac);
    float s = n.x * n.x + n.y * n.y + n.z * n.z;
    if (s == 0.f) s = 1e-6f;
    float inv = 1.0f / s;
    V3 w = sub3(p, t.a);
    V3 uw = cross3(t.ab, w), wv = cross3(w, t.ac);
    float b2 = (uw.x * n.x + uw.y * n.y + uw.z * n.z) * inv;
    float b1 = (wv.x * n.x + wv.y * n.y + wv.z * n.z) * inv;
    float b0 = 1.0f - b1 - b2;
    const float4 *at = m.attr + 6 * (size_t)bi;
    float4 a0 = __ldg(at), a1 = __ldg(at + 1), a2 = __ldg(at + 2), a3 = __ldg(at + 3),
           a4 = __ldg(at + 4), a5 = __ldg(at + 5);
    // normals n0=(a0.x,a0.y,a0.z) n1=(a0.w,a1.x,a1.y) n2=(a1.z,a1.w,a2.x)
    float nx = (a0.x * b0 + a0.w * b1 + a1.z * b2) * -1.f;
    float ny = (a0.y * b0 + a1.x * b1 + a1.w * b2) * 1.f;
    float nz = (a0.z * b0 + a1.y * b1 + a2.x * b2) * -1.f;
    // cmap m0=(a2.y,a2.z,a2.w) m1=(a3.x,a3.y,a3.z) m2=(a3.w,a4.x,a4.y)
    float cx = a2.y * b0 + a3.x * b1 + a3.w * b2;
    float cy = a2.z * b0 + a3.y * b1 + a4.x * b2;
    float cz = a2.w * b0 + a3.z * b1 + a4.y * b2;
    float vv = a4.z * b0 + a4.w * b1 + a5.x * b2;
    float vis = vv >= 0.1f ? 1.f : 0.f;
    float dist = sqrtf(best) / sqrtf(3.0f);
    float sign = 2.0f * ((float)(hits & 1) - 0.5f);
    float sdf = dist * sign;
    float4 *r = (float4 *)(rec + 8 * idx);
    r[0] = make_float4(sdf, cx, cy, cz);
    r[1] = make_float4(nx, ny, nz, vis);
    if (face) face[idx] = bi;
}

struct BrickSmem {
    int c1_f[CAND1_MAX];
    float c1_d[CAND1_MAX];
    int cand[CAND_MAX];
    int rcand[RCAND_MAX];
    float red_f[SDF_T / 32];
    int red_i[SDF_T / 32];
    int n1, ncand, nrc, all_mode, ray_all, f0;
    float ub, dmin;
};

__global__ void __launch_bounds__(SDF_T) k_sdf_brick(const float4 *__restrict__ xyz4,
                                                     const int32_t *__restrict__ perm,
                                                     const int32_t *__restrict__ count,
                                                     const int32_t *__restrict__ offset, MeshView m,
                                                     float *__restrict__ rec, int32_t *__restrict__ face) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BrickSmem &S = *reinterpret_cast<BrickSmem *>(smem_raw);
    const int b = blockIdx.x;
    const int cnt = count[b];
    if (cnt == 0) return;
    const int off = offset[b];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int F = m.F;
    const bool overflow = (b == NBRICK);

    if (tid == 0) {
        S.n1 = 0; S.ncand = 0; S.nrc = 0;
        S.all_mode = overflow ? 1 : 0;
        S.ray_all = overflow ? 1 : 0;
        S.f0 = 0x7fffffff;
    }
    __syncthreads();

    if (!overflow) {
        const int bx = b % NB, by = (b / NB) % NB, bz = b / (NB * NB);
        const V3 c = mk3(-1.f + (bx + 0.5f) * BRICK_H, -1.f + (by + 0.5f) * BRICK_H,
                         -1.f + (bz + 0.5f) * BRICK_H);
        const float hr = 0.5f * BRICK_H + 1e-5f;                 // half edge, with binning slack
        const float r2 = 2.0f * (hr * 1.7320509f) + 1e-5f;       // 2 * brick radius

        // ---- pass A: upper bound on d_c from the bounding spheres
        float ub = FLT_MAX;
        for (int f = tid; f < F; f += SDF_T) {
            float4 s = __ldg(m.sph + f);
            float dx = c.x - s.x, dy = c.y - s.y, dz = c.z - s.z;
            float d = sqrtf(dx * dx + dy * dy + dz * dz) + s.w;
            ub = fminf(ub, d);
        }
        for (int o = 16; o; o >>= 1) ub = fminf(ub, __shfl_xor_sync(0xffffffffu, ub, o));
        if (lane == 0) S.red_f[wid] = ub;
        __syncthreads();
        if (tid == 0) {
            float u = S.red_f[0];
            for (int w = 1; w < SDF_T / 32; ++w) u = fminf(u, S.red_f[w]);
            S.ub = u;
        }
        __syncthreads();
        const float lim1 = S.ub * 1.00001f + r2 + 1e-5f;

        // ---- pass B: exact centre distance for faces whose sphere lower bound passes;
        //      in the same sweep collect +x-ray candidates by yz box
        const float ylo = c.y - hr - 1e-4f, yhi = c.y + hr + 1e-4f;
        const float zlo = c.z - hr - 1e-4f, zhi = c.z + hr + 1e-4f;
        const float xlo = c.x - hr - 1e-4f;
        float dmin = FLT_MAX;
        int fmin = 0x7fffffff;
        for (int f = tid; f < F; f += SDF_T) {
            float4 s = __ldg(m.sph + f);
            float dx = c.x - s.x, dy = c.y - s.y, dz = c.z - s.z;
            float lb = sqrtf(dx * dx + dy * dy + dz * dz) - s.w;
            if (lb <= lim1) {
                Tri t = load_tri(m.tri + 3 * (size_t)f);
                float d2 = tri_sqdist(c, t.a, t.ab, t.ac);
                int slot = atomicAdd(&S.n1, 1);
                if (slot < CAND1_MAX) { S.c1_f[slot] = f; S.c1_d[slot] = d2; }
                if (d2 < dmin || (d2 == dmin && f < fmin)) { dmin = d2; fmin = f; }
            }
            float4 rb = __ldg(m.rbox + 2 * (size_t)f);
            float xmax = __ldg(&m.rbox[2 * (size_t)f + 1].x);
            if (rb.x <= yhi && rb.y >= ylo && rb.z <= zhi && rb.w >= zlo && xmax >= xlo) {
                int slot = atomicAdd(&S.nrc, 1);
                if (slot < RCAND_MAX) S.rcand[slot] = f;
            }
        }
        for (int o = 16; o; o >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, dmin, o);
            int of = __shfl_xor_sync(0xffffffffu, fmin, o);
            if (od < dmin || (od == dmin && of < fmin)) { dmin = od; fmin = of; }
        }
        if (lane == 0) { S.red_f[wid] = dmin; S.red_i[wid] = fmin; }
        __syncthreads();
        if (tid == 0) {
            float d = S.red_f[0]; int f = S.red_i[0];
            for (int w = 1; w < SDF_T / 32; ++w)
                if (S.red_f[w] < d || (S.red_f[w] == d && S.red_i[w] < f)) { d = S.red_f[w]; f = S.red_i[w]; }
            S.dmin = d; S.f0 = f;
            if (S.n1 > CAND1_MAX) S.all_mode = 1;
            if (S.nrc > RCAND_MAX) S.ray_all = 1;
        }
        __syncthreads();

        // ---- pass C: keep faces with d(c,f) <= d_c + 2r
        if (!S.all_mode) {
            float lim = sqrtf(S.dmin) * 1.00001f + r2 + 1e-5f;
            float lim2 = lim * lim;
            int n1 = S.n1;
            for (int k = tid; k < n1; k += SDF_T) {
                if (S.c1_d[k] <= lim2) {
                    int slot = atomicAdd(&S.ncand, 1);
                    S.cand[slot] = S.c1_f[k];     // ncand <= n1 <= CAND1_MAX == CAND_MAX
                }
            }
        }
        __syncthreads();
    }

    const bool all_mode = S.all_mode != 0, ray_all = S.ray_all != 0;
    const int ncand = all_mode ? F : S.ncand;
    const int nrc = ray_all ? F : S.nrc;
    const int f0 = S.f0;

    // ---- phase 2: one point per thread
    for (int i = tid; i < cnt; i += SDF_T) {
        const int64_t idx = perm[off + i];
        float4 q = xyz4[idx];
        V3 p = mk3(q.x, q.y, q.z);
        float best = FLT_MAX;
        int bi = 0x7fffffff;
        if (!all_mode) {
            Tri t = load_tri(m.tri + 3 * (size_t)f0);
            best = tri_sqdist(p, t.a, t.ab, t.ac);
            bi = f0;
        }
        float sb = sqrtf(best);
        for (int k = 0; k < ncand; ++k) {
            int f = all_mode ? k : S.cand[k];
            float4 s = __ldg(m.sph + f);
            float dx = p.x - s.x, dy = p.y - s.y, dz = p.z - s.z;
            float dd = dx * dx + dy * dy + dz * dz;
            float lim = sb + s.w + 1e-6f;
            if (dd > lim * lim * 1.00001f) continue;     // sphere lower bound beats current best
            Tri t = load_tri(m.tri + 3 * (size_t)f);
            float d = tri_sqdist(p, t.a, t.ab, t.ac);
            if (d < best || (d == best && f < bi)) { best = d; bi = f; sb = sqrtf(d); }
        }
        int hits = 0;
        for (int k = 0; k < nrc; ++k) {
            int f = ray_all ? k : S.rcand[k];
            Tri t = load_tri(m.tri + 3 * (size_t)f);
            hits += ray_hit_px(p, t.a, t.ab, t.ac);
        }
        emit_record(p, bi, best, hits, m, rec, face, idx);
    }
}

// brute force: every point against every face, faces staged through shared memory
__global__ void __launch_bounds__(256) k_sdf_brute(const float *__restrict__ pts, int64_t sc, int64_t sn,
                                                   int64_t N, Calib cb, MeshView m,
                                                   float *__restrict__ rec, int32_t *__restrict__ face) {
    __shared__ float4 tile[3 * 256];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < N;
    V3 p = mk3(0.f, 0.f, 0.f);
    if (live) p = load_point(pts, sc, sn, i, cb);
    float best = FLT_MAX;
    int bi = 0, hits = 0;
    for (int f0 = 0; f0 < m.F; f0 += 256) {
        int nf = min(256, m.F - f0);
        __syncthreads();
        for (int k = threadIdx.x; k < 3 * nf; k += 256) tile[k] = m.tri[3 * (size_t)f0 + k];
        __syncthreads();
        if (live) {
            for (int k = 0; k < nf; ++k) {
                float4 r0 = tile[3 * k], r1 = tile[3 * k + 1], r2 = tile[3 * k + 2];
                V3 a = mk3(r0.x, r0.y, r0.z), ab = mk3(r0.w, r1.x, r1.y), ac = mk3(r1.z, r1.w, r2.x);
                float d = tri_sqdist(p, a, ab, ac);
                if (d < best) { best = d; bi = f0 + k; }
                hits += ray_hit_px(p, a, ab, ac);
            }
        }
    }
    if (live) emit_record(p, bi, best, hits, m, rec, face, i);
}

// ---------------------------------------------------------------- host-side pipeline pieces
struct SdfWs {
    float4 *xyz4;
    uint16_t *bid;
    int32_t *perm;
    int32_t *count, *offset, *cursor;
    void *scan_ws;
};

static SdfWs carve_sdf(Carver &c, int64_t N) {
    SdfWs w;
    w.xyz4 = c.take<float4>((size_t)N);
    w.bid = c.take<uint16_t>((size_t)N);
    w.perm = c.take<int32_t>((size_t)N);
    w.count = c.take<int32_t>(NBRICK + 1);
    w.offset = c.take<int32_t>(NBRICK + 1);
    w.cursor = c.take<int32_t>(NBRICK + 1);
    w.scan_ws = c.take<char>(scan_ws_bytes(NBRICK + 1));
    return w;
}

size_t sdf_ws_bytes(int64_t N) {
    Carver c(nullptr);
    carve_sdf(c, N);
    return c.total();
}

Calib make_calib(const float *h) {
    Calib cb;
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) cb.r[3 * r + k] = h[4 * r + k];
        cb.t[r] = h[4 * r + 3];
    }
    return cb;
}

// Runs binning + brick kernel.  Leaves xyz4 (with in_cube in .w) in the workspace for the MLP stage.
int run_sdf(const float *points, int64_t sc, int64_t sn, int64_t N, const float *h_calib,
            const MeshView &m, float *rec, int32_t *face, void *ws, float4 **xyz4_out,
            cudaStream_t stream) {
    Carver c(ws);
    SdfWs w = carve_sdf(c, N);
    Calib cb = make_calib(h_calib);
    profile_mark(0, stream);
    ICON_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int32_t) * (NBRICK + 1), stream));
    ICON_CUDA(cudaMemsetAsync(w.cursor, 0, sizeof(int32_t) * (NBRICK + 1), stream));
    unsigned nblk = (unsigned)((N + 255) / 256);
    k_points_bin<<<nblk, 256, 0, stream>>>(points, sc, sn, N, cb, w.xyz4, w.bid, w.count);
    ICON_LAUNCHED();
    int rc = scan_exclusive_i32(w.count, w.offset, NBRICK + 1, nullptr, w.scan_ws, stream);
    if (rc) return rc;
    k_points_scatter<<<nblk, 256, 0, stream>>>(w.bid, N, w.offset, w.cursor, w.perm);
    ICON_LAUNCHED();
    static bool attr_set = false;
    if (!attr_set) {
        ICON_CUDA(cudaFuncSetAttribute(k_sdf_brick, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(BrickSmem)));
        attr_set = true;
    }
    profile_mark(1, stream);
    k_sdf_brick<<<NBRICK + 1, SDF_T, sizeof(BrickSmem), stream>>>(w.xyz4, w.perm, w.count, w.offset, m,
                                                                  rec, face);
    ICON_LAUNCHED();
    profile_mark(2, stream);
    if (xyz4_out) *xyz4_out = w.xyz4;
    return ICON_OK;
}

__global__ void k_points_only(const float *__restrict__ pts, int64_t sc, int64_t sn, int64_t N, Calib cb,
                              float4 *__restrict__ xyz4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    V3 p = load_point(pts, sc, sn, i, cb);
    float in_cube = (p.x > -1.f && p.x < 1.f && p.y > -1.f && p.y < 1.f && p.z > -1.f && p.z < 1.f) ? 1.f : 0.f;
    xyz4[i] = make_float4(p.x, p.y, p.z, in_cube);
}

// orthogonal() + in_cube only (pifu / pamir priors: no body mesh)
int run_points_only(const float *points, int64_t sc, int64_t sn, int64_t N, const float *h_calib,
                    float4 *xyz4, cudaStream_t stream) {
    Calib cb = make_calib(h_calib);
    k_points_only<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(points, sc, sn, N, cb, xyz4);
    ICON_LAUNCHED();
    return ICON_OK;
}

}  // namespace icon

using namespace icon;

extern "C" int icon_sdf_only(const float *points, int64_t stride_c, int64_t stride_n, int64_t N,
                             const float *h_calib, const void *mesh_ws, int V, int F, float *rec,
                             int32_t *face, void *ws, size_t ws_bytes, icon_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ICON_CHECK_ARG(N >= 0 && N < (int64_t)INT32_MAX, "icon_sdf_only: N=%lld out of range", (long long)N);
    if (N == 0) return ICON_OK;
    ICON_CHECK_ARG(points && h_calib && mesh_ws && rec && ws, "icon_sdf_only: null pointer");
    if (ws_bytes < sdf_ws_bytes(N)) {
        set_error("icon_sdf_only: workspace %zu < %zu", ws_bytes, sdf_ws_bytes(N));
        return ICON_ENOSPC;
    }
    MeshView m = mesh_view(mesh_ws, V, F);
    return run_sdf(points, stride_c, stride_n, N, h_calib, m, rec, face, ws, nullptr, stream);
}

extern "C" int icon_sdf_bruteforce(const float *points, int64_t stride_c, int64_t stride_n, int64_t N,
                                   const float *h_calib, const void *mesh_ws, int V, int F, float *rec,
                                   int32_t *face, icon_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N == 0) return ICON_OK;
    ICON_CHECK_ARG(points && h_calib && mesh_ws && rec, "icon_sdf_bruteforce: null pointer");
    MeshView m = mesh_view(mesh_ws, V, F);
    Calib cb = make_calib(h_calib);
    k_sdf_brute<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(points, stride_c, stride_n, N, cb, m, rec, face);
    ICON_LAUNCHED();
    return ICON_OK;
}
