// SMPL-body SDF block of the occupancy query.  Compiled with -fmad=false.
//
// Replaces, per query point (reference: cal_sdf_batch, lib/dataset/mesh_util.py:357-396):
//   kaolin point_to_mesh_distance  -> nearest face (lowest index on ties) + squared distance
//   kaolin check_sign              -> +x ray parity
//   barycentric_coordinates_of_projection + the four gathers -> cmap / normal / vis
//
// Design: one WARP per group of PPW query points (32 Morton-adjacent lattice points on the dense grid).
//   nearest face  (A) greedy descent of the implicit 4-ary AABB tree over Morton-sorted faces (icon_smpl_prepare) to
//                 the leaf nearest the warp centre: a first bound for every lane; (B) breadth-first cull of the
//                 tree, 32 child boxes per step, ballot-compacted into a 16-bit frontier in shared memory, against
//                 the warp's bound; (C) the surviving leaves' faces are culled 32 at a time by bounding sphere, and
//                 every lane tests the survivors against ITS OWN best through two cheap lower bounds (sphere,
//                 support function) before the exact Ericson distance.  A subtree / face is skipped only when its
//                 bound is strictly farther than the current best (float slack on every bound); ties resolve to
//                 the lowest ORIGINAL face index, exactly like the brute-force scan.
//   sign          the faces listed in the point's yz cell (256 x 256 grid over the mesh's yz
//                 box) are the only ones a +x ray can hit; each is tested with the same
//                 Moller-Trumbore code as the brute-force scan, so the hit COUNT is identical.
// Results are identical to brute force over all faces (icon_sdf_bruteforce, the CPU oracle):
// the pruning is conservative and the per-face arithmetic is the same code (geom.cuh).
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "geom.cuh"

namespace icon {

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

constexpr int SW_T = 128;               // 4 warps per block, one warp = 32 Morton-adjacent points
constexpr int NBIN_AX = 128;            // Morton bins per axis over [-1,1]^3 (+1 overflow bin)
constexpr int NBIN = NBIN_AX * NBIN_AX * NBIN_AX;
// frontier / leaf list capacity per warp: the fewer points a warp carries the tighter its box and the shorter the
// lists, and the smaller footprint lets more warps be resident to hide the tree walk's dependent loads
__host__ __device__ constexpr int fr_cap(int ppw) { return ppw >= 16 ? 1024 : (ppw >= 4 ? 768 : 384); }

__device__ __forceinline__ float box_dist2(V3 p, float4 lo, float4 hi) {
    float dx = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.f);
    float dy = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.f);
    float dz = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.f);
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

__device__ __forceinline__ unsigned spread7(unsigned v) {      // <= 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ int bin_of(V3 p) {
    const bool inside = p.x >= -1.f && p.x <= 1.f && p.y >= -1.f && p.y <= 1.f && p.z >= -1.f && p.z <= 1.f;
    if (!inside) return NBIN;
    const unsigned bx = min(NBIN_AX - 1, (int)((p.x + 1.f) * (NBIN_AX * 0.5f)));
    const unsigned by = min(NBIN_AX - 1, (int)((p.y + 1.f) * (NBIN_AX * 0.5f)));
    const unsigned bz = min(NBIN_AX - 1, (int)((p.z + 1.f) * (NBIN_AX * 0.5f)));
    return (int)(spread7(bx) | (spread7(by) << 1) | (spread7(bz) << 2));
}

// xyz4[i] = (x, y, z, in_cube); bid[i] = Morton bin; count[bin]++ (warp-aggregated atomics)
__global__ void k_points_bin(const float *__restrict__ pts, int64_t sc, int64_t sn, int64_t N, Calib cb,
                             float4 *__restrict__ xyz4, int32_t *__restrict__ bid, int32_t *__restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < N;
    int b = -1;
    if (live) {
        const V3 p = load_point(pts, sc, sn, i, cb);
        // HGPIFuNet.py:270-275: in_cube = all(-1 < xyz < 1), strict
        const float in_cube = (p.x > -1.f && p.x < 1.f && p.y > -1.f && p.y < 1.f && p.z > -1.f && p.z < 1.f) ? 1.f : 0.f;
        xyz4[i] = make_float4(p.x, p.y, p.z, in_cube);
        b = bin_of(p);
        bid[i] = b;
    }
    const unsigned act = __ballot_sync(0xffffffffu, live);
    if (live) {
        const unsigned peers = __match_any_sync(act, b);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&count[b], __popc(peers));
    }
}

__global__ void k_points_scatter(const int32_t *__restrict__ bid, int64_t N, const int32_t *__restrict__ offset,
                                 int32_t *__restrict__ cursor, int32_t *__restrict__ perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < N;
    const int b = live ? bid[i] : -1;
    const unsigned act = __ballot_sync(0xffffffffu, live);
    if (live) {
        const unsigned peers = __match_any_sync(act, b);
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cursor[b], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        perm[offset[b] + base + __popc(peers & ((1u << lane) - 1u))] = (int32_t)i;
    }
}

__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

#ifdef ICON_SDF_STATS
__device__ unsigned long long g_stats[8];   // warps, overflow warps, sum leaves, sum faces, sum exact tests, sum ray tests
#define STAT(i, v) do { if (lane == 0) atomicAdd(&g_stats[i], (unsigned long long)(v)); } while (0)
#else
#define STAT(i, v) do { } while (0)
#endif

template <int FR_CAP>
struct WarpSmem {
    unsigned short fr[2][FR_CAP];      // node / leaf ids (leaf count <= 65535 is checked by the host)
    float4 sph[32];                    // bounding spheres of the surviving faces of the current chunk (compacted)
    float4 tri[32][3];                 // their (a, ab, ac) records
    int kk[32];                        // their sorted positions
};

__device__ __forceinline__ float box_far2(V3 p, float4 lo, float4 hi) {   // squared distance to the farthest corner
    float dx = fmaxf(fabsf(lo.x - p.x), fabsf(hi.x - p.x));
    float dy = fmaxf(fabsf(lo.y - p.y), fabsf(hi.y - p.y));
    float dz = fmaxf(fabsf(lo.z - p.z), fabsf(hi.z - p.z));
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// PPW = query points per warp; each point is replicated on REP = 32/PPW lanes which split the candidate faces
// (and the ray list) between them and merge by shuffle.  32: one point per lane, for dense sets where 32 Morton-
// consecutive points fill a small box.  8 / 1: for the engine's sparse refinement sets, where 32 consecutive points
// span a large box and the shared candidate list explodes; with PPW = 1 the warp box is a point, the lists are the
// per-point minimum and the 32 lanes only share the work.  The kernel is bound by the latency of the dependent
// tree loads, so the shared-memory footprint (fr_cap) is kept small enough for >= 36 resident warps per SM.
template <int PPW>
__global__ void __launch_bounds__(SW_T) k_sdf_warp(const float4 *__restrict__ xyz4, const int32_t *__restrict__ perm,
                                                   int64_t N, MeshView m, float *__restrict__ rec,
                                                   int32_t *__restrict__ face, int order) {
    constexpr int REP = 32 / PPW;
    constexpr int FR_CAP = fr_cap(PPW);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpSmem<FR_CAP> &S = reinterpret_cast<WarpSmem<FR_CAP> *>(smem_raw)[wib];
    const int sub = lane % PPW, rep_id = lane / PPW;           // which point of the warp, which replica
    const int64_t pos0 = ((int64_t)blockIdx.x * (SW_T / 32) + wib) * PPW;
    const int64_t pos = pos0 + sub;
    if (pos0 >= N) return;                                   // whole warp out of range
    const bool live = pos < N;
    const int64_t idx = perm[live ? pos : pos0];
    const float4 q = xyz4[idx];
    const V3 p = mk3(q.x, q.y, q.z);
    // warp bounding sphere (box centre, max distance)
    const V3 c = mk3(0.5f * (warp_min(p.x) + warp_max(p.x)), 0.5f * (warp_min(p.y) + warp_max(p.y)),
                     0.5f * (warp_min(p.z) + warp_max(p.z)));
    const float rw = warp_max(sqrtf(dot3(sub3(p, c), sub3(p, c)))) * 1.00001f + 1e-6f;
    // the warp's bounding box (tighter than the sphere for the flat 4x4x2 blocks of a lattice)
    const float4 wlo = make_float4(warp_min(p.x) - 1e-6f, warp_min(p.y) - 1e-6f, warp_min(p.z) - 1e-6f, 0.f);
    const float4 whi = make_float4(warp_max(p.x) + 1e-6f, warp_max(p.y) + 1e-6f, warp_max(p.z) + 1e-6f, 0.f);

    float best = FLT_MAX;
    int bi = 0x7fffffff;
    auto try_face = [&](int k) {                             // exact test of sorted face k for this lane
        const Tri tr = load_tri(m.tri_s + 3 * (size_t)k);
        const float d = tri_sqdist(p, tr.a, tr.ab, tr.ac);
        const int f = __ldg(m.order + k);
        if (d < best || (d == best && f < bi)) { best = d; bi = f; }
    };

    // ---- phase A: greedy descent towards the warp centre -> a first bound for every lane
    {
        int node = 0;
        for (int lvl = m.nlevels - 1; lvl > 0; --lvl) {
            const int ch = 4 * node + (lane & 3);
            float a = FLT_MAX;
            int ai = ch;
            if (ch < m.lvl_cnt[lvl - 1]) {
                const float4 *nb = m.nodes + 2 * ((size_t)m.lvl_off[lvl - 1] + ch);
                a = box_dist2(c, __ldg(nb), __ldg(nb + 1));
            }
            for (int o = 1; o <= 2; o <<= 1) {               // min over the 4 children (lanes 4j..4j+3)
                const float ob = __shfl_xor_sync(0xffffffffu, a, o);
                const int oi = __shfl_xor_sync(0xffffffffu, ai, o);
                if (ob < a || (ob == a && oi < ai)) { a = ob; ai = oi; }
            }
            node = __shfl_sync(0xffffffffu, ai, 0);
        }
        for (int k = 4 * node; k < min(4 * node + 4, m.F); ++k) try_face(k);
    }
    const float ubw = warp_max(sqrtf(best));                  // every lane's nearest is within ubw
    float ubw2 = ubw;                                         // bound on d(lane, its nearest face), all lanes
    float ub2 = (ubw2 * 1.00001f + 1e-6f) * (ubw2 * 1.00001f + 1e-6f);

    // ---- phase B: breadth-first cull of the tree, 32 child boxes per step.  The bound also
    //      tightens on the way down: some face lies within the nearest far-corner distance of c,
    //      so every lane's nearest face is within that + 2 rw of c.
    int cur = 0, n = 1;
    bool overflow = false;
    if (lane == 0) S.fr[0][0] = 0;
    __syncwarp();
    for (int lvl = m.nlevels - 1; lvl > 0 && !overflow; --lvl) {
        int nn = 0;
        float far2 = FLT_MAX;
        const int ccnt = m.lvl_cnt[lvl - 1];
        const float4 *nodes = m.nodes + 2 * (size_t)m.lvl_off[lvl - 1];
        for (int base = 0; base < n; base += 8) {
            const int slot = base + (lane >> 2);
            bool pass = false;
            int ch = 0;
            if (slot < n) {
                ch = 4 * (int)S.fr[cur][slot] + (lane & 3);
                if (ch < ccnt) {
                    const float4 lo = __ldg(nodes + 2 * (size_t)ch), hi = __ldg(nodes + 2 * (size_t)ch + 1);
                    // distance between the node's box and the warp's box bounds every lane's distance to the node
                    const float gx = fmaxf(fmaxf(lo.x - whi.x, wlo.x - hi.x), 0.f);
                    const float gy = fmaxf(fmaxf(lo.y - whi.y, wlo.y - hi.y), 0.f);
                    const float gz = fmaxf(fmaxf(lo.z - whi.z, wlo.z - hi.z), 0.f);
                    pass = fmaf(gz, gz, fmaf(gy, gy, gx * gx)) <= ub2;
                    far2 = fminf(far2, box_far2(c, lo, hi));
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, pass);
            const int at = nn + __popc(mask & ((1u << lane) - 1u));
            if (pass && at < FR_CAP) S.fr[cur ^ 1][at] = (unsigned short)ch;
            nn += __popc(mask);
        }
        if (nn > FR_CAP) overflow = true;
        n = nn;
        cur ^= 1;
        // some face lies within the nearest far-corner distance of c -> within that + rw of every lane
        const float l2 = sqrtf(warp_min(far2)) + rw;
        if (l2 < ubw2) { ubw2 = l2; ub2 = (ubw2 * 1.00001f + 1e-6f) * (ubw2 * 1.00001f + 1e-6f); }
        __syncwarp();
    }
    // ---- near-first order of the surviving leaves.  The lanes' bounds only tighten while faces are tested, and the
    //      warp-level cull of a chunk uses the LOOSEST lane bound: leaves that can beat even the tightest current
    //      bound (box distance <= min over lanes of sqrt(best)) go first, so that the bounds are (nearly) final
    //      before the long tail of barely-surviving leaves is looked at -- most of the tail then fails the one
    //      warp-level test instead of 32 per-lane tests.  Order does not affect results (ties: lowest face index).
    if (order && !overflow && n > 8) {
        const float ubmin = warp_min(best * rsqrtf(fmaxf(best, 1e-30f))) * 1.00001f + 1e-6f;
        const float t2 = ubmin * ubmin;
        int n1 = 0, n2 = 0;
        for (int base = 0; base < n; base += 32) {
            const int slot = base + lane;
            const bool valid = slot < n;
            int leaf = 0;
            bool near = false;
            if (valid) {
                leaf = (int)S.fr[cur][slot];
                const float4 lo = __ldg(m.nodes + 2 * (size_t)leaf), hi = __ldg(m.nodes + 2 * (size_t)leaf + 1);
                const float gx = fmaxf(fmaxf(lo.x - whi.x, wlo.x - hi.x), 0.f);
                const float gy = fmaxf(fmaxf(lo.y - whi.y, wlo.y - hi.y), 0.f);
                const float gz = fmaxf(fmaxf(lo.z - whi.z, wlo.z - hi.z), 0.f);
                near = fmaf(gz, gz, fmaf(gy, gy, gx * gx)) <= t2;
            }
            const unsigned mn = __ballot_sync(0xffffffffu, valid && near), mf = __ballot_sync(0xffffffffu, valid && !near);
            const unsigned lt = (1u << lane) - 1u;
            if (valid && near) S.fr[cur ^ 1][n1 + __popc(mn & lt)] = (unsigned short)leaf;
            if (valid && !near) S.fr[cur ^ 1][n - 1 - (n2 + __popc(mf & lt))] = (unsigned short)leaf;
            n1 += __popc(mn); n2 += __popc(mf);
        }
        cur ^= 1;
        __syncwarp();
    }
    STAT(0, 1); STAT(1, overflow ? 1 : 0); STAT(2, n);
    // ---- phases C+D: 32 faces (8 leaves) at a time: cull by bounding sphere against the warp's box and bound,
    //      stage the survivors (sphere + triangle) compacted in shared memory, then every lane tests them against
    //      its own best through two cheap lower bounds (bounding sphere, then the support-function bound
    //      d >= |w| - max_k u.(v_k - c_f), u = w/|w|, w = p - c_f, nearly exact for faces seen head-on) before
    //      the exact distance.
    {
        float sbA = best * rsqrtf(best) * 1.00001f + 1e-6f;    // ~sqrt(best), inflated; bounds only
        float ubA = ubw2 * 1.00001f + 1e-6f;
        // `tr` points at the face's three float4 (a, ab, ac); it is only dereferenced once the sphere test passes
        auto lane_test = [&](int k, float4 s, const float4 *tr) {
            const float dx = p.x - s.x, dy = p.y - s.y, dz = p.z - s.z;
            const float dd = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float l = sbA + s.w;
            if (dd > l * l) return;                            // sphere bound beats this lane's best
            const float4 r0 = tr[0], r1 = tr[1], r2 = tr[2];
            const V3 ab = mk3(r0.w, r1.x, r1.y), ac = mk3(r1.z, r1.w, r2.x);
            const float S1 = fmaf(dz, ab.z, fmaf(dy, ab.y, dx * ab.x));
            const float T1 = fmaf(dz, ac.z, fmaf(dy, ac.y, dx * ac.x));
            const float M = fmaxf(fmaxf(-(S1 + T1), fmaf(2.f, S1, -T1)), fmaf(2.f, T1, -S1)) * (1.f / 3.f);
            const float g = dd - M - 1e-7f;                    // |w|^2 - |w| h(u)
            if (g > 0.f && g * g > best * dd * 1.0001f) return;   // support bound beats this lane's best
            const float d = tri_sqdist(p, mk3(r0.x, r0.y, r0.z), ab, ac);
            if (d > best) return;
            const int f = __ldg(m.order + k);
            if (d < best || f < bi) {
                best = d; bi = f;
                sbA = d * rsqrtf(d) * 1.00001f + 1e-6f;
                if (!(d > 0.f)) sbA = 1e-6f;
            }
        };
        if (!overflow) {
            for (int base = 0; base < n; base += 8) {
                const int slot = base + (lane >> 2);
                bool pass = false;
                int k = 0;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                if (slot < n) {
                    k = 4 * (int)S.fr[cur][slot] + (lane & 3);
                    if (k < m.F) {
                        s = __ldg(m.sph_s + k);
                        const float l2 = ubA + s.w;            // sphere vs the warp's box: some lane may be that close
                        pass = box_dist2(mk3(s.x, s.y, s.z), wlo, whi) <= l2 * l2;
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, pass);
                const int cnt = __popc(mask);
                STAT(3, cnt);
                if (pass) {
                    const int at = __popc(mask & ((1u << lane) - 1u));
                    const float4 *tp = m.tri_s + 3 * (size_t)k;
                    S.sph[at] = s;
                    S.kk[at] = k;
                    S.tri[at][0] = __ldg(tp); S.tri[at][1] = __ldg(tp + 1); S.tri[at][2] = __ldg(tp + 2);
                }
                __syncwarp();
                for (int j = rep_id; j < cnt; j += REP) lane_test(S.kk[j], S.sph[j], &S.tri[j][0]);
                if (REP > 1) {                                  // replicas of a point share their best
#pragma unroll
                    for (int o = PPW; o < 32; o <<= 1) {
                        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                    sbA = best > 0.f ? best * rsqrtf(best) * 1.00001f + 1e-6f : 1e-6f;
                }
                ubA = fminf(ubA, warp_max(sbA));               // the lanes' bounds only shrink: cull the next chunk harder
                __syncwarp();
            }
        } else {
            for (int k = rep_id; k < m.F; k += REP) {
                lane_test(k, __ldg(m.sph_s + k), m.tri_s + 3 * (size_t)k);
            }
        }
        if (REP > 1) {                                          // final merge over the replicas (ties: lowest face id)
#pragma unroll
            for (int o = PPW; o < 32; o <<= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
        }
    }
    // ---- +x ray parity
    int hits = 0;
    const MeshHeader h = *m.hdr;
    if (!h.ray_overflow) {
        const float fy = (p.y - h.y0) * h.inv_cy, fz = (p.z - h.z0) * h.inv_cz;
        if (fy >= 0.f && fy < (float)RAY_GRID && fz >= 0.f && fz < (float)RAY_GRID) {
            const int cc = (int)fz * RAY_GRID + (int)fy;
            const int k0 = __ldg(m.roff + cc), k1 = __ldg(m.roff + cc + 1);
            for (int k = k0 + rep_id; k < k1; k += REP) {
                const int f = __ldg(m.rlist + k);
                if (__ldg(&m.rbox[2 * (size_t)f + 1].x) < p.x - 1e-3f) continue;     // wholly behind the ray origin
                const Tri tr = load_tri(m.tri + 3 * (size_t)f);
                hits += ray_hit_px(p, tr.a, tr.ab, tr.ac);
            }
        }
    } else {
        for (int f = rep_id; f < m.F; f += REP) {
            const Tri tr = load_tri(m.tri + 3 * (size_t)f);
            hits += ray_hit_px(p, tr.a, tr.ab, tr.ac);
        }
    }
    if (REP > 1) {
#pragma unroll
        for (int o = PPW; o < 32; o <<= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
    }
    if (live && rep_id == 0) emit_record(p, bi, best, hits, m, rec, face, idx);
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

// points-per-warp policy of k_sdf_warp (see its header comment); icon_set_sdf_policy() overrides it for tuning
static int64_t g_sdf_ppw32_from = 6000000, g_sdf_ppw8_from = 300000;
static int g_sdf_order = -1;          // near-first leaf order: measured neutral-to-slower on the dense lattice (8.43 vs 8.33 ms,
                                      // profiles/r2_summary.md), so OFF unless ICON_B200_SDF_ORDER=1
static int g_sdf_ppw_force = 0;

// ---------------------------------------------------------------- host-side pipeline pieces
struct SdfWs {
    float4 *xyz4;
    int32_t *bid, *perm, *count, *offset;
    void *scan_ws;
};
static SdfWs carve_sdf(Carver &c, int64_t N) {
    SdfWs w;
    w.xyz4 = c.take<float4>((size_t)N);
    w.bid = c.take<int32_t>((size_t)N);
    w.perm = c.take<int32_t>((size_t)N);
    w.count = c.take<int32_t>(NBIN + 1);
    w.offset = c.take<int32_t>(NBIN + 1);
    w.scan_ws = c.take<char>(scan_ws_bytes(NBIN + 1));
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

// orthogonal() + in_cube, Morton binning, warp-cooperative SDF.  Leaves xyz4 (in_cube in .w) in the
// workspace for the MLP stage.
int run_sdf(const float *points, int64_t sc, int64_t sn, int64_t N, const float *h_calib,
            const MeshView &m, float *rec, int32_t *face, void *ws, float4 **xyz4_out,
            cudaStream_t stream) {
    Carver c(ws);
    SdfWs w = carve_sdf(c, N);
    Calib cb = make_calib(h_calib);
    profile_mark(0, stream);
    ICON_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int32_t) * (NBIN + 1), stream));
    const unsigned nblk = (unsigned)((N + 255) / 256);
    k_points_bin<<<nblk, 256, 0, stream>>>(points, sc, sn, N, cb, w.xyz4, w.bid, w.count);
    ICON_LAUNCHED();
    int rc = scan_exclusive_i32(w.count, w.offset, NBIN + 1, nullptr, w.scan_ws, stream);
    if (rc) return rc;
    ICON_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int32_t) * (NBIN + 1), stream));       // reuse as cursor
    k_points_scatter<<<nblk, 256, 0, stream>>>(w.bid, N, w.offset, w.count, w.perm);
    ICON_LAUNCHED();
    static bool attr_set[ICON_MAX_DEVICES] = {};
    if (device_needs_setup(attr_set)) {
#define ICON_SDF_ATTR(P) ICON_CUDA(cudaFuncSetAttribute(k_sdf_warp<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                        (int)(sizeof(WarpSmem<fr_cap(P)>) * (SW_T / 32))))
        ICON_SDF_ATTR(32); ICON_SDF_ATTR(16); ICON_SDF_ATTR(8); ICON_SDF_ATTR(4); ICON_SDF_ATTR(2); ICON_SDF_ATTR(1);
#undef ICON_SDF_ATTR
    }
    profile_mark(1, stream);
    // the kernel needs many warps in flight to hide the latency of its tree walk, so the fewer points a call has the
    // fewer of them a warp carries.  The thresholds are tuned on the engine's sparse refinement sets (36k / 167k
    // points -> PPW 1, 826k .. ~5M -> PPW 8: 3.6 ms of engine time per image instead of 6.6 ms with PPW 32) and the
    // dense 256^3 lattice (PPW 32).  A DENSE mid-sized call would prefer PPW 32 (2.1M-point lattice: 1.7 vs 2.4 ms);
    // the point count alone cannot tell the two apart -- callers that know can pin it (icon_set_sdf_policy).
    int ppw = N >= g_sdf_ppw32_from ? 32 : (N >= g_sdf_ppw8_from ? 8 : 1);
    if (g_sdf_ppw_force) ppw = g_sdf_ppw_force;
    if (g_sdf_order < 0) {
        const char *e = getenv("ICON_B200_SDF_ORDER");
        g_sdf_order = (e && e[0] == '1') ? 1 : 0;
    }
    const int wpb = SW_T / 32;
    const int64_t nwarps = (N + ppw - 1) / ppw;
    const unsigned nblk_w = (unsigned)((nwarps + wpb - 1) / wpb);
#define ICON_SDF_LAUNCH(P) k_sdf_warp<P><<<nblk_w, SW_T, sizeof(WarpSmem<fr_cap(P)>) * (SW_T / 32), stream>>>( \
        w.xyz4, w.perm, N, m, rec, face, g_sdf_order)
    switch (ppw) {
        case 32: ICON_SDF_LAUNCH(32); break;
        case 16: ICON_SDF_LAUNCH(16); break;
        case 8: ICON_SDF_LAUNCH(8); break;
        case 4: ICON_SDF_LAUNCH(4); break;
        case 2: ICON_SDF_LAUNCH(2); break;
        default: ICON_SDF_LAUNCH(1); break;
    }
#undef ICON_SDF_LAUNCH
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

#ifdef ICON_SDF_STATS
extern "C" int icon_debug_sdf_stats(unsigned long long *out, int reset) {
    cudaMemcpyFromSymbol(out, icon::g_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(icon::g_stats, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int icon_set_sdf_policy(int force_ppw, int64_t ppw8_from, int64_t ppw32_from) {
    ICON_CHECK_ARG(force_ppw >= 0 && force_ppw <= 32 && (force_ppw & (force_ppw - 1)) == 0, "icon_set_sdf_policy: ppw in {0,1,2,4,8,16,32}");
    icon::g_sdf_ppw_force = force_ppw;
    if (ppw8_from >= 0) icon::g_sdf_ppw8_from = ppw8_from;
    if (ppw32_from >= 0) icon::g_sdf_ppw32_from = ppw32_from;
    return ICON_OK;
}

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
