// Shared helpers for libicon_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/icon_b200.h"

namespace icon {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define ICON_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            icon::set_error(__VA_ARGS__);         \
            return ICON_EINVAL;                   \
        }                                         \
    } while (0)

#define ICON_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            icon::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,             \
                            cudaGetErrorString(_e));                                 \
            return ICON_ECUDA;                                                       \
        }                                                                            \
    } while (0)

// after a kernel launch
#define ICON_LAUNCHED()                 \
    do {                                \
        icon::count_launch();           \
        ICON_CUDA(cudaGetLastError());  \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cudaFuncSetAttribute is per DEVICE, the process may use several (the reference's cfg.test_gpus): remember which
// devices a kernel's opt-in has been done on.  `flags` = one static array per call site.
constexpr int ICON_MAX_DEVICES = 64;
static inline bool device_needs_setup(bool (&flags)[ICON_MAX_DEVICES], int *dev_out = nullptr) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= ICON_MAX_DEVICES) { if (dev_out) *dev_out = 0; return true; }
    if (dev_out) *dev_out = dev;
    if (flags[dev]) return false;
    flags[dev] = true;
    return true;
}
int device_sm_count();    // multiprocessors of the CURRENT device (cached per device)

// Programmatic dependent launch (PDL): the encoders are chains of ~600 short kernels; launched with the programmatic-
// stream-serialization attribute a kernel's CTAs may be scheduled while its predecessor drains, do their local set-up
// (barrier init, TMEM allocation, descriptor prefetch) and then block in pdl_wait() until the predecessor has completed
// and its writes are visible.  Every kernel launched through launch_pdl() calls pdl_launch_dependents() first thing and
// pdl_wait() before its first global-memory access; both are no-ops for a normal launch.  OFF by default (it measured
// slower than plain CUDA-graph replay: dependents park on the SMs the predecessor's tail still needs); ICON_B200_PDL=1 enables it.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// carve a workspace
struct Carver {
    char *base;
    size_t off;
    explicit Carver(void *p) : base((char *)p), off(0) {}
    template <typename T>
    T *take(size_t n) {
        off = align_up(off, 256);
        T *r = (T *)(base ? base + off : nullptr);
        off += n * sizeof(T);
        return r;
    }
    size_t total() const { return align_up(off, 256); }
};

// ---------------------------------------------------------------- mesh workspace layout
constexpr int BVH_MAX_LEVELS = 10;   // 4-ary implicit tree over leaves of 4 Morton-sorted faces
constexpr int RAY_GRID = 256;        // yz cell grid for the +x ray parity
constexpr int RAY_LIST_PER_FACE = 64;

struct MeshHeader {                  // device-resident, written by icon_smpl_prepare
    float y0, z0, inv_cy, inv_cz;    // ray grid origin / inverse cell size over the mesh yz box
    int ray_overflow;                // 1 -> cell lists overflowed: kernels fall back to all faces
    int pad[3];
};

struct MeshView {
    const float4 *tri;    // [F][3]: (a.xyz, ab.x) (ab.yz, ac.xy) (ac.z, -, -, -), original face order
    const float4 *sph;    // [F]: bounding sphere (centre xyz, radius), conservative
    const float4 *attr;   // [F][6]: normals 9, cmap 9, vis 3, pad 3
    const float4 *rbox;   // [F][2]: (ymin, ymax, zmin, zmax) (xmax, xmin, -, -)
    float *vnormals;      // [V][3] scratch
    // Morton-sorted copy + implicit AABB tree
    unsigned long long *keys;   // [F] morton << 32 | face
    int32_t *order;       // [F] sorted position -> face id
    float4 *tri_s;        // [F][3] sorted
    float4 *sph_s;        // [F] sorted
    float4 *nodes;        // [total_nodes][2]: (min.xyz, -) (max.xyz, -), level 0 (leaves) first
    // ray grid
    int32_t *rcount;      // [RAY_GRID^2 + 1]
    int32_t *roff;        // [RAY_GRID^2 + 1]
    int32_t *rlist;       // [F * RAY_LIST_PER_FACE]
    MeshHeader *hdr;
    void *scan_ws;
    int V, F;
    int nlevels;
    int lvl_cnt[BVH_MAX_LEVELS];
    int lvl_off[BVH_MAX_LEVELS];
};
size_t mesh_ws_bytes(int V, int F);
MeshView mesh_view(const void *ws, int V, int F);

// stage timing for bench.py (icon_profile_*): mark(i) records event i on `stream` when enabled
void profile_mark(int i, cudaStream_t stream);

// exclusive scan of int32 -> int64 totals; in/out may alias.  ws: scan_ws_bytes(n).
size_t scan_ws_bytes(int64_t n);
int scan_exclusive_i32(const int32_t *in, int32_t *out, int64_t n, int64_t *d_total /*may be null*/,
                       void *ws, cudaStream_t stream);

}  // namespace icon
