/*
 * icon_b200.h -- C ABI of libicon_b200.so (sm_100a kernels for ICON's occupancy-query +
 * mesh-extraction hot path).
 *
 * The reference (YuliangXiu/ICON) has no FFI: its "plugin API" for this path is Python
 * nn.Module duck typing.  The Python mirror in icon_b200/ (re-exported under
 * the reference's import paths in lib/) keeps those signatures and calls the entry points
 * below through ctypes with raw device pointers.  Each entry point cites the reference
 * code it replaces.  Conventions:
 *
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named h_*;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises unless stated;
 *   - no hidden allocation: the caller (torch's caching allocator) provides outputs and
 *     workspaces; *_workspace_bytes() tell how much;
 *   - return 0 on success, a negative ICON_E* code otherwise; icon_last_error() gives the
 *     message (thread local).  There is no CPU fallback anywhere.
 */
#ifndef ICON_B200_H
#define ICON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICON_OK 0
#define ICON_EINVAL (-1)   /* bad argument / unsupported size */
#define ICON_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define ICON_ENOSPC (-3)   /* workspace or output buffer too small */

#define ICON_PRIOR_ICON 0
#define ICON_PRIOR_PIFU 1
#define ICON_PRIOR_PAMIR 2

typedef void *icon_stream_t;

int icon_version(void);
const char *icon_last_error(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t icon_launch_count(void);

/* Measurement hook (bench.py roofline): when enabled, icon_query records CUDA events on its
 * stream around its stages; icon_profile_last_query synchronises on the last one and returns
 * the stage durations of the most recent icon_query in milliseconds:
 * h_ms[0] binning+sort, [1] SDF brick kernel, [2] outlier rank, [3] gather+MLP kernel. */
int icon_profile_enable(int on);
int icon_profile_last_query(float *h_ms);

/* ------------------------------------------------------------------ SMPL body preparation
 * Replaces the per-call preamble of cal_sdf_batch, lib/dataset/mesh_util.py:367-372:
 * pytorch3d Meshes.verts_normals_padded (area-weighted vertex normals, deterministic
 * sequential-index_add order) and the four face_vertices gathers
 * (lib/common/render_utils.py:149-163).  Builds, into `mesh_ws`, per-face records
 * (a, ab, ac, bounding sphere), per-face attribute records (normals, cmap, vis at the three
 * corners) and the +x-ray culling boxes.  Done once per body, not once per query.
 * verts [V,3] f32, faces [F,3] i64, cmap [V,3] f32, vis [V] f32 (0/1). */
size_t icon_smpl_workspace_bytes(int V, int F);
int icon_smpl_prepare(const float *verts, const int64_t *faces, const float *cmap,
                      const float *vis, int V, int F, void *mesh_ws, size_t mesh_ws_bytes,
                      icon_stream_t stream);

/* ------------------------------------------------------------------ occupancy MLP weights
 * Replaces MLP.__init__/forward's per-layer Conv1d + BatchNorm1d(eval) (lib/net/MLP.py:26-72):
 * the host folds BN into the 1x1 convs and packs the layers; this copies nothing, it only
 * describes the packed device buffer so the kernels can check it.
 * Packed layout (floats): W0t [16][512] | b0 [512] | W1t [512][256] | b1 [256] |
 *                         W2t [272][128] | b2 [128] | W3 [144] | b3 [1]
 * with k-major ("t") storage, the c0 input channels zero-padded to 16, and skip-concat
 * columns ordered [y | x0]. */
#define ICON_MLP_PACKED_FLOATS (16 * 512 + 512 + 512 * 256 + 256 + 272 * 128 + 128 + 144 + 1)
/* Tensor-core form of the same folded weights (host: icon_b200/ops.py pack_mlp): every matrix
 * split W = hi + lo in fp16 and stored as ready-to-use K-major UMMA tiles (bytes):
 *   W0  hi 16384 | lo 16384    512 rows x 16 k, no swizzle (LBO 8192, SBO 128); k = 15 holds b0: the kernel feeds
 *                              x0 column 15 = 1, so the tensor-core path takes c0 <= 15
 *   W1  8 x (hi 32768 | lo 32768)   256 rows x 64 k per chunk, SWIZZLE_128B
 *   W2  4 x (hi 16384 | lo 16384)   128 rows x 64 k per chunk, SWIZZLE_128B
 *   W2t hi 4096 | lo 4096      128 rows x 16 k (the skip-concat x0 columns), no swizzle; k = 15 holds b2
 *   f32 b0[512] b1[256] b2[128] w3[144] b3[1] pad[3] */
#define ICON_MLP_TC_BYTES (32768 + 8 * 65536 + 4 * 32768 + 8192 + (512 + 256 + 128 + 144 + 4) * 4)
/* 0 = FP32 FMA kernel (mlp.cu), 1 = tcgen05 fp16x3 kernel (mlp_tc.cu, default when mlp_tc != NULL) */
int icon_set_mlp_impl(int impl);
int icon_get_mlp_impl(void);

/* ------------------------------------------------------------------ fused occupancy query
 * Replaces HGPIFuNet.query (lib/net/HGPIFuNet.py:268-367) + cal_sdf_batch
 * (lib/dataset/mesh_util.py:357-396; kaolin point_to_mesh_distance / check_sign) +
 * geometry.index / orthogonal (lib/net/geometry.py:21-61) + feat_select
 * (mesh_util.py:266-277) + MLP.forward (lib/net/MLP.py:49-72) for one feature stack, B=1.
 *
 *   points      : xyz of point i at points[c*stride_c + i*stride_n], c in 0..2
 *   calib       : HOST pointer to 12 floats, rows of [R|t] (calibs[0,:3,:4])
 *   feat        : image feature map [C,H,W] f32 (C = 12 or 6 for icon, 12 pifu, 6 pamir)
 *   vol_feat    : pamir only: [7,D,D,D] f32; else NULL
 *   mesh_ws     : icon only: workspace filled by icon_smpl_prepare (same V,F); else NULL
 *   mlp_packed  : ICON_MLP_PACKED_FLOATS floats
 *   mlp_tc      : ICON_MLP_TC_BYTES bytes (16-byte aligned) or NULL (-> FP32 kernel)
 *   c0          : MLP input channels (13 or 10)
 *   sdf_clip    : cfg.sdf_clip/100 (icon)
 *   out         : [N] f32 occupancy (preds[0,0,:])
 *   ws          : scratch, icon_query_workspace_bytes(N, F) bytes
 */
size_t icon_query_workspace_bytes(int64_t N, int F, int prior);
int icon_query(int prior, const float *points, int64_t stride_c, int64_t stride_n, int64_t N,
               const float *h_calib, const float *feat, int C, int H, int W,
               const float *vol_feat, int VD, const void *mesh_ws, int V, int F,
               const float *mlp_packed, const void *mlp_tc, int c0, float sdf_clip, float *out,
               void *ws, size_t ws_bytes, icon_stream_t stream);

/* Points-per-warp policy of the SDF kernel: force_ppw in {1, 8, 32} pins it (0 = automatic: 1 below ppw8_from
 * points per call, 8 below ppw32_from, else 32; negative thresholds keep the current value).  Results are
 * identical for every setting. */
int icon_set_sdf_policy(int force_ppw, int64_t ppw8_from, int64_t ppw32_from);

/* Debug / parity tap: the SMPL block alone (cal_sdf_batch outputs before the outlier rule).
 * rec [N,8] f32 = sdf, cmap xyz, norm xyz, vis(0/1); face [N] i32 nearest face id. */
int icon_sdf_only(const float *points, int64_t stride_c, int64_t stride_n, int64_t N,
                  const float *h_calib, const void *mesh_ws, int V, int F, float *rec,
                  int32_t *face, void *ws, size_t ws_bytes, icon_stream_t stream);
/* Same outputs by brute force over all faces (no bricks); slow, for cross-checks. */
int icon_sdf_bruteforce(const float *points, int64_t stride_c, int64_t stride_n, int64_t N,
                        const float *h_calib, const void *mesh_ws, int V, int F, float *rec,
                        int32_t *face, icon_stream_t stream);
/* MLP alone on a ready [c0,N] feature matrix (parity tap for lib/net/MLP.py:49-72). */
int icon_mlp_only(const float *feature, int c0, int64_t N, const float *mlp_packed,
                  const void *mlp_tc, float *out, icon_stream_t stream);

/* ------------------------------------------------------------------ reconstruction engine
 * Replace the per-level body of Seg3dLossless._forward_faster
 * (lib/common/seg3d_lossless.py:186-263).  Grids are [R,R,R] indexed [z][y][x]. */

/* seg3d_lossless.py:190-223: both F.interpolate(trilinear, align_corners=True) calls
 * (occupancy and (occ>balance) mask) fused with is_boundary = 0<valid<1 and with the
 * carry-over of the already-evaluated set (coords_accum*2).  R_out = 2*R_in-1.
 * boundary/done_out may be NULL (last level: upsample only, :186-203). */
int icon_display(const float *occ, int R, uint8_t *out /* [R][4R][3] */, icon_stream_t stream);   /* Seg3dLossless.display:
   first-hit + finite-difference normal preview, views front | left | right | back (seg3d_lossless.py:497-581) */
int icon_grid_upsample(const float *occ_in, const uint8_t *done_in, int R_in, float balance,
                       float *occ_out, uint8_t *boundary, uint8_t *done_out,
                       icon_stream_t stream);
/* seg3d_lossless.py:226-234 + seg3d_utils.py:169-181: (SmoothConv3D(k)(mask) > 0) == binary
 * dilation by a k^3 box; separable.  tmp: R^3 bytes.  The result is written TRANSPOSED,
 * out_xyz[x][y][z], which is the order `is_boundary.permute(2,1,0).nonzero()` walks. */
int icon_grid_dilate(const uint8_t *mask, int R, int k, uint8_t *tmp, uint8_t *out_xyz,
                     icon_stream_t stream);
/* seg3d_lossless.py:236-249 + batch_eval :125-138: clear already-evaluated voxels, stable
 * compaction in (x,y,z)-lexicographic order, emit query points
 * p = c*stride/(R_last-1)*(b_max-b_min)+b_min and linear indices z*R*R+y*R+x; marks the
 * selected voxels in `done`.  *d_count (device int64) receives n.  ws: icon_compact_workspace_bytes(R). */
size_t icon_compact_workspace_bytes(int R);
int icon_grid_compact(const uint8_t *mask_xyz, uint8_t *done, int R, int R_last,
                      const float *h_bmin, const float *h_bmax, float *points, int64_t *indices,
                      int64_t capacity, int64_t *d_count, void *ws, size_t ws_bytes,
                      icon_stream_t stream);
/* seg3d_lossless.py:255-258: occupancys.scatter_(2, point_indices, occupancys_topk). */
int icon_grid_scatter(float *occ, const int64_t *indices, const float *values, int64_t n,
                      icon_stream_t stream);
/* seg3d_lossless.py:168-177: level-0 lattice points (create_grid3D + batch_eval) and the
 * "(occupancys > 0.5).sum() == 0" test (d_count receives the number above balance). */
int icon_grid_init_points(int R0, int R_last, const float *h_bmin, const float *h_bmax,
                          float *points, icon_stream_t stream);
int icon_grid_count_above(const float *occ, int64_t n, float balance, int64_t *d_count,
                          icon_stream_t stream);

/* ------------------------------------------------------------------ marching cubes
 * Replaces Seg3dLossless.export_mesh (lib/common/seg3d_lossless.py:583-604): kaolin
 * voxelgrids_to_trianglemeshes (<=256^3) / PyMCubes (>256^3) at iso `balance`, including the
 * occupancys[1:,1:,1:] crop and the [:, [2,1,0]] / [:, [0,2,1]] permutations.  Indexing
 * contract: DESIGN.md "marching cubes" (oracle/mcubes.py restates it).
 *   occ [R,R,R]; padded=1 -> kaolin branch (zero pad, padded-frame coords, f32 verts),
 *   padded=0 -> PyMCubes branch (f64 verts).
 * icon_mc_count fills ws and writes {n_verts, n_tris} to d_counts (device int64[2]);
 * icon_mc_emit writes verts [n_verts,3] (f32 or f64) and faces [n_tris,3] i64. */
size_t icon_mc_workspace_bytes(int R, int padded);
int icon_mc_count(const float *occ, int R, float iso, int padded, void *ws, size_t ws_bytes,
                  int64_t *d_counts, icon_stream_t stream);
int icon_mc_emit(const float *occ, int R, float iso, int padded, const void *ws, void *verts,
                 int64_t *faces, int64_t n_verts, int64_t n_tris, icon_stream_t stream);

/* ------------------------------------------------------------------ PaMIR semantic voxelisation
 * Replaces voxelize_cuda.forward_semantic_voxelization as called by VoxelizationFunction.forward
 * (lib/net/voxelize.py:57-59) plus the bzyxc -> bcdhw permute of Voxelization.forward (:137).
 * verts [NV,3] f32 (surface vertices first, then the tetra-SMPL interior ones, in [-0.5,0.5]^3),
 * codes [NVsurf,3] f32, tets [NT,4] i32 (indices into verts), out [3,res,res,res] f32 (c,z,y,x).
 * Source of voxelize_cuda is absent: the algorithm is restated (csrc/voxelize.cu header), parity unpinned. */
size_t icon_voxelize_workspace_bytes(int res);
int icon_voxelize(const float *verts, int NV, int NVsurf, const float *codes, const int32_t *tets, int NT,
                  int res, float sigma, float *out, void *ws, size_t ws_bytes, icon_stream_t stream);

/* ------------------------------------------------------------------ vertex visibility (producer of smpl_vis)
 * Replaces get_visibility (lib/dataset/mesh_util.py:280-316): z-buffer rasterisation of the body at
 * image_size^2 (4096 in the reference) with pytorch3d's conventions (csrc/visibility.cu header; parity unpinned),
 * vis[v] = 1 for the vertices of every face that owns a pixel (plus the last face's, as faces[-1] does upstream).
 * xyz [V,3] f32 = (cat(xy, -z) + 1) / 2 as the reference builds it, faces [F,3] i64, vis [V] f32. */
size_t icon_visibility_workspace_bytes(int image_size);
int icon_visibility(const float *xyz, int V, const int64_t *faces, int F, int image_size, float *vis,
                    void *ws, size_t ws_bytes, icon_stream_t stream);

/* ------------------------------------------------------------------ encoder operators (NCHW fp32)
 * Replace the cuDNN / torch calls inside HGFilter (lib/net/HGFilters.py:161-197, ConvBlock
 * lib/net/net_util.py:258-280), GlobalGenerator / ResnetBlock (lib/net/FBNet.py:216-319) and
 * NormalNet.forward (lib/net/NormalNet.py:84-97).
 * icon_conv2d: nn.Conv2d (zero padding, or reflect = ReflectionPad2d(pad) folded in) or, with
 * transposed = 1, nn.ConvTranspose2d(stride, padding = pad, output_padding = out_pad); weights in
 * torch layout; optional bias, residual add and activation (0 none, 1 ReLU, 2 tanh) in the epilogue.
 * icon_group_norm: nn.GroupNorm(groups, C) with affine (gamma, beta), or nn.InstanceNorm2d(C) when
 * groups == C and gamma == beta == NULL; optional residual add and ReLU fused. */
int icon_conv2d(const float *x, const float *w, const float *bias, const float *res, float *y, int N, int Cin,
                int H, int W, int Cout, int KH, int KW, int stride, int pad, int out_pad, int reflect,
                int transposed, int act, icon_stream_t stream);
/* Tensor-core (tcgen05, fp16 hi/lo x3, fp32-class accuracy) variant for Cin % 64 == 0; same semantics.
 * wt_packed: per output-channel tile (n_tile rows, zero padded) and per 64-wide K-chunk (k = tap*Cin + ci) a
 * K-major SWIZZLE_128B tile of the fp16 hi parts followed by one of the lo parts (icon_b200/conv_ops.py packs
 * it).  splits > 1 = split-K over the chunks, partial sums in `ws` (icon_conv2d_tc_workspace_bytes). */
size_t icon_conv2d_tc_workspace_bytes(int N, int Cout, int OH, int OW, int splits);
int icon_conv2d_tc(const float *x, const void *wt_packed, const float *bias, const float *res, float *y, int N,
                   int Cin, int H, int W, int Cout, int KH, int KW, int stride, int pad, int out_pad, int reflect,
                   int transposed, int act, int n_tile, int splits, void *ws, size_t ws_bytes, icon_stream_t stream);
/* ---- NHWC encoder path (csrc/conv_nhwc.cu, csrc/act_nhwc.cu): activations between layers are NHWC, pre-split
 * x = hi + lo into two fp16 tensors [N * nplanes][Hp][Wp][Cp] (Cp % 64 == 0) by icon_act_nhwc.
 *
 * icon_conv_nhwc: implicit-GEMM convolution on tcgen05, A operand staged by 4-D tiled TMA loads straight from the
 *   hi / lo tensors (`dims` = tensor-map extents innermost first, `strides` = element strides of dims 1..3), weights
 *   from `wt_packed` (per n_tile output channels and per chunk a K-major SWIZZLE_128B fp16 hi tile followed by the lo
 *   tile; `wt_chunks` chunks per tile, chunk index = wtap * cpt + channel block).  The launch covers a logical
 *   Ht x Wt grid per image; output pixel (a, b) is written to out[n][a * osy + ooy][b * osx + oox][co_off + co] of an
 *   fp32 NHWC tensor with Cs channels (channel slices = concatenation for free; osy = 2 = one phase of a transposed
 *   convolution).  `taps`: ntaps x (dy, dx, plane, wtap) -- the A box of tap t is read at (y0 + dy, x0 + dx) of image
 *   n * nplanes + plane (out-of-range coordinates read zeros = zero padding).  stats (optional, zeroed by the
 *   caller): [N][Cout][2] doubles receive per-channel sum / sum of squares of the result (bias included).
 * icon_norm_finalize: stats -> [N][C] (scale, shift) for InstanceNorm2d (groups = 0) / GroupNorm(groups) with affine.
 * icon_act_nhwc: y = [relu](x * scale + shift) [+ res], scale / shift either from `scale_shift` or folded in from the
 *   producer's `stats` (+ GroupNorm gamma / beta, groups = 0: instance norm; 1, 2, 4 or 8 channels per group)
 *   -> hi / lo operand tensors (halo > 0: reflection halo;
 *   s2d = 1: four parity planes for a stride-2 consumer; channels padded to Cp with zeros) and / or fp32 NHWC.
 * icon_ew_nhwc: mode 0 a + b (+ c), 1 avg_pool2(a), 2 b + bicubic_up2(a, align_corners), 3 relu(a * scale + shift) with
 *   b = the [N][C] (scale, shift) table of icon_norm_finalize; optional stats of the result.
 * icon_nchw_to_nhwc / icon_nhwc_to_nchw: layout adaptors (the former with optional stats).
 * icon_conv7_head: 7 x 7 reflection-padded 64 -> (1..3) channel convolution from fp32 NHWC to NCHW, act 2 = tanh. */
size_t icon_conv_nhwc_workspace_bytes(int N, int Ht, int Wt, int Cout, int splits);
int icon_conv_nhwc(const void *a_hi, const void *a_lo, const int64_t *dims, const int64_t *strides, const void *wt_packed,
                   int wt_chunks, const float *bias, float *out, int OHf, int OWf, int Cs, int co_off, int Cout, int N,
                   int Ht, int Wt, int osy, int osx, int ooy, int oox, int nplanes, int ntaps, const int *taps, int cpt,
                   int n_tile, int splits, double *stats, void *ws, size_t ws_bytes, icon_stream_t stream);
int icon_norm_finalize(const double *stats, const float *gamma, const float *beta, float *scale_shift, int N, int C,
                       int groups, double count, float eps, icon_stream_t stream);
int icon_act_nhwc(const float *x, int Cs_in, int ci_off, const float *scale_shift, const double *stats, const float *gamma,
                  const float *beta, int groups, float eps, const float *res, void *hi, void *lo, float *f32, int N, int H,
                  int W, int C, int Cp, int halo, int s2d, int relu, icon_stream_t stream);
/* icon_splitk_instnorm_act: for a split-K convolution followed by InstanceNorm2d(affine=False) [+ ReLU] [+ residual]
 * (the ResnetBlocks): call icon_conv_nhwc with out == NULL (the split partials stay parked in its workspace
 * [splits][N][H][W][C]), then this -- one kernel sums the partials in split order, computes each channel's mean /
 * variance over the image inside a block (no atomics), normalises and writes the next operand (hi / lo with a
 * reflection halo) and / or the fp32 tensor. */
int icon_splitk_instnorm_act(const float *partial, int splits, const float *bias, const float *res, void *hi, void *lo,
                             float *f32, int N, int H, int W, int C, int Cp, int halo, int relu, float eps,
                             icon_stream_t stream);
/* icon_col2im7: second half of the 7 x 7 output head computed as GEMM + col2im: P [N][H][W][Ps] holds, per INPUT pixel,
 * the products with every tap's weights (column (ky * 7 + kx) * Cout + co; the first half is icon_conv_nhwc with the
 * regrouped 1 x 1 weights); out[n][co][y][x] = act(bias + sum over the 49 reflected neighbours). */
int icon_col2im7(const float *P, const float *bias, float *y, int N, int H, int W, int Cout, int Ps, int act,
                 icon_stream_t stream);
int icon_ew_nhwc(int mode, const float *a, const float *b, const float *c, float *y, double *stats, int N, int H, int W,
                 int C, icon_stream_t stream);
int icon_nchw_to_nhwc(const float *x, float *y, double *stats, int N, int C, int64_t HW, icon_stream_t stream);
int icon_nhwc_to_nchw(const float *x, float *y, int N, int C, int Cs, int c_off, int64_t HW, icon_stream_t stream);
/* icon_stem_pack: NCHW fp32 image -> hi / lo fp16 [N * sy][Hrows][Wp][Cp8] (Cp8 = 8 | 16 >= Cin) with a 3-pixel halo
 * (reflect = 1: ReflectionPad2d(3); 0: zeros) and rows split into sy parity planes: the operand layout of the 7 x 7
 * first layers (K runs over the 8 x Cp8 contiguous values of a filter row; icon_conv_nhwc reads it through a tensor
 * map whose W stride is sx * Cp8 elements). */
int icon_stem_pack(const float *x, void *hi, void *lo, int N, int Cin, int H, int W, int Cp8, int Wp, int Hrows, int sy,
                   int reflect, icon_stream_t stream);
int icon_conv7_head(const float *x, const float *w, const float *bias, float *y, int N, int H, int W, int Cin, int Cout,
                    int act, icon_stream_t stream);
/* ---- clean_mesh (csrc/clean.cu; reference lib/dataset/mesh_util.py:778-791: trimesh split + largest component).
 * faces int64 [nf][3] indexing nv vertices.  _count: union-find over the vertices, picks the component with the most
 * vertices (ties: the one containing the smallest vertex id), leaves the kept counts in d_counts[0..1] and the
 * re-index tables in `ws`; _emit writes the compacted float32 vertices / int32 faces (ascending original order). */
size_t icon_clean_mesh_workspace_bytes(int64_t nv, int64_t nf);
int icon_clean_mesh_count(const int64_t *faces, int64_t nv, int64_t nf, void *ws, size_t ws_bytes, int64_t *d_counts,
                          icon_stream_t stream);
int icon_clean_mesh_emit(const void *verts, int verts_f64, const int64_t *faces, int64_t nv, int64_t nf, const void *ws,
                         float *out_verts, int32_t *out_faces, icon_stream_t stream);
int icon_group_norm(const float *x, const float *gamma, const float *beta, const float *res, float *y, int N,
                    int C, int HW, int groups, float eps, int relu, void *stats_ws /* 16*N*groups bytes or NULL */,
                    icon_stream_t stream);
/* nn.Conv3d (B = 1, Cout <= 8, cubic kernel / stride / padding / dilation) followed by eval-mode BatchNorm3d
 * folded into per-channel (scale, shift), optional residual add and ReLU: the layers of PaMIR's VolumeEncoder
 * (lib/net/VE.py:96-183).  x [Cin,D,H,W], w [Cout,Cin,k,k,k], y [Cout,OD,OH,OW]. */
int icon_conv3d(const float *x, const float *w, const float *scale, const float *shift, const float *res, float *y,
                int Cin, int Cout, int D, int H, int W, int k, int stride, int pad, int dil, int relu,
                icon_stream_t stream);
/* F.avg_pool2d(x, 2, stride=2); planes = N*C */
int icon_avg_pool2(const float *x, float *y, int64_t planes, int H, int W, icon_stream_t stream);
/* y = add + F.interpolate(x, scale_factor=2, mode="bicubic", align_corners=True) (HGFilters.py:70-76) */
int icon_bicubic_up2_add(const float *x, const float *add, float *y, int64_t planes, int H, int W,
                         icon_stream_t stream);
/* y = cat(a, b, c, dim=1) + res (ConvBlock, net_util.py:273-278) */
int icon_cat3_add(const float *a, const float *b, const float *c, const float *res, float *y, int N, int C1, int C2,
                  int C3, int64_t HW, icon_stream_t stream);
int icon_add3(const float *a, const float *b, const float *c, float *y, int64_t n, icon_stream_t stream);
/* y = x / ||x||_2 (over 3 channels, no eps) * (sum_c |image| != 0)  (NormalNet.py:88-97) */
int icon_normalize_mask(const float *x, const float *image, float *y, int N, int Cimg, int64_t HW,
                        icon_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
