/*
 * nbp_b200.h -- C ABI of libnbp_b200.so: the B200-native exploration inner loop of NextBestPath.
 *
 * The reference (shiyao-li/NextBestPath) is pure Python and has no FFI; the interface each entry
 * point replaces is a Python call site, cited below relative to /root/reference.  INTEGRATION.md
 * shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory,
 *     including workspaces (sizes from the *_workspace_bytes queries); the library never allocates,
 *     frees or synchronises;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - return value: 0 = OK, NBP_ERR_INVALID (<0) = bad argument, >0 = cudaError_t; a human-readable
 *     message for the calling thread is available from nbp_last_error();
 *   - ragged batches use CSR offsets; cameras are packed R[n,9] (row-major, world->view is
 *     x_view = x_world * R + T with row vectors, PyTorch3D convention) and T[n,3];
 *   - kernels are compiled for sm_100a only; there is no CPU fallback.
 */
#ifndef NBP_B200_H_
#define NBP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBP_OK 0
#define NBP_ERR_INVALID (-1)
#define NBP_ERR_WORKSPACE (-2)
#define NBP_ERR_UNSUPPORTED (-3)

#define NBP_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------ misc */
int nbp_version(void);
const char* nbp_last_error(void);
/* number of kernel launches this library has enqueued from the calling process (bench gpu_launches) */
uint64_t nbp_launch_count(void);
/* add n to that count: the host side replays captured CUDA graphs of this library's kernels (NBP.forward eval) and accounts
 * for the kernels of each replay here */
void nbp_count_launches(uint64_t n);

/* ------------------------------------------------------------------------------------------ a2
 * Batched depth rasterisation.  Replaces Camera.capture_image's
 *     images, fragments = self.renderer(mesh, cameras=fov_camera); depth = fragments.zbuf
 * (macarons/utility/macarons_utils.py:2759-2762; renderer built at :905-937 with blur_radius 0,
 * faces_per_pixel 1, FoVPerspectiveCameras fov 60 / znear 1 / z_clip znear/2).
 *
 * n_views cameras, view v looks at scene view_scene[v].  verts [sumV,3] fp32 packed;
 * faces [sumF,3] int32 packed, indices LOCAL to the scene's vertex block.
 * zbuf [n_views,H,W] fp32 view-space z, -1 where no face; pix_to_face [n_views,H,W] int32 scene-local
 * face index or -1 (may be NULL).  total_view_faces = sum over views of the face count of its scene.
 * Workspace (caller-owned, size from nbp_raster_workspace_bytes): per view up to 2 triangle records (96 B) + pixel boxes (8 B) per
 * face and 32 coarse-bin lists of 8-byte entries with the same capacity; only the used prefix of each list is touched.  Images whose
 * coarse bins (<= 32 per view, multiples of 16 px) would exceed 256 px are rejected (NBP_ERR_INVALID).
 */
size_t nbp_raster_workspace_bytes(int n_views, int64_t total_view_faces);
int nbp_raster_depth_batched(const float* verts, const int32_t* faces,
                             const int64_t* vert_offsets, const int64_t* face_offsets, int n_scenes,
                             const int32_t* view_scene, const float* R, const float* T, int n_views,
                             int64_t total_view_faces, int max_faces_per_scene,
                             int H, int W, float tan_half_fov, float z_clip,
                             float* zbuf, int32_t* pix_to_face,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------ a4+a5
 * Depth back-projection, validity filter, subsample and append to per-scene point clouds.  Replaces
 * Camera.compute_partial_point_cloud -> Camera.project_depth_in_3D -> unproject_points
 * (macarons_utils.py:2811-2847, :2788-2809, NDC tables :2270-2279) and the growing
 * torch.vstack((full_pc, part_pc)) at next_best_path/testers/nbp_planning.py:105,352.
 *
 * Frame f (zbuf[f], R[f], T[f]) belongs to cloud frame_scene[f].  A pixel is valid when mask != 0 -- mask NULL
 * means the stored frame mask zbuf > -1 (macarons_utils.py:2771) -- and zbuf < fov_range (fov_range <= 0: no
 * range test), i.e. points_mask = mask * (depth < fov_range) of macarons_utils.py:2825.
 * Of the n valid pixels of a frame, k = (int)(n * gathering_factor) are kept:
 *   gathering_factor >= 1 : all of them, in row-major pixel order (the parity path: the caller applies
 *                           torch.randperm(n)[:k] itself, exactly as macarons_utils.py:2837);
 *   otherwise             : the k pixels with the smallest keys of a keyed bijection of the pixel
 *                           index (counter-based, seed + frame_uid[f]) -- a uniform random k-subset,
 *                           written in row-major order.  Same distribution as randperm(n)[:k] as a SET;
 *                           the consumer (the grid histogram) is order-independent.
 * Points are appended to cloud[scene] (layout [n_scenes, cloud_capacity, 3] fp32) at cloud_len[scene],
 * frames of one scene in frame order; cloud_len is updated.  Points beyond cloud_capacity are dropped
 * and counted in *overflow (device int32, may be NULL).
 * frame_valid / frame_kept [n_frames] int32 receive n and k (may be NULL).
 */
size_t nbp_backproject_workspace_bytes(int n_frames);
/* Optional extra scratch: a workspace of nbp_backproject_workspace_bytes(n) + nbp_backproject_key_cache_bytes(n, H, W) bytes lets the
 * sub-sampling path (gathering_factor < 1) evaluate every pixel's selection key once (4 bytes per pixel kept in the scratch) instead of
 * once per selection pass and once more when writing; the result is identical. */
size_t nbp_backproject_key_cache_bytes(int n_frames, int H, int W);
int nbp_backproject_append(const float* zbuf, const uint8_t* mask, const float* R, const float* T,
                           const int32_t* frame_scene, const int32_t* frame_uid, int n_frames,
                           int H, int W, float tan_half_fov, float fov_range,
                           double gathering_factor, uint64_t seed,
                           float* cloud, int32_t* cloud_len, int64_t cloud_capacity, int n_scenes,
                           int32_t* frame_valid, int32_t* frame_kept, int32_t* overflow,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------ a6-a10
 * Height-slab split + egocentric transform + histogram into the model-input grid.  Replaces, per scene,
 *     bins = torch.bucketize(full_pc[:, 1], y_bins[:-1]) - 1                (nbp_planning.py:114-115)
 *     transform_points_to_n_pieces(..., no_rotation=True)                    (next_best_path/utility/utils.py:166-196)
 *     map_points_to_n_imgs(points_2d, (S,S), (lo,hi), device)                (utils.py:198-223)
 *     the trajectory image from camera.X_cam_history                         (nbp_planning.py:130-132)
 *     torch.cat -> (1, n_pieces+1, S, S)                                     (nbp_planning.py:126-127,166)
 *
 * cloud [n_scenes, cloud_capacity, 3], cloud_len [n_scenes]; traj [n_scenes, traj_capacity, 3],
 * traj_len [n_scenes] (traj may be NULL: channel n_pieces stays zero); pose [n_scenes,5] (x,y,z,elev,azim);
 * slab_bounds [n_scenes, max_bounds] = the reference's y_bins[:-1] computed by the caller with the
 * reference's own torch.arange expression (it has n_pieces or n_pieces+1 entries),
 * n_bounds [n_scenes].  grid [n_scenes, n_pieces+1, S, S] fp32 counts; it is zeroed by this call.
 * Cell indices are rint((p - lo) * fp32(S/(hi-lo))) with every operation rounded to fp32: bit-exact
 * with torch.  max_points bounds the longest cloud (<= cloud_capacity) and only sizes the launch.
 */
int nbp_grid_scatter(const float* cloud, const int32_t* cloud_len, int64_t cloud_capacity,
                     const float* traj, const int32_t* traj_len, int64_t traj_capacity,
                     const float* pose, const float* slab_bounds, const int32_t* n_bounds, int max_bounds,
                     int n_scenes, int n_pieces, int S, float range_lo, float range_hi,
                     int64_t max_points, float* grid, void* stream);

/* Plain map_points_to_n_imgs (utils.py:198-223): points_2d [n, m, 2] fp32 -> out [n, S0, S1] fp32
 * counts (zeroed here).  n_valid_host: m may be ragged through lens [n] (NULL: all m). */
int nbp_map_points(const float* points_2d, const int32_t* lens, int n, int64_t m, int S0, int S1,
                   float range_lo, float range_hi, float* out, void* stream);

/* get_point_position_in_the_img (utils.py:160-164): points [n,2] fp32 -> cells [2,n] int64. */
int nbp_point_cells(const float* points_2d, int64_t n, int S0, int S1, float range_lo, float range_hi,
                    int64_t* cells, void* stream);

/* ------------------------------------------------------------------------------------------ a11
 * NBP attention-U-Net forward, eval mode (next_best_path/networks/nbp_model.py:110-160).  The network is
 * a sequence of the calls below over NHWC fp16 activations; BatchNorm (eval) and the conv bias are folded
 * into a per-channel fp32 (scale, shift) applied to the fp32 accumulator.  Channel strides (`ld`) and
 * offsets let a producer write straight into a concatenation buffer (torch.cat at :128,133,...).
 */
typedef struct nbp_conv_desc {
    int precise;                           /* numeric mode = format of the SOURCE tensors and of the weights:
                                              0: single fp16 plane;
                                              1: fp16x2 split format (hi plane + lo*2048 plane), 3 tensor passes, fp32-grade results;
                                              2: fp16 hi plane + an e4m3 pair plane (per 64-channel group: 64 bytes e4m3(x / 8), then 64 bytes
                                                 e4m3((x - hi) * 2048 / 8)); the hi product runs on the fp16 pipe, both correction products as one
                                                 e4m3 reduction (kind::f8f6f4, 2x rate): 2 pass-equivalents, ~15-bit operands */
    const void* src0; int c0; int ld0; int lo0;  /* NHWC fp16: c0 channels per pixel, pixel stride ld0 elements; with
                                              precise, the lo plane starts lo0 elements after src0 */
    const void* src1; int c1; int ld1; int lo1;  /* optional second source, concatenated after src0 along channels */
    int n, h, w;                           /* images and spatial size (stride 1, same-size output) */
    int taps;                              /* 1 = 1x1 conv, 9 = 3x3 conv with zero padding 1, 4 = see up2x */
    int up2x;                              /* 1: the op is nn.Upsample(2, nearest) followed by a 3x3 conv (up_conv, nbp_model.py:23-34)
                                              computed WITHOUT materialising the upsampled tensor: output pixel (2y+py, 2x+px) is a 2x2
                                              conv of the source with parity-specific pre-summed weights.  n,h,w are the SOURCE dims, dst
                                              is [n][2h][2w]; weight holds 4 blocks (parity = py*2+px), each [c_out][4 taps][c_in] packed
                                              like a normal weight; tap t reads source (y + t/2 - 1 + py, x + t%2 - 1 + px). 2.25x fewer MACs. */
    const void* weight;                    /* fp16 [c_out][taps][c0+c1] (K-major); tap = ky*3+kx.  precise: per tile of
                                              BN = 128|64|32 output channels (largest dividing c_out) BN hi rows then BN lo rows; mode 2: the
                                              lo rows hold, per 64-element K slice, 64 bytes e4m3(W_lo*2048*s) then 64 bytes e4m3(W_hi*s) */
    int c_out;                             /* multiple of 32 */
    const float* scale; const float* shift;/* [c_out] fp32: y = acc*scale + shift */
    int relu;
    void* dst; int dst_ld; int dst_c_off;  /* NHWC fp16 output, written at channels [dst_c_off, dst_c_off+c_out).  The epilogue stores 32-byte
                                              pieces: dst and pool_dst 32-byte aligned; dst_ld, dst_c_off, dst_lo_off, pool_ld, pool_lo_off
                                              multiples of 16 (fp32 output: dst_ld, dst_c_off multiples of 8) */
    int dst_lo_off;                        /* precise: lo plane written dst_lo_off elements after the hi channels */
    int out_f32;                           /* 1: dst is plain fp32 NHWC [pix][dst_ld] (used for dgrad: gradients are fp32) */
    int k_chunk;                           /* precise mode: number of 64-element K slices summed inside the tensor core's (truncating)
                                              fp32 accumulator before the partial sum is folded into fp32 round-to-nearest registers;
                                              0 = default (8).  Smaller = more accurate, slower (measured: 8 -> 3.6e-5 network error,
                                              2 -> floor of the 22-bit operands, -10 % throughput) */
    int dst_fmt;                           /* precise != 0, out_f32 == 0: second-plane format WRITTEN to dst: 0 = the mode's own, 1 = fp16 lo*2048 (consumers run
                                              mode 1), 2 = e4m3 pair plane (consumers run mode 2; dst_c_off % 32 == 0, dst_lo_off % 64 == 0).
                                              A mode-1 layer may write format 2 and vice versa: the network mixes both (nbp_model.py) */
    int pool_fmt;                          /* same for pool_dst */
    float w_lo_scale;                      /* mode 2: 1 / (2048 s), s = the power of two the e4m3 weight rows were scaled by */
    uint64_t* sat_count;                   /* optional device counter (NULL = off): += 1 for every (pixel, 32-channel group) written in
                                              format 2 in which a value exceeded the e4m3 window (|x| > 3584: the planes hold x / 8).  Such
                                              elements fall back to fp16 precision; a non-zero count means the inputs are far outside the
                                              range the BatchNorm statistics were calibrated for */
    void* pool_dst; int pool_ld; int pool_lo_off; /* optional (NULL = off): ALSO write nn.MaxPool2d(2,2) of the output (nbp_model.py:68,113-121)
                                              as an NHWC fp16 tensor [n][h/2][w/2][pool_ld] (lo plane pool_lo_off elements after the hi
                                              channels), fused into the epilogue: the encoder needs both the skip tensor and its pooled
                                              copy.  h, w even; not with up2x / out_f32 */
    const float* dot_w; float* dot_out;    /* optional (dot_out NULL = off) "dot" epilogue: the c_out activations of a pixel are NOT stored (dst may be
                                              NULL) but contracted in fp32 with dot_w [c_out] (16-byte aligned):
                                              dot_out[pixel] = f(dot_scale * sum_c y[c] * dot_w[c] + dot_shift), f = sigmoid if dot_sigmoid else
                                              identity; dot_out fp32 [n][h][w] ([n][2h][2w] with up2x).  Fuses the 1-channel 1x1 convolution that
                                              consumes this layer -- Attention_block.psi (nbp_model.py:49-53,60) and Final2 (:106-108) -- so the
                                              layer's output never goes to memory.  c_out = 32, 64 or 128 (one n-tile); the whole K reduction
                                              runs as one in-TMEM chain (k_chunk is ignored); not with out_f32 / pool_dst */
    float dot_scale; float dot_shift; int dot_sigmoid;
    const void* gate_src; int gate_c; int gate_ld; int gate_lo;   /* optional gate on top of the dot epilogue (NULL = off; dot_w required, dot_out
                                              optional): dst[pixel][dst_c_off + c] = gate_src[pixel][c] * f(dot), c < gate_c -- the tail of
                                              Attention_block, `x * psi` (nbp_model.py:62), written by the attention GEMM itself in dst_fmt.
                                              gate_src: NHWC tensor [n][h][w][gate_ld] in the sources' format (`precise`), second plane at gate_lo;
                                              32-byte aligned, gate_c % 32 == 0 (64 in mode 2); not with up2x */
    int dot_n; const float* dot_bias; float* dot_max;   /* dot epilogue with several vectors (0 = 1): dot_w is [dot_n][c_out], dot_n <= 8, and
                                              dot_out [n][dot_n][h][w] = f(dot_scale * <y, dot_w[o]> + dot_shift + dot_bias[o]) -- a fused
                                              1x1 convolution to dot_n channels in NCHW fp32 (Final1, nbp_model.py:89,133); dot_bias
                                              (optional) [dot_n]; dot_max (optional) [n][h][w] = max over o, the heading read-out
                                              torch.max(predicted_value_map, dim=1) of nbp_planning.py:193 */
} nbp_conv_desc;

/* tcgen05/TMEM/TMA implicit-GEMM convolution: conv_block / up_conv / Attention_block W_g,W_x (nbp_model.py:8-62) */
int nbp_conv_fwd(const nbp_conv_desc* desc, void* stream);
/* The pointwise kernels below take, for every NHWC fp16 tensor, the pixel stride `ld` (elements) and the
 * offset `lo` of the second plane (0 = single-plane fp16 tensor).  `fmt` (nbp_conv_first, nbp_att_gate, nbp_conv1x1_head) selects
 * the second-plane format of ALL tensors of the call: 1 = fp16 (x - hi) * 2048, 2 = the e4m3 pair plane of nbp_conv_desc mode 2. */
/* Measurement aid for bench.py (not part of the data path): between _begin and _end every conv_gemm launch is
 * bracketed by CUDA events on its stream; _end waits for them and returns the summed device time, the summed
 * ALGORITHMIC flops (2*M*N*K of the convolution, independent of the numeric mode) and the launch count.
 * These two calls are the only ones in the library that create CUDA events / synchronise. */
int nbp_conv_profile_begin(int max_launches);
int nbp_conv_profile_end(double* total_ms_host, double* total_flops_host, uint64_t* launches_host, uint64_t* dropped_host);
/* Conv1.conv.0: x fp32 NCHW [n,c_in,h,w] counts -> NHWC fp16; weight fp32 [9*c_in][c_out]
 * (tap-major, then input channel), c_out = 64; fused affine + optional ReLU (nbp_model.py:11-13).  Train mode calls it
 * with scale 1 / shift bias / relu 0 to get the raw pre-BatchNorm tensor. */
int nbp_conv_first(const float* x, int n, int c_in, int h, int w, const float* weight, const float* scale,
                   const float* shift, int c_out, int relu, void* dst, int dst_ld, int dst_lo, int fmt, void* stream);
/* nn.MaxPool2d(2,2) (nbp_model.py:68) and nn.Upsample(scale_factor=2) nearest (:27) on NHWC fp16 */
int nbp_maxpool2x2(const void* src, int n, int h, int w, int c, int ld_src, int lo_src, void* dst, int ld_dst, int lo_dst, void* stream);
int nbp_upsample2x(const void* src, int n, int h, int w, int c, int ld_src, int lo_src, void* dst, int ld_dst, int lo_dst, void* stream);
/* Attention_block tail (nbp_model.py:49-62): psi = sigmoid(psi_scale * dot(a, w_psi) + psi_shift); dst = x * psi.
 * `fmt` applies to x and dst; `a` (never a GEMM operand, f_int may be 32) always carries the fp16 lo plane (format 1).
 * a [npix] x f_int = relu(BN(W_g g) + BN(W_x x)); x [npix] x f_l; dst written at channel dst_c_off */
int nbp_att_gate(const void* a, int f_int, int ld_a, int lo_a, const void* x, int f_l, int ld_x, int lo_x,
                 const float* w_psi, float psi_scale, float psi_shift,
                 void* dst, int dst_ld, int dst_c_off, int dst_lo, int64_t npix, int fmt, void* stream);
/* The same tail with psi [npix] fp32 already computed -- by the dot epilogue of the attention GEMM (nbp_conv_desc.dot_out):
 * dst[:, dst_c_off : dst_c_off + f_l] = x * psi (nbp_model.py:62) */
int nbp_att_scale(const float* psi, const void* x, int f_l, int ld_x, int lo_x,
                  void* dst, int dst_ld, int dst_c_off, int dst_lo, int64_t npix, int fmt, void* stream);
/* Final1 (256->8) and Final2 (64->1, sigmoid) (nbp_model.py:89,106-108): NHWC fp16 in, NCHW fp32 out [n,c_out,hw];
 * weight fp32 [c_out][c_in], c_out in {1, 8}.  dst_max (optional, NULL = off): [n,hw] fp32 = max over the c_out channels, the
 * heading read-out `torch.max(predicted_value_map, dim=1)` of next_best_path/testers/nbp_planning.py:193 fused into the head */
int nbp_conv1x1_head(const void* src, int c_in, int ld_src, int lo_src, const float* weight, const float* bias, int c_out,
                     int sigmoid, float* dst, float* dst_max, int n, int64_t hw, int fmt, void* stream);

/* ------------------------------------------------------------------------------------------ section 8f row 3 (second half)
 * Ground-truth obstacle map of the data collection: get_binary_obstacle_array(mesh, camera_pose, view_size)
 * (next_best_path/utility/utils.py:226-262, called at nbp_utils.py:638 -> `current_gt_2d_layout`): the mesh cut by the horizontal
 * plane through the camera (trimesh.intersections.mesh_plane), drawn as lines into a view_size x view_size window centred on the
 * camera and binarised at S x S (rows towards -z, columns towards -x).  One call for n_maps (scene, pose) pairs.
 * meshes: packed as for nbp_raster_depth_batched; map_scene[n_maps] (NULL = identity) = scene of map b; pose [n_maps][5];
 * out [n_maps][S][S] fp32 of 0/1 (zeroed by the call); half_width_px = half the drawn line width in output pixels
 * (1.35 = matplotlib's default 1.5 pt line after the reference's resize).  Geometry restated in oracle/section_oracle.c. */
int nbp_gt_obstacle_map(const float* verts, const int32_t* faces, const int64_t* vert_offsets, const int64_t* face_offsets,
                        const int32_t* map_scene, const float* pose, int n_maps, int64_t max_faces_per_scene, int S,
                        float view_size, float half_width_px, float* out, void* stream);

/* ------------------------------------------------------------------------------------------ a11 (train mode), a13
 * Train-mode BatchNorm2d (batch statistics, running-stat update: momentum 0.1, unbiased running variance, eps 1e-5 --
 * torch.nn.BatchNorm2d as instantiated at nbp_model.py:12,15,28,41,46,51) and the backward pass of NBP.forward
 * (autograd through nbp_model.py:110-160, driven by next_best_path/utility/nbp_utils.py:378-390).
 * Forward activations: NHWC fp16x2 split tensors (ld, lo as above).  Gradients between layers: plain NHWC fp32.
 * The raw pre-BatchNorm tensor `z` of nbp_bn_train_stats / nbp_affine_act / nbp_att_pre / nbp_bn_bwd[_split] (and the destination of
 * nbp_conv_first) may instead be PLAIN FP32 NHWC: pass lo < 0 and ld = row stride in floats (train mode uses this: the
 * cancellation in z - mean needs fp32's 24 bits, see scripts/gradient_study.py).
 * `workspace` arguments are caller-owned fp64 scratch of the stated length; they are zeroed by the call. */
int nbp_bn_train_stats(const void* z, int ld, int lo, int64_t npix, int C, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps,
                       float* mean, float* invstd, float* scale, float* shift, double* workspace /* [2C] */, void* stream);
/* y = act(z*scale + shift) */
int nbp_affine_act(const void* z, int ld_z, int lo_z, int64_t npix, int C, const float* scale, const float* shift, int relu,
                   void* y, int ld_y, int lo_y, void* stream);
/* a = relu(BN_g(zg) + BN_x(zx)) with the two normalisations given as (scale, shift) (nbp_model.py:57-59) */
int nbp_att_pre(const void* zg, const void* zx, int ld, int lo, int64_t npix, int C, const float* sg, const float* tg,
                const float* sx, const float* tx, void* a, int ld_a, int lo_a, void* stream);
/* zpsi = conv1x1(a) (1 channel) and its train-mode BatchNorm2d(1): stat4 = {mean, invstd, scale, shift} on the device */
int nbp_psi_train(const void* a, int ld_a, int lo_a, int64_t npix, int C, const float* w_psi, const float* b_psi,
                  const float* gamma1, const float* beta1, float* running_mean1, float* running_var1, float momentum, float eps,
                  float* zpsi, float* stat4, double* workspace /* [2] */, void* stream);
/* psi = sigmoid(zpsi*scale+shift); dst[:, c_off:c_off+C] = x * psi; psi_out [npix] saved for backward */
int nbp_att_apply(const float* zpsi, const float* psi_scale, const float* psi_shift, const void* x, int ld_x, int lo_x, int64_t npix, int C,
                  void* dst, int ld_d, int c_off, int lo_d, float* psi_out, void* stream);
/* BatchNorm(+ReLU) backward: dz (fp32 NHWC, row stride ld_dz), dgamma += , dbeta += ; *amax = max|dz| (for operand scaling) */
int nbp_bn_bwd(const float* dy, int ld_dy, const void* z, int ld_z, int lo_z, int64_t npix, int C, const float* scale, const float* shift,
               const float* mean, const float* invstd, const float* gamma, int relu, float* dz, int ld_dz, float* amax,
               float* dgamma, float* dbeta, double* workspace /* [2C] */, void* stream);
/* Same backward, but dz leaves directly as the fp16x2 split GEMM operand of dgrad / wgrad (dz_split, pixel stride ld_s, lo plane at
 * lo_s), scaled by a power of two chosen from an upper bound of max|dz| that the reduction pass provides (amax2: 2-float scratch);
 * inv_scale_vec[0..n_vec) = 2^-k.  dz may be NULL then.  Replaces nbp_bn_bwd + nbp_to_split_nhwc (no fp32 dz round trip). */
int nbp_bn_bwd_split(const float* dy, int ld_dy, const void* z, int ld_z, int lo_z, int64_t npix, int C, const float* scale, const float* shift,
                     const float* mean, const float* invstd, const float* gamma, int relu, float* dz, int ld_dz, float* amax,
                     float* dgamma, float* dbeta, double* workspace /* [2C] */, void* dz_split, int ld_s, int lo_s, float* amax2,
                     float* inv_scale_vec, int n_vec, void* stream);
/* fp32 NHWC gradient -> fp16x2 split NHWC scaled by 2^k (amax*2^k in [128,256)); inv_scale_vec[0..n_vec) = 2^-k */
int nbp_to_split_nhwc(const float* src, int ld_s, int64_t npix, int C, const float* amax, void* dst, int ld_d, int lo_d,
                      float* inv_scale_vec, int n_vec, void* stream);
/* fp32 NHWC (scaled like above) or split NHWC (unscaled) -> channel-major split [2][C][n*h][w_pad] for nbp_conv_wgrad;
 * rows padded from w to w_pad pixels; dst[.., x] = src[.., x + dx] (x-shifted copy; entries without a source are not
 * written: the caller pre-zeroes dst when w_pad != w or dx != 0); plane_stride = C*row_stride */
int nbp_to_split_cnhw(const float* src_f32, const void* src_split, int ld_s, int lo_s, int64_t npix, int w, int w_pad, int dx, int C, const float* amax,
                      void* dst, int64_t row_stride, int64_t plane_stride, float* inv_scale_out, void* stream);
/* tcgen05 weight-gradient GEMM: dweight[c_out][taps][c_in] += inv_scale * sum_pixels dz[p][co] * x[p+tap][ci].
 * dz and x are NHWC fp16x2 split tensors (dz from nbp_to_split_nhwc, x the saved forward activation); channel counts
 * multiples of 64 (pad with zero channels).  Consumed as MN-major UMMA operands straight from the NHWC layout.
 * max_k_tiles > 0 bounds the number of 64-pixel slices accumulated inside the tensor core per partial result (the
 * partials are combined with fp32 atomics); 0 = only as many splits as needed to fill the GPU. */
int nbp_conv_wgrad(const void* dz, int c_out, int ld_dz, int lo_dz, const void* x, int c_in, int ld_x, int lo_x,
                   int n, int h, int w, int taps, const float* inv_scale, float* dweight, int max_k_tiles, void* stream);
/* debugging aid: zero-copy host ints [0..3] = {wait tag, block, thread, parity} written when a pipeline wait of the wgrad
 * kernel times out (the kernel then traps instead of hanging) */
int nbp_debug_attach_wgrad(int* device_visible_host_ptr);
int nbp_maxpool2x2_bwd(const float* dy, const void* x, int ld_x, int lo_x, int n, int h, int w, int C, float* dx, int accumulate, void* stream);
int nbp_upsample2x_bwd(const float* dy, int n, int h, int w, int C, float* dx, int accumulate, void* stream);
/* Attention_block backward from d(x*psi) to d(x) (+=), d(relu(g1+x1)) = dpre, and the psi conv / BN parameter grads (+=) */
int nbp_att_bwd(const float* dout, int ld_do, const void* x, int ld_x, int lo_x, const float* psi, const float* zpsi, int64_t npix, int C_l,
                const float* stat4, const float* gamma1, const void* a, int ld_a, int lo_a, int C_int, const float* w_psi,
                float* dx_skip, int accumulate, float* dpre, float* dt, float* dw_psi, float* dgamma1, float* dbeta1,
                double* workspace /* [2 + C_int] */, void* stream);
/* Final1 / Final2 backward: dsrc fp32 NHWC [pix][c_in]; dweight, dbias += ; out_sigmoid != NULL applies s(1-s) */
int nbp_head_bwd(const float* dout, const float* out_sigmoid, const void* src, int c_in, int ld_s, int lo_s, const float* weight, int c_out,
                 int n, int64_t hw, float* dsrc, float* dweight, float* dbias, double* workspace /* [c_out*c_in + c_out] */, void* stream);
/* Conv1.conv.0 weight gradient: dweight[9*c_in][64] += sum_p x[p+tap][ci] * dz[p][co] (x fp32 NCHW, dz fp32 NHWC) */
int nbp_stem_wgrad(const float* x, int n, int c_in, int h, int w, const float* dz, float* dweight, double* workspace /* [9*c_in*64] */, void* stream);
/* y[p][c] += x[p*ld_x + c] */
int nbp_add_f32(float* y, const float* x, int ld_x, int64_t npix, int C, void* stream);

/* ------------------------------------------------------------------------------------------ SURVEY section 8(f) row 1
 * The re-plan branch right after NBP.forward (next_best_path/testers/nbp_planning.py:166-233).
 * nbp_obstacle_fuse (:168-190): fused = (pred >= threshold); where the cloud has any point (all_cnt > 0) fused = (slice_cnt > 0)
 * (slice = points within +-0.1 of the camera height); along the trajectory (traj > 0) fused = 0.  full_proj = (all_cnt > 0).
 * pred/fused/full_proj [n, S, S] fp32; the three count images are [S*S] per scene with the given scene strides (elements), e.g.
 * channels of nbp_grid_scatter outputs. */
int nbp_obstacle_fuse(const float* pred, const float* all_cnt, int64_t all_stride, const float* slice_cnt, int64_t slice_stride,
                      const float* traj, int64_t traj_stride, int n, int S, float threshold, float* fused, float* full_proj, void* stream);
/* nbp_candidate_scores (:193-231 + check_pixel_values macarons_utils.py:86-100): cand [n, max_cand, 3] world positions, n_cand [n],
 * skip [n, max_cand] (known collisions, may be NULL), pose [n,5], value_map [n, n_ch, Sv, Sv], full_proj [n, S, S].
 * Per candidate: cell (value-map row, col; -1 if outside), value = max over channels at that cell, density = full_proj at the
 * S-grid cell, valid = inside the value map AND some full_proj cell == 1 within +-window px.  The reference's score is
 * value - 10*density (computed by the caller in double, as the Python loop does). */
int nbp_candidate_scores(const float* cand, const int32_t* n_cand, int max_cand, const uint8_t* skip, const float* pose,
                         const float* value_map, int n_ch, int Sv, const float* full_proj, int S, int n,
                         float range_lo, float range_hi, int window, float* value, float* density, int32_t* cell, uint8_t* valid, void* stream);

/* ------------------------------------------------------------------------------------------ SURVEY section 8(f) row 2
 * calculate_coverage_percentage (next_best_path/utility/long_term_utils.py:437-468; called every pose at
 * next_best_path/testers/nbp_planning.py:71 and next_best_path/utility/nbp_utils.py:572), for n_scenes scenes in one call:
 *   coverage[b] = mean_i [ min_j || gt_b[i] - sampled_b[j] || < threshold ],
 *   sampled_b = cloud_b if len <= weight*G_b, else a uniform random subset of weight*G_b points (sample_idx [n, sample_stride]
 *   int64 gives the subset explicitly -- e.g. torch.randperm(len)[:k] for the reference's own stream; NULL = keyed permutation).
 * The ground truth comes bucketed into a uniform grid of edge `cell` >= threshold (built once per scene by the caller):
 * gt_sorted [sum G][3] points ordered by cell index (z*ny + y)*nx + x, gt_off [n+1]; cell_start = per scene ncells+1 scene-local
 * starts, cell_off [n+1]; origin [n][3], dims [n][3] = (nx, ny, nz); cell of a point = floor((p - origin) * (1/cell)) in fp32.
 * covered [sum G] bytes = scratch (zeroed here, 1 where the ground-truth point is covered); counts (may be NULL) = covered
 * points per scene.  A scene with an empty cloud reports 0 (long_term_utils.py:461-462). */
int nbp_coverage_percentage(const float* cloud, int64_t cloud_stride, const int32_t* cloud_len, const int64_t* sample_idx,
                            int64_t sample_stride, const float* gt_sorted, const int64_t* gt_off, const int32_t* cell_start,
                            const int64_t* cell_off, const float* origin, const int32_t* dims, int n_scenes, int64_t max_samples,
                            int64_t total_gt, float cell, float threshold, int weight, uint64_t seed, uint8_t* covered,
                            float* coverage, int32_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------ SURVEY section 8(f) row 3
 * Batched Trimesh-free collision tests of the planner: line_segment_mesh_intersection (macarons/utility/macarons_utils.py:120-151;
 * callers next_best_path/utility/long_term_utils.py:347, next_best_path/testers/nbp_planning.py:142,245) and the axis rays of
 * check_camera_in_mesh (long_term_utils.py:158-170).  Meshes packed as for nbp_raster_depth_batched (scene-local face indices).
 * segments [n][6] fp32: rays == 0: (start, end) -> hit[i] = 1 iff some triangle is crossed at a distance < |end - start| from the
 * start; rays != 0: (origin, unit direction) -> hit / count over the whole forward ray.  count (may be NULL) = number of triangles
 * hit (the inside test takes its parity); hit may be NULL when only counts are wanted.  fp64 plane + barycentric formulation
 * (tolerances as Trimesh's ray_triangle: 1e-13 on the barycentrics, -1e-6 on the forward distance). */
int nbp_segments_hit_mesh(const float* verts, const int32_t* faces, const int64_t* vert_offsets, const int64_t* face_offsets,
                          int n_scenes, const float* segments, const int32_t* seg_scene, int n_segments, int rays,
                          uint8_t* hit, int32_t* count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NBP_B200_H_ */
