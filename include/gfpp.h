/*
 * gfpp.h -- C-ABI of libgfpp.so: the B200 (sm_100a) render hot path of GeneFace++.
 *
 * Drop-in boundary.  Plain pointers and sizes only; no torch types.
 * All pointers are DEVICE pointers unless the name ends in _host.  Every call is
 * stream-ordered on `stream` (a cudaStream_t passed as void*), never synchronises the
 * device, never allocates device memory (callers provide packed-model and workspace
 * buffers whose sizes they query first) and never throws: the return value is 0 on success
 * or a negative gfpp_status; gfpp_last_error() returns a thread-local message.
 *
 * Two levels are exported.
 *
 *  (A) per-op entry points that mirror, argument for argument, the pybind functions of the
 *      reference's four CUDA extensions, so the reference's own Python wrappers can bind them
 *      (INTEGRATION.md shows the stub):
 *        gfpp_near_far_from_aabb   <- modules/radnerfs/raymarching/src/raymarching.h:7  (raymarching.cu:148-156)
 *        gfpp_march_rays           <- raymarching.h:19 (raymarching.cu:932-939)
 *        gfpp_composite_rays       <- raymarching.h:20 (raymarching.cu:1032-1037)
 *        gfpp_grid_encode_forward  <- modules/radnerfs/encoders/gridencoder/src/gridencoder.h:12 (gridencoder.cu:446-471)
 *        gfpp_sh_encode_forward    <- modules/radnerfs/encoders/shencoder/src/shencoder.h:9 (shencoder.cu:400)
 *        gfpp_freq_encode_forward  <- modules/radnerfs/encoders/freqencoder/src/freqencoder.h:7 (freqencoder.cu:97-110)
 *
 *  (B) the fused frame renderer that replaces the whole host-driven loop of
 *      NeRFRenderer.render / RADNeRFTorso.render (modules/radnerfs/renderer.py:340-399,
 *      modules/radnerfs/radnerf_torso.py:129-197) for one or many frames per call:
 *        gfpp_model_packed_bytes / gfpp_model_pack / gfpp_render_workspace_bytes / gfpp_render_frames
 *
 *  (C) the SR checkpoints' extra stages: gfpp_sr_* (256 -> 512 super-resolution head) and
 *      gfpp_torso_sr_* (the torso field of radnerf_torso_sr.py), declared at the end of this file.
 *
 *  (D) the training-side native ops, mirroring the reference's pybind functions like (A).
 */
#ifndef GFPP_H_
#define GFPP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFPP_API __attribute__((visibility("default")))

typedef enum {
    GFPP_OK = 0,
    GFPP_ERR_INVALID = -1,     /* bad argument (null pointer, unsupported C/D/degree, size mismatch) */
    GFPP_ERR_CUDA = -2,        /* a CUDA runtime call failed; see gfpp_last_error() */
    GFPP_ERR_WORKSPACE = -3,   /* caller-provided buffer too small */
    GFPP_ERR_UNSUPPORTED = -4  /* configuration outside what the kernels were built for */
} gfpp_status;

GFPP_API const char *gfpp_last_error(void);
GFPP_API int gfpp_version(void);
/* device check: returns GFPP_OK iff the current device is compute capability 10.x */
GFPP_API int gfpp_check_device(void);

/* ------------------------------------------------------------------ (A) per-op mirror */

/* nears/fars [N] <- slab test of rays [N,3] against aabb[6]; miss => both FLT_MAX; near >= min_near */
GFPP_API int gfpp_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N,
                                     float min_near, float *nears, float *fars, void *stream);

/* one round of inference marching: for each of n_alive rays emit up to n_step occupied samples.
 * xyzs/dirs [M,3] and deltas [M,2] must be zero-initialised by the caller (as the reference wrapper does). */
GFPP_API int gfpp_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                             const float *rays_o, const float *rays_d, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid, const float *nears,
                             const float *fars, float *xyzs, float *dirs, float *deltas, const float *noises,
                             void *stream);

/* front-to-back compositing of one round, in place; kills rays by writing -1 into rays_alive */
GFPP_API int gfpp_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive,
                                 float *rays_t, const float *sigmas, const float *rgbs, const float *deltas,
                                 float *weights_sum, float *depth, float *image, void *stream);

/* multi-resolution grid encode, forward only.  inputs [B,D] in [0,1]; embeddings [sum,C] fp32;
 * offsets_host [L+1] (HOST pointer: the level layout is tiny and static); outputs [L,B,C] (L-major,
 * exactly what the reference kernel writes).  D in {2,3}, C == 2 (the only shapes on the path). */
GFPP_API int gfpp_grid_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets_host,
                                      float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                      uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp, void *stream);

/* real spherical harmonics, degree (called C in the reference) 1..4: outputs [B, degree^2] */
GFPP_API int gfpp_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t degree,
                                    void *stream);

/* [x, sin(2^f x), sin(2^f x + pi/2)]_f : outputs [B, C], C = D + 2*D*deg */
GFPP_API int gfpp_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                      float *outputs, void *stream);

/* ------------------------------------------------------------------ (B) fused renderer */

/* One multi-resolution grid in the reference layout (GridEncoder, grid.py:98-136). */
typedef struct {
    const float *embeddings;      /* device [n_entries, 2] fp32 */
    const int32_t *offsets_host;  /* host   [num_levels+1] */
    uint32_t input_dim;           /* 2 or 3 */
    uint32_t num_levels;          /* <= 16 */
    uint32_t base_resolution;     /* H (16) */
    float log2_per_level_scale;   /* S, as float32(log2(per_level_scale)) */
    uint32_t gridtype;            /* 0 hash, 1 tiled */
    uint32_t interp;              /* 0 linear, 1 smoothstep */
    int align_corners;
} gfpp_grid_desc;

/* Weights in the reference's nn.Linear layout [out, in], fp32, device pointers (state_dict tensors). */
typedef struct {
    /* head field (modules/radnerfs/radnerf.py:56-86) */
    gfpp_grid_desc position_grid;       /* 3-D */
    gfpp_grid_desc ambient_grid;        /* ambient_coord_dim-D (3 for the May configs) */
    const float *ambient_w[3];          /* (128,96) (128,128) (3,128) */
    const float *sigma_w[3];            /* (128,64) (128,128) (129,128) */
    const float *color_w[2];            /* (128,148) (3,128) */
    const float *individual_code;       /* [ind_dim] = individual_embeddings[0]; may be NULL if ind_dim == 0 */
    uint32_t cond_dim;                  /* 64 */
    uint32_t ind_dim;                   /* 4 */
    /* occupancy + marching (modules/radnerfs/renderer.py:67-102) */
    const uint8_t *density_bitfield;    /* [cascade * grid_size^3 / 8], Morton order */
    float aabb[6];
    float bound;
    float min_near;
    uint32_t cascade;
    uint32_t grid_size;
    float density_scale;
    /* torso field (modules/radnerfs/radnerf_torso.py:17-49); has_torso == 0 for head-only RADNeRF */
    int has_torso;
    gfpp_grid_desc torso_grid;          /* 2-D */
    const float *torso_deform_w[3];     /* (64,104) (64,64) (2,64) */
    const float *torso_canon_w[3];      /* (32,136) (32,32) (4,32) */
    const float *torso_code;            /* [8] = torso_individual_codes[0] */
    const float *density_grid_torso;    /* [grid_size^2] */
    float density_thresh_torso;         /* min(density_thresh_torso, mean_density_torso) as the reference evaluates it */
    float torso_shrink;
    uint32_t torso_code_dim;            /* 8 */
    /* arithmetic of the head MLP GEMMs: 0 = fp32 FFMA (CUDA cores), 1 = fp16 operands on tcgen05 tensor cores with
     * fp32 accumulation (what the reference runs under autocast), 2 = bf16 hi/lo split x3 on tcgen05 (~fp32 accuracy),
     * 3 = bf16 x1, 4 = "robust": fp16 on tcgen05 with the hi/lo split (3 MMAs per k-step, ~22 mantissa bits) on the ambient
     * net only and a 16-bit fixed-point position table -- the two roundings the field amplifies; holds 1e-3 max-abs on
     * well-conditioned (trained-like) scenes where plain fp16 does not.  Modes 1 and 4 run the row-owner kernel
     * (head_v2_kernel.cu).  Everything else (marching, compositing, torso) is fp32 in every mode. */
    uint32_t mlp_precision;
} gfpp_model_desc;

/* Host-side model handle: plain data, caller-allocated (stack, heap, numpy buffer ...), filled by
 * gfpp_model_pack().  It holds level tables, launch constants and device pointers into `packed` and into
 * the borrowed tables; copying it is fine, it owns nothing. */
typedef struct gfpp_model {
    uint64_t opaque[512];
} gfpp_model;

/* bytes of the device buffer gfpp_model_pack() fills (repacked/transposed weights, occupancy bounds) */
GFPP_API size_t gfpp_model_packed_bytes(const gfpp_model_desc *desc);
/* Repack on `stream` into `packed` (device, >= gfpp_model_packed_bytes) and fill *model.  Weights are copied
 * (k-major, layer chunks in streaming order); the desc's grid tables, bitfield and density_grid_torso are
 * BORROWED and must stay alive while the model is used.  Unsupported layer shapes => GFPP_ERR_UNSUPPORTED
 * (the kernels are built for the May architecture: hidden 128, cond 64, geo 128, 16 levels x 2 features). */
GFPP_API int gfpp_model_pack(const gfpp_model_desc *desc, void *packed, size_t packed_bytes, gfpp_model *model,
                             void *stream);

typedef struct {
    uint32_t n_frames;            /* F */
    uint32_t n_rays;              /* N = H*W rays per frame */
    /* rays: either supplied (parity on identical rays) ... */
    const float *rays_o;          /* [F,N,3] or NULL */
    const float *rays_d;          /* [F,N,3] or NULL */
    /* ... or generated in-kernel from the camera (get_rays, modules/radnerfs/utils.py:283-364) */
    const float *poses_c2w;       /* [F,4,4] row-major; used when rays_o == NULL */
    float fx, fy, cx, cy;
    uint32_t img_h, img_w;
    const float *cond_feat;       /* [F,cond_dim] output of cal_cond_feat (radnerf.py:88-106) */
    const float *torso_pose6;     /* [F,6] euler+trans from convert_poses; NULL if no torso */
    const float *bg_coords;       /* [N,2] (shared by all frames); NULL if no torso */
    const float *bg_color;        /* [N,3] shared by all frames, or NULL => 1.0 (renderer.py:387-388) */
    float dt_gamma;
    uint32_t max_steps;
    float T_thresh;
} gfpp_frames;

typedef struct {
    float *rgb_map;               /* [F,N,3] final image (renderer.py:390-397 / radnerf_torso.py:191-197) */
    float *depth_map;             /* [F,N] */
    float *weights_sum;           /* [F,N] head alpha */
    float *torso_alpha_map;       /* [F,N]   or NULL */
    float *torso_rgb_map;         /* [F,N,3] or NULL (bg mixed with torso, radnerf_torso.py:187-189) */
    float *torso_deform;          /* [F,N,2] deformation for masked pixels, 0 elsewhere; or NULL */
    int32_t *stats;               /* [F,4] = {B_total, n_survivors, S_valid_samples, P_torso_pixels} or NULL */
    uint8_t *rgb_u8;              /* [F,N,3] the frame as the video writer wants it, (uint8)(int)(rgb * 255) of the clamped colour
                                   * (inference/genefacepp_infer.py:469,505), written by the epilogue kernel itself; or NULL.
                                   * rgb_map may be NULL when rgb_u8 is given (then no fp32 frame is written at all) */
} gfpp_outputs;

GFPP_API size_t gfpp_render_workspace_bytes(uint32_t n_frames, uint32_t n_rays, uint32_t max_steps);
/* Enqueue the whole clip (or one frame): pass-1 persistent head kernel, on-device round-schedule
 * replay, pass-2 for rays that outlive max_steps, torso + composite epilogue.  No host sync. */
GFPP_API int gfpp_render_frames(const gfpp_model *model, const gfpp_frames *frames, const gfpp_outputs *out,
                                void *workspace, size_t workspace_bytes, void *stream);
/* Diagnostics for bench.py's roofline leg: when enabled, gfpp_render_frames records CUDA events (on the launching
 * stream) around its three main kernels; gfpp_profile_read() waits for those events and returns the durations of the
 * most recent call in milliseconds: ms[0] = head pass 1, ms[1] = schedule + head pass 2, ms[2] = torso/composite epilogue. */
GFPP_API int gfpp_profile_enable(int on);
GFPP_API int gfpp_profile_read(float ms[4]);   /* ms[3] = everything before the head kernel (memset, torso biases, ray setup) */
/* optional: device buffer of 32 uint64 that the tensor-core head kernel fills with per-phase cycle totals (thread 0 of each
 * CTA; slot 31 = batches processed); NULL switches it off.  Diagnostics only. */
GFPP_API int gfpp_profile_phases(void *dev_u64x32);
/* Self-test of the tcgen05 plumbing (tile layouts, descriptors, MMA issue, TMEM read-back) used by the tensor-core MLP:
 * out[128,N] = A[128,K] @ W[N,K]^T with 16-bit operands, fp32 accumulation.  K = 64*j (+16 if k16_tail), N % 16 == 0,
 * N <= 144, K <= 144.  precision: 1 = fp16, 2 = bf16 hi/lo split (3 MMAs), 3 = bf16.  scratch: >= 6*18432 bytes. */
GFPP_API int gfpp_tc_selftest(const float *A, const float *W, uint32_t N, uint32_t K, int k16_tail, int precision,
                              void *scratch, float *out, void *stream);
/* Debug / test hook: write out the rays the fused path generates IN-KERNEL from poses + intrinsics (pixel centres, normalised,
 * rotated by c2w[:3,:3]: modules/radnerfs/utils.py:302-360), [F, img_h*img_w, 3] each, so that tests can measure their distance to
 * torch get_rays and hand exactly these rays to the oracle. */
GFPP_API int gfpp_debug_generate_rays(const float *poses_c2w, uint32_t n_frames, float fx, float fy, float cx, float cy,
                                      uint32_t img_h, uint32_t img_w, float *rays_o, float *rays_d, void *stream);
/* number of kernels the last gfpp_render_frames call on this thread launched */
GFPP_API int gfpp_last_launch_count(void);

/* ------------------------------------------------------------------ (C) super-resolution head of the SR checkpoints
 *
 * Replaces `Superresolution.forward` (modules/radnerfs/radnerf_sr.py:15-48; StyleGAN2 blocks of
 * modules/eg3ds/models/superresolution.py:159-257 and networks_stylegan2.py:286-475) for the 256 -> 512 head the
 * `with_sr` checkpoints carry (inference/genefacepp_infer.py:464-465,480-481 takes `sr_rgb_map`).  The network is
 * always run with the constant latent w = 1 (radnerf_sr.py:33-34), so the style modulation, the demodulation and the
 * [1,3,3,1] resampling filter are constants of the checkpoint: the caller folds them into plain GEMM matrices (fp32,
 * device) once -- genefaceplusplus_b200/superres.py::folded_weights documents the algebra -- and hands them over here:
 *   conv_in_w  block0.conv0  [27][128]     k-major, k = (ky*3+kx)*3 + ci
 *   conv0_w    block0.conv1  [128][1152]   k = (ky*3+kx)*128 + ci
 *   up_w       block1.conv0  [256][1152]   transposed stride-2 conv + FIR merged into four 3x3 phase kernels on the input
 *                                          grid, row = (py*2+px)*64 + co for output pixel (2y+py, 2x+px)
 *   conv1_w    block1.conv1  [64][576]     k = (ky*3+kx)*64 + ci
 *   bias[4]    conv biases (128, 128, 64, 64);  rgb_w[2] effective toRGB weights [3][128], [3][64];  rgb_b[2] [3]
 * Operands of the three tensor-core layers are rounded to fp16 (fp32 accumulation), as the reference's use_fp16 blocks do. */
typedef struct {
    const float *conv_in_w;
    const float *conv0_w;
    const float *up_w;
    const float *conv1_w;
    const float *bias[4];
    const float *rgb_w[2];
    const float *rgb_b[2];
} gfpp_sr_desc;

typedef struct gfpp_sr_model {
    uint64_t opaque[32];
} gfpp_sr_model;

GFPP_API size_t gfpp_sr_packed_bytes(void);
/* Repack on `stream` into `packed` (device, 1024-byte aligned, >= gfpp_sr_packed_bytes); the desc's tensors are copied and
 * may be freed once the stream has passed this call. */
GFPP_API int gfpp_sr_pack(const gfpp_sr_desc *desc, void *packed, size_t packed_bytes, gfpp_sr_model *model, void *stream);
GFPP_API size_t gfpp_sr_workspace_bytes(uint32_t n_frames, uint32_t in_res);
/* rgb_in [F, in_res*in_res, 3] fp32 in [0,1] (the layout gfpp_render_frames writes `rgb_map` in) -> out [F, 3, 2*in_res,
 * 2*in_res] fp32 (planar, as the reference returns it); clamp01 != 0 clamps it to [0,1] (every caller of the SR head does
 * that next: radnerf_sr.py:208).  noise[i]: per-layer noise planes ALREADY multiplied by the layer's noise_strength
 * (layers 0,1: in_res^2; layers 2,3: (2*in_res)^2), one plane shared by all frames (noise_mode 'const') or, with
 * noise_per_frame != 0, F consecutive planes ('random'); NULL entries / a NULL array mean no noise ('none').
 * in_res must be a multiple of 128 (256 in the reference).  Four kernel launches, no host sync, no allocation. */
GFPP_API int gfpp_sr_forward(const gfpp_sr_model *model, uint32_t n_frames, uint32_t in_res, const float *rgb_in,
                             const float *const noise[4], uint32_t noise_per_frame, float *out, int clamp01, void *workspace,
                             size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ (C) torso field of the torso-SR checkpoints
 *
 * Replaces RADNeRFTorsowithSR.forward_torso + the composite of its render() (modules/radnerfs/radnerf_torso_sr.py:75-113,
 * 196-228): mask by grid_sample of density_grid_torso, the landmark-conditioned (and, `torso_head_aware`, head-colour-aware)
 * deformation + canonical torso field on the masked pixels, torso over background, head over both, clamp.  The head NeRF of
 * these checkpoints is the plain head field: render it with gfpp_render_frames (has_torso = 0, bg_color = zeros, so that
 * `rgb_map` is the premultiplied head colour and `weights_sum` its alpha) and hand both over here; the result feeds
 * gfpp_sr_forward.  Weights in the reference's nn.Linear layout [out, in], fp32 device pointers (state_dict tensors); the
 * grid table, density_grid_torso, the first deform / canonical layers and the code are BORROWED (per-frame folds read them). */
typedef struct {
    gfpp_grid_desc torso_grid;          /* 2-D tiled grid, 16 levels x 2 */
    const float *torso_deform_w[3];     /* (64,din) (64,64) (2,64); din = 42 + torso_code_dim + 126 (+ 16 if head_aware) */
    const float *torso_canon_w[3];      /* (32,32+din) (32,32) (4,32) */
    const float *torso_code;            /* [torso_code_dim] = torso_individual_codes[0] */
    uint32_t torso_code_dim;            /* 8 */
    int head_aware;                     /* hparams['torso_head_aware'] */
    const float *ha_w[3];               /* head_color_weights_encoder: (16,4) (32,16) (16,32) */
    const float *ha_b[3];               /* its biases: 16, 32, 16 */
    const float *density_grid_torso;    /* [grid_size^2] */
    uint32_t grid_size;
    float density_thresh_torso;         /* min(density_thresh_torso, mean_density_torso) as the reference evaluates it */
    float torso_shrink;
} gfpp_torso_sr_desc;

typedef struct gfpp_torso_sr_model {
    uint64_t opaque[128];
} gfpp_torso_sr_model;

typedef struct {
    uint32_t n_frames;            /* F */
    uint32_t n_rays;              /* N pixels per frame (256*256 in the reference) */
    const float *image;           /* [F,N,3] premultiplied head colour */
    const float *weights_sum;     /* [F,N]   head alpha */
    const float *lm68;            /* [F,136] 68 2-D landmarks per frame (points 5..11 = the jaw are used) */
    const float *bg_coords;       /* [N,2]   shared by all frames */
    const float *bg_color;        /* [N,3]   shared by all frames, or NULL => 1.0 */
} gfpp_torso_sr_frames;

GFPP_API size_t gfpp_torso_sr_packed_bytes(void);
GFPP_API int gfpp_torso_sr_pack(const gfpp_torso_sr_desc *desc, void *packed, size_t packed_bytes, gfpp_torso_sr_model *model,
                                void *stream);
GFPP_API size_t gfpp_torso_sr_workspace_bytes(uint32_t n_frames);
/* rgb_map [F,N,3] (required): the clamped composite; torso_alpha_map [F,N], torso_rgb_map [F,N,3] (torso over background),
 * torso_deform [F,N,2] (0 outside the mask), torso_pixels [F] (masked pixels per frame): optional, NULL to skip.
 * Two kernel launches, no host sync, no allocation. */
GFPP_API int gfpp_torso_sr_composite(const gfpp_torso_sr_model *model, const gfpp_torso_sr_frames *frames, float *rgb_map,
                                     float *torso_alpha_map, float *torso_rgb_map, float *torso_deform, int32_t *torso_pixels,
                                     void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ (D) training-side ops
 *
 * Mirror, argument for argument (+ stream; grid offsets as a host pointer, as in (A)), the training-side pybind functions of
 * the reference's extensions, so that its own autograd wrappers (raymarching.py:184-344, grid.py:24-88) can bind them:
 *   gfpp_march_rays_train                <- raymarching.h:14 (raymarching.cu:352-533)
 *   gfpp_march_rays_train_backward       <- raymarching.h:15 (raymarching.cu:535-598)
 *   gfpp_composite_rays_train_forward    <- raymarching.h:16 (raymarching.cu:603-700)
 *   gfpp_composite_rays_train_backward   <- raymarching.h:17 (raymarching.cu:711-822)
 *   gfpp_grid_encode_forward_dydx        <- gridencoder.h:12 with dy_dx given (gridencoder.cu:87-243)
 *   gfpp_grid_encode_backward            <- gridencoder.h:13 (gridencoder.cu:246-368, 413-443)
 *   gfpp_grad_total_variation            <- gridencoder.h:15 (gridencoder.cu:505-609)
 *   gfpp_packbits / gfpp_morton3D / gfpp_morton3D_invert / gfpp_morton3D_dilation / gfpp_sph_from_ray
 *                                        <- raymarching.h:8-12 (raymarching.cu:162-342; update_extra_state)
 * fp32 only.  One deliberate difference: gfpp_march_rays_train is deterministic -- point offsets are an exclusive prefix sum of
 * the per-ray sample counts in ray order (rays[n] = (n, offset, num_steps)) instead of the reference's atomicAdd arrival
 * order; every consumer addresses samples through `rays`, so this is one of the layouts the reference itself can produce.
 * It needs a small scratch buffer for the device-wide scan (query the size first).  `counter` keeps the reference's
 * meaning: counter[0] += points, counter[1] += N. */
GFPP_API size_t gfpp_march_rays_train_scratch_bytes(uint32_t N);
GFPP_API int gfpp_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                                   uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                                   const float *fars, float *xyzs, float *dirs, float *deltas, int32_t *rays, int32_t *counter,
                                   const float *noises, void *scratch, size_t scratch_bytes, void *stream);
/* grad_rays_o / grad_rays_d [N,3] are accumulated into (+=), as in the reference */
GFPP_API int gfpp_march_rays_train_backward(const float *grad_xyzs, const float *grad_dirs, const int32_t *rays, const float *deltas,
                                            uint32_t N, uint32_t M, float *grad_rays_o, float *grad_rays_d, void *stream);
GFPP_API int gfpp_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *ambient, const float *deltas,
                                               const int32_t *rays, uint32_t M, uint32_t N, float T_thresh, float *weights_sum,
                                               float *ambient_sum, float *depth, float *image, void *stream);
/* grad_sigmas [M], grad_rgbs [M,3], grad_ambient [M] must be zero-initialised by the caller (raymarching.py:316-318) */
GFPP_API int gfpp_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_ambient_sum, const float *grad_image,
                                                const float *sigmas, const float *rgbs, const float *ambient, const float *deltas,
                                                const int32_t *rays, const float *weights_sum, const float *ambient_sum,
                                                const float *image, uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas,
                                                float *grad_rgbs, float *grad_ambient, void *stream);
/* outputs [L,B,C] as gfpp_grid_encode_forward, plus dy_dx [B, L, D, C] */
GFPP_API int gfpp_grid_encode_forward_dydx(const float *inputs, const float *embeddings, const int32_t *offsets_host, float *outputs,
                                           uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, float *dy_dx,
                                           uint32_t gridtype, int align_corners, uint32_t interp, void *stream);
/* grad [L,B,C]; grad_embeddings (zero-initialised by the caller, grid.py:76) is accumulated into with vector reductions;
 * dy_dx / grad_inputs [B,D]: both or neither */
GFPP_API int gfpp_grid_encode_backward(const float *grad, const float *inputs, const float *embeddings, const int32_t *offsets_host,
                                       float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                       const float *dy_dx, float *grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                                       void *stream);
GFPP_API int gfpp_grad_total_variation(const float *inputs, const float *embeddings, float *grad, const int32_t *offsets_host,
                                       float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                       uint32_t gridtype, int align_corners, void *stream);
GFPP_API int gfpp_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream);
GFPP_API int gfpp_morton3D(const int32_t *coords, uint32_t N, int32_t *indices, void *stream);
GFPP_API int gfpp_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords, void *stream);
GFPP_API int gfpp_morton3D_dilation(const float *grid, uint32_t C, uint32_t H, float *grid_dilation, void *stream);
GFPP_API int gfpp_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GFPP_H_ */
