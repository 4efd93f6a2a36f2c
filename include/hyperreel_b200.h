/*
 * hyperreel_b200 -- C-ABI of the B200-native HyperReel per-ray rendering hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference has no
 * FFI (it is pure PyTorch); every entry point below names the reference interface it replaces
 * (file:line under the reference checkout).  The Python binding a maintainer would add is the ctypes
 * stub shown in INTEGRATION.md (hyperreel_b200/lib.py is that stub, complete).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; hr_last_error() gives the message
 *     (the reference raises Python exceptions / asserts: nlf/nets/tensorf_dynamic.py:744-745;
 *     the Python shim turns a non-zero status into RuntimeError);
 *   - all work is enqueued on the caller's cudaStream_t (passed as void*): the reference runs on
 *     torch's current stream (nlf/__init__.py:486-502); no hidden synchronisation except in
 *     hr_render_host, which is synchronous by contract;
 *   - a handle is re-entrant per handle, no global state except the thread-local error string;
 *   - there is NO CPU fallback: if no sm_100 device is present hr_create fails.
 */
#ifndef HYPERREEL_B200_H
#define HYPERREEL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HR_ABI_VERSION 10

#define HR_MAX_GROUPS 4   /* ray-parameterisation groups feeding the sample net (ray.py:235-263) */
#define HR_MAX_LAYERS 10  /* Linear layers of the sample net (mlp.py:127-154) */
#define HR_MAX_SAMPLES 256 /* z_channels S (per-ray sample primitives) */
#define HR_MAX_PEERS 8    /* destination buffers of hr_render_scatter (GPUs of one NVSwitch domain) */

/* Activation y = f(x*inner_fac + shift) * outer_fac  (nlf/activations.py:53-69,121-137,163-178).
 * EaseValue (activations.py:462-496) is resolved on the host at render iteration to its inner act. */
enum { HR_ACT_IDENTITY = 0, HR_ACT_SIGMOID = 1, HR_ACT_TANH = 2 };
typedef struct hr_act {
  int32_t kind;
  float inner_fac, shift, outer_fac;
} hr_act;

/* One `params:` group of RayPredictionEmbedding (nlf/embedding/ray.py:235-263,320-326):
 * rays[:, start:end] -> RayParam fn -> WindowedPE (all windows open). */
enum { HR_PARAM_IDENTITY = 0, HR_PARAM_TWO_PLANE = 1, HR_PARAM_PLUECKER = 2 };
typedef struct hr_encode_group {
  int32_t start, end;        /* channel slice of the ray                                    */
  int32_t fn;                /* HR_PARAM_* (nlf/param.py:20-24, 63-118, 223-256)             */
  int32_t n_freqs;           /* WindowedPE bands 2^1..2^n (nlf/pe.py:148,210-221)            */
  int32_t exclude_identity;  /* pe.py:160-168                                               */
  float freq_mult;           /* freq_multiplier (pe.py:147)                                 */
  float base_mult;           /* base_multiplier (pe.py:151)                                 */
  float near, far;           /* two_plane plane offsets (param.py:78-79)                    */
  float dir_mult, mom_mult;  /* pluecker multipliers (param.py:236-237)                     */
} hr_encode_group;

enum { HR_ISECT_Z_PLANE = 0, HR_ISECT_SPHERE = 1, HR_ISECT_CYLINDER = 2, HR_ISECT_SPHERE_NEW = 3,
       HR_ISECT_DISTANCE = 4,  /* z.py:16-97, primitive.py:366-438, :181-250, :440-546, :126-180 (euclidean_distance_unified) */
       HR_ISECT_VOXEL = 5,     /* voxel_grid: axis-aligned planes, sample s = plane s/3 of axis s%3 (voxel.py:19-112)        */
       HR_ISECT_PLANE = 6 };   /* deformable_voxel_grid: predicted plane normals + offsets (voxel.py:115-214)                */
enum { HR_CONTRACT_NONE = 0, HR_CONTRACT_MIPNERF = 1, HR_CONTRACT_AFFINE = 2 };  /* AFFINE: bbox / z_depth (contract.py:65-110) */
enum { HR_SHADE_SH = 0, HR_SHADE_RGB = 1 };
enum { HR_DENSE_RELU = 0, HR_DENSE_SOFTPLUS = 1, HR_DENSE_RELU_ABS = 2 };

/* Sample-net arithmetic.  FP32_SIMT: fp32 FMA on CUDA cores (bit-level closest to the reference's
 * cuBLAS SGEMM).  BF16X3_TC: tcgen05 tensor cores, every fp32 operand split into bf16 hi+lo and the
 * three leading cross products accumulated in fp32 TMEM (error ~2^-16 per product). */
enum { HR_MLP_FP32_SIMT = 0, HR_MLP_BF16X3_TC = 1,
       HR_MLP_ZERO = 2 /* `net: {type: zero}` (ZeroMLP, nlf/nets/mlp.py:14-33): every head is 0, no network runs */ };

/* The recognised pipeline signature; one struct describes what the
 * reference assembles from conf/experiment/model/<name>.yaml.  Anything the YAML asks for that this
 * struct cannot express is rejected by the host binding at construction (no fallback). */
typedef struct hr_config {
  int32_t abi_version;  /* = HR_ABI_VERSION */
  int32_t c_in;         /* ray channels: 6 static, 8 video (datasets/technicolor.py:360-396) */

  /* --- sample-prediction net input (a5,a6) --- */
  int32_t n_groups;
  hr_encode_group groups[HR_MAX_GROUPS];

  /* --- sample-prediction net (a7): nlf/nets/mlp.py:60-178 --- */
  int32_t mlp_in;      /* sum of PE output channels                                   */
  int32_t mlp_width;   /* hidden_channels W                                           */
  int32_t mlp_layers;  /* number of Linear layers = yaml depth (ray.py:283-285)        */
  int32_t mlp_skip;    /* index of the layer whose input is cat([in, h]) or -1         */
  int32_t mlp_out;     /* S * head_stride                                             */
  float leaky_slope;   /* 0.01 (activations.py:14-29)                                 */
  int32_t mlp_mode;    /* HR_MLP_*                                                    */

  /* --- per-sample heads (a8): ray.py:331-337; channel offsets inside one sample, -1 = absent --- */
  int32_t n_samples;    /* S                                                          */
  int32_t head_stride;  /* channels per sample (15 for the shipped targets)            */
  int32_t off_z, n_z;   /* z_vals: 1 channel (z_plane) or 4 (sphere: origin3 + radius) */
  int32_t off_flow, off_sigma, off_point_sigma, off_offset, off_cscale, off_cshift;
  hr_act act_z, act_flow, act_sigma, act_point_sigma, act_offset, act_cscale, act_cshift;

  /* --- intersection (a10-a13): nlf/intersect/base.py:142-259 --- */
  int32_t isect_type;             /* HR_ISECT_*                                        */
  hr_act isect_act;               /* `activation:` of the intersect block (base.py:119) */
  int32_t isect_use_sigma;        /* base.py:155-161                                   */
  int32_t isect_density_off;      /* head channel used as sigma there, -1 = zeros      */
  float z_scale;                  /* |samples[1]-samples[0]| (z.py:60-71)              */
  float isect_near, isect_far;    /* mask bounds (base.py:194-203)                     */
  int32_t isect_sort;             /* base.py:206-210                                   */
  float samples[HR_MAX_SAMPLES];  /* base primitives, host-computed linspace (z.py:50-57) */
  int32_t contract_type;          /* HR_CONTRACT_* (nlf/contract.py:113-192)           */
  int32_t contract_samples;       /* inverse-contract z before intersecting (base.py:132-133) */
  float contract_start_radius, contract_end_radius;      /* end may be +inf          */
  float contract_start_distance, contract_end_distance;
  float sphere_origin_initial[3]; /* primitive.py:388                                  */
  float sphere_origin_scale;      /* origin_scale_factor (primitive.py:387)            */

  /* --- keyframe snap + flow (a14): nlf/embedding/point.py:780-831, utils/flow_utils.py:10-35 --- */
  int32_t use_flow;
  int32_t num_keyframes, num_frames;  /* K, F                                          */
  hr_act flow_act;                    /* spatial_flow_activation (point.py:817)         */

  /* --- point offset (a15): point.py:371-396 --- */
  int32_t use_offset;
  int32_t offset_density_off;  /* head channel of `in_density_field`, -1 = zeros       */
  hr_act offset_act;

  /* --- TensoRF decode + composite (a17-a23) --- */
  int32_t dynamic;      /* 1: tensor_vm_split_time, 0: tensor_vm_split_no_sample         */
  float aabb[6];        /* min xyz, max xyz (tensorf_base.py:148)                        */
  float distance_scale; /* tensorf_base.py:212                                          */
  int32_t n_sigma[3];   /* n_lamb_sigma                                                 */
  int32_t n_app[3];     /* n_lamb_sh                                                    */
  int32_t app_dim;      /* data_dim_color: 27 (SH) or 3 (RGB)                           */
  int32_t shading;      /* HR_SHADE_*                                                   */
  int32_t white_bg, black_bg;
  float weight_thre;    /* rm_weight_mask_thre                                          */
  int32_t fea2dense;    /* HR_DENSE_*                                                   */
  float density_shift;
  int32_t use_color_scale_shift; /* 'color_scale' reaches the colour net (tensorf_dynamic.py:780-784) */
  int32_t clamp_output; /* eval(): clamp(0,1) (tensorf_dynamic.py:805-806)              */

  /* --- ABI 5 / 6 additions --- */
  /* HR_CONTRACT_AFFINE: c(p) = (p - min) / den per axis (BBoxContract :83-84: den = max - min; ZDepthContract :109-110:
   * min = 0, den = fac), distances are divided / multiplied by dist_fac (:77-81, :103-107).                             */
  float contract_affine_min[3], contract_affine_den[3];
  float contract_dist_fac;
  /* per-ray colour heads applied after compositing: rgb_map * (1 + scale[:,0]) + shift[:,0]
   * (scale_shift_color_one, utils/tensorf_utils.py:275-281; tensorf_dynamic.py:798-800); -1 = absent                 */
  int32_t off_cscale_global, off_cshift_global;
  hr_act act_cscale_global, act_cshift_global;
  /* HR_ISECT_SPHERE_NEW (8 z channels: origin 3, resize 3, offset 1, radius 1; primitive.py:489-546):
   * origin = z[0:3] * sphere_origin_scale, resize = z[3:6] * sphere_resize_scale + sphere_resize_initial           */
  float sphere_resize_scale;
  float sphere_resize_initial[3];

  /* --- ABI 10 additions --- */
  /* HR_ISECT_VOXEL / HR_ISECT_PLANE: `samples` holds one linspace per axis interleaved (sample s = plane s / isect_axes of
   * axis s % isect_axes); z_scale3 = spacing per axis (voxel.py:58-63; HR_ISECT_PLANE and the other primitives: z_scale x3) */
  int32_t isect_axes;       /* 3 (voxel_grid), number of start normals (deformable_voxel_grid), 1 otherwise              */
  float z_scale3[3];
  int32_t isect_outward;    /* voxel_grid outward_facing: z *= sign(d_axis) (voxel.py:80-83)                            */
  int32_t isect_max_axis;   /* voxel_grid max_axis: drop the planes of the non-dominant axes (voxel.py:100-110)          */
  float plane_normal[9];    /* deformable_voxel_grid start_normal rows (voxel.py:120-128)                                */
  float plane_normal_scale; /* normal_scale_factor (voxel.py:129)                                                        */
  /* ColorTransformEmbedding (point.py:558-612) + transform_color_one (utils/tensorf_utils.py:308-331): the composited pixel
   * becomes rgb + M rgb + shift with (M, shift) = row round(rays[:, -2]) of hr_params.color_embedding [n_color_views, 12];
   * 0 = off (no such embedding, or the dataset does not validate on every camera) */
  int32_t n_color_views;
  hr_act act_ctransform, act_ctshift;

  /* --- cascaded pipelines (PointPredictionEmbedding, nlf/embedding/point.py:39-219) ---
   * cascade == 1:  ray_prediction -> ray_intersect (pre_samples z-planes) -> point_prediction -> ray_intersect -> ...
   * The fields above then describe the SECOND stage: `groups` / `mlp_*` are the point net (evaluated once per first-stage
   * point on the 8-float row pt_src describes, emitting n_samples / pre_samples samples each, :142-206), the heads and the
   * intersection are those of the second ray_intersect.  The first stage -- a ray net or none, z-planes only -- is: */
  int32_t cascade;
  int32_t pre_samples;                      /* S0 = z_channels of the first stage (<= 32, divides n_samples)             */
  int32_t pre_n_groups;
  hr_encode_group pre_groups[HR_MAX_GROUPS];
  int32_t pre_mlp_in, pre_mlp_width, pre_mlp_layers, pre_mlp_skip;
  int32_t pre_mlp_mode;                     /* HR_MLP_ZERO or the same mode as mlp_mode                                  */
  int32_t pre_head_stride, pre_off_z, pre_off_sigma;  /* first-stage heads: z_vals (1 channel) and optionally sigma      */
  hr_act pre_act_z, pre_act_sigma, pre_isect_act;
  int32_t pre_use_sigma, pre_sort;
  float pre_z_scale, pre_near, pre_far;     /* like z_scale / isect_near / isect_far                                     */
  float pre_samples_tab[32];
  int32_t pt_src[8];                        /* HR_PT_*: what channel k of the point net's input row holds (:151-160)     */
} hr_config;

/* sources of the point net's inputs (`inputs:` of point_prediction; the named tensors are concatenated in YAML order) */
enum { HR_PT_NONE = -1, HR_PT_POINT_X = 0, HR_PT_POINT_Y = 1, HR_PT_POINT_Z = 2, HR_PT_VIEW_X = 3, HR_PT_VIEW_Y = 4,
       HR_PT_VIEW_Z = 5, HR_PT_ORIGIN_X = 6, HR_PT_ORIGIN_Y = 7, HR_PT_ORIGIN_Z = 8, HR_PT_TIME = 9 };

/* Parameters in the reference's own state_dict layout, fp32, contiguous.
 * hr_upload re-lays them out on the device (channel-last tables, packed / split net weights) into
 * memory owned by the handle; the caller's buffers are not referenced after the call returns
 * (device sources: after the stream reaches that point). */
typedef struct hr_params {
  int32_t on_device;                     /* 1: pointers are device pointers, 0: host     */
  const float* mlp_weight[HR_MAX_LAYERS];/* layers.{l}[.0].weight  [out_l, in_l] row-major */
  const float* mlp_bias[HR_MAX_LAYERS];  /* layers.{l}[.0].bias    [out_l]               */
  /* density_plane[_space].{i} / app_plane[_space].{i}: [1, C_i, H_i, W_i]  (C_i may be 0 -> NULL) */
  const float* sigma_plane[3];
  const float* app_plane[3];
  int32_t plane_h[3], plane_w[3];
  /* second factor: dynamic  density_plane_time.{i} / app_plane_time.{i}  [1, C_i, K, L_i]
   *                static   density_line.{i}       / app_line.{i}        [1, C_i, L_i, 1]      */
  const float* sigma_second[3];
  const float* app_second[3];
  int32_t second_len[3];                 /* L_i                                          */
  const float* basis_mat;                /* basis_mat.weight [app_dim, sum(n_app)]       */
  const float* color_embedding;          /* embeddings.{i}.color_embedding [n_color_views, 12] (NULL when n_color_views == 0) */
  /* cascade: the first-stage ray net (embeddings.0.net); mlp_weight / mlp_bias above are then the point net's */
  const float* pre_mlp_weight[HR_MAX_LAYERS];
  const float* pre_mlp_bias[HR_MAX_LAYERS];
} hr_params;

typedef struct hr_handle hr_handle;

/* Version of this ABI (compare with HR_ABI_VERSION). */
int hr_abi_version(void);

/* Last error message of the calling thread ("" if none).  Replaces: Python exceptions. */
const char* hr_last_error(void);

/* Replaces: model_dict['lightfield'](cfg.model, system) + render_fn_dict['lightfield'](...)
 * construction (nlf/__init__.py:351-364, nlf/models/models.py:104-129, nlf/rendering.py:59-70).
 * `device` is the CUDA ordinal.  Fails if the device is not sm_100 or the signature is unsupported. */
int hr_create(const hr_config* cfg, int device, hr_handle** out);

/* Replaces: INRSystem.load_state_dict (nlf/__init__.py:433-479) for the render path: ingest a
 * state_dict (any grid size: sizes come from the tensors, Appendix B), pack for the kernels. */
int hr_upload(hr_handle* h, const hr_params* p, void* stream);

/* Bytes of scratch hr_render needs for n rays (per-ray sample-net outputs).  Bounded: hr_render walks a large batch in
 * sub-batches of 16 sample-net tile waves (16 x num_sms x 128 rays), so the scratch never exceeds ~0.6 GB. */
int64_t hr_workspace_bytes(const hr_handle* h, int64_t n_rays);
/* Scratch of hr_render_heads / hr_render_backward (two full [n, mlp_out] buffers). */
int64_t hr_train_workspace_bytes(const hr_handle* h, int64_t n_rays);
/* Tuning: rays per sub-batch of hr_render (0 = default: 16 tile waves; < 0 = never split, scratch grows with n). */
int hr_set_sub_batch(hr_handle* h, int64_t rays);

/* Replaces: RenderLightfield.forward -> LightfieldModel.forward (nlf/rendering.py:72-77,
 * nlf/models/models.py:135-138) for one chunk: rays [n, c_in] fp32 device -> rgb [n,3] fp32 device.
 * `workspace` is device scratch of at least hr_workspace_bytes(h, n) bytes (16B aligned). */
int hr_render(hr_handle* h, const float* rays, int64_t n_rays, float* rgb,
              void* workspace, int64_t workspace_bytes, void* stream);

/* Ray-sharded rendering (the reference renders a frame on rank 0 only, nlf/__init__.py:810-811).
 * Like hr_render, but the finished pixels of rays [0, n_rays) are stored at rows [row0, row0 + n_rays) of EVERY buffer
 * dst[0..n_dst): the caller passes its own [N_total,3] gather buffer and the peer-mapped gather buffers of the other ranks
 * (CUDA IPC / symmetric memory over NVLink), so the render kernel's epilogue is the gather of the pixel tiles -- 12 bytes
 * per ray and peer cross the fabric, no collective kernel follows.  The caller orders the peers' reads (a barrier after
 * the kernel on `stream`). */
int hr_render_scatter(hr_handle* h, const float* rays, int64_t n_rays, float* const* dst, int32_t n_dst, int64_t row0,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* Debug/bisect variant (SURVEY.md section 4 "stage-boundary tests").  Any output pointer may be
 * NULL.  mlp_out [n, mlp_out] is in the reference's order (sample-major, ray.py:333);
 * distances [n,S] sorted t (base.py:206-210,257); points [n,S,3] final sample points (after flow
 * and offset); sigma [n,S]; weights [n,S] compositing weights (tensorf_utils.py:242-253); rgb_samples [n,S,3] the shaded
 * colour of every sample (renderModule output scattered by app_mask, tensorf_dynamic.py:757-777) before the colour transform. */
int hr_render_stages(hr_handle* h, const float* rays, int64_t n_rays, float* rgb,
                     float* mlp_out, float* distances, float* points, float* sigma, float* weights, float* rgb_samples,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ---- extra outputs of the colour net ----
 * Replaces: the `fields` / `no_over_fields` / `pred_weights_fields` render_kwargs of TensorVMKeyframeTime.forward /
 * TensorVMNoSample.forward (nlf/nets/tensorf_dynamic.py:808-837, nlf/nets/tensorf_no_sample.py:254-278) and the per-sample
 * dict RayPointEmbedding.forward returns to render_fn.embed (nlf/embedding/embedding.py:100-117).  A field is a key of the
 * dict `x` that reaches the colour net after `extract_fields`; the render kernel's epilogue reduces it in the warp that owns
 * the ray, nothing per-sample is written unless HR_FIELD_NO_OVER asks for it. */
enum {
  HR_FIELD_POINTS = 0,      /* x['points']       3 channels (after flow and offset)                   */
  HR_FIELD_DISTANCES = 1,   /* x['distances']    1           (sorted, contracted)                      */
  HR_FIELD_BASE_TIMES = 2,  /* x['base_times']   1           keyframe time (flow_utils.py:18-31)       */
  HR_FIELD_TIME_OFFSET = 3, /* x['time_offset']  1           t - base_time                             */
  HR_FIELD_TIMES = 4,       /* x['times']        1           rays[:, -1] broadcast (point.py:864-867)  */
  HR_FIELD_VIEWDIRS = 5,    /* x['viewdirs']     3           rays[:, 3:6] broadcast                    */
  HR_FIELD_WEIGHTS = 6,     /* x['weights']      1           ones (base.py:183-191, no weight_fn)      */
  /* per-sample heads of the sample net after their activation (ray.py:333-337), reachable through render_kwargs['fields']
   * (ExtractFieldsEmbedding adds them, point.py:236-244) */
  HR_FIELD_COLOR_SCALE = 7,   /* x['color_scale']   3                                                    */
  HR_FIELD_COLOR_SHIFT = 8,   /* x['color_shift']   3                                                    */
  HR_FIELD_SPATIAL_FLOW = 9,  /* x['spatial_flow']  3   as AdvectPoints leaves it (point.py:815-817)           */
  HR_FIELD_SIGMA = 10,        /* x['sigma']         1                                                    */
  HR_FIELD_POINT_SIGMA = 11,  /* x['point_sigma']   1                                                    */
  HR_FIELD_POINT_OFFSET = 12, /* x['point_offset']  3   as PointOffset leaves it: act(.) * (1 - sigma) (point.py:383-389) */
  HR_FIELD_COLOR_SCALE_GLOBAL = 13, /* x['color_scale_global'] 3  (per-sample head; the colour net uses sample 0's, tensorf_utils.py:275-281) */
  HR_FIELD_COLOR_SHIFT_GLOBAL = 14, /* x['color_shift_global'] 3                                                                   */
  HR_N_FIELDS = 15
};
enum {
  HR_FIELD_OVER = 0,         /* out [n, dim]   = sum_s w_s * x_s            (tensorf_dynamic.py:832-836)            */
  HR_FIELD_NO_OVER = 1,      /* out [n, S*dim] = x                          (:824-825)                              */
  HR_FIELD_PRED_WEIGHTS = 2  /* out [n, dim]   = sum_s alpha2weights(x['weights'])_s * x_s   (:818-819, :826-830)   */
};
typedef struct hr_field_request {
  int32_t field;  /* HR_FIELD_*                       */
  int32_t mode;   /* HR_FIELD_OVER / NO_OVER / PRED_WEIGHTS */
  float* out;     /* device buffer, shape per mode    */
} hr_field_request;

/* hr_render plus extra outputs: rgb [n,3]; render_weights [n,S] (the 'render_weights' key, :822-823) when non-NULL; every
 * request in req[0..n_req).  A field the pipeline does not carry (e.g. base_times of a static model) is an error. */
int hr_render_fields(hr_handle* h, const float* rays, int64_t n_rays, float* rgb, float* render_weights,
                     const hr_field_request* req, int32_t n_req, void* workspace, int64_t workspace_bytes, void* stream);

/* Replaces: INRSystem.forward(coords) on HOST buffers, i.e. the `.cuda()` upload, render_chunked
 * (nlf/rendering.py:100-150) and the `.cpu()` read-back of validation_video
 * (nlf/__init__.py:828-855).  rays_host/rgb_host should be pinned; the call splits the batch in
 * `chunk` rays (0 = default), overlaps H2D / kernels / D2H on internal streams and returns when the
 * rgb is on the host. */
int hr_render_host(hr_handle* h, const float* rays_host, int64_t n_rays, float* rgb_host, int64_t chunk);

/* ---- the step before the path: camera -> rays on the device ----
 * Replaces: get_coords_from_camera / get_coords (datasets/base.py:485-518, datasets/technicolor.py:360-396), i.e.
 * get_ray_directions_K + get_rays (+ get_ndc_rays_fx_fy) of utils/ray_utils.py:98-164, executed on the CPU by the
 * reference and uploaded with .cuda() per frame (nlf/__init__.py:828-834).  Pixel p = y*width + x, row-major like
 * kornia.create_meshgrid(H, W, normalized_coordinates=False). */
typedef struct hr_camera {
  float c2w[12];            /* camera-to-world, row-major 3x4 (pose[:3,:4])                         */
  float fx, fy, cx, cy;     /* K[0,0], K[1,1], K[0,2], K[1,2]                                       */
  int32_t width, height;    /* W, H                                                                 */
  int32_t centered_pixels;  /* +0.5 pixel offset (ray_utils.py:103-104)                             */
  int32_t flipped;          /* sign of the y direction (ray_utils.py:108)                           */
  int32_t normalize;        /* get_rays(normalize=True) (ray_utils.py:127-128)                      */
  int32_t use_ndc;          /* get_ndc_rays_fx_fy (ray_utils.py:137-164) with H, W, fx, fy of this camera */
  float ndc_near;           /* dataset.near                                                         */
  float cam_idx, time;      /* channels 6 and 7 when c_in == 8 (technicolor.py:389-393)             */
} hr_camera;

/* rays_out [n_pixels, c_in] fp32 device, for pixels first_pixel .. first_pixel + n_pixels - 1; c_in is 6 or 8. */
int hr_generate_rays(const hr_camera* cam, int32_t c_in, int64_t first_pixel, int64_t n_pixels, float* rays_out,
                     void* stream);

/* ---- the step after the path: 8-bit packing ----
 * Replaces: to8b(x) = (255 * clip(x, 0, 1)).astype(uint8) (utils/__init__.py:47) applied to the rendered frame before
 * it is written or displayed (nlf/__init__.py:857-891, utils/gui_utils.py:174-186).  Same as hr_render, but the fused
 * kernel's epilogue stores rgb8 [n,3] uint8 (3 B/ray instead of 12 B/ray leave the GPU). */
int hr_render_to8b(hr_handle* h, const float* rays, int64_t n_rays, uint8_t* rgb8, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* Whole frame on HOST output: rays generated on the device from `cam`, rendered, packed to 8 bit, copied into the
 * (pinned) host buffer rgb8_host [width*height, 3]; synchronous.  Replaces one iteration of validation_video /
 * NeRFGUI.test_step (nlf/__init__.py:828-891, utils/gui_utils.py:139-212). */
int hr_render_frame_to8b_host(hr_handle* h, const hr_camera* cam, uint8_t* rgb8_host, int64_t chunk);

/* ---- backward pass of the path ----
 * Replaces: what loss.backward() runs for the render path inside INRSystem.training_step (nlf/__init__.py:634-709): the
 * autograd graph of RayPointEmbedding + TensorVMKeyframeTime / TensorVMNoSample.  The sample net's Linear layers stay with
 * the caller (their forward / backward are plain GEMMs on activations the caller keeps); the library provides the three
 * pieces around them, all hand-written kernels:
 *   hr_encode_rays      rays -> encoded sample-net input (RayParam + PE, ray.py:320-326), kernel feature order
 *   hr_render_heads     sample-net output -> rgb, the forward of everything after the net (training or eval semantics)
 *   hr_render_backward  d rgb -> d (sample-net output); gradients of the VM tables and basis_mat accumulate in the handle
 *   hr_grad_zero / hr_grad_read   clear / export the accumulated parameter gradients in the reference's tensor layouts
 * Supported: z_plane / sphere / cylinder primitives with origin_scale_factor == 0, no or mipnerf contraction, per-sample
 * colour heads; other pipelines are rejected (hr_last_error). */
typedef struct hr_train_opts {
  int32_t clamp_output; /* 1: eval() forward, clamp(0,1) (tensorf_dynamic.py:805-806); 0: training forward            */
  int32_t white_bg;     /* rgb_map += 1 - acc_map: cfg.white_bg, or the training coin flip of :795-796 drawn by the caller */
} hr_train_opts;

typedef struct hr_grads {  /* device buffers in the reference's layouts (hr_params), overwritten by hr_grad_read; NULL = skip */
  float* sigma_plane[3];
  float* app_plane[3];
  float* sigma_second[3];
  float* app_second[3];
  float* basis_mat;
} hr_grads;

/* enc [n, mlp_in] fp32 device */
int hr_encode_rays(hr_handle* h, const float* rays, int64_t n_rays, float* enc, void* stream);
/* heads [n, mlp_out] in the reference's order (sample-major: column s*head_stride + c, ray.py:333); workspace of
 * hr_train_workspace_bytes(h, n) */
int hr_render_heads(hr_handle* h, const float* rays, const float* heads, int64_t n_rays, float* rgb, const hr_train_opts* opts,
                    void* workspace, int64_t workspace_bytes, void* stream);
/* d_rgb [n,3] -> d_heads [n, mlp_out] (reference order); workspace of hr_train_workspace_bytes(h, n) */
int hr_render_backward(hr_handle* h, const float* rays, const float* heads, int64_t n_rays, const float* d_rgb, float* d_heads,
                       const hr_train_opts* opts, void* workspace, int64_t workspace_bytes, void* stream);
int hr_grad_zero(hr_handle* h, void* stream);
int hr_grad_read(hr_handle* h, const hr_grads* out, void* stream);

/* Number of kernels hr_render launched since creation (bench.py's gpu_launches). */
int64_t hr_launch_count(const hr_handle* h);

/* Average device time (ms, CUDA events on the launching stream) of the dominant kernel -- the fused
 * gather+decode+composite kernel -- over launches since the last hr_timing_reset; needs
 * hr_timing_enable(h, 1).  Used by bench.py for roofline.achieved. */
int hr_timing_enable(hr_handle* h, int enable);
int hr_timing_reset(hr_handle* h);
int hr_timing_read(hr_handle* h, double* render_ms_avg, double* mlp_ms_avg, int64_t* launches);
/* same for the render-backward kernel of hr_render_backward */
int hr_timing_read_backward(hr_handle* h, double* backward_ms_avg, int64_t* launches);

/* Replaces: module destruction. */
int hr_destroy(hr_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* HYPERREEL_B200_H */
