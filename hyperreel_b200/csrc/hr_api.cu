// C-ABI of libhyperreel_b200.so (declared in include/hyperreel_b200.h).
// Host-side glue only: parameter packing, workspace carving, kernel launches, timing.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "hr_common.cuh"
#include "hr_encode.cuh"
#include "hr_geom.cuh"
#include "hr_mlp.cuh"
#include "hyperreel_b200.h"

namespace hr {
cudaError_t launch_render(const hr_config& cfg, const Derived& dv, const RenderTabs& tabs, const float* rays,
                          const float* heads, const RgbDst& rgb, long long n, const ExtraOut* so, int num_sms,
                          cudaStream_t stream, unsigned char* rgb8);
cudaError_t launch_render_bwd(const hr_config& cfg, const Derived& dv, const RenderTabs& tabs, float* const* g_sig_space,
                              float* const* g_sig_second, float* const* g_app_space, float* const* g_app_second, float* g_basis,
                              const float* rays, const float* heads, const float* d_rgb, float* d_heads, long long n, int clamp_output,
                              int white_bg, int num_sms, cudaStream_t stream);
cudaError_t launch_generate_rays(const hr_camera& cam, int c_in, long long first, long long n, float* out, cudaStream_t st);
}  // namespace hr

static thread_local std::string g_err;

int hr_fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

// Every entry point runs with the handle's device current and restores the caller's device on the way out (the reference
// never changes torch's current device; a stray cudaSetDevice here would silently move the caller's later allocations).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#include "hr_handle.h"

namespace {

__global__ void pack_channel_last(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W) {
  // src [C][H][W] (reference [1,C,H,W]) -> dst [H][W][C]
  long long total = (long long)C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long hw = i / C;
    dst[i] = src[(long long)c * H * W + hw];
  }
}

// (axis, time) planes -> per-keyframe lines.  The reference samples plane_time[1,C,K,L] bilinearly at (u_c, tau) where
// tau = normalize_time_coord(base_t) (tensorf_dynamic.py:615-616) and base_t is the ray's keyframe time
// (utils/flow_utils.py:18-31): tau therefore takes exactly K values, one per keyframe k, and the two rows grid_sample
// blends (rows it_k, it_k+1 with fraction ft_k, align_corners=True) are a property of k alone.  dst[k][l][c] holds that
// blend, so the render kernel's second factor is a 2-tap linear lookup instead of 4 taps.
__global__ void pack_time_lines(const float* __restrict__ src, float* __restrict__ dst, int C, int K, int L, float inv_fac,
                                float time_scale, float time_offset) {
  long long total = (long long)K * L * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int l = (int)((i / C) % L);
    int k = (int)(i / ((long long)C * L));
    float base_t = __fmul_rn((float)k, inv_fac);
    float tau = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(base_t, time_scale), time_offset), 2.0f), 1.0f);
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(tau, 1.0f), 0.5f), (float)(K - 1));
    int it = max(0, min((int)floorf(iy), K - 2));
    float ft = iy - (float)it;
    float a = src[((long long)c * K + it) * L + l];
    float b = src[((long long)c * K + it + 1) * L + l];
    dst[i] = __fmul_rn(1.0f - ft, a) + __fmul_rn(ft, b);
  }
}

// Wt[k][n] (k-major, zero padded) from the reference weight W[out][in] (nn.Linear layout).
//   k -> source column: k < in_pad: (k < in_ch ? k : none) when the layer consumes the encoded input;
//                        hidden rows follow.  n -> source row: last layer channel-major permutation.
__global__ void pack_simt_layer(const float* __restrict__ Wsrc, const float* __restrict__ bsrc, float* __restrict__ Wt,
                                float* __restrict__ bias, int Kp, int Np, int out_ch, int in_ch_src, int in_enc,
                                int in_pad, int has_input, int has_hidden, int width, int perm_S, int perm_stride) {
  long long total = (long long)Kp * Np;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total + Np; i += (long long)gridDim.x * blockDim.x) {
    if (i >= total) {
      int n = (int)(i - total);
      int ns = n;
      if (perm_S > 0 && n < out_ch) ns = (n % perm_S) * perm_stride + (n / perm_S);
      bias[n] = (n < out_ch) ? bsrc[ns] : 0.0f;
      continue;
    }
    int n = (int)(i % Np);
    int k = (int)(i / Np);
    float v = 0.0f;
    if (n < out_ch) {
      int ns = n;
      if (perm_S > 0) ns = (n % perm_S) * perm_stride + (n / perm_S);  // n = c*S+s  <-  s*stride+c
      int ks = -1;
      if (has_input) {
        if (k < in_pad) ks = (k < in_enc) ? k : -1;
        else if (has_hidden) ks = in_enc + (k - in_pad);
      } else {
        ks = k;
      }
      if (ks >= 0 && ks < in_ch_src && (has_input || k < width)) v = Wsrc[(long long)ns * in_ch_src + ks];
    }
    Wt[i] = v;
  }
}

__global__ void unpermute_heads(const float* __restrict__ src, float* __restrict__ dst, long long n, int S, int stride) {
  // src [n][c*S+s] -> dst [n][s*stride+c]
  long long total = n * (long long)S * stride;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long ray = i / (S * stride);
    int rem = (int)(i % (S * stride));
    int s = rem / stride, c = rem % stride;
    dst[i] = src[ray * (long long)S * stride + c * S + s];
  }
}

__global__ void permute_heads(const float* __restrict__ src, float* __restrict__ dst, long long n, int S, int stride) {
  // src [n][s*stride+c] (reference order) -> dst [n][c*S+s] (the kernels' channel-major rows)
  long long total = n * (long long)S * stride;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long ray = i / (S * stride);
    int rem = (int)(i % (S * stride));
    int c = rem / S, s = rem % S;
    dst[i] = src[ray * (long long)S * stride + s * stride + c];
  }
}

__global__ void encode_rays_kernel(const __grid_constant__ hr_config cfg, const float* __restrict__ rays, float* __restrict__ enc,
                                   long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    hr::encode_ray(cfg, rays + i * cfg.c_in, enc + i * cfg.mlp_in, 1);
}

// First stage of a cascaded pipeline (PointPredictionEmbedding, nlf/embedding/point.py:142-160): one warp per ray,
// lane = first-stage sample.  heads0 [n][c*S0+s] are the ray net's outputs (null: `zero` net); the S0 z-planes are
// intersected, masked and sorted like any other (Intersect.forward, base.py:142-226; IntersectZPlane, z.py:77-97), and every
// point o + t d becomes one 8-float input row of the point net (channel k holds cfg.pt_src[k]).
__global__ void cascade_points_kernel(const __grid_constant__ hr_config cfg, const float* __restrict__ rays,
                                      const float* __restrict__ heads0, float* __restrict__ rows, long long n) {
  const int lane = threadIdx.x & 31;
  const int S0 = cfg.pre_samples;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int out0 = S0 * cfg.pre_head_stride;
  for (long long ray = warp0; ray < n; ray += nwarps) {
    const float* r = rays + ray * cfg.c_in;
    const float ox = __ldg(r + 0), oy = __ldg(r + 1), oz = __ldg(r + 2);
    const float dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
    const float time = __ldg(r + cfg.c_in - 1);
    const bool act = lane < S0;
    const int s = act ? lane : 0;
    const float zraw = heads0 ? __ldg(heads0 + ray * out0 + cfg.pre_off_z * S0 + s) : 0.0f;
    const float sraw = (heads0 && cfg.pre_off_sigma >= 0) ? __ldg(heads0 + ray * out0 + cfg.pre_off_sigma * S0 + s) : 0.0f;
    const float sg = cfg.pre_use_sigma ? hr::apply_act(cfg.pre_act_sigma, sraw) : 0.0f;
    const float zr = __fmul_rn(hr::apply_act(cfg.pre_isect_act, hr::apply_act(cfg.pre_act_z, zraw)), __fsub_rn(1.0f, sg));
    const float z = __fadd_rn(__fmul_rn(zr, cfg.pre_z_scale), cfg.pre_samples_tab[s]);
    const float dzg = (fabsf(dz) < 1e-5f) ? 1e12f : dz;  // intersect_utils.py:135-142
    float t = __fdiv_rn(__fsub_rn(z, oz), dzg);
    if ((t <= cfg.pre_near) || (t >= cfg.pre_far)) t = 0.0f;
    float key[1] = {act ? t : __int_as_float(0x7f800000)};
    if (cfg.pre_sort) hr::sort_keys<1>(key, lane);
    t = key[0];
    if (!act) continue;
    const float px = __fadd_rn(ox, __fmul_rn(dx, t)), py = __fadd_rn(oy, __fmul_rn(dy, t)), pz = __fadd_rn(oz, __fmul_rn(dz, t));
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float x = 0.0f;
      switch (cfg.pt_src[k]) {
        case HR_PT_POINT_X: x = px; break;
        case HR_PT_POINT_Y: x = py; break;
        case HR_PT_POINT_Z: x = pz; break;
        case HR_PT_VIEW_X: x = dx; break;
        case HR_PT_VIEW_Y: x = dy; break;
        case HR_PT_VIEW_Z: x = dz; break;
        case HR_PT_ORIGIN_X: x = ox; break;
        case HR_PT_ORIGIN_Y: x = oy; break;
        case HR_PT_ORIGIN_Z: x = oz; break;
        case HR_PT_TIME: x = time; break;
        default: break;
      }
      v[k] = x;
    }
    float4* dst = reinterpret_cast<float4*>(rows + (ray * S0 + lane) * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// gradient table [H][W][C] (channel-last, the kernels' layout) -> reference layout [C][H][W]
__global__ void unpack_channel_last(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W) {
  long long total = (long long)C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long hw = i % ((long long)H * W);
    int c = (int)(i / ((long long)H * W));
    dst[i] = src[hw * C + c];
  }
}

// gradient of the pre-blended keyframe lines [K][L][C] -> gradient of the (axis, time) plane [C][K][L]: row r collects
// (1 - ft_k) of every keyframe k whose lower row is r and ft_k of every keyframe whose upper row is r (see pack_time_lines)
__global__ void unblend_time_lines(const float* __restrict__ src, float* __restrict__ dst, int C, int K, int L, float inv_fac,
                                   float time_scale, float time_offset) {
  long long total = (long long)C * K * L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int l = (int)(i % L);
    int r = (int)((i / L) % K);
    int c = (int)(i / ((long long)K * L));
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) {
      float base_t = __fmul_rn((float)k, inv_fac);
      float tau = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(base_t, time_scale), time_offset), 2.0f), 1.0f);
      float iy = __fmul_rn(__fmul_rn(__fadd_rn(tau, 1.0f), 0.5f), (float)(K - 1));
      int it = max(0, min((int)floorf(iy), K - 2));
      float ft = iy - (float)it;
      float g = src[((long long)k * L + l) * C + c];
      if (it == r) acc += (1.0f - ft) * g;
      if (it + 1 == r) acc += ft * g;
    }
    dst[i] = acc;
  }
}

int grid_for(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}

// Packed-parameter storage.  hr_upload asks for its buffers in a fixed order; a buffer whose size is unchanged since the
// previous upload is reused, so a parameter refresh (same grid, same net) performs no cudaFree / cudaMalloc.
int dev_alloc(hr_handle* h, void** p, size_t bytes) {
  bytes = bytes ? bytes : 16;
  const size_t i = h->slot_cursor++;
  if (i < h->slots.size() && h->slots[i].bytes == bytes) {
    *p = h->slots[i].ptr;
    return 0;
  }
  if (i < h->slots.size()) {
    cudaFree(h->slots[i].ptr);
    h->slots[i] = {nullptr, 0};
  } else {
    h->slots.push_back({nullptr, 0});
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) return fail("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  h->slots[i] = {*p, bytes};
  return 0;
}

// Copy a reference-layout tensor to the device if it is on the host (returns device pointer).
int stage_in(const float* src, size_t count, int on_device, cudaStream_t st, std::vector<void*>& temps, const float** out) {
  if (on_device) {
    *out = src;
    return 0;
  }
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, count * sizeof(float));
  if (e != cudaSuccess) return fail("cudaMalloc(temp %zu) failed: %s", count * sizeof(float), cudaGetErrorString(e));
  temps.push_back(d);
  e = cudaMemcpyAsync(d, src, count * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return fail("H2D of parameters failed: %s", cudaGetErrorString(e));
  *out = (const float*)d;
  return 0;
}

int validate(const hr_config& c) {
  if (c.abi_version != HR_ABI_VERSION) return fail("hr_config.abi_version %d != %d", c.abi_version, HR_ABI_VERSION);
  if (c.c_in != 6 && c.c_in != 8) return fail("unsupported c_in %d (6: static rays, 8: video rays)", c.c_in);
  if (c.n_groups < 1 || c.n_groups > HR_MAX_GROUPS) return fail("unsupported n_groups %d", c.n_groups);
  if (c.mlp_mode != HR_MLP_FP32_SIMT && c.mlp_mode != HR_MLP_BF16X3_TC && c.mlp_mode != HR_MLP_ZERO) return fail("unsupported mlp_mode");
  const bool has_net = c.mlp_mode != HR_MLP_ZERO;
  if (has_net && (c.mlp_layers < 2 || c.mlp_layers > HR_MAX_LAYERS)) return fail("unsupported mlp_layers %d", c.mlp_layers);
  if (has_net && c.mlp_width != 128 && c.mlp_width != 256) return fail("unsupported mlp_width %d (128 or 256)", c.mlp_width);
  if (c.mlp_in < 1 || c.mlp_in > 64) return fail("unsupported mlp_in %d", c.mlp_in);
  if (has_net && c.mlp_skip != -1 && (c.mlp_skip < 1 || c.mlp_skip > c.mlp_layers - 2)) return fail("bad mlp_skip %d", c.mlp_skip);
  if (c.n_samples < 1 || c.n_samples > HR_MAX_SAMPLES) return fail("unsupported n_samples %d (max %d)", c.n_samples, HR_MAX_SAMPLES);
  if (c.mlp_out != c.n_samples * c.head_stride) return fail("mlp_out %d != S*head_stride %d", c.mlp_out, c.n_samples * c.head_stride);
  if (c.off_z < 0) return fail("z_vals head is required");
  if ((c.isect_type == HR_ISECT_Z_PLANE || c.isect_type == HR_ISECT_DISTANCE) && c.n_z != 1) return fail("z_plane / euclidean_distance need 1 z channel");
  if ((c.isect_type == HR_ISECT_SPHERE || c.isect_type == HR_ISECT_CYLINDER) && c.n_z != 4) return fail("sphere / cylinder need 4 z channels");
  if (c.isect_type == HR_ISECT_SPHERE_NEW && c.n_z != 8) return fail("sphere_new needs 8 z channels");
  if (c.isect_type < HR_ISECT_Z_PLANE || c.isect_type > HR_ISECT_PLANE) return fail("unsupported intersect type %d", c.isect_type);
  if (c.isect_type == HR_ISECT_VOXEL && (c.n_z != 1 || c.isect_axes != 3 || c.n_samples % 3 != 0))
    return fail("voxel_grid needs 1 z channel and a multiple of 3 samples");
  if (c.isect_type == HR_ISECT_PLANE && (c.n_z != 4 || c.isect_axes < 1 || c.isect_axes > 3 || c.n_samples % c.isect_axes != 0))
    return fail("deformable_voxel_grid needs 4 z channels and 1-3 axes dividing the sample count");
  if (c.cascade) {
    if (c.pre_samples < 1 || c.pre_samples > 32 || c.n_samples % c.pre_samples != 0) return fail("cascade: bad pre_samples %d", c.pre_samples);
    if (c.mlp_mode == HR_MLP_ZERO) return fail("cascade: the point net cannot be a zero net");
    if (c.pre_mlp_mode != HR_MLP_ZERO && c.pre_mlp_mode != c.mlp_mode) return fail("cascade: pre_mlp_mode must be zero or mlp_mode");
    if (c.pre_head_stride < 1 || c.pre_off_z < 0 || c.pre_off_z >= c.pre_head_stride || c.pre_off_sigma >= c.pre_head_stride)
      return fail("cascade: bad first-stage head layout");
    if (c.pre_mlp_mode != HR_MLP_ZERO) {
      if (c.pre_n_groups < 1 || c.pre_n_groups > HR_MAX_GROUPS) return fail("cascade: unsupported pre_n_groups %d", c.pre_n_groups);
      if (c.pre_mlp_layers < 2 || c.pre_mlp_layers > HR_MAX_LAYERS) return fail("cascade: unsupported pre_mlp_layers %d", c.pre_mlp_layers);
      if (c.pre_mlp_width != 128 && c.pre_mlp_width != 256) return fail("cascade: unsupported pre_mlp_width %d", c.pre_mlp_width);
      if (c.pre_mlp_in < 1 || c.pre_mlp_in > 64) return fail("cascade: unsupported pre_mlp_in %d", c.pre_mlp_in);
      if (c.pre_mlp_skip != -1 && (c.pre_mlp_skip < 1 || c.pre_mlp_skip > c.pre_mlp_layers - 2)) return fail("cascade: bad pre_mlp_skip");
      if ((c.pre_samples * c.pre_head_stride) % 4 != 0) return fail("cascade: first-stage output width must be a multiple of 4");
    }
    if ((c.mlp_out / c.pre_samples) % 4 != 0) return fail("cascade: the point net's output width must be a multiple of 4");
    for (int g = 0; g < c.n_groups; ++g)
      if (c.groups[g].start < 0 || c.groups[g].end > 8) return fail("cascade: point-net param group outside the 8-channel row");
  }
  if (c.n_color_views < 0) return fail("bad n_color_views");
  if (c.n_color_views > 0 && c.c_in != 8) return fail("colour transform needs 8-channel rays (camera id = rays[:, -2])");
  if (c.n_color_views > 0 && c.off_cscale_global >= 0) return fail("colour transform and global colour heads are exclusive");
  if (c.contract_type != HR_CONTRACT_NONE && c.contract_type != HR_CONTRACT_MIPNERF && c.contract_type != HR_CONTRACT_AFFINE)
    return fail("unsupported contract type");
  if (c.contract_type == HR_CONTRACT_AFFINE) {
    for (int i = 0; i < 3; ++i)
      if (c.contract_affine_den[i] == 0.0f) return fail("affine contraction: zero extent on axis %d", i);
    if (c.contract_dist_fac == 0.0f) return fail("affine contraction: zero distance factor");
  }
  if ((c.off_cscale_global >= 0) != (c.off_cshift_global >= 0)) return fail("color_scale_global and color_shift_global come together");
  if (c.off_cscale_global + 3 > c.head_stride || c.off_cshift_global + 3 > c.head_stride) return fail("global colour heads out of range");
  if (c.use_flow && (c.off_flow < 0 || c.num_keyframes < 1 || c.num_frames < 1)) return fail("flow needs spatial_flow head and K,F");
  if (c.use_offset && c.off_offset < 0) return fail("point_offset needs point_offset head");
  if (c.use_color_scale_shift && (c.off_cscale < 0 || c.off_cshift < 0)) return fail("colour scale/shift heads missing");
  if (c.dynamic && (c.num_keyframes < 1 || c.num_frames < 1)) return fail("dynamic net needs K,F");
  for (int i = 0; i < 3; ++i)
    if (c.n_sigma[i] != c.n_app[i]) return fail("n_lamb_sigma != n_lamb_sh not supported");
  const int* s = c.n_sigma;
  bool ok = (s[0] == 8 && s[1] == 0 && s[2] == 0) || (s[0] == 8 && s[1] == 4 && s[2] == 4) || (s[0] == 8 && s[1] == 8 && s[2] == 8);
  if (!ok) return fail("unsupported component layout [%d,%d,%d]", s[0], s[1], s[2]);
  if (c.shading == HR_SHADE_SH && c.app_dim != 27) return fail("SH shading needs app_dim 27");
  if (c.shading == HR_SHADE_RGB && c.app_dim != 3) return fail("RGB shading needs app_dim 3");
  if (c.shading != HR_SHADE_SH && c.shading != HR_SHADE_RGB) return fail("unsupported shading");
  return 0;
}

void derive(const hr_config& c, hr::Derived& d) {
  double K = c.num_keyframes > 0 ? c.num_keyframes : 1, F = c.num_frames > 0 ? c.num_frames : 1;
  double fac = K * (F - 1.0) / F;
  d.time_fac = (float)fac;
  d.time_inv_fac = (fac != 0.0) ? (float)(1.0 / fac) : 0.0f;
  d.time_scale = (float)((F - 1.0) / F);
  d.time_offset = (float)(0.5 / K);
  d.kf_max = (float)(K - 1.0);
  double ied = (double)c.contract_start_distance / (double)c.contract_end_distance;
  d.inv_end_dist = (float)ied;
  d.dist_scale_fac = (float)(1.0 / (1.0 - ied));
  double ier = (double)c.contract_start_radius / (double)c.contract_end_radius;
  d.inv_end_rad = (float)ier;
  d.rad_scale_fac = (float)(1.0 / (1.0 - ier));
}

}  // namespace

extern "C" {

int hr_abi_version(void) { return HR_ABI_VERSION; }

const char* hr_last_error(void) { return g_err.c_str(); }

int hr_create(const hr_config* cfg, int device, hr_handle** out) {
  if (!cfg || !out) return fail("hr_create: null argument");
  *out = nullptr;
  if (validate(*cfg)) return 1;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail("hr_create: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail("hr_create: device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("hr_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  DeviceGuard guard(device);
  hr_handle* h = new (std::nothrow) hr_handle();
  if (!h) return fail("hr_create: out of memory");
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  derive(h->cfg, h->dv);
  memset(&h->tabs, 0, sizeof(h->tabs));
  memset(&h->simt, 0, sizeof(h->simt));
  memset(&h->tc, 0, sizeof(h->tc));
  memset(&h->simt_pre, 0, sizeof(h->simt_pre));
  memset(&h->tc_pre, 0, sizeof(h->tc_pre));
  h->cfg_net = h->cfg;
  h->cfg_pre = h->cfg;
  if (h->cfg.cascade) {
    const hr_config& c = h->cfg;
    // the point net: one 8-float row per first-stage point in, n_samples / pre_samples samples out, columns in the
    // reference's order (n_samples = 1 makes the packers' channel-major permutation the identity)
    hr_config& n = h->cfg_net;
    n.c_in = 8;
    n.mlp_out = c.mlp_out / c.pre_samples;
    n.n_samples = 1;
    n.head_stride = n.mlp_out;
    // the first-stage ray net
    hr_config& q = h->cfg_pre;
    q.n_groups = c.pre_n_groups;
    for (int g = 0; g < HR_MAX_GROUPS; ++g) q.groups[g] = c.pre_groups[g];
    q.mlp_in = c.pre_mlp_in; q.mlp_width = c.pre_mlp_width; q.mlp_layers = c.pre_mlp_layers; q.mlp_skip = c.pre_mlp_skip;
    q.mlp_mode = c.pre_mlp_mode;
    q.n_samples = c.pre_samples;
    q.head_stride = c.pre_head_stride;
    q.mlp_out = c.pre_samples * c.pre_head_stride;
  }
  *out = h;
  return 0;
}

static void drop_host_graph(hr_handle* h);

int hr_upload(hr_handle* h, const hr_params* p, void* stream) {
  if (!h || !p) return fail("hr_upload: null argument");
  DeviceGuard guard(h->device);
  drop_host_graph(h);
  cudaStream_t st = (cudaStream_t)stream;
  const hr_config& c = h->cfg;
  // Buffers of the previous pack are reused slot by slot when their sizes are unchanged (dev_alloc).  Work already
  // enqueued on `st` that reads the old contents is ordered before the pack kernels below (same stream).
  h->slot_cursor = 0;
  h->uploaded = false;
  h->tc_ready = false;
  std::vector<void*> temps;
  int rc = 0;

  // ---- sample net(s) ----
  auto pack_net = [&](const hr_config& nc, const float* const* wsrc, const float* const* bsrc, hr::MlpSimtPack& simt,
                      hr::MlpTcPack& tc, bool& tc_ready, size_t& tc_bytes, int& tc_bias) -> int {
    const int L = nc.mlp_layers, W = nc.mlp_width;
    const int in_pad = (nc.mlp_in + 15) / 16 * 16;
    simt.in_pad = in_pad;
    simt.n_layers = L;
    simt.skip = nc.mlp_skip;
    tc_ready = false;
    const float* w_dev[HR_MAX_LAYERS] = {nullptr};
    const float* b_dev[HR_MAX_LAYERS] = {nullptr};
    int r = 0;
    for (int l = 0; l < L && !r && nc.mlp_mode != HR_MLP_ZERO; ++l) {
      if (!wsrc[l] || !bsrc[l]) return fail("hr_upload: mlp layer %d missing", l);
      const bool first = (l == 0), last = (l == L - 1), skip = (l == nc.mlp_skip);
      const int in_src = first ? nc.mlp_in : (skip ? nc.mlp_in + W : W);
      const int out_ch = last ? nc.mlp_out : W;
      const int Kp = first ? in_pad : (skip ? in_pad + W : W);
      const int Np = last ? (nc.mlp_out + W - 1) / W * W : W;
      if ((r = stage_in(wsrc[l], (size_t)out_ch * in_src, p->on_device, st, temps, &w_dev[l]))) break;
      if ((r = stage_in(bsrc[l], (size_t)out_ch, p->on_device, st, temps, &b_dev[l]))) break;
      float *Wt = nullptr, *bias = nullptr;
      if ((r = dev_alloc(h, (void**)&Wt, (size_t)Kp * Np * sizeof(float)))) break;
      if ((r = dev_alloc(h, (void**)&bias, (size_t)Np * sizeof(float)))) break;
      pack_simt_layer<<<grid_for((long long)Kp * Np + Np), 256, 0, st>>>(
          w_dev[l], b_dev[l], Wt, bias, Kp, Np, out_ch, in_src, nc.mlp_in, in_pad, (first || skip) ? 1 : 0, skip ? 1 : 0, W,
          last ? nc.n_samples : 0, nc.head_stride);
      simt.Wt[l] = Wt;
      simt.bias[l] = bias;
      simt.Kp[l] = Kp;
      simt.Np[l] = Np;
    }
    if (!r && nc.mlp_mode == HR_MLP_BF16X3_TC) {
      r = hr::pack_mlp_tc2(h, nc, tc, tc_bytes, tc_bias, w_dev, b_dev, st);
      if (!r) tc_ready = true;
    }
    return r;
  };
  rc = pack_net(h->cfg_net, p->mlp_weight, p->mlp_bias, h->simt, h->tc, h->tc_ready, h->tc_alloc_bytes, h->tc_alloc_bias);
  if (!rc && c.cascade)
    rc = pack_net(h->cfg_pre, p->pre_mlp_weight, p->pre_mlp_bias, h->simt_pre, h->tc_pre, h->tc_pre_ready, h->tc_pre_alloc_bytes,
                  h->tc_pre_alloc_bias);

  // ---- VM tables, channel-last ----
  auto pack_tab = [&](const float* src, int C, int H, int Wd, const float** out) -> int {
    *out = nullptr;
    if (C == 0) return 0;
    if (!src) return fail("hr_upload: table with C=%d missing", C);
    const float* d = nullptr;
    if (stage_in(src, (size_t)C * H * Wd, p->on_device, st, temps, &d)) return 1;
    float* dst = nullptr;
    if (dev_alloc(h, (void**)&dst, (size_t)C * H * Wd * sizeof(float))) return 1;
    pack_channel_last<<<grid_for((long long)C * H * Wd), 256, 0, st>>>(d, dst, C, H, Wd);
    *out = dst;
    return 0;
  };
  int n_app_total = 0;
  for (int i = 0; i < 3 && !rc; ++i) {
    const int C = c.n_sigma[i];
    n_app_total += c.n_app[i];
    const int H2 = c.dynamic ? c.num_keyframes : 1;
    hr::PlaneTab& ts = h->tabs.sig[i];
    hr::PlaneTab& ta = h->tabs.app[i];
    ts.C = C; ta.C = c.n_app[i];
    ts.H = ta.H = p->plane_h[i];
    ts.W = ta.W = p->plane_w[i];
    ts.H2 = ta.H2 = H2;
    ts.L = ta.L = p->second_len[i];
    if (C > 0 && (ts.H < 2 || ts.W < 2 || ts.L < 2)) { rc = fail("hr_upload: plane %d too small (%dx%d, L=%d)", i, ts.H, ts.W, ts.L); break; }
    if ((rc = pack_tab(p->sigma_plane[i], C, ts.H, ts.W, &ts.space))) break;
    if ((rc = pack_tab(p->app_plane[i], C, ts.H, ts.W, &ta.space))) break;
    // second factor: static [C][L][1] -> [1][L][C]; dynamic [C][K][L] -> K pre-blended keyframe lines [K][L][C]
    if (!c.dynamic) {
      if ((rc = pack_tab(p->sigma_second[i], C, H2, ts.L, &ts.second))) break;
      if ((rc = pack_tab(p->app_second[i], C, H2, ts.L, &ta.second))) break;
    } else if (C > 0) {
      if (H2 < 2) { rc = fail("hr_upload: the keyframe (time) planes need at least 2 keyframes"); break; }
      for (int f = 0; f < 2 && !rc; ++f) {
        const float* srcp = f ? p->app_second[i] : p->sigma_second[i];
        if (!srcp) { rc = fail("hr_upload: time plane %d missing", i); break; }
        const float* d = nullptr;
        if ((rc = stage_in(srcp, (size_t)C * H2 * ts.L, p->on_device, st, temps, &d))) break;
        float* dst = nullptr;
        if ((rc = dev_alloc(h, (void**)&dst, (size_t)C * H2 * ts.L * sizeof(float)))) break;
        pack_time_lines<<<grid_for((long long)C * H2 * ts.L), 256, 0, st>>>(d, dst, C, H2, ts.L, h->dv.time_inv_fac,
                                                                            h->dv.time_scale, h->dv.time_offset);
        (f ? ta.second : ts.second) = dst;
      }
      if (rc) break;
    }
  }
  if (!rc) {
    // every table derives from one gridSize (tensorf_base.py:911-944, tensorf_dynamic.py:126-169): plane i is
    // [C, grid[b_i], grid[a_i]], its second factor runs along grid[v_i]
    const int rx = p->plane_w[0], ry = p->plane_h[0], rz = p->second_len[0];
    h->dv.res[0] = rx; h->dv.res[1] = ry; h->dv.res[2] = rz;
    h->dv.kt = c.dynamic ? c.num_keyframes : 1;
    if (c.dynamic && c.num_keyframes < 2) rc = fail("hr_upload: the keyframe (time) planes need at least 2 keyframes");
    if (c.n_sigma[1] > 0 && (p->plane_w[1] != rx || p->plane_h[1] != rz || p->second_len[1] != ry))
      rc = fail("hr_upload: table group 1 is inconsistent with grid %dx%dx%d", rx, ry, rz);
    if (!rc && c.n_sigma[2] > 0 && (p->plane_w[2] != ry || p->plane_h[2] != rz || p->second_len[2] != rx))
      rc = fail("hr_upload: table group 2 is inconsistent with grid %dx%dx%d", rx, ry, rz);
  }
  if (!rc) {
    if (!p->basis_mat) rc = fail("hr_upload: basis_mat missing");
    else {
      const float* d = nullptr;
      size_t cnt = (size_t)c.app_dim * n_app_total;
      rc = stage_in(p->basis_mat, cnt, p->on_device, st, temps, &d);
      if (!rc) {
        float* dst = nullptr;
        rc = dev_alloc(h, (void**)&dst, cnt * sizeof(float));
        if (!rc) {
          cudaError_t e = cudaMemcpyAsync(dst, d, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st);
          if (e != cudaSuccess) rc = fail("basis copy failed: %s", cudaGetErrorString(e));
          h->tabs.basis = dst;
          h->tabs.n_app_total = n_app_total;
        }
      }
    }
  }
  if (!rc) {
    h->tabs.color_embedding = nullptr;
    if (c.n_color_views > 0) {
      if (!p->color_embedding) rc = fail("hr_upload: color_embedding missing (n_color_views = %d)", c.n_color_views);
      else {
        const float* d = nullptr;
        const size_t cnt = (size_t)c.n_color_views * 12;
        rc = stage_in(p->color_embedding, cnt, p->on_device, st, temps, &d);
        float* dst = nullptr;
        if (!rc) rc = dev_alloc(h, (void**)&dst, cnt * sizeof(float));
        if (!rc) {
          cudaError_t e = cudaMemcpyAsync(dst, d, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st);
          if (e != cudaSuccess) rc = fail("color_embedding copy failed: %s", cudaGetErrorString(e));
          h->tabs.color_embedding = dst;
        }
      }
    }
  }
  cudaError_t le = cudaGetLastError();
  if (!rc && le != cudaSuccess) rc = fail("hr_upload: pack kernel launch failed: %s", cudaGetErrorString(le));
  if (!temps.empty()) {  // host sources only: the staging copies must outlive the pack kernels
    cudaStreamSynchronize(st);
    for (void* t : temps) cudaFree(t);
  }
  // slots past the cursor belong to a previous, larger layout
  while (h->slots.size() > h->slot_cursor) {
    cudaFree(h->slots.back().ptr);
    h->slots.pop_back();
  }
  if (rc) return rc;
  h->uploaded = true;
  return 0;
}

// bytes of a heads scratch for n rays: one [mlp_out] fp32 row per ray
static int64_t heads_bytes(const hr_handle* h, int64_t n_rays) {
  int64_t b = n_rays * (int64_t)h->cfg.mlp_out * (int64_t)sizeof(float);
  return (b + 255) / 256 * 256 + 256;
}

// hr_render walks a large batch in sub-batches (sample net, then render kernel, per sub-batch) so that the heads scratch
// stays bounded (0.58 GB at S*15 = 480) instead of growing to 8-21 GB for 4 M-ray batches / full Neural-3D frames.
// Measured (profiles/r2_notes.md): wave-sized sub-batches that would keep the scratch in L2 cost more in launch ramps than
// the HBM round trip they save (0.341 vs 0.288 ms per 65 536 rays); 16 tile waves per sub-batch cost ~1.5 % at 1 M rays.
static int64_t sub_batch_rays(const hr_handle* h) {
  if (h->sub_rays > 0) return h->sub_rays;
  return (int64_t)h->num_sms * 128 * 16;
}

// scratch of the first stage of a cascaded pipeline for n rays, placed after the heads: first-stage heads [n][S0*stride0],
// point rows [n*S0][8], point-net output [n][mlp_out] (reference order, before the channel-major permutation)
static int64_t cascade_bytes(const hr_handle* h, int64_t n_rays) {
  const hr_config& c = h->cfg;
  if (!c.cascade) return 0;
  auto seg = [](int64_t floats) { return (floats * (int64_t)sizeof(float) + 255) / 256 * 256; };  // each buffer 256-byte aligned
  return seg(n_rays * c.pre_samples * c.pre_head_stride) + seg(n_rays * c.pre_samples * 8) + seg(n_rays * (int64_t)c.mlp_out);
}

// workspace of one render call over n rays that is not split further
static int64_t ws_bytes_for(const hr_handle* h, int64_t n_rays) { return heads_bytes(h, n_rays) + cascade_bytes(h, n_rays); }

int64_t hr_workspace_bytes(const hr_handle* h, int64_t n_rays) {
  if (!h || n_rays < 0) return -1;
  const int64_t sub = sub_batch_rays(h);
  return ws_bytes_for(h, (h->sub_rays < 0 || n_rays < sub) ? n_rays : sub);
}

int64_t hr_train_workspace_bytes(const hr_handle* h, int64_t n_rays) {
  if (!h || n_rays < 0) return -1;
  return 2 * heads_bytes(h, n_rays);
}

int hr_set_sub_batch(hr_handle* h, int64_t rays) {
  if (!h) return fail("hr_set_sub_batch: null handle");
  h->sub_rays = rays;
  return 0;
}

// The cached hr_render_host graph holds kernel parameters by value: drop it whenever they may have changed.
static void drop_host_graph(hr_handle* h) {
  if (h->pipe.graph) cudaGraphExecDestroy(h->pipe.graph);
  h->pipe.graph = nullptr;
  h->pipe.g_rays = nullptr; h->pipe.g_rgb = nullptr; h->pipe.g_n = 0; h->pipe.g_chunk = 0;
}

// one net (ray net or point net) over `rows` input rows -> out [rows][nc.mlp_out]
static int launch_net(hr_handle* h, const hr_config& nc, const hr::MlpSimtPack& simt, const hr::MlpTcPack& tc, bool tc_ready,
                      const float* in, int64_t rows, float* out, cudaStream_t st) {
  cudaError_t e;
  if (nc.mlp_mode == HR_MLP_ZERO) {  // ZeroMLP (nlf/nets/mlp.py:29-30): x.new_zeros(N, out_channels)
    e = cudaMemsetAsync(out, 0, (size_t)rows * nc.mlp_out * sizeof(float), st);
    if (e != cudaSuccess) return fail("heads memset failed: %s", cudaGetErrorString(e));
    return 0;
  }
  if (nc.mlp_mode == HR_MLP_BF16X3_TC) {
    if (!tc_ready) return fail("hr_render: tensor-core pack missing");
    e = hr::launch_mlp_tc2(nc, tc, h->tma_encode, in, out, rows, h->num_sms, st);
  } else {
    e = hr::launch_mlp_simt(nc, simt, in, out, rows, h->num_sms, st);
  }
  if (e != cudaSuccess) return fail("sample-net launch failed: %s", cudaGetErrorString(e));
  h->launches += 1;
  return 0;
}

// rays [n, c_in] -> heads scratch [n, mlp_out] (channel-major per ray).  `scratch` (cascade_bytes(h, n), only read for a
// cascaded pipeline) holds the first stage's intermediates.
static int launch_sample_net(hr_handle* h, const float* rays, int64_t n, float* heads, cudaStream_t st, void* scratch = nullptr) {
  const hr_config& c = h->cfg;
  if (!c.cascade) return launch_net(h, h->cfg_net, h->simt, h->tc, h->tc_ready, rays, n, heads, st);
  if (!scratch) return fail("hr_render: cascade scratch missing");
  // PointPredictionEmbedding (nlf/embedding/point.py:142-206): ray net -> S0 z-planes -> one point-net row per point
  auto seg = [](int64_t floats) { return (floats * (int64_t)sizeof(float) + 255) / 256 * 256; };  // as in cascade_bytes
  float* heads0 = (float*)scratch;                                                                // [n][S0*stride0] channel-major
  float* rows = (float*)((char*)heads0 + seg(n * c.pre_samples * c.pre_head_stride));              // [n*S0][8]
  float* out = (float*)((char*)rows + seg(n * c.pre_samples * 8));                                 // [n*S0][mlp_out/S0] = [n][S][stride]
  const bool has_pre = c.pre_mlp_mode != HR_MLP_ZERO;
  if (has_pre) {
    int rc = launch_net(h, h->cfg_pre, h->simt_pre, h->tc_pre, h->tc_pre_ready, rays, n, heads0, st);
    if (rc) return rc;
  }
  cascade_points_kernel<<<grid_for(n * 32), 256, 0, st>>>(c, rays, has_pre ? heads0 : nullptr, rows, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("cascade point kernel launch failed: %s", cudaGetErrorString(e));
  int rc = launch_net(h, h->cfg_net, h->simt, h->tc, h->tc_ready, rows, n * c.pre_samples, out, st);
  if (rc) return rc;
  permute_heads<<<grid_for(n * (long long)c.mlp_out), 256, 0, st>>>(out, heads, n, c.n_samples, c.head_stride);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("heads permutation launch failed: %s", cudaGetErrorString(e));
  h->launches += 2;
  return 0;
}

static hr::RgbDst one_dst(float* rgb) {
  hr::RgbDst d{};
  d.p[0] = rgb;
  d.n = 1;
  d.row0 = 0;
  return d;
}

static int render_impl(hr_handle* h, const float* rays, int64_t n, float* rgb, float* mlp_out, const hr::ExtraOut* so,
                       void* workspace, int64_t ws_bytes, cudaStream_t st, unsigned char* rgb8 = nullptr,
                       const hr::RgbDst* scatter = nullptr) {
  if (!h) return fail("hr_render: null handle");
  if (!h->uploaded) return fail("hr_render: parameters not uploaded (call hr_upload)");
  if (n == 0) return 0;
  if (!rays || (!rgb && !rgb8 && !scatter) || !workspace) return fail("hr_render: null buffer");
  if (ws_bytes < hr_workspace_bytes(h, n)) return fail("hr_render: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)hr_workspace_bytes(h, n));
  if (((uintptr_t)workspace & 15) != 0) return fail("hr_render: workspace must be 16-byte aligned");
  float* heads = (float*)workspace;
  const hr_config& c = h->cfg;
  const bool timing = h->timing && h->ev_render.size() < 8192;
  const int64_t sub = (h->sub_rays < 0) ? n : sub_batch_rays(h);
  for (int64_t off = 0; off < n; off += sub) {
    const int64_t m = (n - off < sub) ? (n - off) : sub;
    const float* r = rays + off * c.c_in;
    EventPair em{nullptr, nullptr}, er{nullptr, nullptr};
    if (timing) {
      CK(cudaEventCreate(&em.a)); CK(cudaEventCreate(&em.b)); CK(cudaEventCreate(&er.a)); CK(cudaEventCreate(&er.b));
      CK(cudaEventRecord(em.a, st));
    }
    int rc0 = launch_sample_net(h, r, m, heads, st, (char*)workspace + heads_bytes(h, (n < sub) ? n : sub));
    if (rc0) return rc0;
    if (timing) { CK(cudaEventRecord(em.b, st)); CK(cudaEventRecord(er.a, st)); }
    hr::RgbDst d = scatter ? *scatter : one_dst(rgb);
    d.row0 += off;
    hr::ExtraOut so_off;
    if (so) {  // every extra output is indexed by ray: advance the pointers to this sub-batch
      so_off = *so;
      const int64_t S = c.n_samples;
      if (so_off.distances) so_off.distances += off * S;
      if (so_off.points) so_off.points += off * S * 3;
      if (so_off.sigma) so_off.sigma += off * S;
      if (so_off.weights) so_off.weights += off * S;
      if (so_off.rgb_samples) so_off.rgb_samples += off * S * 3;
      for (int f = 0; f < HR_N_FIELDS; ++f) {
        if (!so_off.field_out[f]) continue;
        const bool three = f == HR_FIELD_POINTS || f == HR_FIELD_VIEWDIRS || f == HR_FIELD_COLOR_SCALE || f == HR_FIELD_COLOR_SHIFT ||
                           f == HR_FIELD_SPATIAL_FLOW || f == HR_FIELD_POINT_OFFSET || f == HR_FIELD_COLOR_SCALE_GLOBAL ||
                           f == HR_FIELD_COLOR_SHIFT_GLOBAL;
        so_off.field_out[f] += off * (three ? 3 : 1) * (so_off.field_mode[f] == HR_FIELD_NO_OVER ? S : 1);
      }
    }
    cudaError_t e = hr::launch_render(c, h->dv, h->tabs, r, heads, d, m, so ? &so_off : nullptr, h->num_sms, st, rgb8 ? rgb8 + off * 3 : nullptr);
    if (e != cudaSuccess) return fail("render launch failed: %s", cudaGetErrorString(e));
    if (timing) {
      CK(cudaEventRecord(er.b, st));
      h->ev_mlp.push_back(em);
      h->ev_render.push_back(er);
    }
    h->launches += 1;
    if (mlp_out) {
      unpermute_heads<<<grid_for(m * (long long)c.mlp_out), 256, 0, st>>>(heads, mlp_out + off * c.mlp_out, m, c.n_samples, c.head_stride);
      e = cudaGetLastError();
      if (e != cudaSuccess) return fail("unpermute launch failed: %s", cudaGetErrorString(e));
      h->launches += 1;
    }
  }
  if (timing) h->timed_calls += 1;
  return 0;
}

int hr_render(hr_handle* h, const float* rays, int64_t n_rays, float* rgb, void* workspace, int64_t workspace_bytes,
              void* stream) {
  DeviceGuard guard(h ? h->device : 0);
  return render_impl(h, rays, n_rays, rgb, nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

int hr_render_scatter(hr_handle* h, const float* rays, int64_t n_rays, float* const* dst, int32_t n_dst, int64_t row0,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h) return fail("hr_render_scatter: null handle");
  if (!dst || n_dst < 1 || n_dst > HR_MAX_PEERS) return fail("hr_render_scatter: 1..%d destination buffers", HR_MAX_PEERS);
  if (row0 < 0) return fail("hr_render_scatter: negative row offset");
  DeviceGuard guard(h->device);
  hr::RgbDst d{};
  for (int i = 0; i < n_dst; ++i) {
    if (!dst[i]) return fail("hr_render_scatter: null destination %d", i);
    d.p[i] = dst[i];
  }
  d.n = n_dst;
  d.row0 = row0;
  return render_impl(h, rays, n_rays, nullptr, nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream, nullptr, &d);
}

int hr_render_stages(hr_handle* h, const float* rays, int64_t n_rays, float* rgb, float* mlp_out, float* distances,
                     float* points, float* sigma, float* weights, float* rgb_samples, void* workspace, int64_t workspace_bytes,
                     void* stream) {
  DeviceGuard guard(h ? h->device : 0);
  hr::ExtraOut so{};
  so.distances = distances; so.points = points; so.sigma = sigma; so.weights = weights; so.rgb_samples = rgb_samples;
  return render_impl(h, rays, n_rays, rgb, mlp_out, &so, workspace, workspace_bytes, (cudaStream_t)stream);
}

int hr_render_fields(hr_handle* h, const float* rays, int64_t n_rays, float* rgb, float* render_weights,
                     const hr_field_request* req, int32_t n_req, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h) return fail("hr_render_fields: null handle");
  if (n_req < 0 || (n_req > 0 && !req)) return fail("hr_render_fields: bad request list");
  DeviceGuard guard(h->device);
  const hr_config& c = h->cfg;
  hr::ExtraOut so{};
  so.weights = render_weights;
  for (int i = 0; i < n_req; ++i) {
    const int f = req[i].field, m = req[i].mode;
    if (f < 0 || f >= HR_N_FIELDS) return fail("hr_render_fields: unknown field %d", f);
    if (m != HR_FIELD_OVER && m != HR_FIELD_NO_OVER && m != HR_FIELD_PRED_WEIGHTS) return fail("hr_render_fields: unknown mode %d", m);
    if (!req[i].out) return fail("hr_render_fields: null output for field %d", f);
    if (so.field_out[f]) return fail("hr_render_fields: field %d requested twice", f);
    // fields the pipeline does not carry (reference: KeyError on x[key])
    if ((f == HR_FIELD_BASE_TIMES || f == HR_FIELD_TIME_OFFSET) && !(c.dynamic || c.use_flow))
      return fail("hr_render_fields: this pipeline has no keyframe times");
    const int head_off[HR_N_FIELDS] = {0, 0, 0, 0, 0, 0, 0, c.off_cscale, c.off_cshift, c.off_flow, c.off_sigma,
                                       c.off_point_sigma, c.off_offset, c.off_cscale_global, c.off_cshift_global};
    if (f >= HR_FIELD_COLOR_SCALE && head_off[f] < 0) return fail("hr_render_fields: the sample net has no head for field %d", f);
    so.field_out[f] = req[i].out;
    so.field_mode[f] = m;
  }
  return render_impl(h, rays, n_rays, rgb, nullptr, &so, workspace, workspace_bytes, (cudaStream_t)stream);
}

int hr_render_to8b(hr_handle* h, const float* rays, int64_t n_rays, uint8_t* rgb8, void* workspace, int64_t workspace_bytes,
                   void* stream) {
  DeviceGuard guard(h ? h->device : 0);
  if (!rgb8) return fail("hr_render_to8b: null output");
  return render_impl(h, rays, n_rays, nullptr, nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream, rgb8);
}

int hr_generate_rays(const hr_camera* cam, int32_t c_in, int64_t first_pixel, int64_t n_pixels, float* rays_out, void* stream) {
  if (!cam || !rays_out) return fail("hr_generate_rays: null argument");
  if (c_in != 6 && c_in != 8) return fail("hr_generate_rays: c_in must be 6 or 8");
  if (cam->width < 1 || cam->height < 1) return fail("hr_generate_rays: bad image size");
  if (first_pixel < 0 || n_pixels < 0 || first_pixel + n_pixels > (int64_t)cam->width * cam->height)
    return fail("hr_generate_rays: pixel range outside the image");
  cudaError_t e = hr::launch_generate_rays(*cam, c_in, first_pixel, n_pixels, rays_out, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail("hr_generate_rays: %s", cudaGetErrorString(e));
  return 0;
}

int hr_render_frame_to8b_host(hr_handle* h, const hr_camera* cam, uint8_t* rgb8_host, int64_t chunk) {
  if (!h || !cam || !rgb8_host) return fail("hr_render_frame_to8b_host: null argument");
  if (!h->uploaded) return fail("hr_render_frame_to8b_host: parameters not uploaded");
  DeviceGuard guard(h->device);
  const hr_config& c = h->cfg;
  const int64_t n_rays = (int64_t)cam->width * cam->height;
  if (chunk <= 0) chunk = (h->cfg.mlp_mode == HR_MLP_BF16X3_TC) ? (int64_t)h->num_sms * 128 * 14 : 262144;  // whole tile waves
  if (chunk > n_rays) chunk = n_rays;
  HostPipe& P = h->pipe;
  // device scratch per slot: rays, 8-bit tile (stored in the rgb slot), workspace
  if (P.chunk < chunk) {
    drop_host_graph(h);
    for (int i = 0; i < 3; ++i) {
      if (P.d_rays[i]) cudaFree(P.d_rays[i]);
      if (P.d_rgb[i]) cudaFree(P.d_rgb[i]);
      if (P.d_ws[i]) cudaFree(P.d_ws[i]);
      P.d_rays[i] = P.d_rgb[i] = nullptr; P.d_ws[i] = nullptr;
      if (!P.streams[i]) CK(cudaStreamCreateWithFlags(&P.streams[i], cudaStreamNonBlocking));
    }
    P.ws_bytes = ws_bytes_for(h, chunk);
    for (int i = 0; i < 3; ++i) {
      CK(cudaMalloc((void**)&P.d_rays[i], (size_t)chunk * c.c_in * sizeof(float)));
      CK(cudaMalloc((void**)&P.d_rgb[i], (size_t)chunk * 3 * sizeof(float)));
      CK(cudaMalloc(&P.d_ws[i], (size_t)P.ws_bytes));
    }
    P.chunk = chunk;
  }
  int slot = 0;
  for (int64_t off = 0; off < n_rays; off += chunk, slot = (slot + 1) % 3) {
    const int64_t m = (n_rays - off < chunk) ? (n_rays - off) : chunk;
    cudaStream_t st = P.streams[slot];
    cudaError_t e = hr::launch_generate_rays(*cam, c.c_in, off, m, P.d_rays[slot], st);
    if (e != cudaSuccess) return fail("ray generation failed: %s", cudaGetErrorString(e));
    h->launches += 1;
    int rc = render_impl(h, P.d_rays[slot], m, nullptr, nullptr, nullptr, P.d_ws[slot], P.ws_bytes, st, (unsigned char*)P.d_rgb[slot]);
    if (rc) return rc;
    CK(cudaMemcpyAsync(rgb8_host + off * 3, P.d_rgb[slot], (size_t)m * 3, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < 3; ++i) CK(cudaStreamSynchronize(P.streams[i]));
  return 0;
}

int hr_render_host(hr_handle* h, const float* rays_host, int64_t n_rays, float* rgb_host, int64_t chunk) {
  if (!h) return fail("hr_render_host: null handle");
  if (!h->uploaded) return fail("hr_render_host: parameters not uploaded");
  if (n_rays == 0) return 0;
  if (!rays_host || !rgb_host) return fail("hr_render_host: null buffer");
  DeviceGuard guard(h->device);
  const hr_config& c = h->cfg;
  HostPipe& P = h->pipe;
  const int64_t wave = (int64_t)h->num_sms * 128;  // one full wave of 128-ray tiles of the tensor-core sample net
  // Default (chunk <= 0), tensor-core net, batches of a few waves: the "wave split" pipeline below.  Otherwise chunks of
  // `chunk` rays (default: whole waves for the tensor-core net, 32 768 rays for the CUDA-core net) on three streams.
  // (a cascaded pipeline runs its nets on point rows: it takes the plain chunked pipeline)
  const bool whole = chunk <= 0 && c.mlp_mode == HR_MLP_BF16X3_TC && n_rays <= 16 * wave && !c.cascade;  // the batch stays whole on the device
  const float* rays_dev_view = nullptr;
  if (whole && h->tc_ready && !h->timing) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, rays_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr)
      rays_dev_view = (const float*)pa.devicePointer;
    else
      cudaGetLastError();  // pageable memory: not an error, just not device-addressable
  }
  const bool zero_copy = rays_dev_view != nullptr;
  float* rgb_dev_view = nullptr;
  if (zero_copy) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, rgb_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr)
      rgb_dev_view = (float*)pa.devicePointer;
    else
      cudaGetLastError();
  }
  const bool split = !zero_copy && whole && n_rays > wave;
  if (chunk <= 0) chunk = (c.mlp_mode == HR_MLP_BF16X3_TC) ? wave : 32768;
  if (chunk > n_rays) chunk = n_rays;
  const int64_t alloc = (split || zero_copy) ? n_rays : chunk;  // rays per device slot
  if (P.chunk < alloc) {
    drop_host_graph(h);
    for (int i = 0; i < 3; ++i) {
      if (P.d_rays[i]) cudaFree(P.d_rays[i]);
      if (P.d_rgb[i]) cudaFree(P.d_rgb[i]);
      if (P.d_ws[i]) cudaFree(P.d_ws[i]);
      P.d_rays[i] = P.d_rgb[i] = nullptr; P.d_ws[i] = nullptr;
      if (!P.streams[i]) CK(cudaStreamCreateWithFlags(&P.streams[i], cudaStreamNonBlocking));
    }
    P.ws_bytes = ws_bytes_for(h, alloc);
    for (int i = 0; i < 3; ++i) {
      CK(cudaMalloc((void**)&P.d_rays[i], (size_t)alloc * c.c_in * sizeof(float)));
      CK(cudaMalloc((void**)&P.d_rgb[i], (size_t)alloc * 3 * sizeof(float)));
      CK(cudaMalloc(&P.d_ws[i], (size_t)P.ws_bytes));
    }
    P.chunk = alloc;
  }
  if (!P.fork_ev) {
    CK(cudaEventCreateWithFlags(&P.fork_ev, cudaEventDisableTiming));
    for (int i = 0; i < 3; ++i) CK(cudaEventCreateWithFlags(&P.join_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 8; ++i) CK(cudaEventCreateWithFlags(&P.dep_ev[i], cudaEventDisableTiming));
  }
  // Chunked pipeline: chunk i runs H2D -> sample net -> render -> D2H on stream i % 3.
  auto enqueue_chunks = [&]() -> int {
    int slot = 0;
    for (int64_t off = 0; off < n_rays; off += chunk, slot = (slot + 1) % 3) {
      int64_t m = (n_rays - off < chunk) ? (n_rays - off) : chunk;
      cudaStream_t st = P.streams[slot];
      CK(cudaMemcpyAsync(P.d_rays[slot], rays_host + off * c.c_in, (size_t)m * c.c_in * sizeof(float), cudaMemcpyHostToDevice, st));
      int rc = render_impl(h, P.d_rays[slot], m, P.d_rgb[slot], nullptr, nullptr, P.d_ws[slot], P.ws_bytes, st);
      if (rc) return rc;
      CK(cudaMemcpyAsync(rgb_host + off * 3, P.d_rgb[slot], (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    return 0;
  };
  // Wave-split pipeline: splitting a batch into independent chunks costs one cold sample-net launch and one render tail
  // per chunk, which eats what the copy overlap wins (measured: 0.404 ms unsplit vs 0.410 ms in four chunks).  Instead the
  // batch stays whole on the device and only the edges are split:
  //   copy-in  (stream 1): rays of the first tile wave, then the rest          -> events A, B
  //   compute  (stream 0): sample net on the first wave after A, on the rest after B (together exactly the waves of the
  //                        unsplit batch), then the render kernel in four pieces -> events R0..R3
  //   copy-out (stream 2): rgb of piece j after Rj
  // so only the first wave's H2D and the last piece's D2H are exposed.
  auto enqueue_split = [&]() -> int {
    cudaStream_t s0 = P.streams[0], s1 = P.streams[1], s2 = P.streams[2];
    float* d_rays = P.d_rays[0];
    float* d_rgb = P.d_rgb[0];
    float* heads = (float*)P.d_ws[0];
    const int64_t nA = wave, nB = n_rays - wave;
    CK(cudaMemcpyAsync(d_rays, rays_host, (size_t)nA * c.c_in * sizeof(float), cudaMemcpyHostToDevice, s1));
    CK(cudaEventRecord(P.dep_ev[0], s1));
    CK(cudaMemcpyAsync(d_rays + nA * c.c_in, rays_host + nA * c.c_in, (size_t)nB * c.c_in * sizeof(float), cudaMemcpyHostToDevice, s1));
    CK(cudaEventRecord(P.dep_ev[1], s1));
    CK(cudaStreamWaitEvent(s0, P.dep_ev[0], 0));
    int rc = launch_sample_net(h, d_rays, nA, heads, s0);
    if (rc) return rc;
    CK(cudaStreamWaitEvent(s0, P.dep_ev[1], 0));
    rc = launch_sample_net(h, d_rays + nA * c.c_in, nB, heads + nA * (int64_t)c.mlp_out, s0);
    if (rc) return rc;
    const int pieces = 4;
    const int64_t per = ((n_rays + pieces - 1) / pieces + 255) / 256 * 256;
    int j = 0;
    for (int64_t off = 0; off < n_rays; off += per, ++j) {
      const int64_t m = (n_rays - off < per) ? (n_rays - off) : per;
      cudaError_t e = hr::launch_render(c, h->dv, h->tabs, d_rays + off * c.c_in, heads + off * (int64_t)c.mlp_out, one_dst(d_rgb + off * 3), m,
                                        nullptr, h->num_sms, s0, nullptr);
      if (e != cudaSuccess) return fail("render launch failed: %s", cudaGetErrorString(e));
      h->launches += 1;
      CK(cudaEventRecord(P.dep_ev[2 + j], s0));
      CK(cudaStreamWaitEvent(s2, P.dep_ev[2 + j], 0));
      CK(cudaMemcpyAsync(rgb_host + off * 3, d_rgb + off * 3, (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, s2));
    }
    return 0;
  };
  // Zero-copy input: when the caller's rays are pinned (device-addressable) host memory and the second tensor-core layout
  // is in use, the sample net reads them straight over PCIe -- its two encoder warps work one tile ahead of the tensor
  // pipe, so the transfer hides under the math without splitting any launch -- and leaves a device copy for the render
  // kernel.  The render kernel then runs in two pieces so that half of the D2H overlaps it.
  auto enqueue_zero_copy = [&]() -> int {
    cudaStream_t s0 = P.streams[0], s2 = P.streams[2];
    float* d_rays = P.d_rays[0];
    float* d_rgb = P.d_rgb[0];
    float* heads = (float*)P.d_ws[0];
    cudaError_t e = hr::launch_mlp_tc2(c, h->tc, h->tma_encode, rays_dev_view, heads, n_rays, h->num_sms, s0, d_rays);
    if (e != cudaSuccess) return fail("sample-net launch failed: %s", cudaGetErrorString(e));
    h->launches += 1;
    if (rgb_dev_view != nullptr) {
      // Zero-copy output: the caller's rgb buffer is device-addressable pinned memory too -- the render kernel's epilogue
      // stores the pixels straight into it (12 bytes per ray as posted writes over PCIe, spread over the kernel's whole run),
      // so there is no D2H copy and no reason to split the launch.  Completion of the stream makes the writes visible.
      e = hr::launch_render(c, h->dv, h->tabs, d_rays, heads, one_dst(rgb_dev_view), n_rays, nullptr, h->num_sms, s0, nullptr);
      if (e != cudaSuccess) return fail("render launch failed: %s", cudaGetErrorString(e));
      h->launches += 1;
      return 0;
    }
    const int pieces = 2;
    const int64_t per = ((n_rays + pieces - 1) / pieces + 255) / 256 * 256;
    int j = 0;
    for (int64_t off = 0; off < n_rays; off += per, ++j) {
      const int64_t m = (n_rays - off < per) ? (n_rays - off) : per;
      e = hr::launch_render(c, h->dv, h->tabs, d_rays + off * c.c_in, heads + off * (int64_t)c.mlp_out, one_dst(d_rgb + off * 3), m, nullptr,
                            h->num_sms, s0, nullptr);
      if (e != cudaSuccess) return fail("render launch failed: %s", cudaGetErrorString(e));
      h->launches += 1;
      CK(cudaEventRecord(P.dep_ev[2 + j], s0));
      CK(cudaStreamWaitEvent(s2, P.dep_ev[2 + j], 0));
      CK(cudaMemcpyAsync(rgb_host + off * 3, d_rgb + off * 3, (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, s2));
    }
    return 0;
  };
  auto enqueue = [&]() -> int { return zero_copy ? enqueue_zero_copy() : (split ? enqueue_split() : enqueue_chunks()); };
  const int64_t n_chunks = (n_rays + chunk - 1) / chunk;
  const int64_t key_chunk = zero_copy ? -2 : (split ? -1 : chunk);
  if (!h->timing && (zero_copy || split || n_chunks > 1)) {
    // Launch-bound when issued call by call: capture the whole multi-stream pipeline once per (buffers, size) signature
    // and replay it with a single graph launch.
    if (!(P.graph && P.g_rays == rays_host && P.g_rgb == rgb_host && P.g_n == n_rays && P.g_chunk == key_chunk)) {
      drop_host_graph(h);
      const int64_t launches_before = h->launches;
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(P.streams[0], cudaStreamCaptureModeRelaxed));
      CK(cudaEventRecord(P.fork_ev, P.streams[0]));
      for (int i = 1; i < 3; ++i) CK(cudaStreamWaitEvent(P.streams[i], P.fork_ev, 0));
      int rc = enqueue();
      for (int i = 1; i < 3; ++i) {
        cudaEventRecord(P.join_ev[i], P.streams[i]);
        cudaStreamWaitEvent(P.streams[0], P.join_ev[i], 0);
      }
      cudaError_t ce = cudaStreamEndCapture(P.streams[0], &g);
      P.g_launches = h->launches - launches_before;
      h->launches = launches_before;  // capture enqueued nothing; the replay below is what runs
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (ce != cudaSuccess) return fail("hr_render_host: graph capture failed: %s", cudaGetErrorString(ce));
      ce = cudaGraphInstantiate(&P.graph, g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) { P.graph = nullptr; return fail("hr_render_host: graph instantiate failed: %s", cudaGetErrorString(ce)); }
      P.g_rays = rays_host; P.g_rgb = rgb_host; P.g_n = n_rays; P.g_chunk = key_chunk;
    }
    CK(cudaGraphLaunch(P.graph, P.streams[0]));
    h->launches += P.g_launches;
    CK(cudaStreamSynchronize(P.streams[0]));
    return 0;
  }
  int rc = enqueue();
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) CK(cudaStreamSynchronize(P.streams[i]));
  return 0;
}

// ---- backward pass ----
static int train_supported(const hr_config& c) {
  if (c.isect_type == HR_ISECT_SPHERE_NEW) return fail("backward: the sphere_new primitive is not supported yet");
  if (c.mlp_mode == HR_MLP_ZERO) { /* no sample net: d heads is simply unused by the caller */ }
  if ((c.isect_type == HR_ISECT_SPHERE || c.isect_type == HR_ISECT_CYLINDER) && c.sphere_origin_scale != 0.0f)
    return fail("backward: learned primitive origins (origin_scale_factor != 0) are not supported yet");
  if (c.contract_type == HR_CONTRACT_AFFINE) return fail("backward: bbox / z_depth contraction is not supported yet");
  if (c.off_cscale_global >= 0) return fail("backward: per-ray colour heads are not supported yet");
  if (c.isect_type == HR_ISECT_VOXEL || c.isect_type == HR_ISECT_PLANE) return fail("backward: voxel-grid primitives are not supported yet");
  if (c.n_color_views > 0) return fail("backward: the per-camera colour transform is not supported yet");
  if (c.n_samples > 64) return fail("backward: more than 64 samples per ray are not supported yet");
  if (c.cascade) return fail("backward: cascaded (point_prediction) pipelines are not supported yet");
  return 0;
}

static int ensure_grad_tables(hr_handle* h, cudaStream_t st) {
  const hr_config& c = h->cfg;
  size_t want[13];
  for (int i = 0; i < 3; ++i) {
    const hr::PlaneTab& t = h->tabs.sig[i];
    const size_t sp = (size_t)t.C * t.H * t.W, se = (size_t)t.C * t.H2 * t.L;
    want[i] = sp; want[3 + i] = se; want[6 + i] = sp; want[9 + i] = se;
  }
  want[12] = (size_t)c.app_dim * h->tabs.n_app_total;
  float** bufs[13] = {&h->g_sig_space[0], &h->g_sig_space[1], &h->g_sig_space[2], &h->g_sig_second[0], &h->g_sig_second[1],
                      &h->g_sig_second[2], &h->g_app_space[0], &h->g_app_space[1], &h->g_app_space[2], &h->g_app_second[0],
                      &h->g_app_second[1], &h->g_app_second[2], &h->g_basis};
  for (int i = 0; i < 13; ++i) {
    if (h->g_sizes[i] == want[i] && (*bufs[i] || want[i] == 0)) continue;
    if (*bufs[i]) cudaFree(*bufs[i]);
    *bufs[i] = nullptr;
    h->g_sizes[i] = 0;
    if (want[i] == 0) continue;
    cudaError_t e = cudaMalloc((void**)bufs[i], want[i] * sizeof(float));
    if (e != cudaSuccess) return fail("cudaMalloc(gradient table %zu floats): %s", want[i], cudaGetErrorString(e));
    e = cudaMemsetAsync(*bufs[i], 0, want[i] * sizeof(float), st);
    if (e != cudaSuccess) return fail("cudaMemsetAsync: %s", cudaGetErrorString(e));
    h->g_sizes[i] = want[i];
  }
  return 0;
}

int hr_encode_rays(hr_handle* h, const float* rays, int64_t n_rays, float* enc, void* stream) {
  if (!h || !rays || !enc) return fail("hr_encode_rays: null argument");
  if (h->cfg.cascade) return fail("hr_encode_rays: a cascaded pipeline has two nets; its training path is not built");
  if (n_rays == 0) return 0;
  DeviceGuard guard(h->device);
  encode_rays_kernel<<<grid_for(n_rays), 256, 0, (cudaStream_t)stream>>>(h->cfg, rays, enc, n_rays);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("hr_encode_rays: %s", cudaGetErrorString(e));
  h->launches += 1;
  return 0;
}

int hr_render_heads(hr_handle* h, const float* rays, const float* heads, int64_t n, float* rgb, const hr_train_opts* opts,
                    void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !rays || !heads || !rgb || !opts || !workspace) return fail("hr_render_heads: null argument");
  if (!h->uploaded) return fail("hr_render_heads: parameters not uploaded");
  if (n == 0) return 0;
  if (workspace_bytes < heads_bytes(h, n)) return fail("hr_render_heads: workspace too small");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  hr_config c = h->cfg;
  c.clamp_output = opts->clamp_output ? 1 : 0;
  c.white_bg = opts->white_bg ? 1 : 0;
  if (opts->white_bg) c.black_bg = 0;
  float* hcm = (float*)workspace;
  permute_heads<<<grid_for(n * (long long)c.mlp_out), 256, 0, st>>>(heads, hcm, n, c.n_samples, c.head_stride);
  cudaError_t e = hr::launch_render(c, h->dv, h->tabs, rays, hcm, one_dst(rgb), n, nullptr, h->num_sms, st, nullptr);
  if (e != cudaSuccess) return fail("hr_render_heads: %s", cudaGetErrorString(e));
  h->launches += 2;
  return 0;
}

int hr_render_backward(hr_handle* h, const float* rays, const float* heads, int64_t n, const float* d_rgb, float* d_heads,
                       const hr_train_opts* opts, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !rays || !heads || !d_rgb || !d_heads || !opts || !workspace) return fail("hr_render_backward: null argument");
  if (!h->uploaded) return fail("hr_render_backward: parameters not uploaded");
  if (train_supported(h->cfg)) return 1;
  if (n == 0) return 0;
  if (workspace_bytes < 2 * heads_bytes(h, n)) return fail("hr_render_backward: workspace too small (hr_train_workspace_bytes)");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const hr_config& c = h->cfg;
  if (ensure_grad_tables(h, st)) return 1;
  float* hcm = (float*)workspace;
  float* gcm = (float*)((char*)workspace + heads_bytes(h, n));
  permute_heads<<<grid_for(n * (long long)c.mlp_out), 256, 0, st>>>(heads, hcm, n, c.n_samples, c.head_stride);
  const int white = opts->white_bg ? 1 : 0;
  EventPair eb{nullptr, nullptr};
  const bool timing = h->timing && h->ev_bwd.size() < 8192;
  if (timing) { CK(cudaEventCreate(&eb.a)); CK(cudaEventCreate(&eb.b)); CK(cudaEventRecord(eb.a, st)); }
  cudaError_t e = hr::launch_render_bwd(c, h->dv, h->tabs, h->g_sig_space, h->g_sig_second, h->g_app_space, h->g_app_second, h->g_basis,
                                        rays, hcm, d_rgb, gcm, n, opts->clamp_output ? 1 : 0, white, h->num_sms, st);
  if (e != cudaSuccess) return fail("hr_render_backward: %s", cudaGetErrorString(e));
  if (timing) { CK(cudaEventRecord(eb.b, st)); h->ev_bwd.push_back(eb); }
  unpermute_heads<<<grid_for(n * (long long)c.mlp_out), 256, 0, st>>>(gcm, d_heads, n, c.n_samples, c.head_stride);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("hr_render_backward: %s", cudaGetErrorString(e));
  h->launches += 3;
  return 0;
}

int hr_grad_zero(hr_handle* h, void* stream) {
  if (!h) return fail("hr_grad_zero: null handle");
  if (!h->uploaded) return fail("hr_grad_zero: parameters not uploaded");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (ensure_grad_tables(h, st)) return 1;
  float* bufs[13] = {h->g_sig_space[0], h->g_sig_space[1], h->g_sig_space[2], h->g_sig_second[0], h->g_sig_second[1], h->g_sig_second[2],
                     h->g_app_space[0], h->g_app_space[1], h->g_app_space[2], h->g_app_second[0], h->g_app_second[1], h->g_app_second[2],
                     h->g_basis};
  for (int i = 0; i < 13; ++i)
    if (bufs[i]) CK(cudaMemsetAsync(bufs[i], 0, h->g_sizes[i] * sizeof(float), st));
  return 0;
}

int hr_grad_read(hr_handle* h, const hr_grads* out, void* stream) {
  if (!h || !out) return fail("hr_grad_read: null argument");
  if (!h->uploaded) return fail("hr_grad_read: parameters not uploaded");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (ensure_grad_tables(h, st)) return 1;
  const hr_config& c = h->cfg;
  for (int i = 0; i < 3; ++i) {
    const hr::PlaneTab& t = h->tabs.sig[i];
    if (t.C == 0) continue;
    for (int f = 0; f < 2; ++f) {
      float* dsp = f ? out->app_plane[i] : out->sigma_plane[i];
      float* dse = f ? out->app_second[i] : out->sigma_second[i];
      const float* gsp = f ? h->g_app_space[i] : h->g_sig_space[i];
      const float* gse = f ? h->g_app_second[i] : h->g_sig_second[i];
      if (dsp) unpack_channel_last<<<grid_for((long long)t.C * t.H * t.W), 256, 0, st>>>(gsp, dsp, t.C, t.H, t.W);
      if (dse) {
        if (c.dynamic)
          unblend_time_lines<<<grid_for((long long)t.C * t.H2 * t.L), 256, 0, st>>>(gse, dse, t.C, t.H2, t.L, h->dv.time_inv_fac,
                                                                                   h->dv.time_scale, h->dv.time_offset);
        else
          unpack_channel_last<<<grid_for((long long)t.C * t.H2 * t.L), 256, 0, st>>>(gse, dse, t.C, t.H2, t.L);
      }
    }
  }
  if (out->basis_mat) CK(cudaMemcpyAsync(out->basis_mat, h->g_basis, h->g_sizes[12] * sizeof(float), cudaMemcpyDeviceToDevice, st));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("hr_grad_read: %s", cudaGetErrorString(e));
  return 0;
}

int64_t hr_launch_count(const hr_handle* h) { return h ? h->launches : -1; }

int hr_timing_enable(hr_handle* h, int enable) {
  if (!h) return fail("null handle");
  h->timing = enable != 0;
  return 0;
}

static void drop_events(std::vector<EventPair>& v) {
  for (auto& p : v) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  v.clear();
}

int hr_timing_reset(hr_handle* h) {
  if (!h) return fail("null handle");
  drop_events(h->ev_render);
  drop_events(h->ev_mlp);
  drop_events(h->ev_bwd);
  h->timed_calls = 0;
  return 0;
}

int hr_timing_read_backward(hr_handle* h, double* backward_ms_avg, int64_t* launches) {
  if (!h) return fail("null handle");
  double sb = 0;
  for (auto& p : h->ev_bwd) {
    CK(cudaEventSynchronize(p.b));
    float ms = 0; CK(cudaEventElapsedTime(&ms, p.a, p.b)); sb += ms;
  }
  const size_t k = h->ev_bwd.size();
  if (backward_ms_avg) *backward_ms_avg = k ? sb / k : 0.0;
  if (launches) *launches = (int64_t)k;
  return 0;
}

int hr_timing_read(hr_handle* h, double* render_ms_avg, double* mlp_ms_avg, int64_t* launches) {
  if (!h) return fail("null handle");
  double sr = 0, sm = 0;
  for (auto& p : h->ev_render) {
    CK(cudaEventSynchronize(p.b));
    float ms = 0; CK(cudaEventElapsedTime(&ms, p.a, p.b)); sr += ms;
  }
  for (auto& p : h->ev_mlp) {
    CK(cudaEventSynchronize(p.b));
    float ms = 0; CK(cudaEventElapsedTime(&ms, p.a, p.b)); sm += ms;
  }
  // per hr_render call: a call may run several sub-batches, i.e. several launches of each kernel
  const size_t k = h->timed_calls;
  if (render_ms_avg) *render_ms_avg = k ? sr / k : 0.0;
  if (mlp_ms_avg) *mlp_ms_avg = k ? sm / k : 0.0;
  if (launches) *launches = (int64_t)h->ev_render.size();
  return 0;
}

int hr_destroy(hr_handle* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  for (auto& sl : h->slots) cudaFree(sl.ptr);
  for (int i = 0; i < 3; ++i) {
    cudaFree(h->g_sig_space[i]); cudaFree(h->g_sig_second[i]); cudaFree(h->g_app_space[i]); cudaFree(h->g_app_second[i]);
  }
  cudaFree(h->g_basis);
  hr::free_mlp_tc2(h);
  drop_events(h->ev_render);
  drop_events(h->ev_mlp);
  drop_events(h->ev_bwd);
  for (int i = 0; i < 3; ++i) {
    if (i == 0) {
      drop_host_graph(h);
      if (h->pipe.fork_ev) cudaEventDestroy(h->pipe.fork_ev);
      for (int k = 0; k < 8; ++k)
        if (h->pipe.dep_ev[k]) cudaEventDestroy(h->pipe.dep_ev[k]);
    }
    if (h->pipe.join_ev[i]) cudaEventDestroy(h->pipe.join_ev[i]);
    if (h->pipe.d_rays[i]) cudaFree(h->pipe.d_rays[i]);
    if (h->pipe.d_rgb[i]) cudaFree(h->pipe.d_rgb[i]);
    if (h->pipe.d_ws[i]) cudaFree(h->pipe.d_ws[i]);
    if (h->pipe.streams[i]) cudaStreamDestroy(h->pipe.streams[i]);
  }
  delete h;
  return 0;
}

}  // extern "C"
