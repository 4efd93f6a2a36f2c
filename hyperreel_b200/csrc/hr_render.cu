// Fused per-ray render kernel: heads -> intersect -> sort -> points -> VM gather -> decode -> composite.
//
// One warp renders one ray.  Two lane mappings are used:
//   * "lane = sample"  (S <= 32*SPL samples, SPL registers per lane) for everything that is per-sample
//     scalar math: head activations, intersection, the bitonic sort of t, points, validity, alpha,
//     the transmittance scan (reference: nlf/intersect/base.py:142-259, nlf/embedding/point.py:780-831,
//     371-396, utils/tensorf_utils.py:242-253);
//   * "quad = sample"  (4 lanes per sample, 8 samples per round) for the VM gather: the 4 lanes of a
//     quad fetch the bilinear footprint of one sample as 16-byte slices of channel-last texels so that
//     the two x-neighbouring taps (64 contiguous bytes for C=8) are served by adjacent lanes of one
//     LDG.128 (reference: F.grid_sample calls in nlf/nets/tensorf_dynamic.py:287-371 and
//     nlf/nets/tensorf_no_sample.py:47-126).
// Nothing per-sample ever goes to HBM: rays (4*c_in B) + sample-net heads in, rgb (12 B) out.
#include "hr_common.cuh"

namespace hr {

static constexpr int kWarpsPerCta = 8;
static constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int C>
struct TapSet {
  float4 a, b;
  float w0, w1;
};

// Issue the loads of one bilinear footprint (reference: grid_sample, bilinear, zeros padding,
// align_corners=True; unnormalise ((g+1)/2)*(size-1)).  gx indexes W, gy indexes H.
// x0/y0 are clamped to [0, size-2] and the fraction recomputed, which is exact for in-range
// coordinates (the out-of-range neighbour of a point on the max face has weight 0 in the reference).
template <int C>
__device__ __forceinline__ void issue(TapSet<C>& t, const float* __restrict__ tab, int H, int W, float gx, float gy,
                                      int xt, int alt, bool pred) {
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
  int x0 = (int)floorf(ix);
  x0 = max(0, min(x0, W - 2));
  float fx = ix - (float)x0;
  int y0 = 0, y1 = 0;
  float fy = 0.0f;
  if (H > 1) {
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
    y0 = (int)floorf(iy);
    y0 = max(0, min(y0, H - 2));
    fy = iy - (float)y0;
    y1 = y0 + 1;
  }
  float wx = xt ? fx : 1.0f - fx;
  int x = min(x0 + xt, W - 1);
  t.a = make_float4(0.f, 0.f, 0.f, 0.f);
  t.b = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (C == 8) {
    t.w0 = wx * (1.0f - fy);
    t.w1 = wx * fy;
    if (pred) {
      t.a = ldg4(tab + ((size_t)y0 * W + x) * 8 + alt * 4);
      if (H > 1) t.b = ldg4(tab + ((size_t)y1 * W + x) * 8 + alt * 4);
    }
  } else {
    int y = alt ? y1 : y0;
    t.w0 = wx * (alt ? fy : 1.0f - fy);
    t.w1 = 0.0f;
    if (pred) t.a = ldg4(tab + ((size_t)y * W + x) * 4);
  }
}

// Finish the interpolation across the quad.  C=8: out = this lane's 4-channel half (alt selects
// channels 4*alt..4*alt+3).  C=4: out = all 4 channels, replicated on the 4 lanes.
template <int C>
__device__ __forceinline__ void combine(const TapSet<C>& t, float (&out)[4]) {
  if constexpr (C == 8) {
    out[0] = t.w0 * t.a.x + t.w1 * t.b.x;
    out[1] = t.w0 * t.a.y + t.w1 * t.b.y;
    out[2] = t.w0 * t.a.z + t.w1 * t.b.z;
    out[3] = t.w0 * t.a.w + t.w1 * t.b.w;
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] += __shfl_xor_sync(kFull, out[c], 2);
  } else {
    out[0] = t.w0 * t.a.x;
    out[1] = t.w0 * t.a.y;
    out[2] = t.w0 * t.a.z;
    out[3] = t.w0 * t.a.w;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      out[c] += __shfl_xor_sync(kFull, out[c], 1);
      out[c] += __shfl_xor_sync(kFull, out[c], 2);
    }
  }
}

// Ascending bitonic sort of 32*SPL keys, element e = reg*32 + lane (reference: torch.argsort +
// gather of the distances only, utils/intersect_utils.py:12-16; ties are equal values).
template <int SPL>
__device__ __forceinline__ void sort_keys(float (&k)[SPL], int lane) {
  constexpr int NE = 32 * SPL;
#pragma unroll
  for (int size = 2; size <= NE; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        // partner lives in the other register of the same lane (SPL == 2, stride == 32, size == 64)
        float lo = fminf(k[0], k[SPL - 1]), hi = fmaxf(k[0], k[SPL - 1]);
        k[0] = lo;
        k[SPL - 1] = hi;
      } else {
#pragma unroll
        for (int r = 0; r < SPL; ++r) {
          int e = r * 32 + lane;
          float other = __shfl_xor_sync(kFull, k[r], stride);
          bool up = ((e & size) == 0);
          bool lower = ((lane & stride) == 0);
          k[r] = (lower == up) ? fminf(k[r], other) : fmaxf(k[r], other);
        }
      }
    }
  }
}

// mipnerf inverse contraction of a scalar distance (reference: nlf/contract.py:143-158).
__device__ __forceinline__ float inv_contract_distance(const hr_config& cfg, const Derived& dv, float d) {
  d = __fmul_rn(__fmul_rn(d, 0.5f), 2.0f);  // distance_activation = identity: (d/2)*2
  d = fminf(fmaxf(d, -2.0f), 2.0f);
  float t = __fsub_rn(2.0f, fabsf(d));
  float inv = __fadd_rn(__fdiv_rn(t, dv.dist_scale_fac), dv.inv_end_dist);
  float sgn = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
  float far_v = __fmul_rn(sgn, __fdiv_rn(1.0f, inv));
  float v = (fabsf(d) < 1.0f) ? d : far_v;
  return __fmul_rn(v, cfg.contract_start_distance);
}

// mipnerf point contraction (reference: nlf/contract.py:178-192).
__device__ __forceinline__ void contract_point(const hr_config& cfg, const Derived& dv, float& x, float& y, float& z) {
  x = __fdiv_rn(x, cfg.contract_start_radius);
  y = __fdiv_rn(y, cfg.contract_start_radius);
  z = __fdiv_rn(z, cfg.contract_start_radius);
  float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  float inv = __fdiv_rn(1.0f, fabsf(dist));
  float t = __fmul_rn(__fsub_rn(inv, dv.inv_end_rad), dv.rad_scale_fac);
  if (!(dist < 1.0f)) {
    float s = __fsub_rn(2.0f, t);
    x = __fmul_rn(__fdiv_rn(x, dist), s);
    y = __fmul_rn(__fdiv_rn(y, dist), s);
    z = __fmul_rn(__fdiv_rn(z, dist), s);
  }
}

// Real SH basis, degree 2 (reference: utils/sh_utils.py:94-119).
__device__ __forceinline__ void sh_basis9(float x, float y, float z, float (&Y)[9]) {
  const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
  const float C20 = 1.0925484305920792f, C21 = -1.0925484305920792f, C22 = 0.31539156525252005f,
              C23 = -1.0925484305920792f, C24 = 0.5462742152960396f;
  float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  Y[0] = C0;
  Y[1] = -C1 * y;
  Y[2] = C1 * z;
  Y[3] = -C1 * x;
  Y[4] = C20 * xy;
  Y[5] = C21 * yz;
  Y[6] = C22 * (2.0f * zz - xx - yy);
  Y[7] = C23 * xz;
  Y[8] = C24 * (xx - yy);
}

__host__ __device__ inline int basis_block_stride(int rows, int nt) {
  int bs = rows * nt;
  while ((bs & 31) != 8) ++bs;
  return bs;
}

template <int SPL, bool DYN, int C0, int C1, int C2, int SHADE, bool STAGES>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
render_kernel(const __grid_constant__ hr_config cfg, const __grid_constant__ Derived dv,
              const __grid_constant__ RenderTabs tabs, const float* __restrict__ rays,
              const float* __restrict__ heads, float* __restrict__ rgb_out, long long n_rays, StageOut so) {
  constexpr int NT = C0 + C1 + C2;
  constexpr int ROWS = (SHADE == HR_SHADE_SH) ? 9 : 1;
  constexpr int ROUNDS = 4 * SPL;
  extern __shared__ float smem[];
  float* s_basis = smem;  // [3][BS]
  const int BS = basis_block_stride(ROWS, NT);
  for (int i = threadIdx.x; i < 3 * ROWS * NT; i += blockDim.x) {
    int ch = i / (ROWS * NT), rem = i % (ROWS * NT);
    s_basis[ch * BS + rem] = tabs.basis[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int q = lane & 3, xt = q >> 1, alt = q & 1, quad = lane >> 2;
  const int S = cfg.n_samples;
  const int out_stride = cfg.mlp_out;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerCta;

  for (long long ray = warp0; ray < n_rays; ray += nwarps) {
    const float* r = rays + ray * cfg.c_in;
    const float ox = __ldg(r + 0), oy = __ldg(r + 1), oz = __ldg(r + 2);
    const float dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
    const float time = __ldg(r + cfg.c_in - 1);
    const float* hrow = heads + ray * (long long)out_stride;

    // ---- per-ray keyframe snap (utils/flow_utils.py:18-31) and time coordinate ----
    float base_t = 0.0f, toff = 0.0f, tau = 0.0f;
    if (DYN || cfg.use_flow) {
      float tt = __fmul_rn(time, dv.time_fac);
      tt = fminf(fmaxf(tt, 0.0f), dv.kf_max);
      tt = rintf(__fsub_rn(tt, 1e-5f));
      base_t = __fmul_rn(tt, dv.time_inv_fac);
      toff = __fsub_rn(time, base_t);
      // normalize_time_coord (tensorf_dynamic.py:615-616)
      tau = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(base_t, dv.time_scale), dv.time_offset), 2.0f), 1.0f);
    }

    float tkey[SPL];
    float cs[SPL][3], csh[SPL][3], padd[SPL][3];  // colour scale/shift, total point displacement
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = lane + 32 * j;
      const bool act = s < S;
      auto H = [&](int c) -> float { return act ? __ldg(hrow + c * S + s) : 0.0f; };
      float sg = 0.0f, sgp = 0.0f;
      if (cfg.off_sigma >= 0) sg = apply_act(cfg.act_sigma, H(cfg.off_sigma));
      if (cfg.off_point_sigma >= 0) sgp = apply_act(cfg.act_point_sigma, H(cfg.off_point_sigma));
      auto density = [&](int off) -> float {
        return (off < 0) ? 0.0f : ((off == cfg.off_sigma) ? sg : sgp);
      };
      // ---- intersection (base.py:155-203) ----
      const float one_m = __fsub_rn(1.0f, cfg.isect_use_sigma ? density(cfg.isect_density_off) : 0.0f);
      float t;
      if (cfg.isect_type == HR_ISECT_Z_PLANE) {
        float zr = __fmul_rn(apply_act(cfg.isect_act, apply_act(cfg.act_z, H(cfg.off_z))), one_m);
        float z = __fadd_rn(__fmul_rn(zr, cfg.z_scale), cfg.samples[act ? s : 0]);
        if (cfg.contract_samples) z = inv_contract_distance(cfg, dv, z);
        float dzg = (fabsf(dz) < 1e-5f) ? 1e12f : dz;  // intersect_utils.py:135-142
        t = __fdiv_rn(__fsub_rn(z, oz), dzg);
      } else {
        float zc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          zc[c] = __fmul_rn(apply_act(cfg.isect_act, apply_act(cfg.act_z, H(cfg.off_z + c))), one_m);
        // primitive.py:410-418
        float gx = __fadd_rn(__fmul_rn(zc[0], cfg.sphere_origin_scale), cfg.sphere_origin_initial[0]);
        float gy = __fadd_rn(__fmul_rn(zc[1], cfg.sphere_origin_scale), cfg.sphere_origin_initial[1]);
        float gz = __fadd_rn(__fmul_rn(zc[2], cfg.sphere_origin_scale), cfg.sphere_origin_initial[2]);
        float rad = __fadd_rn(__fmul_rn(zc[3], cfg.z_scale), cfg.samples[act ? s : 0]);
        if (cfg.contract_samples) rad = inv_contract_distance(cfg, dv, rad);
        // primitive.py:420-438 + intersect_utils.py:45-84
        float sox = __fmul_rn(ox, gx), soy = __fmul_rn(oy, gy), soz = __fmul_rn(oz, gz);
        float sdx = __fmul_rn(dx, gx), sdy = __fmul_rn(dy, gy), sdz = __fmul_rn(dz, gz);
        float oo = __fadd_rn(__fadd_rn(__fmul_rn(sox, sox), __fmul_rn(soy, soy)), __fmul_rn(soz, soz));
        float dd = __fadd_rn(__fadd_rn(__fmul_rn(sdx, sdx), __fmul_rn(sdy, sdy)), __fmul_rn(sdz, sdz));
        float od = __fadd_rn(__fadd_rn(__fmul_rn(sox, sdx), __fmul_rn(soy, sdy)), __fmul_rn(soz, sdz));
        float a = dd, b = __fmul_rn(2.0f, od), c = __fsub_rn(oo, __fmul_rn(rad, rad));
        float disc = __fsub_rn(__fmul_rn(b, b), __fmul_rn(__fmul_rn(4.0f, a), c));
        disc = (disc < 0.0f) ? 0.0f : disc;
        float sq = sqrtf(__fadd_rn(disc, 1e-8f));
        float a2 = __fmul_rn(2.0f, a);
        float t1 = __fdiv_rn(__fadd_rn(-b, sq), a2);
        float t2 = __fdiv_rn(__fsub_rn(-b, sq), a2);
        if (disc <= 0.0f) { t1 = 0.0f; t2 = 0.0f; }
        t = ((t2 < 0.0f) || (rad < 0.0f)) ? t1 : t2;
      }
      if ((t <= cfg.isect_near) || (t >= cfg.isect_far)) t = 0.0f;
      tkey[j] = act ? t : __int_as_float(0x7f800000);

      // ---- per-sample point displacement: flow * dt (point.py:816-820) + offset (point.py:383-391) ----
      padd[j][0] = padd[j][1] = padd[j][2] = 0.0f;
      float fl[3] = {0.f, 0.f, 0.f}, of[3] = {0.f, 0.f, 0.f};
      if (cfg.use_flow) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          fl[c] = __fmul_rn(apply_act(cfg.flow_act, apply_act(cfg.act_flow, H(cfg.off_flow + c))), toff);
      }
      if (cfg.use_offset) {
        float om = __fsub_rn(1.0f, density(cfg.offset_density_off));
#pragma unroll
        for (int c = 0; c < 3; ++c)
          of[c] = __fmul_rn(apply_act(cfg.offset_act, apply_act(cfg.act_offset, H(cfg.off_offset + c))), om);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { padd[j][c] = fl[c]; cs[j][c] = of[c]; }
      // (of[] parked in cs[] until the points are formed; overwritten by the colour heads below)
#pragma unroll
      for (int c = 0; c < 3; ++c) csh[j][c] = 0.0f;
      if (cfg.use_color_scale_shift) {
#pragma unroll
        for (int c = 0; c < 3; ++c) csh[j][c] = apply_act(cfg.act_cshift, H(cfg.off_cshift + c));
      }
    }

    // ---- sort distances only (base.py:206-210) ----
    if (cfg.isect_sort) sort_keys<SPL>(tkey, lane);

    // ---- points, contraction, flow, offset, validity, normalised coordinates ----
    float dist[SPL], un[SPL][3];
    bool valid[SPL];
    float cocx = ox, cocy = oy, cocz = oz;
    if (cfg.contract_type == HR_CONTRACT_MIPNERF) contract_point(cfg, dv, cocx, cocy, cocz);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = lane + 32 * j;
      const bool act = s < S;
      float t = act ? tkey[j] : 0.0f;
      const bool zero = (t == 0.0f);
      float px = __fadd_rn(ox, __fmul_rn(dx, t));
      float py = __fadd_rn(oy, __fmul_rn(dy, t));
      float pz = __fadd_rn(oz, __fmul_rn(dz, t));
      if (cfg.contract_type == HR_CONTRACT_MIPNERF) {  // base.py:242-246, contract.py:43-50
        contract_point(cfg, dv, px, py, pz);
        float ex = __fsub_rn(px, cocx), ey = __fsub_rn(py, cocy), ez = __fsub_rn(pz, cocz);
        t = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
        if (zero) t = 0.0f;
      }
      // flow then offset (two separate adds in the reference)
      px = __fadd_rn(__fadd_rn(px, padd[j][0]), cs[j][0]);
      py = __fadd_rn(__fadd_rn(py, padd[j][1]), cs[j][1]);
      pz = __fadd_rn(__fadd_rn(pz, padd[j][2]), cs[j][2]);
      dist[j] = t;
      // valid_mask (tensorf_base.py:349-353) & distance > 0 (tensorf_dynamic.py:690)
      bool inside = !((cfg.aabb[0] > px) || (px > cfg.aabb[3]) || (cfg.aabb[1] > py) || (py > cfg.aabb[4]) ||
                      (cfg.aabb[2] > pz) || (pz > cfg.aabb[5]));
      valid[j] = act && inside && (t > 0.0f);
      // normalize_coord (tensorf_base.py:308-309)
      un[j][0] = __fsub_rn(__fmul_rn(__fsub_rn(px, cfg.aabb[0]), __fdiv_rn(2.0f, __fsub_rn(cfg.aabb[3], cfg.aabb[0]))), 1.0f);
      un[j][1] = __fsub_rn(__fmul_rn(__fsub_rn(py, cfg.aabb[1]), __fdiv_rn(2.0f, __fsub_rn(cfg.aabb[4], cfg.aabb[1]))), 1.0f);
      un[j][2] = __fsub_rn(__fmul_rn(__fsub_rn(pz, cfg.aabb[2]), __fdiv_rn(2.0f, __fsub_rn(cfg.aabb[5], cfg.aabb[2]))), 1.0f);
      if (STAGES && act) {
        if (so.distances) so.distances[ray * S + s] = t;
        if (so.points) {
          so.points[(ray * S + s) * 3 + 0] = px;
          so.points[(ray * S + s) * 3 + 1] = py;
          so.points[(ray * S + s) * 3 + 2] = pz;
        }
      }
      // colour scale head (needed only at the end)
#pragma unroll
      for (int c = 0; c < 3; ++c) cs[j][c] = 0.0f;
      if (cfg.use_color_scale_shift) {
        const float* hp = hrow + s;
#pragma unroll
        for (int c = 0; c < 3; ++c) cs[j][c] = act ? apply_act(cfg.act_cscale, __ldg(hp + (cfg.off_cscale + c) * S)) : 0.0f;
      }
    }

    // ---- VM gather, 8 samples per round, 4 lanes per sample ----
    float sig_r[ROUNDS];  // sigma feature of (round, quad), replicated in the quad
    float rgb_r[ROUNDS];  // shaded colour channel q of (round, quad) (lanes q<3)
    float Y[9];
    if constexpr (SHADE == HR_SHADE_SH) sh_basis9(dx, dy, dz, Y);  // viewdirs = rays[:,3:6] as given (point.py:866-867)
    const int qc = min(q, 2);
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int j = rd >> 2;
      const int src = (rd & 3) * 8 + quad;
      const float u0 = __shfl_sync(kFull, un[j][0], src);
      const float u1 = __shfl_sync(kFull, un[j][1], src);
      const float u2 = __shfl_sync(kFull, un[j][2], src);
      const bool ok = __shfl_sync(kFull, valid[j] ? 1 : 0, src) != 0;
      sig_r[rd] = 0.0f;
      rgb_r[rd] = 0.0f;
      if (rd * 8 >= S) continue;  // warp-uniform
      // matMode = [[0,1],[0,2],[1,2]], vecMode / time axis = [2,1,0]
      TapSet<C0> s0s, s0t, a0s, a0t;
      TapSet<(C1 ? C1 : 4)> s1s, s1t, a1s, a1t;
      TapSet<(C2 ? C2 : 4)> s2s, s2t, a2s, a2t;
      issue<C0>(s0s, tabs.sig[0].space, tabs.sig[0].H, tabs.sig[0].W, u0, u1, xt, alt, ok);
      issue<C0>(s0t, tabs.sig[0].second, tabs.sig[0].H2, tabs.sig[0].L, u2, tau, xt, alt, ok);
      issue<C0>(a0s, tabs.app[0].space, tabs.app[0].H, tabs.app[0].W, u0, u1, xt, alt, ok);
      issue<C0>(a0t, tabs.app[0].second, tabs.app[0].H2, tabs.app[0].L, u2, tau, xt, alt, ok);
      if constexpr (C1 > 0) {
        issue<C1>(s1s, tabs.sig[1].space, tabs.sig[1].H, tabs.sig[1].W, u0, u2, xt, alt, ok);
        issue<C1>(s1t, tabs.sig[1].second, tabs.sig[1].H2, tabs.sig[1].L, u1, tau, xt, alt, ok);
        issue<C1>(a1s, tabs.app[1].space, tabs.app[1].H, tabs.app[1].W, u0, u2, xt, alt, ok);
        issue<C1>(a1t, tabs.app[1].second, tabs.app[1].H2, tabs.app[1].L, u1, tau, xt, alt, ok);
      }
      if constexpr (C2 > 0) {
        issue<C2>(s2s, tabs.sig[2].space, tabs.sig[2].H, tabs.sig[2].W, u1, u2, xt, alt, ok);
        issue<C2>(s2t, tabs.sig[2].second, tabs.sig[2].H2, tabs.sig[2].L, u0, tau, xt, alt, ok);
        issue<C2>(a2s, tabs.app[2].space, tabs.app[2].H, tabs.app[2].W, u1, u2, xt, alt, ok);
        issue<C2>(a2t, tabs.app[2].second, tabs.app[2].H2, tabs.app[2].L, u0, tau, xt, alt, ok);
      }
      // ---- density feature: sum_c space_c * second_c over all groups (tensorf_dynamic.py:330) ----
      float f[NT];  // appearance product features, torch.cat order over groups
      float sf = 0.0f;
      {
        float A[4], B[4];
        combine<C0>(s0s, A);
        combine<C0>(s0t, B);
        float part = A[0] * B[0] + A[1] * B[1] + A[2] * B[2] + A[3] * B[3];
        if constexpr (C0 == 8) part += __shfl_xor_sync(kFull, part, 1);
        sf = part;
        combine<C0>(a0s, A);
        combine<C0>(a0t, B);
        if constexpr (C0 == 8) {
          float mine[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) mine[c] = A[c] * B[c];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float oth = __shfl_xor_sync(kFull, mine[c], 1);
            f[c] = alt ? oth : mine[c];
            f[4 + c] = alt ? mine[c] : oth;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) f[c] = A[c] * B[c];
        }
      }
      if constexpr (C1 > 0) {
        float A[4], B[4];
        combine<(C1 ? C1 : 4)>(s1s, A);
        combine<(C1 ? C1 : 4)>(s1t, B);
        float part = A[0] * B[0] + A[1] * B[1] + A[2] * B[2] + A[3] * B[3];
        if constexpr (C1 == 8) part += __shfl_xor_sync(kFull, part, 1);
        sf += part;
        combine<(C1 ? C1 : 4)>(a1s, A);
        combine<(C1 ? C1 : 4)>(a1t, B);
        if constexpr (C1 == 8) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float m = A[c] * B[c];
            float oth = __shfl_xor_sync(kFull, m, 1);
            f[C0 + c] = alt ? oth : m;
            f[C0 + 4 + c] = alt ? m : oth;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) f[C0 + c] = A[c] * B[c];
        }
      }
      if constexpr (C2 > 0) {
        float A[4], B[4];
        combine<(C2 ? C2 : 4)>(s2s, A);
        combine<(C2 ? C2 : 4)>(s2t, B);
        float part = A[0] * B[0] + A[1] * B[1] + A[2] * B[2] + A[3] * B[3];
        if constexpr (C2 == 8) part += __shfl_xor_sync(kFull, part, 1);
        sf += part;
        combine<(C2 ? C2 : 4)>(a2s, A);
        combine<(C2 ? C2 : 4)>(a2t, B);
        if constexpr (C2 == 8) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float m = A[c] * B[c];
            float oth = __shfl_xor_sync(kFull, m, 1);
            f[C0 + C1 + c] = alt ? oth : m;
            f[C0 + C1 + 4 + c] = alt ? m : oth;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) f[C0 + C1 + c] = A[c] * B[c];
        }
      }
      sig_r[rd] = ok ? sf : 0.0f;
      // ---- appearance: basis_mat (tensorf_dynamic.py:371) + shading (tensorf_utils.py:334-343) ----
      const float* bq = s_basis + qc * BS;
      float col;
      if constexpr (SHADE == HR_SHADE_SH) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          float F = 0.0f;
#pragma unroll
          for (int i = 0; i < NT; ++i) F = fmaf(bq[k * NT + i], f[i], F);
          acc = fmaf(Y[k], F, acc);
        }
        col = fmaxf(acc + 0.5f, 0.0f);
      } else {
        float F = 0.0f;
#pragma unroll
        for (int i = 0; i < NT; ++i) F = fmaf(bq[i], f[i], F);
        col = 1.0f / (1.0f + expf(-F));
      }
      rgb_r[rd] = ok ? col : 0.0f;
    }

    // ---- back to lane = sample: sigma, alpha, transmittance, weights (tensorf_utils.py:242-253) ----
    float wgt[SPL];
    float carryT = 1.0f;
    float accw = 0.0f, accB[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = lane + 32 * j;
      float feat = 0.0f;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        float v = __shfl_sync(kFull, sig_r[4 * j + rr], 4 * (lane & 7));
        if ((lane >> 3) == rr) feat = v;
      }
      // feature2density (tensorf_dynamic.py:373-392; static tensorf_no_sample.py:82-88,187: weights == 1)
      float sigma;
      if (cfg.fea2dense == HR_DENSE_RELU) sigma = fmaxf(feat, 0.0f);
      else if (cfg.fea2dense == HR_DENSE_RELU_ABS) sigma = fabsf(feat);
      else {
        float xs = feat + cfg.density_shift;
        sigma = (xs > 20.0f) ? xs : log1pf(expf(xs));
      }
      if (!valid[j]) sigma = 0.0f;
      // deltas: dist[i+1]-dist[i], last = 1e10 (tensorf_dynamic.py:663-670)
      float nxt = __shfl_down_sync(kFull, dist[j], 1);
      if (j + 1 < SPL) {
        float first_next = __shfl_sync(kFull, dist[(j + 1 < SPL) ? j + 1 : j], 0);
        if (lane == 31) nxt = first_next;
      }
      float delta = (s == S - 1) ? 1e10f : __fsub_rn(nxt, dist[j]);
      float alpha = __fsub_rn(1.0f, expf(-__fmul_rn(sigma, __fmul_rn(delta, cfg.distance_scale))));
      if (s >= S) alpha = 0.0f;
      float a1 = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      if (s >= S) a1 = 1.0f;
      // inclusive product scan
      float inc = a1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        float o = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc *= o;
      }
      float exc = __shfl_up_sync(kFull, inc, 1);
      if (lane == 0) exc = 1.0f;
      float T = carryT * exc;
      carryT = carryT * __shfl_sync(kFull, inc, 31);
      float w = alpha * T;
      wgt[j] = w;
      if (STAGES && s < S) {
        if (so.sigma) so.sigma[ray * S + s] = sigma;
        if (so.weights) so.weights[ray * S + s] = w;
      }
      accw += w;
#pragma unroll
      for (int c = 0; c < 3; ++c) accB[c] += w * csh[j][c];
    }

    // ---- composite: sum_s w_s * (rgb_s*(1+cs_s) + csh_s) (tensorf_dynamic.py:780-792) ----
    float accq = 0.0f;
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int j = rd >> 2;
      const int src = (rd & 3) * 8 + quad;
      float m = (wgt[j] > cfg.weight_thre) ? wgt[j] : 0.0f;  // app_mask (tensorf_dynamic.py:750)
      float A0 = m * (cs[j][0] + 1.0f), A1 = m * (cs[j][1] + 1.0f), A2 = m * (cs[j][2] + 1.0f);
      float g0 = __shfl_sync(kFull, A0, src);
      float g1 = __shfl_sync(kFull, A1, src);
      float g2 = __shfl_sync(kFull, A2, src);
      float Aq = (q == 0) ? g0 : ((q == 1) ? g1 : g2);
      accq = fmaf(Aq, rgb_r[rd], accq);
    }
#pragma unroll
    for (int d = 4; d < 32; d <<= 1) accq += __shfl_xor_sync(kFull, accq, d);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      accw += __shfl_xor_sync(kFull, accw, d);
      accB[0] += __shfl_xor_sync(kFull, accB[0], d);
      accB[1] += __shfl_xor_sync(kFull, accB[1], d);
      accB[2] += __shfl_xor_sync(kFull, accB[2], d);
    }
    if (lane < 3) {
      float v = accq + ((lane == 0) ? accB[0] : ((lane == 1) ? accB[1] : accB[2]));
      if (cfg.white_bg && !cfg.black_bg) v = v + (1.0f - accw);
      if (cfg.clamp_output) v = fminf(fmaxf(v, 0.0f), 1.0f);
      rgb_out[ray * 3 + lane] = v;
    }
  }
}

template <int SPL, bool DYN, int C0, int C1, int C2, int SHADE>
static cudaError_t launch_one(const hr_config& cfg, const Derived& dv, const RenderTabs& tabs, const float* rays,
                              const float* heads, float* rgb, long long n, const StageOut* so, int num_sms,
                              cudaStream_t stream) {
  constexpr int ROWS = (SHADE == HR_SHADE_SH) ? 9 : 1;
  constexpr int NT = C0 + C1 + C2;
  size_t smem = 3 * (size_t)basis_block_stride(ROWS, NT) * sizeof(float);
  long long ctas_needed = (n + kWarpsPerCta - 1) / kWarpsPerCta;
  long long grid = ctas_needed < (long long)num_sms * 16 ? ctas_needed : (long long)num_sms * 16;
  if (grid < 1) grid = 1;
  if (so) {
    render_kernel<SPL, DYN, C0, C1, C2, SHADE, true><<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(
        cfg, dv, tabs, rays, heads, rgb, n, *so);
  } else {
    StageOut none{nullptr, nullptr, nullptr, nullptr};
    render_kernel<SPL, DYN, C0, C1, C2, SHADE, false><<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(
        cfg, dv, tabs, rays, heads, rgb, n, none);
  }
  return cudaGetLastError();
}

template <int SPL, bool DYN, int C0, int C1, int C2>
static cudaError_t launch_shade(const hr_config& cfg, const Derived& dv, const RenderTabs& tabs, const float* rays,
                                const float* heads, float* rgb, long long n, const StageOut* so, int num_sms,
                                cudaStream_t stream) {
  if (cfg.shading == HR_SHADE_SH)
    return launch_one<SPL, DYN, C0, C1, C2, HR_SHADE_SH>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream);
  return launch_one<SPL, DYN, C0, C1, C2, HR_SHADE_RGB>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream);
}

template <int SPL, bool DYN>
static cudaError_t launch_comps(const hr_config& cfg, const Derived& dv, const RenderTabs& tabs, const float* rays,
                                const float* heads, float* rgb, long long n, const StageOut* so, int num_sms,
                                cudaStream_t stream) {
  const int c0 = cfg.n_sigma[0], c1 = cfg.n_sigma[1], c2 = cfg.n_sigma[2];
  if (c0 == 8 && c1 == 0 && c2 == 0)
    return launch_shade<SPL, DYN, 8, 0, 0>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream);
  if (c0 == 8 && c1 == 4 && c2 == 4)
    return launch_shade<SPL, DYN, 8, 4, 4>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream);
  if (c0 == 8 && c1 == 8 && c2 == 8)
    return launch_shade<SPL, DYN, 8, 8, 8>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream);
  return cudaErrorInvalidValue;
}

// Entry used by hr_api.cu.  Returns cudaErrorInvalidValue for an unsupported component layout.
cudaError_t launch_render(const hr_config& cfg, const Derived& dv, const RenderTabs& tabs, const float* rays,
                          const float* heads, float* rgb, long long n, const StageOut* so, int num_sms,
                          cudaStream_t stream) {
  const bool two = cfg.n_samples > 32;
  if (cfg.dynamic) {
    return two ? launch_comps<2, true>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream)
               : launch_comps<1, true>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream);
  }
  return two ? launch_comps<2, false>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream)
             : launch_comps<1, false>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream);
}

}  // namespace hr
