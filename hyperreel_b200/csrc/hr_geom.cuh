// Per-sample geometry shared by the forward and backward render kernels: the sort of the distances, the mipnerf / affine
// contractions and the SH basis (references cited per function).
#pragma once
#include "hr_common.cuh"

namespace hr {

static constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

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
        // partner lives in another register of the same lane: r ^ (stride / 32); size >= 64, so the direction of element
        // e = r*32 + lane depends on r only
#pragma unroll
        for (int r = 0; r < SPL; ++r) {
          const int rs = stride >> 5;
          if ((r & rs) == 0) {
            const int r2 = (r | rs) < SPL ? (r | rs) : r;
            const bool up = (((r * 32) & size) == 0);
            const float lo = fminf(k[r], k[r2]), hi = fmaxf(k[r], k[r2]);
            k[r] = up ? lo : hi;
            k[r2] = up ? hi : lo;
          }
        }
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

// Same for one key per lane within groups of LW consecutive lanes (two rays per warp: LW = 16), e = lane within the group.
template <int LW>
__device__ __forceinline__ void sort_keys_sub(float& k, int e) {
#pragma unroll
  for (int size = 2; size <= LW; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const float other = __shfl_xor_sync(kFull, k, stride);
      const bool up = ((e & size) == 0);
      const bool lower = ((e & stride) == 0);
      k = (lower == up) ? fminf(k, other) : fmaxf(k, other);
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

// bbox / z_depth contraction of a point: (p - min) / den per axis (reference: nlf/contract.py:83-84, :109-110).
__device__ __forceinline__ void contract_point_affine(const hr_config& cfg, float& x, float& y, float& z) {
  x = __fdiv_rn(__fsub_rn(x, cfg.contract_affine_min[0]), cfg.contract_affine_den[0]);
  y = __fdiv_rn(__fsub_rn(y, cfg.contract_affine_min[1]), cfg.contract_affine_den[1]);
  z = __fdiv_rn(__fsub_rn(z, cfg.contract_affine_min[2]), cfg.contract_affine_den[2]);
}
// inverse contraction of a sample position (base.py:132-133): mipnerf (:143-158) or distance * fac (:77-78, :103-104)
__device__ __forceinline__ float inv_contract_sample(const hr_config& cfg, const Derived& dv, float d) {
  return (cfg.contract_type == HR_CONTRACT_AFFINE) ? __fmul_rn(d, cfg.contract_dist_fac) : inv_contract_distance(cfg, dv, d);
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


}  // namespace hr
