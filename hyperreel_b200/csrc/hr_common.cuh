// Shared device-side types for the hyperreel_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hyperreel_b200.h"

namespace hr {

// One VM factor pair, channel-last (DESIGN.md "Data layout in HBM"):
//   space  [H][W][C]   (reference [1,C,H,W]: x -> W (axis a), y -> H (axis b))
//   second [H2][L][C]  dynamic: H2 = K keyframe rows, x -> L (axis c), y -> time
//                      static : H2 = 1, the line [L][C]
struct PlaneTab {
  const float* space;
  const float* second;
  int H, W;   // space plane rows / cols
  int H2, L;  // second factor rows / cols
  int C;      // channels (0 = group absent, 4 or 8)
};

struct RenderTabs {
  PlaneTab sig[3];
  PlaneTab app[3];
  const float* basis;  // [app_dim][sum C_app] row-major
  int n_app_total;
  const float* color_embedding;  // [n_color_views][12] per-camera colour transform + shift (point.py:558-592), or null
};

// Host-derived scalars (computed in double on the host, then rounded once to fp32, the way the
// reference's Python-float constants meet fp32 tensors).
struct Derived {
  float time_fac;       // K*(F-1)/F            (utils/flow_utils.py:19)
  float time_inv_fac;   // 1/fac                (flow_utils.py:31)
  float time_scale;     // (F-1)/F              (tensorf_dynamic.py:58)
  float time_offset;    // 0.5/K                (tensorf_dynamic.py:59)
  float kf_max;         // K-1
  float inv_end_dist;   // contract_start_distance / contract_end_distance (contract.py:145)
  float dist_scale_fac; // 1/(1-inv_end_dist)
  float inv_end_rad;    // contract_start_radius / contract_end_radius (contract.py:184)
  float rad_scale_fac;  // 1/(1-inv_end_rad)
  int res[3];           // grid resolution along x, y, z (taken from the uploaded tables)
  int kt;               // rows of the (axis, time) planes = number of keyframes
};

// Where the finished pixels go: one or more [n_total,3] fp32 buffers (the local output; or, for ray-sharded rendering, the
// same row range of every rank's gather buffer, peer pointers mapped over NVLink -- the render kernel's epilogue is the
// gather).  `row0` is added to the ray index.
struct RgbDst {
  float* p[HR_MAX_PEERS];
  int n;
  long long row0;
};

// Outputs beyond rgb, produced by the EXTRA variant of the render kernel (all pointers may be null):
//   * per-sample dumps for stage-boundary parity tests (hr_render_stages);
//   * the extra composited fields of the reference's colour nets (tensorf_dynamic.py:808-837, tensorf_no_sample.py:254-278):
//     field_out[f] receives sum_s w_s * x[f]_s ([n, dim], HR_FIELD_OVER), sum_s pred_w_s * x[f]_s (HR_FIELD_PRED_WEIGHTS) or
//     the per-sample values themselves ([n, S*dim], HR_FIELD_NO_OVER).
struct ExtraOut {
  float* distances;    // [n,S]
  float* points;       // [n,S,3]
  float* sigma;        // [n,S]
  float* weights;      // [n,S]
  float* rgb_samples;  // [n,S,3] shaded colour of every sample before the colour transform, 0 where w <= rm_weight_mask_thre
  float* field_out[HR_N_FIELDS];
  int field_mode[HR_N_FIELDS];
};

__device__ __forceinline__ float apply_act(const hr_act& a, float x) {
  // y = f(x*inner + shift) * outer, each op rounded separately like the eager reference.
  float v = __fadd_rn(__fmul_rn(x, a.inner_fac), a.shift);
  // ex2-based forms: relative error ~2e-6, far below the 1e-4 RGB gate and cheaper than expf / tanhf by ~4x
  if (a.kind == HR_ACT_SIGMOID) {
    v = __fdividef(1.0f, 1.0f + __expf(-v));
  } else if (a.kind == HR_ACT_TANH) {
    const float av = fminf(fabsf(v), 15.0f);
    const float t = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * av));
    v = copysignf(t, v);
  }
  return __fmul_rn(v, a.outer_fac);
}

}  // namespace hr
