// Instantiations of the fused render kernel (hr_render_kernel.cuh) for S <= 64 samples per ray with every primitive compiled
// in (RARE): sphere_new, euclidean_distance_unified, voxel_grid, deformable_voxel_grid, and the per-camera colour transform.
// The z-plane / sphere / cylinder pipelines use the leaner variants of hr_render.cu.
#include "hr_render_kernel.cuh"

namespace hr {

cudaError_t launch_render_rare(const hr_config& cfg, const Derived& dv, const RenderTabs& tabs, const float* rays,
                               const float* heads, const RgbDst& rgb, long long n, const ExtraOut* so, int num_sms,
                               cudaStream_t stream, unsigned char* rgb8) {
  const bool two = cfg.n_samples > 32;
  if (cfg.dynamic) {
    return two ? launch_comps<2, true, true>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream, rgb8)
               : launch_comps<1, true, true>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream, rgb8);
  }
  return two ? launch_comps<2, false, true>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream, rgb8)
             : launch_comps<1, false, true>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream, rgb8);
}

}  // namespace hr
