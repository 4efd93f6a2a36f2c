// Instantiations of the fused render kernel (hr_render_kernel.cuh) for 64 < S <= 256 samples per ray: 4 or 8 samples per
// lane (neural_3d_z_plane_static: z_channels 256, technicolor_z_plane_no_sample: 128, catacaustics_voxel: 96).  Same code as
// the S <= 64 variants (every primitive compiled in: RARE) with 4-8 times larger per-sample register arrays (one CTA per SM).  The file is compiled four times
// (HR_BIG_SPL in {4, 8} x HR_BIG_DYN in {0, 1}, see the Makefile) so the variants build in parallel.
#include "hr_render_kernel.cuh"

#ifndef HR_BIG_SPL
#error "compile with -DHR_BIG_SPL=4|8 -DHR_BIG_DYN=0|1"
#endif

namespace hr {

#define HR_BIG_NAME2(spl, dyn) launch_render_big_##spl##_##dyn
#define HR_BIG_NAME(spl, dyn) HR_BIG_NAME2(spl, dyn)

cudaError_t HR_BIG_NAME(HR_BIG_SPL, HR_BIG_DYN)(const hr_config& cfg, const Derived& dv, const RenderTabs& tabs, const float* rays,
                                                const float* heads, const RgbDst& rgb, long long n, const ExtraOut* so,
                                                int num_sms, cudaStream_t stream, unsigned char* rgb8) {
  return launch_comps<HR_BIG_SPL, (HR_BIG_DYN != 0), true>(cfg, dv, tabs, rays, heads, rgb, n, so, num_sms, stream, rgb8);
}

}  // namespace hr
