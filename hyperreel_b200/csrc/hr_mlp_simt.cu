// Sample-prediction network, fp32 CUDA-core path (HR_MLP_FP32_SIMT).
//
// Computes, per tile of 128 rays held in shared memory for the whole network,
//   ray -> RayParam + WindowedPE (nlf/param.py:87-115,244-253; nlf/pe.py:210-221)
//       -> BaseMLP (nlf/nets/mlp.py:159-172): Linear + LeakyReLU stack, skip = cat([input, h])
// and writes the per-sample heads in channel-major order (column c*S+s) so the render kernel's
// "lane = sample" loads are coalesced.  fp32 FMA accumulation, i.e. the same arithmetic class as
// the reference's SGEMM; this is the parity anchor for the tensor-core path.
#include "hr_common.cuh"
#include "hr_mlp.cuh"
#include "hr_encode.cuh"

namespace hr {

static constexpr int BM = 128;   // rays per CTA tile
static constexpr int KC = 16;    // k-chunk of the weight ring
static constexpr int NTHREADS = 256;
static constexpr int RPT = 16;   // rows (rays) per thread
static constexpr int LDA = BM + 4;  // padded row stride of the k-major activation tile (bank spread)

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int W>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_simt_kernel(const __grid_constant__ hr_config cfg, const __grid_constant__ MlpSimtPack pk,
                const float* __restrict__ rays, float* __restrict__ heads, long long n_rays) {
  constexpr int CPT = W / 32;       // columns per thread: col = c*32 + lane
  extern __shared__ __align__(16) float smem[];
  const int in_pad = pk.in_pad;
  float* A_s = smem;                               // [(in_pad + W)][LDA]
  float* B_s = smem + (size_t)(in_pad + W) * LDA;  // [2][KC][W]

  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const long long tile0 = (long long)blockIdx.x * BM;

  for (long long base = tile0; base < n_rays; base += (long long)gridDim.x * BM) {
    __syncthreads();
    // ---- encode ----
    if (tid < BM) {
      for (int k = 0; k < in_pad; ++k) A_s[k * LDA + tid] = 0.0f;
      long long ray = base + tid;
      if (ray < n_rays) encode_ray(cfg, rays + ray * cfg.c_in, A_s + tid, LDA);
    }
    __syncthreads();

    for (int l = 0; l < pk.n_layers; ++l) {
      const bool last = (l == pk.n_layers - 1);
      const int k_begin = (l == 0 || l == pk.skip) ? 0 : in_pad;
      const int Kp = pk.Kp[l];
      const int Np = pk.Np[l];
      const int nk = Kp / KC;
      const float* __restrict__ Wt = pk.Wt[l];
      const float* __restrict__ bias = pk.bias[l];
      for (int nb = 0; nb < Np; nb += W) {
        float acc[RPT][CPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
          for (int c = 0; c < CPT; ++c) acc[i][c] = 0.0f;

        auto load_chunk = [&](int kc, int buf) {
          // KC x W floats = KC*W/4 float4, NTHREADS threads
          const float* src = Wt + (size_t)(kc * KC) * Np + nb;
          float* dst = B_s + buf * (KC * W);
          for (int i = tid; i < KC * W / 4; i += NTHREADS) {
            int row = i / (W / 4), c4 = i % (W / 4);
            cp_async16(dst + row * W + c4 * 4, src + (size_t)row * Np + c4 * 4);
          }
        };
        load_chunk(0, 0);
        cp_async_commit();
        for (int kc = 0; kc < nk; ++kc) {
          if (kc + 1 < nk) {
            load_chunk(kc + 1, (kc + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          __syncthreads();
          const float* Bb = B_s + (kc & 1) * (KC * W);
          const float* Ab = A_s + (size_t)(k_begin + kc * KC) * LDA + ty * RPT;
#pragma unroll 4
          for (int kk = 0; kk < KC; ++kk) {
            float a[RPT];
#pragma unroll
            for (int i4 = 0; i4 < RPT / 4; ++i4) {
              float4 t = *reinterpret_cast<const float4*>(Ab + kk * LDA + i4 * 4);
              a[i4 * 4 + 0] = t.x; a[i4 * 4 + 1] = t.y; a[i4 * 4 + 2] = t.z; a[i4 * 4 + 3] = t.w;
            }
            float b[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) b[c] = Bb[kk * W + c * 32 + tx];
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
              for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(a[i], b[c], acc[i][c]);
          }
          __syncthreads();
        }
        // ---- epilogue ----
        if (!last) {
          // hidden layer: bias + LeakyReLU (activations.py:14-29), written k-major for the next layer
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            int col = c * 32 + tx;
            float bz = __ldg(bias + col);
            float* dst = A_s + (size_t)(in_pad + col) * LDA + ty * RPT;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              float x = acc[i][c] + bz;
              dst[i] = (x > 0.0f) ? x : x * cfg.leaky_slope;
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            int col = nb + c * 32 + tx;
            if (col >= cfg.mlp_out) continue;
            float bz = __ldg(bias + col);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              long long ray = base + ty * RPT + i;
              if (ray < n_rays) heads[ray * cfg.mlp_out + col] = acc[i][c] + bz;
            }
          }
        }
        __syncthreads();
      }
    }
  }
}

size_t mlp_simt_smem_bytes(const MlpSimtPack& pk, int W) {
  return ((size_t)(pk.in_pad + W) * LDA + 2 * (size_t)KC * W) * sizeof(float);
}

cudaError_t launch_mlp_simt(const hr_config& cfg, const MlpSimtPack& pk, const float* rays, float* heads,
                            long long n, int num_sms, cudaStream_t stream) {
  long long tiles = (n + BM - 1) / BM;
  int grid = (int)(tiles < num_sms ? tiles : num_sms);
  if (grid < 1) grid = 1;
  size_t smem = mlp_simt_smem_bytes(pk, cfg.mlp_width);
  cudaError_t e;
  if (cfg.mlp_width == 256) {
    e = cudaFuncSetAttribute(mlp_simt_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mlp_simt_kernel<256><<<grid, NTHREADS, smem, stream>>>(cfg, pk, rays, heads, n);
  } else if (cfg.mlp_width == 128) {
    e = cudaFuncSetAttribute(mlp_simt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mlp_simt_kernel<128><<<grid, NTHREADS, smem, stream>>>(cfg, pk, rays, heads, n);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace hr
