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

// RayPredictionEmbedding input encoding for one ray (nlf/embedding/ray.py:320-326).
// Writes cfg.mlp_in values with stride `stride` starting at dst.
__device__ void encode_ray(const hr_config& cfg, const float* __restrict__ ray, float* dst, int stride) {
  int k = 0;
  for (int g = 0; g < cfg.n_groups; ++g) {
    const hr_encode_group& G = cfg.groups[g];
    float v[8];
    int dims;
    const float* r = ray + G.start;
    if (G.fn == HR_PARAM_TWO_PLANE) {
      // TwoPlaneParam (param.py:87-115) + intersect_axis_plane (intersect_utils.py:127-150)
      float oz = r[2], dz = r[5];
      float dzg = (fabsf(dz) < 1e-5f) ? 1e12f : dz;
      float t1 = __fdiv_rn(__fsub_rn(G.near, oz), dzg);
      float t2 = __fdiv_rn(__fsub_rn(G.far, oz), dzg);
      v[0] = __fadd_rn(r[0], __fmul_rn(r[3], t1));
      v[1] = __fadd_rn(r[1], __fmul_rn(r[4], t1));
      v[2] = __fadd_rn(r[0], __fmul_rn(r[3], t2));
      v[3] = __fadd_rn(r[1], __fmul_rn(r[4], t2));
      dims = 4;
    } else if (G.fn == HR_PARAM_PLUECKER) {
      // PlueckerParam (param.py:244-253)
      float ox = r[0], oy = r[1], oz = r[2];
      float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r[3], r[3]), __fmul_rn(r[4], r[4])), __fmul_rn(r[5], r[5])));
      nrm = fmaxf(nrm, 1e-12f);
      float dx = __fdiv_rn(r[3], nrm), dy = __fdiv_rn(r[4], nrm), dz = __fdiv_rn(r[5], nrm);
      float mx = __fsub_rn(__fmul_rn(oy, dz), __fmul_rn(oz, dy));
      float my = __fsub_rn(__fmul_rn(oz, dx), __fmul_rn(ox, dz));
      float mz = __fsub_rn(__fmul_rn(ox, dy), __fmul_rn(oy, dx));
      v[0] = __fmul_rn(dx, G.dir_mult);
      v[1] = __fmul_rn(dy, G.dir_mult);
      v[2] = __fmul_rn(dz, G.dir_mult);
      v[3] = __fmul_rn(mx, G.mom_mult);
      v[4] = __fmul_rn(my, G.mom_mult);
      v[5] = __fmul_rn(mz, G.mom_mult);
      dims = 6;
    } else {
      dims = G.end - G.start;
      for (int i = 0; i < dims; ++i) v[i] = r[i];
    }
    // WindowedPE with all windows open (pe.py:210-221): [x | sin(f1 x) | cos(f1 x) | sin(f2 x) | ...]
    if (!G.exclude_identity)
      for (int i = 0; i < dims; ++i) dst[(k++) * stride] = v[i];
    float freq = 1.0f;
    for (int f = 0; f < G.n_freqs; ++f) {
      freq = __fmul_rn(freq, G.freq_mult);  // freq_multiplier ** (f+1), exact for 2.0
      float bf = __fmul_rn(G.base_mult, freq);
      for (int i = 0; i < dims; ++i) dst[(k++) * stride] = sinf(__fmul_rn(bf, v[i]));
      for (int i = 0; i < dims; ++i) dst[(k++) * stride] = cosf(__fmul_rn(bf, v[i]));
    }
  }
}

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
