// Sample-net weight packs and launchers shared between hr_api.cu and the kernels.
#pragma once
#include <cuda_runtime.h>
#include "hyperreel_b200.h"

namespace hr {

// fp32 CUDA-core path: per layer a k-major matrix Wt[Kp][Np] (zero padded) and a bias[Np].
//   layer 0      : rows = encoded input, padded to in_pad
//   skip layer   : rows = [input (in_pad) ; hidden (W)]   (mlp.py:167-168: cat([input_x, x]))
//   other layers : rows = hidden (W)
//   last layer   : columns permuted to channel-major (column c*S+s  <-  reference row s*stride+c)
struct MlpSimtPack {
  const float* Wt[HR_MAX_LAYERS];
  const float* bias[HR_MAX_LAYERS];
  int Kp[HR_MAX_LAYERS];
  int Np[HR_MAX_LAYERS];
  int in_pad;
  int n_layers;
  int skip;
};

// tcgen05 path (HR_MLP_BF16X3_TC): see hr_mlp_tc2.cu.  A "pass" is one accumulator's worth of output columns
// (128 columns of a hidden layer or of the last layer); its weights are stored as n_chunks*2 k-step images.
#define HR_TC_MAX_PASSES 40  // 10 hidden half passes + 28 last-layer parts (S = 256 x 14 channels)
struct TcPass {
  int layer;        // Linear layer index
  int n;            // output columns of this pass (multiple of 16, <= 256)
  int first_chunk;  // first A chunk consumed (0 = encoded input, in_chunks = first hidden chunk)
  int n_chunks;     // chunks of 32 k
  int bias_off;     // offset into the bias table
  int is_final;     // last layer: results go to HBM instead of the next A operand
  int out_col0;     // first output column (channel-major order) of a last-layer pass
  int wait_a;       // the issuer must wait for the A chunks (first pass of a layer)
};
struct MlpTcPack {
  const void* wpack;   // bf16 hi/lo weight images, UMMA K-major no-swizzle layout, consumption order
  const float* bias;   // [bias_count]
  long long wpack_bytes;
  int n_passes;
  int bias_count;
  int in_chunks;       // 32-wide k-chunks of the encoded input (1: mlp_in <= 32, 2: mlp_in <= 64)
  TcPass passes[HR_TC_MAX_PASSES];
};

size_t mlp_simt_smem_bytes(const MlpSimtPack& pk, int W);
cudaError_t launch_mlp_simt(const hr_config& cfg, const MlpSimtPack& pk, const float* rays, float* heads,
                            long long n, int num_sms, cudaStream_t stream);

// rays may point to pinned host memory (read once, by the encoder warps); rays_copy (optional) receives a device copy;
// tma_encode = the driver's cuTensorMapEncodeTiled (hr_handle::tma_encode)
cudaError_t launch_mlp_tc2(const hr_config& cfg, const MlpTcPack& pk, void* tma_encode, const float* rays, float* heads,
                           long long n, int num_sms, cudaStream_t stream, float* rays_copy = nullptr);
}  // namespace hr

struct hr_handle;
struct hr_params;
namespace hr {
// packs the net `c` describes into `pk`; alloc_bytes / alloc_bias remember the allocation behind pk (reused while unchanged)
int pack_mlp_tc2(hr_handle* h, const hr_config& c, MlpTcPack& pk, size_t& alloc_bytes, int& alloc_bias,
                 const float* const* w_dev, const float* const* b_dev, cudaStream_t st);
void free_mlp_tc2(hr_handle* h);
}
