// Host helpers of the tensor-core sample net (hr_mlp_tc2.cu): weight images and the TMA tensor map of the heads scratch.
//
// Weight images: every fp32 weight w of the reference's nn.Linear (nlf/nets/mlp.py:127-154) is split into bf16
// hi = rn(w), lo = rn(w - hi) and laid out the way tcgen05.mma reads its B operand from shared memory: UMMA K-major,
// no swizzle, core matrices of 8 rows x 16 bytes (LBO = N * 16 B between the two 8-wide k-groups of a 16-wide k-step,
// SBO = 128 B between 8-row groups).  One image = one k-step of one pass = N x 16 bf16 hi followed by N x 16 bf16 lo;
// images are concatenated in consumption order so the producer warp streams them with cp.async.bulk.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>

#include <cstring>

#include "hr_tc_prims.cuh"

namespace hr {

// A pass consumes `n_chunks` 32-wide k-chunks starting at chunk `first_chunk`; chunks [0, in_chunks) are the encoded
// input (zero padded past mlp_in), chunks in_chunks.. the hidden activations.  The skip layer's weight is
// [out, mlp_in + W] with the input columns first (mlp.py:167-168: cat([input_x, x])).
__global__ void pack_tc_pass(const float* __restrict__ W, const float* __restrict__ b, uint8_t* __restrict__ dst,
                             float* __restrict__ bias_dst, int n, int first_chunk, int n_chunks, int in_src, int mlp_in,
                             int is_skip, int in_chunks, int out_rows, int perm_S, int perm_stride, int out_col0) {
  // one thread per (image, n, kk)
  const long long total = (long long)n_chunks * 2 * n * 16;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total + n; i += (long long)gridDim.x * blockDim.x) {
    if (i >= total) {
      int nn = (int)(i - total);
      int ncol = out_col0 + nn;  // output column (channel-major for the last layer)
      float v = 0.0f;
      if (ncol < out_rows) {
        int ns = perm_S > 0 ? (ncol % perm_S) * perm_stride + (ncol / perm_S) : ncol;
        v = b[ns];
      }
      bias_dst[nn] = v;
      continue;
    }
    int kk = (int)(i % 16);
    int nn = (int)((i / 16) % n);
    int img = (int)(i / (16LL * n));
    int c = first_chunk + img / 2, ks = img % 2;
    int kc = ks * 16 + kk;  // k inside the chunk
    // source column of the reference weight
    int ksrc = -1;
    if (c < in_chunks) {
      const int kin = c * 32 + kc;
      if (kin < mlp_in) ksrc = kin;  // encoded input (first layer, or the input part of the skip layer)
    } else {
      int hcol = (c - in_chunks) * 32 + kc;
      ksrc = is_skip ? mlp_in + hcol : hcol;
    }
    int ncol = out_col0 + nn;
    float w = 0.0f;
    if (ncol < out_rows && ksrc >= 0 && ksrc < in_src) {
      int ns = perm_S > 0 ? (ncol % perm_S) * perm_stride + (ncol / perm_S) : ncol;
      w = W[(long long)ns * in_src + ksrc];
    }
    __nv_bfloat16 hi = __float2bfloat16_rn(w);
    __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    size_t img_off = (size_t)img * n * 64;
    size_t slot = (size_t)(((kk >> 3) * (n >> 3) + (nn >> 3)) * 128 + (nn & 7) * 16 + (kk & 7) * 2);
    *reinterpret_cast<__nv_bfloat16*>(dst + img_off + slot) = hi;
    *reinterpret_cast<__nv_bfloat16*>(dst + img_off + (size_t)n * 32 + slot) = lo;
  }
}

void launch_pack_tc_pass(const float* W, const float* b, uint8_t* dst, float* bias_dst, int n, int first_chunk, int n_chunks,
                         int in_src, int mlp_in, int is_skip, int in_chunks, int out_rows, int perm_S, int perm_stride,
                         int out_col0, cudaStream_t st) {
  long long total = (long long)n_chunks * 2 * n * 16 + n;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  pack_tc_pass<<<grid, 256, 0, st>>>(W, b, dst, bias_dst, n, first_chunk, n_chunks, in_src, mlp_in, is_skip, in_chunks, out_rows,
                                     perm_S, perm_stride, out_col0);
}

// Tensor map of the heads scratch [n rays][mlp_out] fp32, box = box_cols columns x 32 rows (one epilogue warp's slice).
// encode_fn = cuTensorMapEncodeTiled, fetched once per handle through cudaGetDriverEntryPoint (no link-time libcuda
// dependency).  False = the map could not be built.
bool make_heads_map(CUtensorMap* hmap, void* encode_fn, const float* heads, int mlp_out, long long n, int box_cols) {
  memset(hmap, 0, sizeof(*hmap));
  if ((mlp_out % 4) != 0 || ((uintptr_t)heads % 16) != 0) return false;
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  if (encode_fn == nullptr) return false;
  EncodeFn encode = (EncodeFn)encode_fn;
  cuuint64_t gdim[2] = {(cuuint64_t)mlp_out, (cuuint64_t)n};
  cuuint64_t gstride[1] = {(cuuint64_t)mlp_out * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(hmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)heads, gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace hr
