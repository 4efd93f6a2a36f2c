// Microbenchmark: cost of tcgen05.mma (kind::f16, bf16, M=128) issued by one thread, for N in {128,256}, A from shared
// memory (SS) or tensor memory (TS).  Reports cycles per MMA at issue (clock after the issue loop) and at completion
// (after tcgen05.commit + mbarrier wait).  Operands are garbage; only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue umma_issue.cu && ./umma_issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int reps) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  asm volatile("fence.proxy.async.shared::cta;");
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 32) {
    const uint32_t sb = smem_u32(smem);
    const uint64_t a = umma_desc(sb, 2048, 128);                 // 128 rows x 16 k
    const uint64_t b = umma_desc(sb + 8192, (uint32_t)N * 16, 128);  // N rows x 16 k
    const uint32_t id = idesc(N);
    const uint32_t d = tm + 256;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (TS) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                     "r"(tm + (r & 7) * 8), "l"(b), "r"(id), "r"(r)
                     : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(a), "l"(b), "r"(id), "r"(r)
                     : "memory");
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar))
        : "memory");
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int N, bool TS>
void run(const char* name, long long* d_out, int sms) {
  const int reps = 256;
  cudaFuncSetAttribute(bench<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  for (int grid : {1, sms}) {
    bench<N, TS><<<grid, 128, 48 * 1024>>>(d_out, reps);
    bench<N, TS><<<grid, 128, 48 * 1024>>>(d_out, reps);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-22s grid=%3d  issue %.1f cyc/MMA   complete %.1f cyc/MMA   (%s)\n", name, grid, (double)h[0] / reps, (double)h[1] / reps,
           cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  run<256, false>("M128 N256 SS", d_out, p.multiProcessorCount);
  run<128, false>("M128 N128 SS", d_out, p.multiProcessorCount);
  run<64, false>("M128 N64  SS", d_out, p.multiProcessorCount);
  run<256, true>("M128 N256 TS", d_out, p.multiProcessorCount);
  run<128, true>("M128 N128 TS", d_out, p.multiProcessorCount);
  run<64, true>("M128 N64  TS", d_out, p.multiProcessorCount);
  return 0;
}
