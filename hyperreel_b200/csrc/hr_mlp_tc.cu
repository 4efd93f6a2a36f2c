// Sample-prediction network on the 5th-generation tensor cores (HR_MLP_BF16X3_TC).
//
// Same math as hr_mlp_simt.cu (reference: nlf/nets/mlp.py:159-172 behind nlf/embedding/ray.py:320-326), but
// every fp32 operand x is split into bf16 hi = rn(x) and lo = rn(x - hi) and each Linear layer is
//   D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo          (fp32 accumulation in TMEM)
// i.e. three tcgen05.mma (kind::f16, bf16 inputs) per k-step; the dropped A_lo*B_lo term and the split
// residuals are O(2^-16) relative per product (DESIGN.md "precision of the tensor-core sample net").
//
// FIRST LAYOUT, kept selectable with HR_TC_V=1 for A/B measurements; the product path is hr_mlp_tc2.cu.
// One persistent CTA per SM, one 128-ray tile at a time (UMMA M = 128, one TMEM lane per ray), 320 threads:
//   warps 0-7  epilogue (two groups of four warps, thread = ray): encode the ray, then per layer read the accumulator from
//              TMEM (tcgen05.ld 32x32b), add bias, LeakyReLU, split to bf16 hi/lo and write the next layer's A operand
//              (hi AND lo, in place) into shared memory in the UMMA K-major no-swizzle layout; last layer: 16-column slices
//              staged per warp and written with TMA tensor stores.
//   warp 8     producer: streams the pre-packed weight images (one 16-wide k-step of one full-width pass, N <= 256)
//              through a 3-stage ring with cp.async.bulk + mbarrier complete_tx.
//   warp 9     MMA issuer: one lane issues tcgen05.mma / tcgen05.commit, trailing the epilogue chunk by chunk (a_ready
//              barriers per 32-column chunk), two 256-column TMEM accumulators ping-ponged across layers.
// Measured: 81 K cycles per tile (the in-place activation operand serialises layer l+1's MMAs behind layer l's epilogue
// and the one-lane issue loop costs ~150 cycles per k-step of overhead); see profiles/r1_notes.md.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hr_encode.cuh"
#include "hr_handle.h"
#include "hr_tc_prims.cuh"

namespace hr {

namespace tc {

constexpr int NSTAGE = 3;          // weight ring depth (weights were never late with 4; 3 frees 16 KB for staging)
constexpr int STAGE_BYTES = 16384; // one k-step image: N<=256 rows x 16 k x (hi+lo) bf16
constexpr int CHUNK_BYTES = 8192;  // one A chunk: 128 rows x 32 k bf16
constexpr int NCHUNK = 9;          // chunk 0 = encoded input (32, zero padded), chunks 1..8 = hidden 256
constexpr int EPI_GROUPS = 2;     // epilogue warp groups (4 warps each, one per TMEM lane quadrant)
constexpr int EPI_WARPS = 4 * EPI_GROUPS;
constexpr int NTHREADS = (EPI_WARPS + 2) * 32;
constexpr int BIAS_FLOATS = 2560;

// shared memory map (bytes)
constexpr int OFF_A_HI = 0;
constexpr int OFF_A_LO = OFF_A_HI + NCHUNK * CHUNK_BYTES;   // 73728
constexpr int OFF_B = OFF_A_LO + NCHUNK * CHUNK_BYTES;      // 147456
constexpr int OFF_BIAS = OFF_B + NSTAGE * STAGE_BYTES;      // 212992
constexpr int OFF_BAR = OFF_BIAS + BIAS_FLOATS * 4;         // 223232
constexpr int OFF_STG = OFF_BAR + 256;                      // last-layer transpose staging, 2 KB per epilogue warp
constexpr int SMEM_BYTES = OFF_STG + EPI_WARPS * 2048;      // 223488 (of 232448 available)

// barrier slots (8 bytes each) inside OFF_BAR
constexpr int BAR_FULL = 0;                 // [NSTAGE]
constexpr int BAR_EMPTY = BAR_FULL + NSTAGE;   // [NSTAGE]
constexpr int BAR_AREADY = BAR_EMPTY + NSTAGE; // [NCHUNK]
constexpr int BAR_DFULL = BAR_AREADY + NCHUNK; // [2]
constexpr int BAR_DEMPTY = BAR_DFULL + 2;      // [2]
constexpr int BAR_TMEMPTR = BAR_DEMPTY + 2;    // 4-byte TMEM base address lives in this slot

}  // namespace tc

// CS = cluster size: the CS CTAs of a cluster walk the weight stream in lock step; each loads 1/CS of every image
// and multicasts it to all of them, so L2 -> SM weight traffic drops by CS.
template <int CS>
__global__ void __launch_bounds__(tc::NTHREADS, 1)
mlp_tc_kernel(const __grid_constant__ hr_config cfg, const __grid_constant__ MlpTcPack pk, const float* __restrict__ rays,
              float* __restrict__ heads, long long n_rays, int dbg_products, int dbg_load_lo, unsigned long long* trace,
              const __grid_constant__ CUtensorMap heads_map, int use_tma_store) {
  using namespace tc;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  auto bar = [&](int slot) -> uint32_t { return sbase + OFF_BAR + slot * 8; };
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + BAR_TMEMPTR * 8);

  // ---- one-time setup ----
  for (int i = tid; i < pk.bias_count; i += NTHREADS) s_bias[i] = pk.bias[i];
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar(BAR_FULL + s), 1); mbar_init(bar(BAR_EMPTY + s), CS); }
    mbar_init(bar(BAR_AREADY + 0), 128);  // chunk 0: the encoder (group 0)
    for (int c = 1; c < NCHUNK; ++c) mbar_init(bar(BAR_AREADY + c), 128 * EPI_GROUPS);  // every group writes a slice
    for (int d = 0; d < 2; ++d) { mbar_init(bar(BAR_DFULL + d), 1); mbar_init(bar(BAR_DEMPTY + d), 128 * EPI_GROUPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPI_WARPS + 1) {
    uint32_t dst = sbase + OFF_BAR + BAR_TMEMPTR * 8;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast can target them
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // Let the dependent (render) grid start launching now: its CTAs run their prologue and park in griddepcontrol.wait
  // until this grid has completed and flushed, so their launch latency is hidden behind this kernel.
  if (dbg_products & 0x100) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  dbg_products &= 0xff;

  const long long n_tiles = (n_rays + BM - 1) / BM;
  const int n_passes = pk.n_passes;
  // diagnostic timeline (HR_TC_TRACE): CTA 0, second tile; slot = (pass * 8 + event)
  const bool tracing = (trace != nullptr) && (blockIdx.x == 0);
  auto TR = [&](long long iter, int pass, int ev) {
    if (tracing && iter == 1) trace[pass * 8 + ev] = clock64();
  };
  // every CTA of a cluster runs the same number of iterations (tiles past the end are fully masked)
  const uint32_t crank = (CS > 1) ? cluster_ctarank() : 0u;
  const long long n_clusters = gridDim.x / CS;
  const long long cluster_id = blockIdx.x / CS;
  const long long n_iters = (n_tiles + n_clusters * CS - 1) / (n_clusters * CS);
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);

  if (warp == EPI_WARPS) {
    // =========================== producer: weight images, in consumption order ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long iter = 0; iter < n_iters; ++iter) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(pk.wpack);
        for (int p = 0; p < n_passes; ++p) {
          const uint32_t bytes = (uint32_t)pk.passes[p].n * 64u;
          const int n_img = pk.passes[p].n_chunks * 2;
          for (int i = 0; i < n_img; ++i, ++it) {
            const uint32_t s = it % NSTAGE;
            mbar_wait(bar(BAR_EMPTY + s), ((it / NSTAGE) & 1) ^ 1);
            const uint32_t ld = dbg_load_lo ? bytes : bytes / 2;  // (diagnostic knob: skip the lo halves)
            mbar_expect_tx(bar(BAR_FULL + s), ld);
            if constexpr (CS == 1) {
              bulk_g2s(sbase + OFF_B + s * STAGE_BYTES, src, ld, bar(BAR_FULL + s));
            } else {
              const uint32_t slice = ld / CS;
              bulk_g2s_mc(sbase + OFF_B + s * STAGE_BYTES + crank * slice, src + crank * slice, slice, bar(BAR_FULL + s), kMask);
            }
            src += bytes;
          }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // =========================== MMA issuer ===========================
    // One thread feeds the tensor core, so its own instruction stream is on the critical path: descriptors are
    // built once and advanced with adds, stage/phase are kept incrementally, and the three split products plus the
    // stage-release commit of a k-step go out in a single asm block.
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, gp = 0, titer = 0;
      const int n_hidden = cfg.mlp_layers - 1;
      // descriptor of the first activation k-step / first weight stage; later ones differ only in the start address
      const uint64_t hdesc_hi0 = umma_desc(sbase + OFF_A_HI, 2048, 128);
      const uint64_t hdesc_lo0 = umma_desc(sbase + OFF_A_LO, 2048, 128);
      for (long long iter = 0; iter < n_iters; ++iter, ++titer) {
        for (int p = 0; p < n_passes; ++p, ++gp) {
          const TcPass& P = pk.passes[p];
          const uint32_t db = gp & 1, use = gp >> 1;
          const uint32_t d_tmem = tmem_base + db * 256;
          const uint32_t idesc = umma_idesc(P.n);
          const uint64_t wdesc_hi0 = umma_desc(sbase + OFF_B, (uint32_t)P.n * 16, 128);
          const uint64_t wdesc_lo0 = umma_desc(sbase + OFF_B + (uint32_t)P.n * 32, (uint32_t)P.n * 16, 128);
          mbar_wait(bar(BAR_DEMPTY + db), (use & 1) ^ 1);  // accumulator drained by its previous reader
          TR(iter, p, 0);
          uint32_t acc = 0;
          for (int ci = 0; ci < P.n_chunks; ++ci) {
            const int c = P.first_chunk + ci;
            if (P.wait_a) {
              // chunk 0: written once per tile by the encoder; chunks 1..8: once per hidden layer
              const uint32_t done = (c == 0) ? titer : (titer * n_hidden + (uint32_t)(P.layer - 1));
              mbar_wait(bar(BAR_AREADY + c), done & 1);
              if (ci == 0) TR(iter, p, 1);
              if (ci == P.n_chunks - 1) TR(iter, p, 2);
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const int kidx = ci * 2 + ks;
              const bool ktr = tracing && iter == 1 && (p == 2 || p == 6) && kidx < 16;
              const int kbase = 128 + (p == 6 ? 64 : 0) + kidx * 4;
              if (ktr) trace[kbase + 0] = clock64();
              mbar_wait(bar(BAR_FULL + stage), phase);
              if (ktr) trace[kbase + 1] = clock64();
              tc_fence_after();
              const uint64_t hoff = (uint64_t)((c * CHUNK_BYTES + ks * 4096) >> 4);
              const uint64_t woff = (uint64_t)((stage * STAGE_BYTES) >> 4);
              const uint64_t h_hi = hdesc_hi0 + hoff, h_lo = hdesc_lo0 + hoff;  // activations, 128 rays x 16 k
              const uint64_t w_hi = wdesc_hi0 + woff, w_lo = wdesc_lo0 + woff;  // weights, P.n rows x 16 k
              // D[ray, col] += H * W^T : A = activations (M = 128 rays), B = weights (N = P.n columns)
              const uint64_t a0 = h_hi, b0 = w_hi;  // hi * hi
              const uint64_t a1 = h_lo, b1 = w_hi;  // (act lo) * (weight hi)
              const uint64_t a2 = h_hi, b2 = w_lo;  // (act hi) * (weight lo)
              const uint32_t ebar = bar(BAR_EMPTY + stage);
              if (dbg_products >= 3) {
                if constexpr (CS == 1) {
                  asm volatile(
                      "{\n\t"
                      ".reg .pred p, q;\n\t"
                      "setp.ne.b32 p, %8, 0;\n\t"
                      "setp.eq.b32 q, %8, %8;\n\t"
                      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %7, p;\n\t"
                      "tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %7, q;\n\t"
                      "tcgen05.mma.cta_group::1.kind::f16 [%0], %5, %6, %7, q;\n\t"
                      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
                      "}" ::"r"(d_tmem),
                      "l"(a0), "l"(b0), "l"(a1), "l"(b1), "l"(a2), "l"(b2), "r"(idesc), "r"(acc), "r"(ebar)
                      : "memory");
                } else {
                  asm volatile(
                      "{\n\t"
                      ".reg .pred p, q;\n\t"
                      "setp.ne.b32 p, %8, 0;\n\t"
                      "setp.eq.b32 q, %8, %8;\n\t"
                      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %7, p;\n\t"
                      "tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %7, q;\n\t"
                      "tcgen05.mma.cta_group::1.kind::f16 [%0], %5, %6, %7, q;\n\t"
                      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%9], %10;\n\t"
                      "}" ::"r"(d_tmem),
                      "l"(a0), "l"(b0), "l"(a1), "l"(b1), "l"(a2), "l"(b2), "r"(idesc), "r"(acc), "r"(ebar), "h"(kMask)
                      : "memory");
                }
              } else {  // diagnostic: fewer products
                umma_bf16(d_tmem, a0, b0, idesc, acc);
                if (dbg_products >= 2) umma_bf16(d_tmem, a1, b1, idesc, 1u);
                if constexpr (CS == 1) umma_commit(ebar);
                else umma_commit_mc(ebar, kMask);
              }
              acc = 1;
              if (ktr) trace[kbase + 2] = clock64();
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
          }
          umma_commit(bar(BAR_DFULL + db));  // accumulator complete -> epilogue
          TR(iter, p, 3);
        }
      }
    }
  } else {
    // =========================== epilogue warps: thread = ray ===========================
    // Two warps share each TMEM lane quadrant (warp w and w+4) and alternate over the 32-column chunks, so two
    // epilogue warps are resident per scheduler and each chunk's latency chain overlaps the other group's.
    const int grp = warp >> 2;                  // 0 .. EPI_GROUPS-1
    const int row = (warp & 3) * 32 + lane;     // TMEM lane == ray within the tile
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    uint32_t gp = 0;
    auto encode_tile = [&](long long ray_) {
      float enc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) enc[i] = 0.0f;
      if (ray_ < n_rays) encode_ray(cfg, rays + ray_ * cfg.c_in, enc, 1);
#pragma unroll
      for (int kg = 0; kg < 4; ++kg) {
        uint4 hi, lo;
        split8(enc + kg * 8, hi, lo);
        *reinterpret_cast<uint4*>(smem + OFF_A_HI + a_slot(row, kg)) = hi;
        *reinterpret_cast<uint4*>(smem + OFF_A_LO + a_slot(row, kg)) = lo;
      }
      fence_async_smem();
      mbar_arrive(bar(BAR_AREADY + 0));
    };
    const int last_hidden = cfg.mlp_layers - 2;
    for (long long iter = 0; iter < n_iters; ++iter) {
      const long long tile = (iter * n_clusters + cluster_id) * CS + crank;  // may be >= n_tiles: fully masked
      const long long ray = tile * BM + row;
      // ---- encode (RayParam + WindowedPE), split, write chunk 0 (group 0 only) ----
      // The first tile is encoded here; every later tile was already encoded during the previous tile's last layer.
      if (grp == 0 && iter == 0) encode_tile(ray);
      for (int p = 0; p < n_passes; ++p, ++gp) {
        const TcPass& P = pk.passes[p];
        const uint32_t db = gp & 1, use = gp >> 1;
        mbar_wait(bar(BAR_DFULL + db), use & 1);
        if (tid == 0) TR(iter, p, 4);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + lane_base + db * 256;
        const float* bias = s_bias + P.bias_off;
        if (!P.is_final) {
          // Hidden layer: the groups sweep the eight 32-column chunks together, each taking 32/EPI_GROUPS columns of every
          // chunk, so chunk j of the next A operand is complete (a_ready[1+j]) after 1/8 of the epilogue and the next
          // layer's MMAs trail the epilogue chunk by chunk.
          constexpr int CW = 32 / EPI_GROUPS;  // columns per group per chunk (16)
          for (int j = 0; j < 8; ++j) {
            uint32_t v[16];
            tmem_ld16(t_addr + j * 32 + grp * CW, v);
            if (j == 7) { tc_fence_before(); mbar_arrive(bar(BAR_DEMPTY + db)); }
            const float4* b4 = reinterpret_cast<const float4*>(bias + j * 32 + grp * CW);
#pragma unroll
            for (int kg = 0; kg < CW / 8; ++kg) {
              const float4 ba = b4[kg * 2], bb = b4[kg * 2 + 1];
              const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float t = __uint_as_float(v[kg * 8 + i]) + bv[i];
                x[i] = fmaxf(t, t * cfg.leaky_slope);  // LeakyReLU, slope in (0,1)
              }
              uint4 hi, lo;
              split8(x, hi, lo);
              const uint32_t off = (1 + j) * CHUNK_BYTES + a_slot(row, grp * (CW / 8) + kg);
              *reinterpret_cast<uint4*>(smem + OFF_A_HI + off) = hi;
              *reinterpret_cast<uint4*>(smem + OFF_A_LO + off) = lo;
            }
            fence_async_smem();
            mbar_arrive(bar(BAR_AREADY + 1 + j));
            if (tid == 0 && j == 0) TR(iter, p, 5);
            if (tid == 0 && j == 7) TR(iter, p, 6);
          }
          // Chunk 0 (the encoded input) is read by the first and the skip layer only; once the last hidden layer's
          // accumulator is complete both have retired, so the next tile's rays are encoded now, under the last layer's
          // MMAs, instead of after its epilogue.
          if (P.layer == last_hidden && grp == 0 && iter + 1 < n_iters) {
            const long long ntile = ((iter + 1) * n_clusters + cluster_id) * CS + crank;
            encode_tile(ntile * BM + row);
          }
        } else {
          // Last layer: 16-column slices are transposed through a 2 KB per-warp staging area so that every global store
          // instruction writes two 64-byte row segments.
          float* stg = reinterpret_cast<float*>(smem + OFF_STG) + warp * 512;
          const int nslice = (P.n + 15) / 16;
          const int last_h = (grp < nslice) ? ((nslice - 1 - grp) / EPI_GROUPS) * EPI_GROUPS + grp : -1;
          if (last_h < 0) { tc_fence_before(); mbar_arrive(bar(BAR_DEMPTY + db)); }
          const int half = lane >> 4, cidx = lane & 15;  // store mapping: 2 rows x 16 columns per instruction
          const long long row0 = tile * BM + (warp & 3) * 32;
          if (use_tma_store) {
            // TMEM -> registers -> 32 x 16 box in shared memory (row = lane, 64-byte rows) -> one TMA tensor store per box;
            // rows past n_rays / columns past mlp_out are clipped by the tensor map.
            const uint32_t stg_s = sbase + OFF_STG + warp * 2048;
            for (int h = grp; h < nslice; h += EPI_GROUPS) {
              uint32_t v[16];
              tmem_ld16(t_addr + h * 16, v);
              if (h == last_h) { tc_fence_before(); mbar_arrive(bar(BAR_DEMPTY + db)); }
              const float4* b4 = reinterpret_cast<const float4*>(bias + h * 16);
              // the previous box of this warp must have been read out of the staging buffer
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 b = b4[i4];
                float4 o;
                o.x = __uint_as_float(v[i4 * 4 + 0]) + b.x;
                o.y = __uint_as_float(v[i4 * 4 + 1]) + b.y;
                o.z = __uint_as_float(v[i4 * 4 + 2]) + b.z;
                o.w = __uint_as_float(v[i4 * 4 + 3]) + b.w;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg_s + lane * 64 + i4 * 16), "f"(o.x), "f"(o.y),
                             "f"(o.z), "f"(o.w)
                             : "memory");
              }
              fence_async_smem();
              __syncwarp();
              if (lane == 0) {
                const int x = P.out_col0 + h * 16;
                const int y = (int)(tile * BM + (warp & 3) * 32);
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                 reinterpret_cast<uint64_t>(&heads_map)),
                             "r"(x), "r"(y), "r"(stg_s)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
          } else {
          for (int h = grp; h < nslice; h += EPI_GROUPS) {
            uint32_t v[16];
            tmem_ld16(t_addr + h * 16, v);
            if (h == last_h) { tc_fence_before(); mbar_arrive(bar(BAR_DEMPTY + db)); }
            const float4* b4 = reinterpret_cast<const float4*>(bias + h * 16);
            // row `lane` keeps element i at stg[lane*16 + (i ^ ((lane>>1)&15))]: conflict-free for both phases
            const int sw = (lane >> 1) & 15;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 b = b4[i4];
              stg[lane * 16 + ((i4 * 4 + 0) ^ sw)] = __uint_as_float(v[i4 * 4 + 0]) + b.x;
              stg[lane * 16 + ((i4 * 4 + 1) ^ sw)] = __uint_as_float(v[i4 * 4 + 1]) + b.y;
              stg[lane * 16 + ((i4 * 4 + 2) ^ sw)] = __uint_as_float(v[i4 * 4 + 2]) + b.z;
              stg[lane * 16 + ((i4 * 4 + 3) ^ sw)] = __uint_as_float(v[i4 * 4 + 3]) + b.w;
            }
            __syncwarp();
            const int col = P.out_col0 + h * 16 + cidx;
            const bool col_ok = (h * 16 + cidx < P.n) && (col < cfg.mlp_out);
            float* dst = heads + (row0 + half) * (long long)cfg.mlp_out + col;
#pragma unroll
            for (int rr = 0; rr < 32; rr += 2) {
              const int r = rr + half;
              const float o = stg[r * 16 + (cidx ^ ((r >> 1) & 15))];
              if (col_ok && row0 + r < n_rays) dst[(long long)rr * cfg.mlp_out] = o;
            }
            __syncwarp();
          }
          }
        }
      }
    }
  }

  // ---- teardown ----
  if (warp < EPI_WARPS && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // TMA stores landed
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();  // no CTA leaves while a peer can still signal its barriers
  if (warp == EPI_WARPS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight packing: fp32 reference weights -> bf16 hi/lo images in UMMA K-major no-swizzle layout, consumption order
// ------------------------------------------------------------------------------------------------------------------
__global__ void pack_tc_pass(const float* __restrict__ W, const float* __restrict__ b, uint8_t* __restrict__ dst,
                             float* __restrict__ bias_dst, int n, int first_chunk, int n_chunks, int in_src, int mlp_in,
                             int is_skip, int is_first, int out_rows, int perm_S, int perm_stride, int out_col0) {
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
    if (c == 0) {
      if (kc < mlp_in) ksrc = kc;  // encoded input (first layer, or the input part of the skip layer)
    } else {
      int hcol = (c - 1) * 32 + kc;
      ksrc = is_skip ? mlp_in + hcol : hcol;
    }
    (void)is_first;
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


// Tensor map of the heads scratch [n rays][mlp_out] fp32, box = box_cols columns x 32 rows (one epilogue warp's slice),
// optionally with the 128-byte shared-memory swizzle (box_cols = 32).
// The encoder comes from the driver through the runtime (no link-time libcuda dependency).  False = not available.
bool make_heads_map(CUtensorMap* hmap, const float* heads, int mlp_out, long long n, int box_cols, bool swizzle128) {
  memset(hmap, 0, sizeof(*hmap));
  if ((mlp_out % 4) != 0 || ((uintptr_t)heads % 16) != 0) return false;
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = (EncodeFn)fn;
  }
  if (!encode) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)mlp_out, (cuuint64_t)n};
  cuuint64_t gstride[1] = {(cuuint64_t)mlp_out * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(hmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)heads, gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

void launch_pack_tc_pass(const float* W, const float* b, uint8_t* dst, float* bias_dst, int n, int first_chunk, int n_chunks,
                         int in_src, int mlp_in, int is_skip, int out_rows, int perm_S, int perm_stride, int out_col0,
                         cudaStream_t st) {
  long long total = (long long)n_chunks * 2 * n * 16 + n;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  pack_tc_pass<<<grid, 256, 0, st>>>(W, b, dst, bias_dst, n, first_chunk, n_chunks, in_src, mlp_in, is_skip, 0, out_rows, perm_S,
                                     perm_stride, out_col0);
}

int pack_mlp_tc(hr_handle* h, const hr_params*, const float* const* w_dev, const float* const* b_dev, cudaStream_t st) {
  const hr_config& c = h->cfg;
  MlpTcPack& pk = h->tc;
  memset(&pk, 0, sizeof(pk));
  pk.version = 1;
  if (c.mlp_width != 256) return hr_fail("tensor-core sample net: width must be 256");
  if (c.mlp_in > 32) return hr_fail("tensor-core sample net: encoded input wider than 32");
  const int L = c.mlp_layers;
  int np = 0, bias_off = 0;
  size_t bytes = 0;
  for (int l = 0; l < L; ++l) {
    const bool last = (l == L - 1);
    const int n_parts = last ? (c.mlp_out + 255) / 256 : 1;
    const int n_each = last ? (((c.mlp_out + n_parts - 1) / n_parts + 15) / 16 * 16) : 256;
    for (int part = 0; part < n_parts; ++part) {
      if (np >= HR_TC_MAX_PASSES) return hr_fail("tensor-core sample net: too many passes");
      TcPass& P = pk.passes[np++];
      P.layer = l;
      P.n = n_each;
      P.first_chunk = (l == 0 || l == c.mlp_skip) ? 0 : 1;
      P.n_chunks = (l == 0) ? 1 : (l == c.mlp_skip ? 9 : 8);
      P.bias_off = bias_off;
      P.is_final = last ? 1 : 0;
      P.out_col0 = part * n_each;
      P.wait_a = (part == 0) ? 1 : 0;
      bias_off += n_each;
      bytes += (size_t)P.n_chunks * 2 * P.n * 64;
    }
  }
  if (bias_off > tc::BIAS_FLOATS) return hr_fail("tensor-core sample net: bias table too large");
  pk.n_passes = np;
  pk.bias_count = bias_off;
  uint8_t* wp = nullptr;
  float* bp = nullptr;
  cudaError_t e = cudaMalloc((void**)&wp, bytes);
  if (e != cudaSuccess) return hr_fail("cudaMalloc(tc weights %zu): %s", bytes, cudaGetErrorString(e));
  h->owned.push_back(wp);
  e = cudaMalloc((void**)&bp, (size_t)bias_off * sizeof(float));
  if (e != cudaSuccess) return hr_fail("cudaMalloc(tc bias): %s", cudaGetErrorString(e));
  h->owned.push_back(bp);
  size_t off = 0;
  for (int p = 0; p < np; ++p) {
    const TcPass& P = pk.passes[p];
    const int l = P.layer;
    const bool last = (l == L - 1), skip = (l == c.mlp_skip), first = (l == 0);
    const int in_src = first ? c.mlp_in : (skip ? c.mlp_in + 256 : 256);
    const int out_rows = last ? c.mlp_out : 256;
    launch_pack_tc_pass(w_dev[l], b_dev[l], wp + off, bp + P.bias_off, P.n, P.first_chunk, P.n_chunks, in_src, c.mlp_in,
                        skip ? 1 : 0, out_rows, last ? c.n_samples : 0, c.head_stride, P.out_col0, st);
    (void)first;
    off += (size_t)P.n_chunks * 2 * P.n * 64;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return hr_fail("tc pack launch failed: %s", cudaGetErrorString(e));
  pk.wpack = wp;
  pk.bias = bp;
  pk.wpack_bytes = (long long)bytes;
  return 0;
}

template <int CS>
static cudaError_t launch_mlp_tc_cs(const hr_config& cfg, const MlpTcPack& pk, const float* rays, float* heads, long long n,
                                    int num_sms, cudaStream_t stream, int dbg_products, int dbg_load_lo) {
  unsigned long long* trace = nullptr;
  const bool want_trace = getenv("HR_TC_TRACE") != nullptr;
  if (want_trace) {
    cudaMalloc((void**)&trace, 256 * sizeof(unsigned long long));
    cudaMemset(trace, 0, 256 * sizeof(unsigned long long));
  }
  long long tiles = (n + tc::BM - 1) / tc::BM;
  long long want = (tiles + CS - 1) / CS * CS;
  long long cap = (long long)(num_sms / CS) * CS;
  int grid = (int)(want < cap ? want : cap);
  if (grid < CS) grid = CS;
  static bool attr_set[64] = {false};  // per template instance and device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(tc::NTHREADS);
  lc.dynamicSmemBytes = tc::SMEM_BYTES;
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  // tensor map of the heads scratch [n rays][mlp_out] fp32 for the last layer's TMA stores (box 16 cols x 32 rows)
  CUtensorMap hmap;
  static const int want_tma = getenv("HR_TC_TMA_STORE") ? atoi(getenv("HR_TC_TMA_STORE")) : 1;
  const int use_tma = (want_tma && make_heads_map(&hmap, heads, cfg.mlp_out, n, 16, false)) ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&lc, mlp_tc_kernel<CS>, cfg, pk, rays, heads, n, dbg_products, dbg_load_lo, trace, hmap, use_tma);
  if (want_trace) {
    unsigned long long h[256];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(trace);
    unsigned long long t0 = h[0];
    fprintf(stderr, "[tc-trace] pass: demp_ok a_first a_last commit | dfull_seen first_chunk last_chunk (cycles rel. to pass 0)\n");
    for (int p = 0; p < pk.n_passes; ++p) {
      fprintf(stderr, "[tc-trace] %2d:", p);
      for (int e = 0; e < 7; ++e) fprintf(stderr, " %8lld", h[p * 8 + e] ? (long long)(h[p * 8 + e] - t0) : -1LL);
      fprintf(stderr, "\n");
    }
    for (int blk = 0; blk < 2; ++blk) {
      fprintf(stderr, "[tc-trace] k-steps of pass %d: (before_wait_full, wait_cost, issue_cost) rel. to first\n[tc-trace]  ", blk ? 6 : 2);
      unsigned long long b0 = h[128 + blk * 64];
      for (int k = 0; k < 16; ++k) {
        const unsigned long long* e = h + 128 + blk * 64 + k * 4;
        fprintf(stderr, "(%lld,%lld,%lld) ", (long long)(e[0] - b0), (long long)(e[1] - e[0]), (long long)(e[2] - e[1]));
      }
      fprintf(stderr, "\n");
    }
  }
  return le;
}

cudaError_t launch_mlp_tc(const hr_config& cfg, const MlpTcPack& pk, const float* rays, float* heads, long long n,
                          int num_sms, cudaStream_t stream) {
  // diagnostic knobs (profiling only; defaults = the product path)
  static const int dbg_products = getenv("HR_TC_PRODUCTS") ? atoi(getenv("HR_TC_PRODUCTS")) : 3;
  static const int dbg_load_lo = getenv("HR_TC_LOAD_LO") ? atoi(getenv("HR_TC_LOAD_LO")) : 1;
  static const int pdl_early = getenv("HR_PDL_EARLY") ? atoi(getenv("HR_PDL_EARLY")) : 0;  // measured: parked dependents slow this kernel
  const int prod_flags = dbg_products | (pdl_early ? 0x100 : 0);
  static const int cluster = getenv("HR_TC_CLUSTER") ? atoi(getenv("HR_TC_CLUSTER")) : 1;
  if (cluster == 4) return launch_mlp_tc_cs<4>(cfg, pk, rays, heads, n, num_sms, stream, prod_flags, dbg_load_lo);
  if (cluster == 2) return launch_mlp_tc_cs<2>(cfg, pk, rays, heads, n, num_sms, stream, prod_flags, dbg_load_lo);
  return launch_mlp_tc_cs<1>(cfg, pk, rays, heads, n, num_sms, stream, prod_flags, dbg_load_lo);
}

}  // namespace hr
