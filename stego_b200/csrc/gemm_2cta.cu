// 2-CTA (tcgen05 cta_group::2) GEMM for the wide ViT linears (qkv, fc1):   out[M,N] = act(A[M,K] . B[N,K]^T + bias)
//
// A thread-block cluster of two CTAs (one SM pair) computes one 256 x BN output tile with UMMA M = 256: each CTA
// stages its OWN 128 rows of A and only HALF of the B tile (BN/2 rows); the tensor cores of the pair read both halves.
// Per k-step a CTA therefore receives 16 KB (A) + 16 KB (B half) instead of the 16 + 32 KB of the 1-CTA 128 x 256
// kernel — the 1-CTA kernel measured ~0.65 us per k-step regardless of stage count or multicast, i.e. bound by what
// one SM can take in, not by the tensor pipe (MMA floor 0.26 us per k-step) — and the freed shared memory buys a
// deeper ring (5 stages x 32 KB).
//
//   warp 0      TMA producer (both CTAs): A rows of this CTA + this CTA's half of B; completion is signalled on the
//               LEADER CTA's full barrier (cp.async.bulk.tensor ... .cta_group::2, mbarrier address mapped with mapa)
//   warp 1      leader CTA only: issues tcgen05.mma.cta_group::2 (one thread for the pair); tcgen05.commit multicast
//               releases the smem stage in BOTH CTAs and publishes the accumulator to BOTH CTAs' epilogues
//   warps 2..9  epilogue (both CTAs): own 128 x BN accumulator half from own TMEM -> bias/act -> TMA store;
//               all 16 epilogue warps of the pair arrive on the leader's tmem-empty barrier (remote mbarrier arrive)
#include "common.cuh"
#include "epilogue.cuh"
#include "host_util.h"

namespace stego {

constexpr int G2_BM = 128;  // rows per CTA (256 per cluster)
constexpr int G2_BK = 64;
constexpr int G2_THREADS = 320;

struct Gemm2Params {
  int M, N, K;
  EpiArgs epi;
};

__device__ __forceinline__ uint32_t mapa_cluster(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion (complete_tx) lands on an mbarrier that may live in the PEER CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t tmem_base) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(kCols) : "memory");
}

template <int BN, int kStages>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, Gemm2Params p) {
  static_assert(BN == 256, "one 256-column accumulator per buffer (two buffers = 512 TMEM columns)");
  constexpr uint32_t A_BYTES = G2_BM * G2_BK * 2;      // 16 KB: this CTA's rows
  constexpr uint32_t B_BYTES = (BN / 2) * G2_BK * 2;   // 16 KB: this CTA's half of the B tile
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t kEpiBufs = 2;
  constexpr uint32_t EPI_BYTES = 8 * kEpiBufs * 4096;
  constexpr uint32_t IDESC = make_idesc_bf16(256, BN, 0, 0);  // UMMA M = 256 across the CTA pair

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + kStages * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);  // used in the leader CTA
  uint64_t* empty_bar = full_bar + kStages;                                 // per CTA
  uint64_t* tfull_bar = empty_bar + kStages;                                // [2] per CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                     // [2] used in the leader CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  const int pairs_m = (p.M + 2 * G2_BM - 1) / (2 * G2_BM);
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + G2_BK - 1) / G2_BK;
  const int total_tiles = pairs_m * tiles_n;
  const int sched_start = static_cast<int>(blockIdx.x >> 1);
  const int sched_step = static_cast<int>(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // one arrive.expect_tx by the leader's producer; bytes arrive from both CTAs
      mbar_init(&empty_bar[s], 1);  // one multicast tcgen05.commit per use
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);  // the 8 epilogue warps of each CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are live before any remote complete_tx / arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = sched_start; t < total_tiles; t += sched_step) {
        const int tn = t % tiles_n;
        const int row0 = ((t / tiles_n) * 2 + static_cast<int>(crank)) * G2_BM;
        const int brow0 = tn * BN + static_cast<int>(crank) * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t full_leader = mapa_cluster(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_2sm(sa, &tmA, full_leader, kb * G2_BK, row0);
          tma_load_2d_2sm(sb, &tmB, full_leader, kb * G2_BK, brow0);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // whole warp, uniform control flow (descriptors stay in uniform registers); one elected lane issues
    if (leader) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t DESC_HI = smem_desc_hi_sw128(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), 16);
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem) + A_BYTES, 16);
      for (int t = sched_start; t < total_tiles; t += sched_step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);  // both CTAs have drained this accumulator buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);  // A rows + B halves of BOTH CTAs have landed
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (STAGE_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (uint32_t k = 0; k < G2_BK / 16; ++k)
              umma_bf16_2sm(tmem_d, smem_desc_join(a_lo + 2 * k, DESC_HI), smem_desc_join(b_lo + 2 * k, DESC_HI), IDESC,
                            (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage], 0x3);  // the stage is reusable in both CTAs once these MMAs retire
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_2sm(&tfull_bar[acc], 0x3);  // accumulator complete -> both epilogues
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps (2..9, both CTAs) =====================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t acc = 0, acc_phase = 0, epi_groups = 0;
    uint8_t* buf0 = epi_smem + (warp - 2) * (kEpiBufs * 4096);
    for (int t = sched_start; t < total_tiles; t += sched_step) {
      const int tn = t % tiles_n;
      const int row_base = ((t / tiles_n) * 2 + static_cast<int>(crank)) * G2_BM + quarter * 32;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      epilogue_tma_tile<BN, kEpiBufs, 2>(&tmO, p.epi, taddr, buf0, epi_groups, half, lane, tn * BN, row_base);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty_bar[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_wait_group_read<0>();  // staging smem must outlive the bulk stores
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer can still address it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

}  // namespace stego

using namespace stego;

// Internal entry (called from stego_gemm_bf16 in gemm.cu): K-major A and B, bias/activation, plain TMA store or fp32
// reduce-add epilogue.  Preconditions are checked by the caller.
int stego_launch_gemm_2cta(const void* A, int lda, const void* B, int ldb, int M, int N, int K, void* out, int ldo,
                           int out_bf16, const float* bias, int act, int reduce_add, cudaStream_t stream) {
  constexpr int BN = 256, kStages = 5;
  CUtensorMap tmA, tmB, tmO;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {64, (uint32_t)G2_BM};
    if ((rc = make_tmap_bf16(&tmA, A, 2, dims, str, box)) != STEGO_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {64, (uint32_t)(BN / 2)};
    if ((rc = make_tmap_bf16(&tmB, B, 2, dims, str, box)) != STEGO_OK) return rc;
  }
  {
    const size_t esz = out_bf16 ? 2 : 4;
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)ldo * esz};
    uint32_t box[2] = {out_bf16 ? 64u : 32u, 32u};
    rc = out_bf16 ? make_tmap_bf16(&tmO, out, 2, dims, str, box) : make_tmap_f32(&tmO, out, 2, dims, str, box);
    if (rc != STEGO_OK) return rc;
  }
  Gemm2Params p;
  p.M = M; p.N = N; p.K = K;
  p.epi.bias = bias; p.epi.act = act; p.epi.out_bf16 = out_bf16; p.epi.reduce_add = reduce_add; p.epi.N = N;
  constexpr size_t smem = size_t(kStages) * (G2_BM * G2_BK * 2 + (BN / 2) * G2_BK * 2) + 8 * 2 * 4096 + 1024 + 256;
  auto kern = gemm2_bf16_kernel<BN, kStages>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm2)");
    configured = true;
  }
  const int pairs = ((M + 2 * G2_BM - 1) / (2 * G2_BM)) * ((N + BN - 1) / BN);
  int clusters = num_sms() / 2;
  if (pairs < clusters) clusters = pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm 2-CTA)");
  count_launch();
  return STEGO_OK;
}
