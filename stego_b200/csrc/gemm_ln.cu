// Fused  x += A . W^T + bias ;  y = LayerNorm(x) * gamma + beta   (sm_100a, tcgen05 / TMA / TMEM)
//
// The two residual GEMMs of a pre-norm ViT block (reference: src/dino/vision_transformer.py:92-104 Block.forward:
//   x = x + attn(norm1(x));  x = x + mlp(norm2(x)) )  are each followed by the LayerNorm that feeds the next linear
// (norm2 after proj, the next block's norm1 after fc2).  A separate LayerNorm kernel re-reads the 77 MB fp32 residual
// stream it was just written to; here one CTA owns a whole 128 x 384 row block, so each epilogue thread (= one TMEM
// lane = one token row) sees its full row and the normalised bf16 copy leaves in the same pass:
//
//   warp 0      TMA producer   A [128 x 64] + W [384 x 64] bf16 tiles -> 2-stage 128B-swizzled ring
//   warp 1      MMA issuer     tcgen05.mma M=128, N=256+128, fp32 accumulator = 384 TMEM columns
//   warps 2..9  epilogue       pass 1: TMA-load the residual tile chunk (32 rows x 32 fp32) into a per-warp ring,
//                                      x = acc + bias + residual -> written back into TMEM (the row stays on chip),
//                                      staged in place and TMA-stored to x; row sum
//                              pass 2: sum (x - mean)^2 from TMEM (two-pass variance, like the standalone kernel)
//                              pass 3: normalise from TMEM, gamma/beta, bf16, TMA store to y
// The two warps that share a TMEM lane quarter split the 384 columns and exchange their partial row statistics
// through shared memory (named barriers, 64 threads).  HBM traffic per call: A + x read + x write + y write — the
// LayerNorm's own read of x is gone.  N is fixed at 384 (ViT-S); the 768-wide ViT-B rows do not fit one CTA's TMEM.
#include "common.cuh"
#include "host_util.h"

namespace stego {

constexpr int GL_BM = 128, GL_BN = 384, GL_BK = 64, GL_STAGES = 2, GL_THREADS = 320;
constexpr uint32_t GL_A_BYTES = GL_BM * GL_BK * 2;          // 16 KB
constexpr uint32_t GL_B_BYTES = GL_BN * GL_BK * 2;          // 48 KB
constexpr uint32_t GL_STAGE_BYTES = GL_A_BYTES + GL_B_BYTES;
constexpr int GL_EBUFS = 3;                                 // 4 KB staging tiles per epilogue warp
constexpr uint32_t GL_EPI_BYTES = 8 * GL_EBUFS * 4096;      // 96 KB
constexpr size_t GL_SMEM = size_t(GL_STAGES) * GL_STAGE_BYTES + GL_EPI_BYTES + 2048 /*row stats*/ + 512 /*barriers*/;

struct GemmLnParams {
  int M, K;
  const float* bias;   // [384] or null
  const float* gamma;  // [384]
  const float* beta;   // [384]
  float eps;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(GL_THREADS, 1)
gemm_residual_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                        GemmLnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* epi_smem = smem + GL_STAGES * GL_STAGE_BYTES;
  float* stat_sum = reinterpret_cast<float*>(epi_smem + GL_EPI_BYTES);  // [4 quarters][2 halves][32 lanes]
  float* stat_sq = stat_sum + 256;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stat_sq + 256);
  uint64_t* empty_bar = full_bar + GL_STAGES;
  uint64_t* tfull_bar = empty_bar + GL_STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint64_t* res_bar = tempty_bar + 1;  // [8 warps][GL_EBUFS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8 * GL_EBUFS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.M + GL_BM - 1) / GL_BM;
  const int num_kb = (p.K + GL_BK - 1) / GL_BK;
  constexpr uint32_t IDESC = make_idesc_bf16(GL_BM, 256, 0, 0);
  constexpr uint32_t IDESC_TAIL = make_idesc_bf16(GL_BM, 128, 0, 0);

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < GL_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 8);
    for (int i = 0; i < 8 * GL_EBUFS; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tm = blockIdx.x; tm < tiles_m; tm += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * GL_STAGE_BYTES;
          uint8_t* sb = sa + GL_A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], GL_STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GL_BK, tm * GL_BM);
#pragma unroll
          for (int blk = 0; blk < 3; ++blk)
            tma_load_2d(sb + blk * 16384, &tmB, &full_bar[stage], kb * GL_BK, blk * 128);
          if (++stage == GL_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp, uniform control flow (descriptors stay in uniform registers); one elected lane issues
    {
      uint32_t stage = 0, phase = 0, tphase = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t DESC_HI = smem_desc_hi_sw128(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), 16);
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem) + GL_A_BYTES, 16);
      for (int tm = blockIdx.x; tm < tiles_m; tm += gridDim.x) {
        mbar_wait(tempty_bar, tphase ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (GL_STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (GL_STAGE_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (uint32_t k = 0; k < GL_BK / 16; ++k) {
              const uint64_t da = smem_desc_join(a_lo + 2 * k, DESC_HI);
              umma_bf16(tmem_u, da, smem_desc_join(b_lo + 2 * k, DESC_HI), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
              umma_bf16(tmem_u + 256, da, smem_desc_join(b_lo + ((256u * 128u) >> 4) + 2 * k, DESC_HI), IDESC_TAIL,
                        (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == GL_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(tfull_bar);
        __syncwarp();
        tphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;      // TMEM lane quarter this warp may access
    const int half = ew >> 2;          // columns [half*192, half*192 + 192)
    uint8_t* bufs = epi_smem + ew * (GL_EBUFS * 4096);
    uint64_t* rbar = res_bar + ew * GL_EBUFS;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * 192;
    const int colbase = half * 192;
    const int peer = (quarter * 2 + (half ^ 1)) * 32 + lane, mine = (quarter * 2 + half) * 32 + lane;
    uint32_t tphase = 0;
    bool first = true;
    for (int tm = blockIdx.x; tm < tiles_m; tm += gridDim.x) {
      const int row_base = tm * GL_BM + quarter * 32;
      if (first) {
        // residual chunks 0 and 1 of the first tile (later tiles: issued at the end of the previous tile)
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            mbar_arrive_expect_tx(&rbar[i], 4096);
            tma_load_2d(bufs + i * 4096, &tmX, &rbar[i], colbase + i * 32, row_base);
          }
        }
        first = false;
      }
      mbar_wait(tfull_bar, tphase);
      tc_fence_after();
      // ---- pass 1: x = acc + bias + residual; back to TMEM, out to global, row sum
      float rsum = 0.f;
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int b = i % GL_EBUFS;
        uint8_t* buf = bufs + b * 4096;
        uint32_t v[32];
        tmem_ld32(taddr + i * 32, v);
        mbar_wait(&rbar[b], (i / GL_EBUFS) & 1u);  // each buffer receives exactly two loads per tile
        float x[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r = *reinterpret_cast<const float4*>(buf + sw128_offset(lane, j));
          x[4 * j + 0] = r.x; x[4 * j + 1] = r.y; x[4 * j + 2] = r.z; x[4 * j + 3] = r.w;
        }
        if (i + 2 < 6 && lane == 0) {
          // chunk i+2 goes into the buffer chunk i-1 was stored from: its bulk store has long finished reading
          tma_wait_group_read<0>();
          const int nb = (i + 2) % GL_EBUFS;
          mbar_arrive_expect_tx(&rbar[nb], 4096);
          tma_load_2d(bufs + nb * 4096, &tmX, &rbar[nb], colbase + (i + 2) * 32, row_base);
        }
        tmem_ld_wait();
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + colbase + i * 32) + j);
            x[4 * j + 0] += b4.x; x[4 * j + 1] += b4.y; x[4 * j + 2] += b4.z; x[4 * j + 3] += b4.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x[j] += __uint_as_float(v[j]);
          v[j] = __float_as_uint(x[j]);
          rsum += x[j];
        }
        tmem_st32(taddr + i * 32, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(buf + sw128_offset(lane, j)) =
              make_float4(x[4 * j + 0], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(buf, &tmX, colbase + i * 32, row_base);
          tma_commit_group();
        }
      }
      tmem_st_wait();
      stat_sum[mine] = rsum;
      named_bar_sync(1 + quarter, 64);
      const float mean = (rsum + stat_sum[peer]) * (1.0f / GL_BN);
      // ---- pass 2: centred second moment from TMEM
      float rsq = 0.f;
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        uint32_t v[32];
        tmem_ld32(taddr + i * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(v[j]) - mean;
          rsq = fmaf(d, d, rsq);
        }
      }
      stat_sq[mine] = rsq;
      named_bar_sync(1 + quarter, 64);
      const float rstd = rsqrtf((rsq + stat_sq[peer]) * (1.0f / GL_BN) + p.eps);
      const float nmr = -mean * rstd;
      // ---- pass 3: normalise, affine, bf16, 32 x 64 tiles out through TMA
#pragma unroll 1
      for (int t = 0; t < 3; ++t) {
        uint8_t* buf = bufs + t * 4096;
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + t * 64, v0);
        tmem_ld32(taddr + t * 64 + 32, v1);
        if (lane == 0) tma_wait_group_read<0>();  // pass-1 stores (and earlier y tiles) have released the staging tiles
        __syncwarp();
        tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t* v = hh ? v1 : v0;
          const int cb = colbase + t * 64 + hh * 32;
          float y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + cb) + j);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + cb) + j);
            y[4 * j + 0] = fmaf(fmaf(__uint_as_float(v[4 * j + 0]), rstd, nmr), g4.x, b4.x);
            y[4 * j + 1] = fmaf(fmaf(__uint_as_float(v[4 * j + 1]), rstd, nmr), g4.y, b4.y);
            y[4 * j + 2] = fmaf(fmaf(__uint_as_float(v[4 * j + 2]), rstd, nmr), g4.z, b4.z);
            y[4 * j + 3] = fmaf(fmaf(__uint_as_float(v[4 * j + 3]), rstd, nmr), g4.w, b4.w);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 w;
            w.x = pack_bf16x2(y[8 * j + 0], y[8 * j + 1]);
            w.y = pack_bf16x2(y[8 * j + 2], y[8 * j + 3]);
            w.z = pack_bf16x2(y[8 * j + 4], y[8 * j + 5]);
            w.w = pack_bf16x2(y[8 * j + 6], y[8 * j + 7]);
            *reinterpret_cast<uint4*>(buf + sw128_offset(lane, hh * 4 + j)) = w;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(buf, &tmY, colbase + t * 64, row_base);
          tma_commit_group();
        }
      }
      // accumulator columns are free again: the MMAs of this CTA's next tile may start
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(tempty_bar);
        const int tn = tm + gridDim.x;
        if (tn < tiles_m) {
          // residual chunks 0, 1 of the next tile fly in while its MMAs run
          tma_wait_group_read<0>();
          const int nrow = tn * GL_BM + quarter * 32;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            mbar_arrive_expect_tx(&rbar[i], 4096);
            tma_load_2d(bufs + i * 4096, &tmX, &rbar[i], colbase + i * 32, nrow);
          }
        }
      }
      tphase ^= 1u;
    }
    if (lane == 0) tma_wait_group_read<0>();  // staging smem must outlive the bulk stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace stego

using namespace stego;

// C-ABI: see include/stego_b200.h for the contract.
extern "C" int stego_gemm_residual_ln_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* x,
                                           int ldx, const float* bias, const float* gamma, const float* beta, float eps,
                                           void* y, int ldy, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STEGO_CHECK_ARG(A && W && x && gamma && beta && y, "stego_gemm_residual_ln_bf16: null pointer");
  STEGO_CHECK_ARG(M > 0 && K > 0, "stego_gemm_residual_ln_bf16: bad sizes M=%d K=%d", M, K);
  if (N != GL_BN) {
    set_error("stego_gemm_residual_ln_bf16: N=%d unsupported (one CTA holds a whole row in TMEM: N must be 384)", N);
    return STEGO_ERR_UNSUPPORTED;
  }
  STEGO_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && ldy % 8 == 0 && ldx % 4 == 0,
                  "stego_gemm_residual_ln_bf16: lda/ldw/ldy must be multiples of 8 elements, ldx of 4");
  STEGO_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(A) |
                    reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(gamma) |
                    reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(bias)) & 15u) == 0,
                  "stego_gemm_residual_ln_bf16: pointers must be 16-byte aligned");
  CUtensorMap tmA, tmB, tmX, tmY;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {64, (uint32_t)GL_BM};
    if ((rc = make_tmap_bf16(&tmA, A, 2, dims, str, box)) != STEGO_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)GL_BN};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&tmB, W, 2, dims, str, box)) != STEGO_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)GL_BN, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)ldx * 4};
    uint32_t box[2] = {32, 32};
    if ((rc = make_tmap_f32(&tmX, x, 2, dims, str, box)) != STEGO_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)GL_BN, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)ldy * 2};
    uint32_t box[2] = {64, 32};
    if ((rc = make_tmap_bf16(&tmY, y, 2, dims, str, box)) != STEGO_OK) return rc;
  }
  GemmLnParams p;
  p.M = M; p.K = K; p.bias = bias; p.gamma = gamma; p.beta = beta; p.eps = eps;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_residual_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GL_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_residual_ln)");
    configured = true;
  }
  const int tiles = (M + GL_BM - 1) / GL_BM;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_residual_ln_kernel<<<grid, GL_THREADS, GL_SMEM, stream>>>(tmA, tmB, tmX, tmY, p);
  STEGO_CHECK_LAUNCH("gemm_residual_ln_kernel launch");
  return STEGO_OK;
}
