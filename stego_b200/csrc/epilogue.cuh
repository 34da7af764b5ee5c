// Activation helpers shared by the GEMM epilogues (gemm.cu).
#pragma once
#include "common.cuh"

namespace stego {

// Exact (erf) GELU, nn.GELU default.  erf by Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7): one rcp + one ex2 on
// the MUFU pipe and ~10 FMAs instead of the ~25-instruction erff — the fc1 epilogue is issue-bound otherwise.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));  // MUFU.RCP, no Newton fix-up
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, ex2_approx(-1.4426950408889634f * z * z), 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// MUFU-free variant for bf16 outputs: erf(z) = z * P(z^2) (degree-9 minimax fit on |z| <= 3.2, clamped beyond;
// |abs err| < 8.2e-6, i.e. < 2e-5 on GELU — two orders below bf16 rounding).  The fc1 epilogue applies GELU to
// 77 M elements per layer; two MUFU ops per element made it MUFU-bound (16 ops/clk/SM).
__device__ __forceinline__ float gelu_erf_poly(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752f, 3.2f);
  const float t = z * z;
  float p = fmaf(t, -2.4003365851451727e-09f, 1.4192566410626377e-07f);
  p = fmaf(p, t, -3.73997355423602e-06f);
  p = fmaf(p, t, 5.846926586228758e-05f);
  p = fmaf(p, t, -0.0006113043563036988f);
  p = fmaf(p, t, 0.004584169635313263f);
  p = fmaf(p, t, -0.025814482266624247f);
  p = fmaf(p, t, 0.11186436329524356f);
  p = fmaf(p, t, -0.37570728585235524f);
  p = fmaf(p, t, 1.1283256165012454f);
  const float e = fminf(p * z, 1.0f);
  return 0.5f * x * (1.0f + copysignf(e, x));
}

// bf16-output GELU, eight elements in lock-step as four packed fp32x2 lanes (FFMA2 / FMUL2: two fp32 results per
// issue slot), ~8 instructions per element instead of 13:
//   erf(|x|/sqrt2) = xc * Q(xc^2), xc = min(|x|, 3.2*sqrt2), Q of degree 8 (minimax fit, |abs err| < 4.3e-5 in
//   fp32 evaluation — two orders of magnitude below the bf16 rounding of the result), and
//   gelu(x) = 0.5 x (1 + sign(x) erf(|x|/sqrt2)) = h + |h| * e with h = x/2.
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void gelu_erf_poly8(float* x) {
  uint64_t xc[4], u[4], q[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    xc[j] = pack_f32x2(fminf(fabsf(x[2 * j]), 4.525483399593904f), fminf(fabsf(x[2 * j + 1]), 4.525483399593904f));
    u[j] = mul_f32x2(xc[j], xc[j]);
    q[j] = fma_f32x2(u[j], pack_f32x2(7.28493733954992e-11f, 7.28493733954992e-11f),
                     pack_f32x2(-7.739619932988917e-09f, -7.739619932988917e-09f));
  }
#define STEGO_POLY_STEP(C) \
  _Pragma("unroll") for (int j = 0; j < 4; ++j) q[j] = fma_f32x2(q[j], u[j], pack_f32x2(C, C));
  STEGO_POLY_STEP(3.6041332307651457e-07f)
  STEGO_POLY_STEP(-9.764514095986007e-06f)
  STEGO_POLY_STEP(0.0001730121070631224f)
  STEGO_POLY_STEP(-0.0021448454598048446f)
  STEGO_POLY_STEP(0.01943352726774955f)
  STEGO_POLY_STEP(-0.13244709440462024f)
  STEGO_POLY_STEP(0.7977185244870058f)
#undef STEGO_POLY_STEP
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t e = mul_f32x2(q[j], xc[j]);
    const float h0 = 0.5f * x[2 * j], h1 = 0.5f * x[2 * j + 1];
    const uint64_t r = fma_f32x2(pack_f32x2(fabsf(h0), fabsf(h1)), e, pack_f32x2(h0, h1));
    unpack_f32x2(r, x[2 * j], x[2 * j + 1]);
  }
}

}  // namespace stego

namespace stego {

// TMA epilogue of one accumulator tile, executed by ONE epilogue warp for its 32 TMEM lanes (rows) and every second
// column group (`half` selects which): TMEM -> registers (bias / activation) -> 128B-swizzled staging tile in shared
// memory -> one bulk tensor store (or fp32 reduce-add into the in-place residual) per 32-row x 128-byte tile.
// Shared by the 1-CTA and 2-CTA GEMM kernels; edges are clipped by the tensor map.
struct EpiArgs {
  const float* bias;  // [N] or null
  int act;            // 0 none, 1 GELU(erf), 2 ReLU
  int out_bf16;
  int reduce_add;     // fp32 only: cp.reduce.async.bulk.add instead of a plain store (x += ...)
  int N;
};

template <int BN, uint32_t kEpiBufs, int kParts>
__device__ __forceinline__ void epilogue_tma_tile(const CUtensorMap* tmO, const EpiArgs& p, uint32_t taddr, uint8_t* buf0,
                                                  uint32_t& epi_groups, int part, int lane, int col_tile0, int row_base,
                                                  int batch_idx) {
  // kParts warps share one TMEM lane quarter and take every kParts-th column group.  32 accumulator columns are in
  // registers at a time (the kernel runs 18 warps: 112 registers per thread).
  if (p.out_bf16) {
#pragma unroll 1
    for (int c = part; c < BN / 64; c += kParts) {
      const int col0 = col_tile0 + c * 64;
      if (col0 >= p.N) break;
      uint8_t* buf = buf0 + (epi_groups & (kEpiBufs - 1u)) * 4096;
      if (lane == 0) tma_wait_group_read<kEpiBufs - 1>();  // the store that last read this staging tile is done
      __syncwarp();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 64 + hh * 32, v);
        tmem_ld_wait();
        const int cb = col0 + hh * 32;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr && cb < p.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cb) + j);
            x[4 * j + 0] += b4.x; x[4 * j + 1] += b4.y; x[4 * j + 2] += b4.z; x[4 * j + 3] += b4.w;
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) gelu_erf_poly8(x + j);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.0f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 w;
          w.x = pack_bf16x2(x[8 * j + 0], x[8 * j + 1]);
          w.y = pack_bf16x2(x[8 * j + 2], x[8 * j + 3]);
          w.z = pack_bf16x2(x[8 * j + 4], x[8 * j + 5]);
          w.w = pack_bf16x2(x[8 * j + 6], x[8 * j + 7]);
          *reinterpret_cast<uint4*>(buf + sw128_offset(lane, hh * 4 + j)) = w;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(buf, tmO, col0, row_base, batch_idx);
        tma_commit_group();
      }
      ++epi_groups;
    }
  } else {
#pragma unroll 1
    for (int c = part; c < BN / 32; c += kParts) {
      const int col0 = col_tile0 + c * 32;
      if (col0 >= p.N) break;
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      uint8_t* buf = buf0 + (epi_groups & (kEpiBufs - 1u)) * 4096;
      if (lane == 0) tma_wait_group_read<kEpiBufs - 1>();
      __syncwarp();
      tmem_ld_wait();
      float x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
      if (p.bias != nullptr) {  // (callers only route N % 32 != 0 here without a bias: the store clips, a bias load would not)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
          x[4 * j + 0] += b4.x; x[4 * j + 1] += b4.y; x[4 * j + 2] += b4.z; x[4 * j + 3] += b4.w;
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
      } else if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.0f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(buf + sw128_offset(lane, j)) =
            make_float4(x[4 * j + 0], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (p.reduce_add) tma_reduce_add_3d(buf, tmO, col0, row_base, batch_idx);
        else tma_store_3d(buf, tmO, col0, row_base, batch_idx);
        tma_commit_group();
      }
      ++epi_groups;
    }
  }
}

}  // namespace stego
