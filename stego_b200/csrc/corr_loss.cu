// Fused feature-correspondence loss for STEGO (sm_100a): sampling + L2-norm, the tensor_correlation
// einsum on tcgen05, the centre/shift/clamp/product reduction, and the full backward.
//
// Reference: src/modules.py
//   :275-276 norm, :283-284 tensor_correlation (einsum nchw,ncij->nhwij), :287-288 sample (grid_sample),
//   :325-347 ContrastiveCorrelationLoss.helper, :349-398 ContrastiveCorrelationLoss.forward.
//
// Pipeline (K = 2 + neg_samples helper "calls", S = feature_samples^2 <= 128 sampled points per image):
//   1. sample_norm_*  : bilinear 4-tap gather (border, align_corners=True) of S points per image for every
//                       distinct operand ("slot": 0 = img@coords1, 1 = pos@coords2, 2.. = img[perm_i]@coords2),
//                       L2-normalise over channels in fp32, write [128][C] operand tiles as a bf16 hi/lo SPLIT
//                       (x = hi + lo).  The einsum is then three bf16 tensor-core passes
//                       hi.hi + lo.hi + hi.lo, i.e. ~2^-16 relative accuracy with fp32 accumulation.
//   2. corr_fwd       : one CTA per (image, call): TMA-staged tiles -> tcgen05.mma -> fd and cd accumulators in
//                       TMEM (128 columns each); epilogue straight out of TMEM: row means (pointwise centring),
//                       clamp, shift, product, five partial sums per CTA (no atomics, deterministic).
//                       The batch-global `old_mean` enters the loss linearly, so it is applied in the finish step.
//   3. corr_finish    : per call: old_mean, loss mean, cd mean.
//   4. corr_bwd       : recomputes fd, cd (cheaper than stashing them), forms G = dL/dcd in registers, writes it
//                       as a swizzled bf16 hi/lo tile to smem and runs the two backward GEMMs on the tensor core
//                       from the SAME smem image:  dA = G . Bc (G as K-major A) and dB = G^T . Ac (G as MN-major A);
//                       the code tiles Ac/Bc stay resident in smem from the forward part and are reused as
//                       MN-major B operands.  Results are accumulated into per-slot gradient tiles.
//   5. sample_norm_bwd: normalisation backward + grid_sample backward (4-tap scatter-add into d_code).
#include "common.cuh"
#include "host_util.h"

namespace stego {

constexpr int CL_ROWS = 128;     // tile rows (S padded)
constexpr int CL_CODE_PAD = 128; // code channels padded to 2 k-blocks
constexpr int CL_MAX_CALLS = 16;
constexpr int CL_DT_LD = 72;     // row stride (floats) of the per-slot gradient tiles

// ---------------------------------------------------------------------------------------------
// 1. sampling + normalisation
// ---------------------------------------------------------------------------------------------
struct SampleParams {
  const void* src;      // slot 0 and slots >= 2 (through perm)
  const void* src_pos;  // slot 1
  int src_bf16;         // element type of both sources
  long long sb, sc, sy, sx;  // element strides (batch, channel, y, x) — shared by src and src_pos
  const float* chan_scale;      // optional [B][C] per-(image,channel) multiplier (Dropout2d noise), slot 0/2+
  const float* chan_scale_pos;  // same for src_pos
  const float* coords1;  // [B][fs][fs][2]
  const float* coords2;
  const long long* perms;  // [nslots-2][B]
  int perms_raw;           // 1: perms are raw randperm draws; apply super_perm's fix-up (p == b -> (p+1) % B) here
  bf16* tiles;             // [2 planes][nslots][B][128][Cpad]
  int B, C, Cpad, H, W, fs, S, nslots;
  float eps;
};

struct Taps {
  int i00, i01, i10, i11;      // pixel offsets (y*W + x), clamped in-bounds
  float w00, w01, w10, w11;    // nw, ne, sw, se weights (0 for out-of-bounds taps)
};

// grid_sample(bilinear, padding_mode='border', align_corners=True) source taps for sample index s = i*fs + j,
// which reads coords[b][j][i] because `sample` permutes the grid (modules.py:288).
__device__ __forceinline__ Taps make_taps(const float* coords, int b, int s, int fs, int H, int W) {
  const int i = s / fs, j = s % fs;
  const float* cp = coords + ((static_cast<long long>(b) * fs + j) * fs + i) * 2;
  float x = ((cp[0] + 1.f) / 2.f) * (W - 1);
  float y = ((cp[1] + 1.f) / 2.f) * (H - 1);
  x = fminf(fmaxf(x, 0.f), static_cast<float>(W - 1));
  y = fminf(fmaxf(y, 0.f), static_cast<float>(H - 1));
  const float x0 = floorf(x), y0 = floorf(y);
  const float x1 = x0 + 1.f, y1 = y0 + 1.f;
  Taps t;
  t.w00 = (x1 - x) * (y1 - y);
  t.w01 = (x - x0) * (y1 - y);
  t.w10 = (x1 - x) * (y - y0);
  t.w11 = (x - x0) * (y - y0);
  const int ix0 = static_cast<int>(x0), iy0 = static_cast<int>(y0);
  int ix1 = ix0 + 1, iy1 = iy0 + 1;
  if (ix1 > W - 1) { ix1 = W - 1; t.w01 = 0.f; t.w11 = 0.f; }
  if (iy1 > H - 1) { iy1 = H - 1; t.w10 = 0.f; t.w11 = 0.f; }
  t.i00 = iy0 * W + ix0;
  t.i01 = iy0 * W + ix1;
  t.i10 = iy1 * W + ix0;
  t.i11 = iy1 * W + ix1;
  return t;
}

__device__ __forceinline__ void slot_source(const SampleParams& p, int slot, int b, const void*& src,
                                            const float*& cscale, const float*& coords, int& img) {
  if (slot == 0) { src = p.src; cscale = p.chan_scale; coords = p.coords1; img = b; }
  else if (slot == 1) { src = p.src_pos; cscale = p.chan_scale_pos; coords = p.coords2; img = b; }
  else {
    src = p.src; cscale = p.chan_scale; coords = p.coords2;
    img = static_cast<int>(p.perms[(slot - 2) * p.B + b]);
    if (p.perms_raw && img == b) img = (img + 1) % p.B;  // modules.py:291-295
  }
}

// One warp per tile row. Generic strides / dtypes; each lane owns channels lane, lane+32, ...
template <int NV>  // NV = Cpad / 32
__global__ void __launch_bounds__(256)
sample_norm_kernel(SampleParams p) {
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int total = p.nslots * p.B * CL_ROWS;
  if (warp_global >= total) return;
  const int s = warp_global % CL_ROWS;
  const int b = (warp_global / CL_ROWS) % p.B;
  const int slot = warp_global / (CL_ROWS * p.B);
  const size_t plane = static_cast<size_t>(p.nslots) * p.B * CL_ROWS * p.Cpad;
  bf16* hi = p.tiles + (static_cast<size_t>(slot) * p.B + b) * CL_ROWS * p.Cpad + static_cast<size_t>(s) * p.Cpad;
  bf16* lo = hi + plane;
  float v[NV];
  if (s >= p.S) {
#pragma unroll
    for (int k = 0; k < NV; ++k) { hi[lane + 32 * k] = __float2bfloat16_rn(0.f); lo[lane + 32 * k] = __float2bfloat16_rn(0.f); }
    return;
  }
  const void* src; const float* cscale; const float* coords; int img;
  slot_source(p, slot, b, src, cscale, coords, img);
  const Taps t = make_taps(coords, b, s, p.fs, p.H, p.W);
  // pixel offset -> element offset
  auto off = [&](int pix) { return static_cast<long long>(pix / p.W) * p.sy + static_cast<long long>(pix % p.W) * p.sx; };
  const long long o00 = off(t.i00), o01 = off(t.i01), o10 = off(t.i10), o11 = off(t.i11);
  const long long base = static_cast<long long>(img) * p.sb;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    float val = 0.f;
    if (c < p.C) {
      const long long cb = base + static_cast<long long>(c) * p.sc;
      float a00, a01, a10, a11;
      if (p.src_bf16) {
        const bf16* q = reinterpret_cast<const bf16*>(src) + cb;
        a00 = __bfloat162float(q[o00]); a01 = __bfloat162float(q[o01]);
        a10 = __bfloat162float(q[o10]); a11 = __bfloat162float(q[o11]);
      } else {
        const float* q = reinterpret_cast<const float*>(src) + cb;
        a00 = q[o00]; a01 = q[o01]; a10 = q[o10]; a11 = q[o11];
      }
      // same accumulation order as ATen's grid_sampler_2d: nw, ne, sw, se
      val = a00 * t.w00;
      val += a01 * t.w01;
      val += a10 * t.w10;
      val += a11 * t.w11;
      if (cscale) val *= cscale[static_cast<long long>(img) * p.C + c];
    }
    v[k] = val;
    ss += val * val;
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), p.eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float n = v[k] * inv;
    const bf16 h = __float2bfloat16_rn(n);
    hi[lane + 32 * k] = h;
    lo[lane + 32 * k] = __float2bfloat16_rn(n - __bfloat162float(h));
  }
}

// Vectorised variant for the layout the training step uses: bf16 source, channel stride 1 (tokens-major),
// C % 8 == 0, Cpad == C.  Each lane owns 8 consecutive channels per step: four 16-byte tap loads, two 16-byte tile
// stores (hi, lo) — 8x fewer memory instructions than the generic kernel.
template <int NI>  // NI = ceil(C / 256)
__global__ void __launch_bounds__(256)
sample_norm_vec8_kernel(SampleParams p) {
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int total = p.nslots * p.B * CL_ROWS;
  if (warp_global >= total) return;
  const int s = warp_global % CL_ROWS;
  const int b = (warp_global / CL_ROWS) % p.B;
  const int slot = warp_global / (CL_ROWS * p.B);
  const size_t plane = static_cast<size_t>(p.nslots) * p.B * CL_ROWS * p.Cpad;
  bf16* hi = p.tiles + (static_cast<size_t>(slot) * p.B + b) * CL_ROWS * p.Cpad + static_cast<size_t>(s) * p.Cpad;
  bf16* lo = hi + plane;
  if (s >= p.S) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < p.C) {
        *reinterpret_cast<uint4*>(hi + c) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(lo + c) = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }
  const void* src; const float* cscale; const float* coords; int img;
  slot_source(p, slot, b, src, cscale, coords, img);
  const Taps t = make_taps(coords, b, s, p.fs, p.H, p.W);
  auto off = [&](int pix) { return static_cast<long long>(pix / p.W) * p.sy + static_cast<long long>(pix % p.W) * p.sx; };
  const bf16* q = reinterpret_cast<const bf16*>(src) + static_cast<long long>(img) * p.sb;
  const bf16* q00 = q + off(t.i00);
  const bf16* q01 = q + off(t.i01);
  const bf16* q10 = q + off(t.i10);
  const bf16* q11 = q + off(t.i11);
  float v[NI][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = (lane + 32 * i) * 8;
    if (c < p.C) {
      const uint4 a00 = *reinterpret_cast<const uint4*>(q00 + c), a01 = *reinterpret_cast<const uint4*>(q01 + c);
      const uint4 a10 = *reinterpret_cast<const uint4*>(q10 + c), a11 = *reinterpret_cast<const uint4*>(q11 + c);
      const __nv_bfloat162* h00 = reinterpret_cast<const __nv_bfloat162*>(&a00);
      const __nv_bfloat162* h01 = reinterpret_cast<const __nv_bfloat162*>(&a01);
      const __nv_bfloat162* h10 = reinterpret_cast<const __nv_bfloat162*>(&a10);
      const __nv_bfloat162* h11 = reinterpret_cast<const __nv_bfloat162*>(&a11);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f00 = __bfloat1622float2(h00[k]), f01 = __bfloat1622float2(h01[k]);
        const float2 f10 = __bfloat1622float2(h10[k]), f11 = __bfloat1622float2(h11[k]);
        float x = f00.x * t.w00; x += f01.x * t.w01; x += f10.x * t.w10; x += f11.x * t.w11;
        float y = f00.y * t.w00; y += f01.y * t.w01; y += f10.y * t.w10; y += f11.y * t.w11;
        if (cscale) {
          x *= cscale[static_cast<long long>(img) * p.C + c + 2 * k];
          y *= cscale[static_cast<long long>(img) * p.C + c + 2 * k + 1];
        }
        v[i][2 * k] = x;
        v[i][2 * k + 1] = y;
        ss += x * x + y * y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
    }
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), p.eps);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = (lane + 32 * i) * 8;
    if (c < p.C) {
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float n0 = v[i][2 * k] * inv, n1 = v[i][2 * k + 1] * inv;
        const float h0 = __bfloat162float(__float2bfloat16_rn(n0)), h1 = __bfloat162float(__float2bfloat16_rn(n1));
        wh[k] = pack_bf16x2(h0, h1);
        wl[k] = pack_bf16x2(n0 - h0, n1 - h1);
      }
      *reinterpret_cast<uint4*>(hi + c) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      *reinterpret_cast<uint4*>(lo + c) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 2./4. correlation + loss forward / backward
// ---------------------------------------------------------------------------------------------
struct CorrParams {
  int B, S, E, nslots, ncalls;
  int slot_of_call[CL_MAX_CALLS];   // B-operand slot per call (A operand is always slot 0)
  float shift[CL_MAX_CALLS];
  int pointwise;
  float clamp_lo;    // 0 (zero_clamp) or -9999
  float clamp_hi;    // 0.8 (stabalize) or +inf
  // forward outputs
  float* partials;   // [ncalls][B][8]: P1..P5 (see below)
  float* cd_out;     // optional [ncalls][B][S][S]
  float* fdc_out;    // optional [ncalls][B][S][S] (row-centred fd)
  // backward inputs
  const float* stats;   // [ncalls][4]: loss_mean, cd_mean, old_mean, mean_c
  const float* gscale;  // [ncalls] upstream grad of each call's mean loss
  const float* gelem;   // optional [ncalls][B][S][S] upstream grad of unreduced loss elements
  const float* gcd;     // optional [ncalls][B][S][S] upstream grad of cd elements
  float* dtiles;        // [nslots][B][128][CL_DT_LD] fp32, zero-initialised by the caller
  int D;                // real code channels (<= CL_CODE_PAD)
};

constexpr int CL_THREADS = 192;
constexpr uint32_t CL_TILE = 128 * 64 * 2;  // 16 KB

__device__ __forceinline__ uint32_t tile_row(int plane, int slot, int b, int nslots, int B) {
  return static_cast<uint32_t>(((plane * nslots + slot) * B + b) * CL_ROWS);
}

// shared by forward and backward: issue order of the (A plane, B plane) split passes
__device__ __forceinline__ void split_pass(int pass, int& pa, int& pb) {
  pa = (pass == 1) ? 1 : 0;  // (hi,hi), (lo,hi), (hi,lo)
  pb = (pass == 2) ? 1 : 0;
}

template <bool kBackward>
__global__ void __launch_bounds__(CL_THREADS, 1)
corr_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmC, CorrParams p) {
  // smem: ring of kRing stages x (A 16K + B 16K) for the feature GEMM; backward adds resident code tiles.
  constexpr int kRing = kBackward ? 2 : 3;
  constexpr uint32_t RING_BYTES = kRing * 2 * CL_TILE;
  constexpr uint32_t CODE_BYTES = 8 * CL_TILE;  // Ac: [plane][kb] 4 tiles, Bc: 4 tiles
  constexpr uint32_t OFF_CODE = RING_BYTES;
  constexpr uint32_t OFF_BAR = OFF_CODE + CODE_BYTES;
  constexpr uint32_t TMEM_COLS = kBackward ? 512 : 256;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full_bar = bars;              // [kRing]
  uint64_t* empty_bar = bars + kRing;     // [kRing]
  uint64_t* code_full = empty_bar + kRing;  // [1]
  uint64_t* acc_full = code_full + 1;       // [1]
  uint64_t* g_full = acc_full + 1;          // [1] (backward)
  uint64_t* d_full = g_full + 1;            // [1] (backward)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);
  float* red = reinterpret_cast<float*>(tmem_slot + 4);  // [4 warps][8]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int call = blockIdx.y;
  const int slotB = p.slot_of_call[call];
  const int nkb_f = p.E / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmF);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < kRing; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(code_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(g_full, 4);
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_FD = tmem_base, TM_CD = tmem_base + 128, TM_DA = tmem_base + 256, TM_DB = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      // code tiles first (small, resident): Ac planes then Bc planes, 2 k-blocks each
      mbar_arrive_expect_tx(code_full, CODE_BYTES);
      for (int pl = 0; pl < 2; ++pl)
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(smem + OFF_CODE + (pl * 2 + kb) * CL_TILE, &tmC, code_full, kb * 64,
                      tile_row(pl, 0, b, p.nslots, p.B));
          tma_load_2d(smem + OFF_CODE + (4 + pl * 2 + kb) * CL_TILE, &tmC, code_full, kb * 64,
                      tile_row(pl, slotB, b, p.nslots, p.B));
        }
      uint32_t stage = 0, phase = 0;
      for (int pass = 0; pass < 3; ++pass) {
        int pa, pb;
        split_pass(pass, pa, pb);
        for (int kb = 0; kb < nkb_f; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * 2 * CL_TILE;
          mbar_arrive_expect_tx(&full_bar[stage], 2 * CL_TILE);
          tma_load_2d(sa, &tmF, &full_bar[stage], kb * 64, tile_row(pa, 0, b, p.nslots, p.B));
          tma_load_2d(sa + CL_TILE, &tmF, &full_bar[stage], kb * 64, tile_row(pb, slotB, b, p.nslots, p.B));
          if (++stage == kRing) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp walks the schedule (uniform control flow -> descriptors in uniform registers), one
    // elected lane issues.  All descriptors are base-low-word + constant steps (see smem_desc_lo).
    {
      constexpr uint32_t IDESC_KK = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t DESC_HI = smem_desc_hi_sw128(1024);
      constexpr uint32_t TILE16 = CL_TILE >> 4;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t tm_fd = tmem_u, tm_cd = tmem_u + 128, tm_da = tmem_u + 256, tm_db = tmem_u + 384;
      const uint32_t ring_lo = smem_desc_lo(smem_u32(smem), 16);
      uint32_t stage = 0, phase = 0;
      // fd = An_f . Bn_f^T  (3 split passes over E)
      for (int pass = 0; pass < 3; ++pass) {
        for (int kb = 0; kb < nkb_f; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = ring_lo + stage * 2 * TILE16;
          if (elect_one()) {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
              umma_bf16(tm_fd, smem_desc_join(a_lo + 2 * k, DESC_HI), smem_desc_join(a_lo + TILE16 + 2 * k, DESC_HI), IDESC_KK,
                        (pass | kb | static_cast<int>(k)) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kRing) { stage = 0; phase ^= 1u; }
        }
      }
      // cd = An_c . Bn_c^T from the resident code tiles
      mbar_wait(code_full, 0);
      tc_fence_after();
      const uint32_t code_lo = smem_desc_lo(smem_u32(smem + OFF_CODE), 16);
      if (elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          int pa, pb;
          split_pass(pass, pa, pb);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t a_lo = code_lo + (pa * 2 + kb) * TILE16;
            const uint32_t b_lo = code_lo + (4 + pb * 2 + kb) * TILE16;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
              umma_bf16(tm_cd, smem_desc_join(a_lo + 2 * k, DESC_HI), smem_desc_join(b_lo + 2 * k, DESC_HI), IDESC_KK,
                        (pass | kb | static_cast<int>(k)) ? 1u : 0u);
          }
        }
        umma_commit(acc_full);
      }
      __syncwarp();
      if (kBackward) {
        // G (bf16 hi/lo, [i][j] swizzled K-major image) is written by the epilogue warps into the ring area.
        mbar_wait(g_full, 0);
        tc_fence_after();
        constexpr uint32_t IDESC_K_MN = make_idesc_bf16(128, 128, 0, 1);   // A K-major, B MN-major
        constexpr uint32_t IDESC_MN_MN = make_idesc_bf16(128, 128, 1, 1);  // A MN-major, B MN-major
        // MN-major operands: LBO = CL_TILE (64-wide blocks 16 KB apart); G planes: hi at +0, lo at +32 KB
        const uint32_t g_k_lo = ring_lo;                                          // G as K-major A operand
        const uint32_t g_mn_lo = smem_desc_lo(smem_u32(smem), CL_TILE);           // G^T as MN-major A operand
        const uint32_t code_mn_lo = smem_desc_lo(smem_u32(smem + OFF_CODE), CL_TILE);
        if (elect_one()) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            int pg, pc;
            split_pass(pass, pg, pc);
            const uint32_t g_off = pg * 2 * TILE16;
            const uint32_t bc_lo = code_mn_lo + (4 + pc * 2) * TILE16;  // Bc plane: c-blocks 16 KB apart, rows = j
            const uint32_t ac_lo = code_mn_lo + (pc * 2) * TILE16;      // Ac plane: rows = i
#pragma unroll
            for (uint32_t kk = 0; kk < 8; ++kk) {
              // dA[i][c] += sum_j G[i][j] Bc[j][c]: A = G K-major (k = j), B = Bc MN-major (n = c contiguous)
              umma_bf16(tm_da, smem_desc_join(g_k_lo + g_off + (kk >> 2) * TILE16 + (kk & 3) * 2, DESC_HI),
                        smem_desc_join(bc_lo + kk * (2048u >> 4), DESC_HI), IDESC_K_MN, (pass | static_cast<int>(kk)) ? 1u : 0u);
            }
#pragma unroll
            for (uint32_t kk = 0; kk < 8; ++kk) {
              // dB[j][c] += sum_i G[i][j] Ac[i][c]: A = G^T as MN-major (m = j contiguous, k = i rows)
              umma_bf16(tm_db, smem_desc_join(g_mn_lo + g_off + kk * (2048u >> 4), DESC_HI),
                        smem_desc_join(ac_lo + kk * (2048u >> 4), DESC_HI), IDESC_MN_MN, (pass | static_cast<int>(kk)) ? 1u : 0u);
            }
          }
          umma_commit(d_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;
    const int i = quarter * 32 + lane;  // sample row
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int S = p.S;
    const bool row_ok = i < S;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // pass 1: row sum of fd over the valid columns
    float rsum = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      tmem_ld32(TM_FD + lane_off + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t)
        if (ch * 32 + t < S) rsum += __uint_as_float(v[t]);
    }
    const float rmean = p.pointwise ? rsum / static_cast<float>(S) : 0.f;
    const float shift = p.shift[call];
    float P1 = 0.f, P2 = 0.f, P3 = row_ok ? rsum : 0.f, P4 = 0.f, P5 = 0.f;
    float offset = 0.f, gs = 0.f;
    if (kBackward) {
      const float* st = p.stats + call * 4;
      offset = p.pointwise ? (st[2] - st[3]) : 0.f;  // old_mean - mean(centred fd)
      gs = p.gscale[call] / (static_cast<float>(p.B) * S * S);
    }
    const size_t eoff = ((static_cast<size_t>(call) * p.B + b) * S + i) * S;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t vf[32], vc[32];
      tmem_ld32(TM_FD + lane_off + ch * 32, vf);
      tmem_ld32(TM_CD + lane_off + ch * 32, vc);
      tmem_ld_wait();
      float g[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int j = ch * 32 + t;
        const bool ok = row_ok && j < S;
        const float fdc = __uint_as_float(vf[t]) - rmean;
        const float cd = __uint_as_float(vc[t]);
        if (!kBackward) {
          if (ok) {
            const float cl = fminf(fmaxf(cd, p.clamp_lo), p.clamp_hi);
            P1 += cl * (fdc - shift);
            P2 += cl;
            P4 += fdc;
            P5 += cd;
            if (p.cd_out) p.cd_out[eoff + j] = cd;
            if (p.fdc_out) p.fdc_out[eoff + j] = fdc;
          }
        } else {
          float gv = 0.f;
          if (ok) {
            float up = gs;
            if (p.gelem) up += p.gelem[eoff + j];
            const bool pass_grad = (cd >= p.clamp_lo) && (cd <= p.clamp_hi);
            gv = pass_grad ? -up * (fdc + offset - shift) : 0.f;
            if (p.gcd) gv += p.gcd[eoff + j];
          }
          g[t] = gv;
        }
      }
      if (kBackward) {
        // G[i][j] -> smem (ring area, free: every feature MMA has retired before acc_full fired)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 wh, wl;
          float h[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) h[t] = __bfloat162float(__float2bfloat16_rn(g[8 * q + t]));
          wh.x = pack_bf16x2(h[0], h[1]); wh.y = pack_bf16x2(h[2], h[3]);
          wh.z = pack_bf16x2(h[4], h[5]); wh.w = pack_bf16x2(h[6], h[7]);
          wl.x = pack_bf16x2(g[8 * q + 0] - h[0], g[8 * q + 1] - h[1]);
          wl.y = pack_bf16x2(g[8 * q + 2] - h[2], g[8 * q + 3] - h[3]);
          wl.z = pack_bf16x2(g[8 * q + 4] - h[4], g[8 * q + 5] - h[5]);
          wl.w = pack_bf16x2(g[8 * q + 6] - h[6], g[8 * q + 7] - h[7]);
          const uint32_t o = (ch >> 1) * CL_TILE + sw128_offset(i, (ch & 1) * 4 + q);
          *reinterpret_cast<uint4*>(smem + o) = wh;
          *reinterpret_cast<uint4*>(smem + 2 * CL_TILE + o) = wl;
        }
      }
    }
    if (!kBackward) {
      // deterministic block reduction of the five partial sums
      P1 = warp_sum(P1); P2 = warp_sum(P2); P3 = warp_sum(P3); P4 = warp_sum(P4); P5 = warp_sum(P5);
      if (lane == 0) {
        float* r = red + quarter * 8;
        r[0] = P1; r[1] = P2; r[2] = P3; r[3] = P4; r[4] = P5;
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      if (warp == 2 && lane < 5) {
        const float tot = red[lane] + red[8 + lane] + red[16 + lane] + red[24 + lane];
        p.partials[(static_cast<size_t>(call) * p.B + b) * 8 + lane] = tot;
      }
    } else {
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_full);
      mbar_wait(d_full, 0);
      tc_fence_after();
      // dA -> slot 0, dB -> slot of this call (slot 0 again for the intra call: code is on both GEMM sides)
      float* dA = p.dtiles + ((static_cast<size_t>(0) * p.B + b) * CL_ROWS + i) * CL_DT_LD;
      float* dB = p.dtiles + ((static_cast<size_t>(slotB) * p.B + b) * CL_ROWS + i) * CL_DT_LD;
#pragma unroll 1
      for (int ch = 0; ch < 3; ++ch) {  // 96 >= D columns
        if (ch * 32 >= p.D) break;
        uint32_t va[32], vb[32];
        tmem_ld32(TM_DA + lane_off + ch * 32, va);
        tmem_ld32(TM_DB + lane_off + ch * 32, vb);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const int c = ch * 32 + t;
            if (c < p.D) {
              atomicAdd(dA + c, __uint_as_float(va[t]));
              atomicAdd(dB + c, __uint_as_float(vb[t]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// 3. finish: per call statistics.  stats[call] = {loss_mean, cd_mean, old_mean, mean_c}
// ---------------------------------------------------------------------------------------------
__global__ void corr_finish_kernel(const float* __restrict__ partials, float* __restrict__ stats, int ncalls, int B,
                                   int S, int pointwise) {
  const int call = blockIdx.x;
  const int lane = threadIdx.x;  // 32 threads
  double acc[5] = {0, 0, 0, 0, 0};
  for (int b = lane; b < B; b += 32) {
    const float* r = partials + (static_cast<size_t>(call) * B + b) * 8;
    for (int k = 0; k < 5; ++k) acc[k] += static_cast<double>(r[k]);
  }
  for (int k = 0; k < 5; ++k)
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (lane == 0) {
    const double n = static_cast<double>(B) * S * S;
    const double old_mean = acc[2] / n;
    const double mean_c = acc[3] / n;
    const double offset = pointwise ? (old_mean - mean_c) : 0.0;
    stats[call * 4 + 0] = static_cast<float>(-(acc[0] + offset * acc[1]) / n);
    stats[call * 4 + 1] = static_cast<float>(acc[4] / n);
    stats[call * 4 + 2] = static_cast<float>(old_mean);
    stats[call * 4 + 3] = static_cast<float>(mean_c);
  }
}

// unreduced loss elements for API compatibility (modules.py:337-345 returns them for the negatives)
__global__ void corr_loss_elems_kernel(const float* __restrict__ cd, const float* __restrict__ fdc,
                                       const float* __restrict__ stats, float* __restrict__ loss, long long per_call,
                                       CorrParams p) {
  const long long idx = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_call * p.ncalls) return;
  const int call = static_cast<int>(idx / per_call);
  const float offset = p.pointwise ? (stats[call * 4 + 2] - stats[call * 4 + 3]) : 0.f;
  const float cl = fminf(fmaxf(cd[idx], p.clamp_lo), p.clamp_hi);
  loss[idx] = -cl * (fdc[idx] + offset - p.shift[call]);
}

// ---------------------------------------------------------------------------------------------
// 5. normalisation + grid_sample backward: one warp per (slot, image, sample)
// ---------------------------------------------------------------------------------------------
struct SampleBwdParams {
  SampleParams f;        // forward description of the code sampling (src = code, src_pos = code_pos)
  const float* dtiles;   // [nslots][B][128][CL_DT_LD]
  float* dsrc;           // gradient wrt src, same strides as src, fp32, zero-initialised / accumulated into
  float* dsrc_pos;
};

template <int NV>
__global__ void __launch_bounds__(256)
sample_norm_bwd_kernel(SampleBwdParams q) {
  const SampleParams& p = q.f;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int total = p.nslots * p.B * p.S;
  if (warp_global >= total) return;
  const int s = warp_global % p.S;
  const int b = (warp_global / p.S) % p.B;
  const int slot = warp_global / (p.S * p.B);
  const void* src; const float* cscale; const float* coords; int img;
  slot_source(p, slot, b, src, cscale, coords, img);
  float* dsrc = (slot == 1) ? q.dsrc_pos : q.dsrc;
  const Taps t = make_taps(coords, b, s, p.fs, p.H, p.W);
  auto off = [&](int pix) { return static_cast<long long>(pix / p.W) * p.sy + static_cast<long long>(pix % p.W) * p.sx; };
  const long long o00 = off(t.i00), o01 = off(t.i01), o10 = off(t.i10), o11 = off(t.i11);
  const long long base = static_cast<long long>(img) * p.sb;
  const float* g = q.dtiles + ((static_cast<size_t>(slot) * p.B + b) * CL_ROWS + s) * CL_DT_LD;
  float v[NV], gr[NV];
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    float val = 0.f, gg = 0.f;
    if (c < p.C) {
      const float* qf = reinterpret_cast<const float*>(src) + base + static_cast<long long>(c) * p.sc;
      val = qf[o00] * t.w00;
      val += qf[o01] * t.w01;
      val += qf[o10] * t.w10;
      val += qf[o11] * t.w11;
      gg = g[c];
    }
    v[k] = val; gr[k] = gg;
    ss += val * val;
    dot += val * gg;
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float nrm = sqrtf(ss);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    if (c >= p.C) continue;
    float dv;
    if (nrm > p.eps) {
      const float inv = 1.0f / nrm;
      dv = (gr[k] - v[k] * (dot * inv * inv)) * inv;  // d/dv of v/||v||
    } else {
      dv = gr[k] / p.eps;
    }
    float* d = dsrc + base + static_cast<long long>(c) * p.sc;
    if (t.w00 != 0.f) atomicAdd(d + o00, dv * t.w00);
    if (t.w01 != 0.f) atomicAdd(d + o01, dv * t.w01);
    if (t.w10 != 0.f) atomicAdd(d + o10, dv * t.w10);
    if (t.w11 != 0.f) atomicAdd(d + o11, dv * t.w11);
  }
}

static int encode_tile_maps(const void* ftiles, const void* ctiles, int nslots, int B, int E, CUtensorMap* tmF,
                            CUtensorMap* tmC) {
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)E, (uint64_t)2 * nslots * B * CL_ROWS};
    uint64_t str[1] = {(uint64_t)E * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(tmF, ftiles, 2, dims, str, box)) != STEGO_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)CL_CODE_PAD, (uint64_t)2 * nslots * B * CL_ROWS};
    uint64_t str[1] = {(uint64_t)CL_CODE_PAD * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(tmC, ctiles, 2, dims, str, box)) != STEGO_OK) return rc;
  }
  return STEGO_OK;
}

static int fill_sample_params(SampleParams& sp, const void* src, const void* src_pos, int src_bf16, long long sb,
                              long long sc, long long sy, long long sx, const float* chan_scale,
                              const float* chan_scale_pos, const float* coords1, const float* coords2,
                              const long long* perms, void* tiles, int B, int C, int Cpad, int H, int W, int fs,
                              int nslots) {
  STEGO_CHECK_ARG(src && src_pos && coords1 && coords2, "sample_norm: null pointer");
  STEGO_CHECK_ARG(nslots >= 2 && (nslots == 2 || perms), "sample_norm: nslots=%d needs perms", nslots);
  STEGO_CHECK_ARG(fs * fs <= CL_ROWS, "sample_norm: feature_samples^2 = %d exceeds the 128-row tile", fs * fs);
  STEGO_CHECK_ARG(C > 0 && C <= Cpad && Cpad % 64 == 0 && Cpad <= 768, "sample_norm: C=%d Cpad=%d", C, Cpad);
  STEGO_CHECK_ARG(B > 0 && H > 1 && W > 1, "sample_norm: B=%d H=%d W=%d", B, H, W);
  sp.src = src; sp.src_pos = src_pos; sp.src_bf16 = src_bf16;
  sp.sb = sb; sp.sc = sc; sp.sy = sy; sp.sx = sx;
  sp.chan_scale = chan_scale; sp.chan_scale_pos = chan_scale_pos;
  sp.coords1 = coords1; sp.coords2 = coords2; sp.perms = perms; sp.perms_raw = 0;
  sp.tiles = reinterpret_cast<bf16*>(tiles);
  sp.B = B; sp.C = C; sp.Cpad = Cpad; sp.H = H; sp.W = W; sp.fs = fs; sp.S = fs * fs; sp.nslots = nslots;
  sp.eps = 1e-10f;
  return STEGO_OK;
}

}  // namespace stego

using namespace stego;

extern "C" int stego_sample_norm_fwd(const void* src, const void* src_pos, int src_is_bf16, long long stride_b,
                                     long long stride_c, long long stride_y, long long stride_x,
                                     const float* chan_scale, const float* chan_scale_pos, const float* coords1,
                                     const float* coords2, const long long* perms, void* tiles, int B, int C,
                                     int Cpad, int H, int W, int feature_samples, int nslots, int perms_are_raw_randperm,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SampleParams sp;
  int rc = fill_sample_params(sp, src, src_pos, src_is_bf16, stride_b, stride_c, stride_y, stride_x, chan_scale,
                              chan_scale_pos, coords1, coords2, perms, tiles, B, C, Cpad, H, W, feature_samples, nslots);
  if (rc != STEGO_OK) return rc;
  sp.perms_raw = perms_are_raw_randperm;
  STEGO_CHECK_ARG(tiles, "stego_sample_norm_fwd: null tiles");
  const int warps = nslots * B * CL_ROWS;
  const int blocks = (warps + 7) / 8;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(src_pos) |
                         reinterpret_cast<uintptr_t>(tiles)) & 15u) == 0 &&
                       stride_b % 8 == 0 && stride_y % 8 == 0 && stride_x % 8 == 0;
  if (src_is_bf16 && stride_c == 1 && C % 8 == 0 && Cpad == C && aligned && C <= 768) {
    const int ni = (C + 255) / 256;
    if (ni == 1) sample_norm_vec8_kernel<1><<<blocks, 256, 0, stream>>>(sp);
    else if (ni == 2) sample_norm_vec8_kernel<2><<<blocks, 256, 0, stream>>>(sp);
    else sample_norm_vec8_kernel<3><<<blocks, 256, 0, stream>>>(sp);
    STEGO_CHECK_LAUNCH("sample_norm_vec8_kernel");
    return STEGO_OK;
  }
  switch (Cpad / 32) {
    case 2: sample_norm_kernel<2><<<blocks, 256, 0, stream>>>(sp); break;
    case 4: sample_norm_kernel<4><<<blocks, 256, 0, stream>>>(sp); break;
    case 6: sample_norm_kernel<6><<<blocks, 256, 0, stream>>>(sp); break;
    case 8: sample_norm_kernel<8><<<blocks, 256, 0, stream>>>(sp); break;
    case 12: sample_norm_kernel<12><<<blocks, 256, 0, stream>>>(sp); break;
    case 24: sample_norm_kernel<24><<<blocks, 256, 0, stream>>>(sp); break;
    default:
      set_error("stego_sample_norm_fwd: Cpad=%d unsupported (64,128,192,256,384,768)", Cpad);
      return STEGO_ERR_UNSUPPORTED;
  }
  STEGO_CHECK_LAUNCH("sample_norm_kernel");
  return STEGO_OK;
}

static int fill_corr_params(CorrParams& p, int B, int fs, int E, int D, int nslots, int ncalls, const int* slot_of_call,
                            const float* shifts, int pointwise, int zero_clamp, int stabilize) {
  STEGO_CHECK_ARG(B > 0 && fs * fs <= CL_ROWS && E % 64 == 0 && E >= 64, "corr: B=%d fs=%d E=%d", B, fs, E);
  STEGO_CHECK_ARG(D > 0 && D <= 96, "corr: code dim %d unsupported (<= 96)", D);
  STEGO_CHECK_ARG(ncalls > 0 && ncalls <= CL_MAX_CALLS && nslots >= 2, "corr: ncalls=%d nslots=%d", ncalls, nslots);
  p.B = B; p.S = fs * fs; p.E = E; p.D = D; p.nslots = nslots; p.ncalls = ncalls;
  for (int k = 0; k < ncalls; ++k) {
    STEGO_CHECK_ARG(slot_of_call[k] >= 0 && slot_of_call[k] < nslots, "corr: slot_of_call[%d]=%d", k, slot_of_call[k]);
    p.slot_of_call[k] = slot_of_call[k];
    p.shift[k] = shifts[k];
  }
  p.pointwise = pointwise;
  p.clamp_lo = zero_clamp ? 0.0f : -9999.0f;
  p.clamp_hi = stabilize ? 0.8f : INFINITY;
  p.partials = nullptr; p.cd_out = nullptr; p.fdc_out = nullptr;
  p.stats = nullptr; p.gscale = nullptr; p.gelem = nullptr; p.gcd = nullptr; p.dtiles = nullptr;
  return STEGO_OK;
}

template <bool kBwd>
static int launch_corr(const CUtensorMap& tmF, const CUtensorMap& tmC, const CorrParams& p, cudaStream_t stream) {
  constexpr int kRing = kBwd ? 2 : 3;
  constexpr size_t smem = size_t(kRing) * 2 * CL_TILE + 8 * CL_TILE + 512 + 1024;
  auto kern = corr_kernel<kBwd>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(corr)");
    configured = true;
  }
  kern<<<dim3(p.B, p.ncalls), CL_THREADS, smem, stream>>>(tmF, tmC, p);
  STEGO_CHECK_LAUNCH("corr_kernel");
  return STEGO_OK;
}

// host arrays (slot_of_call, shifts) are plain host pointers: they are copied into kernel parameters.
extern "C" int stego_corr_loss_fwd(const void* feat_tiles, const void* code_tiles, int B, int feature_samples, int E,
                                   int D, int nslots, int ncalls, const int* slot_of_call_host,
                                   const float* shifts_host, int pointwise, int zero_clamp, int stabilize,
                                   float* partials, float* stats, float* cd_out, float* fdc_out, float* loss_out,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STEGO_CHECK_ARG(feat_tiles && code_tiles && partials && stats && slot_of_call_host && shifts_host,
                  "stego_corr_loss_fwd: null pointer");
  STEGO_CHECK_ARG(!loss_out || (cd_out && fdc_out), "stego_corr_loss_fwd: loss_out needs cd_out and fdc_out");
  CorrParams p;
  int rc = fill_corr_params(p, B, feature_samples, E, D, nslots, ncalls, slot_of_call_host, shifts_host, pointwise,
                            zero_clamp, stabilize);
  if (rc != STEGO_OK) return rc;
  p.partials = partials; p.cd_out = cd_out; p.fdc_out = fdc_out;
  CUtensorMap tmF, tmC;
  if ((rc = encode_tile_maps(feat_tiles, code_tiles, nslots, B, E, &tmF, &tmC)) != STEGO_OK) return rc;
  if ((rc = launch_corr<false>(tmF, tmC, p, stream)) != STEGO_OK) return rc;
  corr_finish_kernel<<<ncalls, 32, 0, stream>>>(partials, stats, ncalls, B, p.S, pointwise);
  STEGO_CHECK_LAUNCH("corr_finish_kernel");
  if (loss_out) {
    const long long per_call = 1ll * B * p.S * p.S;
    const long long n = per_call * ncalls;
    corr_loss_elems_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cd_out, fdc_out, stats, loss_out, per_call, p);
    STEGO_CHECK_LAUNCH("corr_loss_elems_kernel");
  }
  return STEGO_OK;
}

extern "C" int stego_corr_loss_bwd(const void* feat_tiles, const void* code_tiles, int B, int feature_samples, int E,
                                   int D, int nslots, int ncalls, const int* slot_of_call_host,
                                   const float* shifts_host, int pointwise, int zero_clamp, int stabilize,
                                   const float* stats, const float* gscale, const float* gelem, const float* gcd,
                                   float* dtiles, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STEGO_CHECK_ARG(feat_tiles && code_tiles && stats && gscale && dtiles && slot_of_call_host && shifts_host,
                  "stego_corr_loss_bwd: null pointer");
  CorrParams p;
  int rc = fill_corr_params(p, B, feature_samples, E, D, nslots, ncalls, slot_of_call_host, shifts_host, pointwise,
                            zero_clamp, stabilize);
  if (rc != STEGO_OK) return rc;
  p.stats = stats; p.gscale = gscale; p.gelem = gelem; p.gcd = gcd; p.dtiles = dtiles;
  CUtensorMap tmF, tmC;
  if ((rc = encode_tile_maps(feat_tiles, code_tiles, nslots, B, E, &tmF, &tmC)) != STEGO_OK) return rc;
  return launch_corr<true>(tmF, tmC, p, stream);
}

extern "C" int stego_sample_norm_bwd(const float* code, const float* code_pos, long long stride_b, long long stride_c,
                                     long long stride_y, long long stride_x, const float* coords1,
                                     const float* coords2, const long long* perms, const float* dtiles, float* dcode,
                                     float* dcode_pos, int B, int C, int H, int W, int feature_samples, int nslots,
                                     int perms_are_raw_randperm, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STEGO_CHECK_ARG(dtiles && dcode && dcode_pos, "stego_sample_norm_bwd: null pointer");
  STEGO_CHECK_ARG(C <= 96, "stego_sample_norm_bwd: C=%d unsupported (<= 96)", C);
  SampleBwdParams q;
  int rc = fill_sample_params(q.f, code, code_pos, 0, stride_b, stride_c, stride_y, stride_x, nullptr, nullptr,
                              coords1, coords2, perms, nullptr, B, C, CL_CODE_PAD, H, W, feature_samples, nslots);
  if (rc != STEGO_OK) return rc;
  q.f.perms_raw = perms_are_raw_randperm;
  q.dtiles = dtiles; q.dsrc = dcode; q.dsrc_pos = dcode_pos;
  const int warps = nslots * B * q.f.S;
  sample_norm_bwd_kernel<3><<<(warps + 7) / 8, 256, 0, stream>>>(q);
  STEGO_CHECK_LAUNCH("sample_norm_bwd_kernel");
  return STEGO_OK;
}
