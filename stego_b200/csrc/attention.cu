// Fused multi-head self-attention forward for the frozen DINO ViT (sm_100a, tcgen05 + TMA + TMEM).
//
// Reference: src/dino/vision_transformer.py:78-90 (Attention.forward):
//     attn = softmax(q k^T * head_dim^-0.5);  x = attn v
// The reference materialises the [B, heads, N, N] fp32 score tensor three times per layer; here it never leaves the
// SM: S = Q K^T accumulates in TMEM, softmax runs out of TMEM in registers, P (bf16) is written BACK INTO TMEM over
// the scores it came from and feeds the tensor core directly as the A operand of O += P V (no shared-memory trip).
//
// Persistent kernel, one CTA per SM, head_dim = 64, 384 threads = three warpgroups (each SM sub-partition hosts one
// warp of each; the CTA's register pool is re-split with setmaxnreg: 56 per thread for the TMA / MMA warpgroup — 4 x 112
// released = 8 x 56 acquired — and 224 for the two softmax warpgroups).  A work item is a PAIR of 128-query tiles of
// one (head, image) — both tiles share one K/V stream (96-key tiles), so K/V shared-memory traffic per flop is half
// that of one tile per CTA — or, when the number of query tiles is odd, the single last tile:
//   warp 0      TMA producer: Q tiles, then a 4-stage ring of K tiles and a 4-stage ring of V tiles, running ahead
//               across work items (the next item's operands arrive while the current one finishes)
//   warp 1      MMA issuer (whole warp in uniform control flow, one elected lane): per KV tile j and query tile t
//               O^t += P^t_j V_j  then  S^t_{j+2} = Q^t K_{j+2}^T  — S runs TWO tiles ahead of P V
//   warps 2, 3  idle (they only pad the first warpgroup)
//   warps 4..7  softmax warpgroup 0 = query tile 0   } one thread per query row: the 96-key score row is read from
//   warps 8..11 softmax warpgroup 1 = query tile 1   } TMEM once and stays in registers
// TMEM (512 columns): per query tile two 96-column score buffers (S_j in buffer j & 1) + a 64-column O accumulator.
// P_j is written over the scores it came from (first 48 columns of buffer j & 1); tcgen05.mma instructions of one
// thread execute in issue order, so S_{j+2} — issued right behind P_j V_j — may overwrite them.  Because S_{j+1} already
// sits in the other buffer when a warpgroup finishes tile j, the softmax warps never wait for the tensor core or for the
// MMA warp's wake-up latency: the first version of this kernel had ONE 128-column score buffer per tile
// (softmax -> P V -> S -> softmax was a serial chain per tile, 2.2 us per 128 keys) and was slower than round 1's.
// The running max is updated lazily (FA4-style): the reference max only moves when the new row max exceeds it by more
// than 2^8 — then, and only then, the warpgroup waits for its previous P V and rescales O in TMEM; otherwise
// P = exp2(s - m_ref) <= 256, harmless in bf16 / fp32.
// Exponentials: MUFU.EX2 caps a head_dim-64 attention at 16 exp/clk/SM = 1.19 PFLOP/s, so ATT_POLY of every 8
// exponentials are computed on the FMA pipe instead (Cody-Waite split + cubic, packed f32x2: rel. error 1e-4, far
// below the bf16 rounding of P).
// Ragged edges (N = hw + 1 is never a multiple of the tile sizes): the last KV tile's S uses an N = round16(valid keys)
// MMA and its P V a K of the same size; out-of-range keys are zero-filled by TMA and masked; rows past N of the last
// query tile are zero-filled, their warps skip the math when all 32 rows are padding, and the TMA store clips them.
// Input is the packed qkv GEMM output [B, N, 3E] bf16 (q | k | v, head-major inside each), read through ONE 3-D
// tensor map.
#include "common.cuh"
#include "host_util.h"

namespace stego {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 96;
constexpr int ATT_D = 64;
constexpr int ATT_KST = 4;  // K ring stages
constexpr int ATT_VST = 4;  // V ring stages
constexpr int ATT_THREADS = 384;  // 3 warpgroups: {TMA warp, MMA warp, 2 idle}, softmax tile 0, softmax tile 1
#ifndef ATT_POLY
#define ATT_POLY 2  // of every 8 exponentials, how many run on the FMA pipe (even, 0..8)
#endif

constexpr uint32_t ATT_TILE_BYTES = 128 * 64 * 2;   // Q / output tile, 16 KB: [128 rows][64 bf16], 128-byte swizzle
constexpr uint32_t ATT_KV_BYTES = ATT_BKV * 64 * 2;  // K / V tile, 12 KB
constexpr uint32_t ATT_SMEM_Q = 0;                                        // 2 tiles
constexpr uint32_t ATT_SMEM_K = ATT_SMEM_Q + 2 * ATT_TILE_BYTES;          // ring
constexpr uint32_t ATT_SMEM_V = ATT_SMEM_K + ATT_KST * ATT_KV_BYTES;      // ring
constexpr uint32_t ATT_SMEM_OUT = ATT_SMEM_V + ATT_VST * ATT_KV_BYTES;    // one output staging tile per warpgroup
constexpr uint32_t ATT_SMEM_BAR = ATT_SMEM_OUT + 2 * ATT_TILE_BYTES;
constexpr uint32_t ATT_SMEM_TOTAL = ATT_SMEM_BAR + 512;
constexpr uint32_t ATT_TMEM_COLS = 512;  // S^t buffer b at (2 t + b) * 96 (P aliases its first 48 columns); O^t at 384 + 64 t
constexpr uint32_t ATT_TM_S = 0, ATT_TM_O = 4 * ATT_BKV;

struct AttnParams {
  int N, E, heads;
  int npair;        // full pairs of query tiles per (head, image)
  int n_heavy;      // B * heads * npair pair items (scheduled first)
  int n_items;      // + B * heads single-tile items if the tile count is odd
  int nkv;          // KV tiles
  int last_valid;   // real keys in the last KV tile
  int nk_last;      // MMA N / K of the last KV tile: round_up(last_valid, 16)
  float scale_log2e;
};

// 2^x for x <= ~9 on the FMA pipe, two lanes at a time: x = n + r, n = round(x), r in [-0.5, 0.5];
// 2^r by a cubic (max rel. error 1.0e-4), 2^n by adding n to the exponent field.
__device__ __forceinline__ void exp2_poly2(float a0, float a1, float& e0, float& e1) {
  constexpr float MAGIC = 12582912.0f;  // 1.5 * 2^23: (x + MAGIC) holds round(x) in its low mantissa bits
  a0 = fmaxf(a0, -126.0f);
  a1 = fmaxf(a1, -126.0f);
  const uint64_t x2 = pack_f32x2(a0, a1);
  const uint64_t t2 = add_f32x2(x2, pack_f32x2(MAGIC, MAGIC));
  const uint64_t n2 = add_f32x2(t2, pack_f32x2(-MAGIC, -MAGIC));
  const uint64_t r2 = fma_f32x2(n2, pack_f32x2(-1.0f, -1.0f), x2);
  uint64_t p2 = fma_f32x2(r2, pack_f32x2(0.05592203512787819f, 0.05592203512787819f),
                          pack_f32x2(0.24264007806777954f, 0.24264007806777954f));
  p2 = fma_f32x2(p2, r2, pack_f32x2(0.6931210160255432f, 0.6931210160255432f));
  p2 = fma_f32x2(p2, r2, pack_f32x2(0.9999244809150696f, 0.9999244809150696f));
  float p0, p1, t0, t1;
  unpack_f32x2(p2, p0, p1);
  unpack_f32x2(t2, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// 32 scores -> 32 exponentials (bf16 pairs in w) + packed partial row sums.
__device__ __forceinline__ void exp_chunk(const uint32_t (&v)[32], float c, float nmc, uint32_t (&w)[16], uint64_t& rsa,
                                          uint64_t& rsb) {
  const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(nmc, nmc);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = 8 * g + 2 * t;
      float a0, a1, e0, e1;
      unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nmc2), a0, a1);
      if (2 * t < ATT_POLY) {
        exp2_poly2(a0, a1, e0, e1);
      } else {
        e0 = ex2_approx(a0);
        e1 = ex2_approx(a1);
      }
      w[4 * g + t] = pack_bf16x2(e0, e1);
      if (t & 1) rsb = add_f32x2(rsb, pack_f32x2(e0, e1));
      else rsa = add_f32x2(rsa, pack_f32x2(e0, e1));
    }
  }
}

// ragged last KV tile: keys past N (zero-filled by TMA, or never written when beyond the N of this S MMA) become -inf
// scores -> exponential 0, so the common path needs no per-element mask
__device__ __forceinline__ void mask_tail(uint32_t (&x)[32], int col0, int valid) {
#pragma unroll
  for (int t = 0; t < 32; ++t)
    if (col0 + t >= valid) x[t] = 0xff800000u;
}
__device__ __forceinline__ void ld32_as_2x16(uint32_t taddr, uint32_t (&x)[32]) {
  uint32_t a[16], b[16];
  tmem_ld16(taddr, a);
  tmem_ld16(taddr + 16, b);
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    x[t] = a[t];
    x[16 + t] = b[t];
  }
}
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float m) {
  float a = m, b = -INFINITY;  // two chains
#pragma unroll
  for (int t = 0; t < 32; t += 2) {
    a = fmaxf(a, __uint_as_float(v[t]));
    b = fmaxf(b, __uint_as_float(v[t + 1]));
  }
  return fmaxf(a, b);
}
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmOut, AttnParams p) {
  // no static shared memory in this kernel: the dynamic window starts at offset 0 of the CTA's allocation and the
  // __align__(1024) below is honoured (128B-swizzled tiles need 1024-byte alignment); checked at run time.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
  uint64_t* q_full = bars;                 // [2]  TMA -> MMA
  uint64_t* q_empty = q_full + 2;          // [2]  MMA (all S of the item retired) -> TMA
  uint64_t* k_full = q_empty + 2;          // [KST]
  uint64_t* k_empty = k_full + ATT_KST;    // [KST]
  uint64_t* v_full = k_empty + ATT_KST;    // [VST]
  uint64_t* v_empty = v_full + ATT_VST;    // [VST]
  uint64_t* s_full = v_empty + ATT_VST;    // [2][2]  MMA -> softmax warpgroup t: score buffer b holds S_j (j & 1 == b)
  uint64_t* p_full = s_full + 4;           // [2]  softmax warpgroup t (4 warps) -> MMA: P_j is in TMEM
  uint64_t* pv_done = p_full + 2;          // [2]  MMA -> softmax warpgroup t: one phase per P_j V_j retired (rescale path)
  uint64_t* o_done = pv_done + 2;          // [2]  MMA -> softmax warpgroup t: the LAST P V of the item has retired
  uint64_t* o_free = o_done + 2;           // [2]  softmax warpgroup t holds O in registers -> MMA may start the next item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&q_empty[t], 1);
      mbar_init(&s_full[2 * t], 1);
      mbar_init(&s_full[2 * t + 1], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_done[t], 1);
      mbar_init(&o_free[t], 4);
    }
    for (int s = 0; s < ATT_KST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < ATT_VST; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkv = p.nkv;
  // item -> (image, head, first query row, number of query tiles)
  auto decode = [&](int item, int& img, int& head, int& q0, int& nqt) {
    int bh, pair;
    if (item < p.n_heavy) {
      pair = item % p.npair;
      bh = item / p.npair;
      nqt = 2;
    } else {
      pair = p.npair;
      bh = item - p.n_heavy;
      nqt = 1;
    }
    head = bh % p.heads;
    img = bh / p.heads;
    q0 = pair * 2 * ATT_BQ;
  };

  // The setmaxnreg of a role sits INSIDE its branch: ptxas sizes the register allocation of the code it dominates
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t uq[2] = {0, 0}, kc = 0, vc = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int img, head, q0, nqt;
        decode(item, img, head, q0, nqt);
        for (int t = 0; t < nqt; ++t) {
          mbar_wait(&q_empty[t], (uq[t] & 1u) ^ 1u);
          mbar_arrive_expect_tx(&q_full[t], ATT_TILE_BYTES);
          tma_load_3d(smem + ATT_SMEM_Q + t * ATT_TILE_BYTES, &tmQKV, &q_full[t], head * ATT_D, q0 + t * ATT_BQ, img);
          ++uq[t];
        }
        for (int j = 0; j < nkv; ++j) {
          const uint32_t ks = kc % ATT_KST, kph = (kc / ATT_KST) & 1u;
          mbar_wait(&k_empty[ks], kph ^ 1u);
          mbar_arrive_expect_tx(&k_full[ks], ATT_KV_BYTES);
          tma_load_3d(smem + ATT_SMEM_K + ks * ATT_KV_BYTES, &tmKV, &k_full[ks], p.E + head * ATT_D, j * ATT_BKV, img);
          ++kc;
          const uint32_t vs = vc % ATT_VST, vph = (vc / ATT_VST) & 1u;
          mbar_wait(&v_empty[vs], vph ^ 1u);
          mbar_arrive_expect_tx(&v_full[vs], ATT_KV_BYTES);
          tma_load_3d(smem + ATT_SMEM_V + vs * ATT_KV_BYTES, &tmKV, &v_full[vs], 2 * p.E + head * ATT_D, j * ATT_BKV, img);
          ++vc;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the schedule (warp-uniform control flow keeps descriptors in uniform registers) and one
    // elected lane issues.
    constexpr uint32_t DESC_HI = smem_desc_hi_sw128(1024);
    constexpr uint32_t IDESC_O = make_idesc_bf16(128, ATT_D, 0, 1);  // P (TMEM, K-major) x V (MN-major: d contiguous)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t q_lo0 = smem_desc_lo(smem_u32(smem + ATT_SMEM_Q), 16);
    const uint32_t k_lo0 = smem_desc_lo(smem_u32(smem + ATT_SMEM_K), 16);
    const uint32_t v_lo0 = smem_desc_lo(smem_u32(smem + ATT_SMEM_V), 8192);
    uint32_t uq[2] = {0, 0}, up[2] = {0, 0}, uo[2] = {0, 0}, kc = 0, vc = 0;
    auto issue_s = [&](int t, int buf, uint32_t ks, int ncols) {  // S^t = Q^t K^T (128 x ncols x 64) into score buffer buf
      const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(ncols), 0, 0);
      const uint32_t q_lo = q_lo0 + t * (ATT_TILE_BYTES >> 4);
      const uint32_t k_lo = k_lo0 + ks * (ATT_KV_BYTES >> 4);
      if (elect_one()) {
#pragma unroll
        for (uint32_t k = 0; k < ATT_D / 16; ++k)
          umma_bf16(tm + ATT_TM_S + (2 * t + buf) * ATT_BKV, smem_desc_join(q_lo + k * 2, DESC_HI),
                    smem_desc_join(k_lo + k * 2, DESC_HI), idesc, k > 0 ? 1u : 0u);
        umma_commit(&s_full[2 * t + buf]);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int nqt = (item < p.n_heavy) ? 2 : 1;
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (t < nqt) {
          mbar_wait(&q_full[t], uq[t] & 1u);
          ++uq[t];
        }
      auto ncols_of = [&](int j) { return (j == nkv - 1) ? p.nk_last : ATT_BKV; };
      // every S of the item issued -> Q may be overwritten once they retire
      auto release_q = [&]() {
        if (elect_one())
          for (int t = 0; t < nqt; ++t) umma_commit(&q_empty[t]);
        __syncwarp();
      };
      for (int jj = 0; jj < 2 && jj < nkv; ++jj) {  // S_0 and S_1: the two score buffers of each query tile
        const uint32_t ks = kc % ATT_KST;
        mbar_wait(&k_full[ks], (kc / ATT_KST) & 1u);
        tc_fence_after();
        for (int t = 0; t < nqt; ++t) issue_s(t, jj, ks, ncols_of(jj));
        if (elect_one()) umma_commit(&k_empty[ks]);
        __syncwarp();
        ++kc;
        if (jj == nkv - 1) release_q();
      }
      for (int j = 0; j < nkv; ++j) {
        const uint32_t vs = vc % ATT_VST;
        const uint32_t ks = kc % ATT_KST;  // stage of K_{j+2}
        const int ksteps = ncols_of(j) / 16;
        const bool more = (j + 2 < nkv);
        const int buf = j & 1;
        mbar_wait(&v_full[vs], (vc / ATT_VST) & 1u);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t >= nqt) break;
          mbar_wait(&p_full[t], up[t] & 1u);  // P^t_j is in TMEM (and O^t rescaled, if it had to be)
          ++up[t];
          if (j == 0) {
            mbar_wait(&o_free[t], (uo[t] & 1u) ^ 1u);  // the previous item's O^t has been read out
            ++uo[t];
          }
          if (more && t == 0) mbar_wait(&k_full[ks], (kc / ATT_KST) & 1u);
          tc_fence_after();
          const uint32_t v_lo = v_lo0 + vs * (ATT_KV_BYTES >> 4);
          if (elect_one()) {
            for (int kk = 0; kk < ksteps; ++kk)
              umma_bf16_ts(tm + ATT_TM_O + t * ATT_D, tm + ATT_TM_S + (2 * t + buf) * ATT_BKV + kk * 8,
                           smem_desc_join(v_lo + kk * (2048u >> 4), DESC_HI), IDESC_O, (j > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&pv_done[t]);
            if (j == nkv - 1) umma_commit(&o_done[t]);
          }
          __syncwarp();
          if (more) issue_s(t, buf, ks, ncols_of(j + 2));  // into the buffer P^t_j vacates (in-order execution)
        }
        if (elect_one()) {
          umma_commit(&v_empty[vs]);
          if (more) umma_commit(&k_empty[ks]);
        }
        __syncwarp();
        ++vc;
        if (more) {
          ++kc;
          if (j + 2 == nkv - 1) release_q();
        }
      }
    }
  }
  } else {
    // ===================== softmax warpgroups =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;\n");
    const int wg = (warp - 4) >> 2;     // query tile of the pair
    const int quarter = warp & 3;       // TMEM lane quarter accessible to this warp
    const int r = quarter * 32 + lane;  // query row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t ts0 = tmem_base + ATT_TM_S + wg * 2 * ATT_BKV + lane_off;  // score buffer 0 (buffer 1: + 96 columns)
    const uint32_t to = tmem_base + ATT_TM_O + wg * ATT_D + lane_off;
    uint8_t* stage = smem + ATT_SMEM_OUT + wg * ATT_TILE_BYTES;
    const bool storer = (warp & 3) == 0 && lane == 0;
    const float c = p.scale_log2e;
    uint32_t us0 = 0, us1 = 0;  // uses of the two score-buffer barriers
    uint32_t uod = 0;           // items finished (phases of o_done)
    uint32_t npv = 0;         // P V products this warpgroup has fed so far = phases of pv_done it may wait on
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int img, head, q0, nqt;
      decode(item, img, head, q0, nqt);
      if (wg >= nqt) continue;  // single-tile item: warpgroup 1 has nothing to do
      const int qt0 = q0 + wg * ATT_BQ;
      const bool warp_has_rows = (qt0 + quarter * 32) < p.N;
      float m_run = -INFINITY, l_run = 0.f;  // m_run: reference max the exponentials are taken against
      for (int j = 0; j < nkv; ++j) {
        const bool last = (j == nkv - 1);
        const int ncols = last ? p.nk_last : ATT_BKV;
        const int valid = last ? p.last_valid : ATT_BKV;
        const int buf = j & 1;
        const uint32_t ts = ts0 + buf * ATT_BKV;  // S_j; P_j aliases its first 48 columns
        mbar_wait(&s_full[2 * wg + buf], (buf ? us1 : us0) & 1u);
        if (buf) ++us1; else ++us0;
        tc_fence_after();
        if (!warp_has_rows) {
          // ragged last query tile: this warp's 32 rows are all padding — keep the barrier protocol, skip the math
          // (their P rows feed only O rows that are never stored)
          tc_fence_before();
          __syncwarp();
          if (npv > 0) mbar_wait(&p_full[wg], (npv - 1u) & 1u);  // see below: never arrive twice on one phase
          if (lane == 0) mbar_arrive(&p_full[wg]);
          ++npv;
          continue;
        }
        // the whole 96-key score row comes out of TMEM once and stays in registers for max, exp and packing
        uint32_t v[3][32];
        const bool ragged = last && valid < ATT_BKV;  // warp-uniform
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          if (ch * 32 < ncols) ld32_as_2x16(ts + ch * 32, v[ch]);
        tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          if (ch * 32 < ncols) {
            if (ragged) mask_tail(v[ch], ch * 32, valid);
            mx = max32(v[ch], mx);
          }
        // lazy reference-max update: move it only on the first tile or when it is off by more than 2^8
        const bool first = (j == 0);
        const bool move = first || ((mx - m_run) * c > 8.0f);
        float alpha = 1.0f;
        if (move) {
          alpha = first ? 0.0f : ex2_approx((m_run - mx) * c);
          m_run = mx;
          l_run *= alpha;
        }
        // Rare path: O must be rescaled.  S runs two tiles ahead, so P_{j-1} V_{j-1} may still be in flight: wait for its
        // pv_done phase (phase npv - 1; P_j V_j cannot be issued before this warpgroup delivers P_j, so the barrier is at
        // most one phase ahead of what is awaited).  tcgen05.ld / st are warp-collective: every lane takes the path, with
        // alpha = 1 if its row stays.
        if (!first && __any_sync(0xffffffffu, move)) {
          mbar_wait(&pv_done[wg], (npv - 1u) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int h = 0; h < ATT_D / 8; ++h) {  // 8 columns at a time keeps the register budget
            uint32_t o[8];
            tmem_ld8(to + h * 8, o);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
            tmem_st8(to + h * 8, o);
          }
        }
        const float nmc = -m_run * c;
        uint64_t rsa = 0ull, rsb = 0ull;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          if (ch * 32 < ncols) {
            uint32_t w[16];
            exp_chunk(v[ch], c, nmc, w, rsa, rsb);
            tmem_st16(ts + ch * 16, w);  // P: two bf16 per 32-bit column, columns [16 ch, 16 ch + 16)
          }
        }
        float s0, s1;
        unpack_f32x2(add_f32x2(rsa, rsb), s0, s1);
        l_run += s0 + s1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        // S_{j+1} is ready long before P_j V_j, so nothing stops a fast warp from finishing tile j+1 while a slower warp
        // of its warpgroup is still on tile j: its second arrival would complete the 4-arrival phase of tile j without
        // the slow warp's rows.  Arrive for tile g only once the phase of tile g-1 has completed (usually long ago).
        if (npv > 0) mbar_wait(&p_full[wg], (npv - 1u) & 1u);
        if (lane == 0) mbar_arrive(&p_full[wg]);
        ++npv;
      }
      // ---- epilogue of the item: O / l -> bf16 -> swizzled staging tile -> one TMA store (rows >= N clipped)
      // (a parity wait on pv_done cannot serve here: with S two tiles ahead the barrier may be one OR two phases behind)
      mbar_wait(&o_done[wg], uod & 1u);  // the last P V of the item has retired
      ++uod;
      tc_fence_after();
      uint32_t o0[32], o1[32];
      if (warp_has_rows) {
        tmem_ld32(to, o0);
        tmem_ld32(to + 32, o1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[wg]);  // the MMA warp may start accumulating the next item into O
      if (storer) tma_wait_group_read<0>();     // the previous item's store has finished reading the staging tile
      asm volatile("bar.sync %0, 128;\n" ::"r"(1 + wg) : "memory");
      if (warp_has_rows) {
        const float inv = 1.0f / l_run;
        uint8_t* srow = stage + r * 128;
        const uint32_t rx = (static_cast<uint32_t>(r) & 7u) << 4;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t wv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t a = h ? o1[8 * g + 2 * t] : o0[8 * g + 2 * t];
              const uint32_t b = h ? o1[8 * g + 2 * t + 1] : o0[8 * g + 2 * t + 1];
              wv[t] = pack_bf16x2(__uint_as_float(a) * inv, __uint_as_float(b) * inv);
            }
            *reinterpret_cast<uint4*>(srow + ((static_cast<uint32_t>(h * 4 + g) << 4) ^ rx)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          }
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;\n" ::"r"(1 + wg) : "memory");
      if (storer) {
        tma_store_3d(stage, &tmOut, head * ATT_D, qt0, img);
        tma_commit_group();
      }
    }
    if (storer) tma_wait_group_read<0>();  // the staging tile must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

}  // namespace stego

using namespace stego;

// qkv: [B][N][3E] bf16 packed (q|k|v), out: [B][N][E] bf16.
extern "C" int stego_attention_fwd(const void* qkv, void* out, int B, int N, int E, int heads, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STEGO_CHECK_ARG(qkv && out, "stego_attention_fwd: null pointer");
  STEGO_CHECK_ARG(B > 0 && N > 0 && heads > 0, "stego_attention_fwd: bad sizes");
  STEGO_CHECK_ARG(E == heads * ATT_D, "stego_attention_fwd: head_dim must be 64 (E=%d heads=%d)", E, heads);
  STEGO_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "stego_attention_fwd: out not 16-byte aligned");
  CUtensorMap tm;
  uint64_t dims[3] = {(uint64_t)3 * E, (uint64_t)N, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)3 * E * 2, (uint64_t)N * 3 * E * 2};
  uint32_t box[3] = {64, 128, 1};       // one 128-row x 64-column box per Q tile (and per output tile)
  uint32_t kvbox[3] = {64, ATT_BKV, 1};  // K / V tiles: 96 rows
  int rc = make_tmap_bf16(&tm, qkv, 3, dims, str, box);
  if (rc != STEGO_OK) return rc;
  CUtensorMap tmkv;
  if ((rc = make_tmap_bf16(&tmkv, qkv, 3, dims, str, kvbox)) != STEGO_OK) return rc;
  CUtensorMap tmo;  // output [B][N][E]: per-image row clipping for the ragged last query tile
  uint64_t odims[3] = {(uint64_t)E, (uint64_t)N, (uint64_t)B};
  uint64_t ostr[2] = {(uint64_t)E * 2, (uint64_t)N * E * 2};
  if ((rc = make_tmap_bf16(&tmo, out, 3, odims, ostr, box)) != STEGO_OK) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ATT_SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attention)");
    configured = true;
  }
  AttnParams p;
  p.N = N;
  p.E = E;
  p.heads = heads;
  const int nq = (N + ATT_BQ - 1) / ATT_BQ;
  p.npair = nq / 2;
  p.n_heavy = B * heads * p.npair;
  p.n_items = p.n_heavy + ((nq & 1) ? B * heads : 0);
  p.nkv = (N + ATT_BKV - 1) / ATT_BKV;
  p.last_valid = N - (p.nkv - 1) * ATT_BKV;
  p.nk_last = (p.last_valid + 15) & ~15;
  p.scale_log2e = 0.125f * 1.4426950408889634f;
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  attention_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM_TOTAL, stream>>>(tm, tmkv, tmo, p);
  STEGO_CHECK_LAUNCH("attention_fwd_kernel");
  return STEGO_OK;
}
