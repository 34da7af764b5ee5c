// Contrastive CRF loss (optional training term): src/modules.py:437-469
//     coords        = randint(h) x randint(w), n_samples points shared by the whole batch
//     sim_kernel    = w1 exp(-|dp|^2 / 2 alpha - |dI|^2 / 2 beta) + w2 exp(-|dp|^2 / 2 gamma) - shift      [B, n, n]
//     cluster_sims  = einsum("nka,nkb->nab", clusters[..., coords], clusters[..., coords])                 [B, n, n]
//     return -(cluster_sims * sim_kernel)
// The reference materialises coord_diff, guidance_diff ([B, 3, n, n]), two exponentials, the Gram matrix and the product
// as separate [B, n, n] tensors (n = 1000: 4 MB each per image).  Here: one gather kernel that lays the selected code
// vectors out k-major ([B][C][n], n contiguous), then one tile kernel that computes the 64 x 64 Gram tile with fp32 FMAs
// out of shared memory (fp32 parity with the reference; 4.5 GFLOP per step at B = 32 — not worth a tensor-core path),
// evaluates the pairwise kernel in registers and writes the product once.  The backward recomputes the (symmetric)
// pairwise kernel instead of saving it:  d sel[a] = sum_b -(g[a,b] + g[b,a]) sim[a,b] sel[b],  scattered back to the
// code gradient with atomics (coords may repeat).
#include "common.cuh"
#include "host_util.h"

namespace stego {

constexpr int CL_T = 64;   // output tile edge
constexpr int CL_WS = 65;  // row stride of the transposed weight tile (conflict-free both ways)

struct CrfLossParams {
  const float* sel;      // [B][C][NP] gathered code vectors, k-major, zero padded to NP = round_up(n, 64)
  const float4* gsel;    // [B][NP] gathered guidance (up to 3 channels + 0)
  const int2* pos;       // [NP] (y, x) of every sample
  int B, C, n, NP;
  float inv2a, inv2b, inv2g, w1, w2, shift;
  float* out;            // fwd: [B][n][n]
  const float* gout;     // bwd: [B][n][n] upstream gradient
  float* dsel;           // bwd: [B][C][NP]
};

__device__ __forceinline__ float crf_pair_kernel(const CrfLossParams& p, int2 pa, int2 pb, float4 ga, float4 gb) {
  const int dy = pa.x - pb.x, dx = pa.y - pb.y;
  const float cd = static_cast<float>(dy * dy + dx * dx);
  const float d0 = ga.x - gb.x, d1 = ga.y - gb.y, d2 = ga.z - gb.z;
  const float gd = d0 * d0 + d1 * d1 + d2 * d2;
  return p.w1 * expf(-cd * p.inv2a - gd * p.inv2b) + p.w2 * expf(-cd * p.inv2g) - p.shift;
}

// gather: one thread per (image, channel-or-guidance, sample)
struct CrfGatherParams {
  const float* clusters; long long c_sb, c_sc, c_sy, c_sx;   // element strides of [B, C, H, W]
  const float* guidance; long long g_sb, g_sc, g_sy, g_sx;   // [B, Cg, H, W]
  const long long* coords;                                    // [2][n]: row 0 indexes H, row 1 indexes W
  int B, C, Cg, n, NP, H, W;
  float* sel; float4* gsel; int2* pos;
};

__global__ void __launch_bounds__(256) crf_loss_gather_kernel(CrfGatherParams p) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  const int k = blockIdx.y, b = blockIdx.z;
  if (a >= p.NP) return;
  const bool real = a < p.n;
  int y = 0, x = 0;
  if (real) {
    y = static_cast<int>(p.coords[a]);
    x = static_cast<int>(p.coords[p.n + a]);
  }
  if (k < p.C) {
    p.sel[(1ll * b * p.C + k) * p.NP + a] = real ? p.clusters[b * p.c_sb + k * p.c_sc + y * p.c_sy + x * p.c_sx] : 0.f;
  } else {  // k == C: guidance + positions
    float g[3] = {0.f, 0.f, 0.f};
    if (real)
      for (int c = 0; c < p.Cg; ++c) g[c] = p.guidance[b * p.g_sb + c * p.g_sc + y * p.g_sy + x * p.g_sx];
    p.gsel[1ll * b * p.NP + a] = make_float4(g[0], g[1], g[2], 0.f);
    if (b == 0) p.pos[a] = make_int2(y, x);
  }
}

// forward: grid (NP/64, NP/64, B), 256 threads as 16 x 16, 4 x 4 outputs each
__global__ void __launch_bounds__(256) crf_loss_fwd_kernel(CrfLossParams p) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                  // [C][64]
  float* Bs = sm + p.C * CL_T;     // [C][64]
  __shared__ float4 ga_s[CL_T], gb_s[CL_T];
  __shared__ int2 pa_s[CL_T], pb_s[CL_T];
  const int b = blockIdx.z, a0 = blockIdx.y * CL_T, b0 = blockIdx.x * CL_T;
  const float* sel = p.sel + 1ll * b * p.C * p.NP;
  for (int i = threadIdx.x; i < p.C * (CL_T / 4); i += 256) {
    const int k = i / (CL_T / 4), q = i % (CL_T / 4);
    reinterpret_cast<float4*>(As)[i] = *reinterpret_cast<const float4*>(sel + 1ll * k * p.NP + a0 + 4 * q);
    reinterpret_cast<float4*>(Bs)[i] = *reinterpret_cast<const float4*>(sel + 1ll * k * p.NP + b0 + 4 * q);
  }
  if (threadIdx.x < CL_T) {
    ga_s[threadIdx.x] = p.gsel[1ll * b * p.NP + a0 + threadIdx.x];
    pa_s[threadIdx.x] = p.pos[a0 + threadIdx.x];
  } else if (threadIdx.x < 2 * CL_T) {
    const int t = threadIdx.x - CL_T;
    gb_s[t] = p.gsel[1ll * b * p.NP + b0 + t];
    pb_s[t] = p.pos[b0 + t];
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < p.C; ++k) {
    const float4 av = *reinterpret_cast<const float4*>(As + k * CL_T + 4 * ty);
    const float4 bv = *reinterpret_cast<const float4*>(Bs + k * CL_T + 4 * tx);
    const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
  }
  const bool vec = (p.n & 3) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + 4 * ty + i;
    if (a >= p.n) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = -(acc[i][j] * crf_pair_kernel(p, pa_s[4 * ty + i], pb_s[4 * tx + j], ga_s[4 * ty + i], gb_s[4 * tx + j]));
    float* dst = p.out + (1ll * b * p.n + a) * p.n + b0 + 4 * tx;
    if (vec && b0 + 4 * tx + 3 < p.n) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (b0 + 4 * tx + j < p.n) dst[j] = o[j];
    }
  }
}

// backward: grid (NP/64, B); the CTA owns 64 samples a and walks all b tiles
constexpr int CL_KMAX = 20;  // channels per thread group (C <= 80)
__global__ void __launch_bounds__(256) crf_loss_bwd_kernel(CrfLossParams p) {
  extern __shared__ __align__(16) float sm[];
  float* Bs = sm;                    // [C][64]   sel of the current b tile
  float* Wt = sm + p.C * CL_T;       // [64 b][64 a]  -(g[a,b] + g[b,a]) sim[a,b]
  __shared__ float4 ga_s[CL_T], gb_s[CL_T];
  __shared__ int2 pa_s[CL_T], pb_s[CL_T];
  const int b = blockIdx.y, a0 = blockIdx.x * CL_T;
  const float* sel = p.sel + 1ll * b * p.C * p.NP;
  const float* g = p.gout + 1ll * b * p.n * p.n;
  if (threadIdx.x < CL_T) {
    ga_s[threadIdx.x] = p.gsel[1ll * b * p.NP + a0 + threadIdx.x];
    pa_s[threadIdx.x] = p.pos[a0 + threadIdx.x];
  }
  const int a = threadIdx.x & 63, kq = threadIdx.x >> 6;
  const int kper = (p.C + 3) / 4, kbeg = kq * kper;
  float acc[CL_KMAX];
#pragma unroll
  for (int i = 0; i < CL_KMAX; ++i) acc[i] = 0.f;
  for (int b0 = 0; b0 < p.NP; b0 += CL_T) {
    __syncthreads();  // previous tile consumed (and ga_s / pa_s visible on the first pass)
    for (int i = threadIdx.x; i < p.C * (CL_T / 4); i += 256) {
      const int k = i / (CL_T / 4), q = i % (CL_T / 4);
      reinterpret_cast<float4*>(Bs)[i] = *reinterpret_cast<const float4*>(sel + 1ll * k * p.NP + b0 + 4 * q);
    }
    if (threadIdx.x < CL_T) {
      gb_s[threadIdx.x] = p.gsel[1ll * b * p.NP + b0 + threadIdx.x];
      pb_s[threadIdx.x] = p.pos[b0 + threadIdx.x];
    }
    // g[a0 + r][b0 + c] read row-wise (coalesced) into Wt[c][r]; g[b0 + r][a0 + c] read row-wise and added at Wt[r][c]
    for (int i = threadIdx.x; i < CL_T * CL_T; i += 256) {
      const int r = i >> 6, c = i & 63;
      const int ga = a0 + r, gb = b0 + c;
      Wt[c * CL_WS + r] = (ga < p.n && gb < p.n) ? g[1ll * ga * p.n + gb] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CL_T * CL_T; i += 256) {
      const int r = i >> 6, c = i & 63;  // r: b index, c: a index
      const int gb = b0 + r, ga = a0 + c;
      const float gt = (ga < p.n && gb < p.n) ? g[1ll * gb * p.n + ga] : 0.f;
      const float s = crf_pair_kernel(p, pa_s[c], pb_s[r], ga_s[c], gb_s[r]);
      Wt[r * CL_WS + c] = -(Wt[r * CL_WS + c] + gt) * s;
    }
    __syncthreads();
#pragma unroll 1
    for (int b4 = 0; b4 < CL_T / 4; ++b4) {
      const float w0 = Wt[(4 * b4 + 0) * CL_WS + a], w1 = Wt[(4 * b4 + 1) * CL_WS + a];
      const float w2 = Wt[(4 * b4 + 2) * CL_WS + a], w3 = Wt[(4 * b4 + 3) * CL_WS + a];
#pragma unroll
      for (int kk = 0; kk < CL_KMAX; ++kk) {
        if (kk < kper && kbeg + kk < p.C) {
          const float4 v = *reinterpret_cast<const float4*>(Bs + (kbeg + kk) * CL_T + 4 * b4);
          acc[kk] = fmaf(w0, v.x, fmaf(w1, v.y, fmaf(w2, v.z, fmaf(w3, v.w, acc[kk]))));
        }
      }
    }
  }
#pragma unroll
  for (int kk = 0; kk < CL_KMAX; ++kk)
    if (kk < kper && kbeg + kk < p.C) p.dsel[(1ll * b * p.C + kbeg + kk) * p.NP + a0 + a] = acc[kk];
}

struct CrfScatterParams {
  const float* dsel; const long long* coords; int B, C, n, NP;
  float* dclusters; long long sb, sc, sy, sx;
};
__global__ void __launch_bounds__(256) crf_loss_scatter_kernel(CrfScatterParams p) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  const int k = blockIdx.y, b = blockIdx.z;
  if (a >= p.n) return;
  const long long y = p.coords[a], x = p.coords[p.n + a];
  atomicAdd(p.dclusters + b * p.sb + k * p.sc + y * p.sy + x * p.sx, p.dsel[(1ll * b * p.C + k) * p.NP + a]);
}

static int crf_loss_check(int B, int C, int Cg, int n, int H, int W) {
  STEGO_CHECK_ARG(B > 0 && C > 0 && C <= 80 && Cg > 0 && Cg <= 3 && n > 0 && H > 0 && W > 0,
                  "stego_crf_loss: B=%d C=%d Cg=%d n=%d unsupported (C <= 80, guidance channels <= 3)", B, C, Cg, n);
  return STEGO_OK;
}

}  // namespace stego

using namespace stego;

// Workspace (caller-allocated): sel [B][C][NP] floats, gsel [B][NP] float4, pos [NP] int2, NP = round_up(n, 64).
// Strides are in elements; coords is the reference's [2][n] int64 tensor (row 0 indexes H, row 1 indexes W).
extern "C" int stego_crf_loss_fwd(const float* guidance, long long g_sb, long long g_sc, long long g_sy, long long g_sx, int Cg,
                                  const float* clusters, long long c_sb, long long c_sc, long long c_sy, long long c_sx, int C,
                                  const long long* coords, int B, int n, int H, int W, float alpha, float beta, float gamma,
                                  float w1, float w2, float shift, float* sel, float* gsel, int* pos, float* out,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STEGO_CHECK_ARG(guidance && clusters && coords && sel && gsel && pos && out, "stego_crf_loss_fwd: null pointer");
  if (int rc = crf_loss_check(B, C, Cg, n, H, W)) return rc;
  const int NP = (n + CL_T - 1) / CL_T * CL_T;
  CrfGatherParams gp;
  gp.clusters = clusters; gp.c_sb = c_sb; gp.c_sc = c_sc; gp.c_sy = c_sy; gp.c_sx = c_sx;
  gp.guidance = guidance; gp.g_sb = g_sb; gp.g_sc = g_sc; gp.g_sy = g_sy; gp.g_sx = g_sx;
  gp.coords = coords; gp.B = B; gp.C = C; gp.Cg = Cg; gp.n = n; gp.NP = NP; gp.H = H; gp.W = W;
  gp.sel = sel; gp.gsel = reinterpret_cast<float4*>(gsel); gp.pos = reinterpret_cast<int2*>(pos);
  crf_loss_gather_kernel<<<dim3((NP + 255) / 256, C + 1, B), 256, 0, stream>>>(gp);
  STEGO_CHECK_LAUNCH("crf_loss_gather_kernel");
  CrfLossParams p;
  p.sel = sel; p.gsel = reinterpret_cast<const float4*>(gsel); p.pos = reinterpret_cast<const int2*>(pos);
  p.B = B; p.C = C; p.n = n; p.NP = NP;
  p.inv2a = 1.0f / (2.0f * alpha); p.inv2b = 1.0f / (2.0f * beta); p.inv2g = 1.0f / (2.0f * gamma);
  p.w1 = w1; p.w2 = w2; p.shift = shift; p.out = out; p.gout = nullptr; p.dsel = nullptr;
  crf_loss_fwd_kernel<<<dim3(NP / CL_T, NP / CL_T, B), 256, (size_t)2 * C * CL_T * sizeof(float), stream>>>(p);
  STEGO_CHECK_LAUNCH("crf_loss_fwd_kernel");
  return STEGO_OK;
}

// Backward with the workspace the forward filled (sel, gsel, pos).  grad_out [B][n][n] contiguous; dsel [B][C][NP] scratch;
// dclusters (same strides as clusters) is ACCUMULATED into: the caller zero-fills it.
extern "C" int stego_crf_loss_bwd(const float* grad_out, const float* sel, const float* gsel, const int* pos,
                                  const long long* coords, int B, int C, int n, float alpha, float beta, float gamma, float w1,
                                  float w2, float shift, float* dsel, float* dclusters, long long c_sb, long long c_sc,
                                  long long c_sy, long long c_sx, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STEGO_CHECK_ARG(grad_out && sel && gsel && pos && coords && dsel && dclusters, "stego_crf_loss_bwd: null pointer");
  if (int rc = crf_loss_check(B, C, 1, n, 1, 1)) return rc;
  const int NP = (n + CL_T - 1) / CL_T * CL_T;
  CrfLossParams p;
  p.sel = sel; p.gsel = reinterpret_cast<const float4*>(gsel); p.pos = reinterpret_cast<const int2*>(pos);
  p.B = B; p.C = C; p.n = n; p.NP = NP;
  p.inv2a = 1.0f / (2.0f * alpha); p.inv2b = 1.0f / (2.0f * beta); p.inv2g = 1.0f / (2.0f * gamma);
  p.w1 = w1; p.w2 = w2; p.shift = shift; p.out = nullptr; p.gout = grad_out; p.dsel = dsel;
  const size_t smem = ((size_t)C * CL_T + CL_T * CL_WS) * sizeof(float);
  crf_loss_bwd_kernel<<<dim3(NP / CL_T, B), 256, smem, stream>>>(p);
  STEGO_CHECK_LAUNCH("crf_loss_bwd_kernel");
  CrfScatterParams sp;
  sp.dsel = dsel; sp.coords = coords; sp.B = B; sp.C = C; sp.n = n; sp.NP = NP;
  sp.dclusters = dclusters; sp.sb = c_sb; sp.sc = c_sc; sp.sy = c_sy; sp.sx = c_sx;
  crf_loss_scatter_kernel<<<dim3((n + 255) / 256, C, B), 256, 0, stream>>>(sp);
  STEGO_CHECK_LAUNCH("crf_loss_scatter_kernel");
  return STEGO_OK;
}
