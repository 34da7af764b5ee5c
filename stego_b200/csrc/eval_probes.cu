// Fused evaluation probes (BASELINE.json configs[4]: 1024 x 2048 frames): the reference's eval loop
// (src/eval_segmentation.py:128-131) does
//     code = F.interpolate(code, label.shape[-2:], mode='bilinear', align_corners=False)   # [B,70,H,W]: 587 MB/img
//     linear_probs  = torch.log_softmax(model.linear_probe(code), dim=1)                   # [B,27,H,W]
//     cluster_probs = model.cluster_probe(code, 2, log_probs=True)                          # [B,27,H,W]
// Here the upsampled code is never materialised.  Both probes are evaluated per output pixel from a few
// low-resolution quantities, using the linearity of bilinear interpolation:
//   * linear probe:  conv1x1(interp(code)) == interp(conv1x1(code))  -> interpolate the 27 low-res logits;
//   * cluster probe: <interp(code), c_k> == interp(<code, c_k>)      -> interpolate the low-res dot products with the
//     normalised centroids; ||interp(code)||^2 = sum_{t,t'} w_t w_t' <code_t, code_t'> needs only the Gram
//     entries between neighbouring low-res pixels (self, right, down, down-right, down-left).
// Two kernels: a per-low-res-pixel preparation (warp per pixel) and the per-output-pixel evaluation, which is
// bound by writing the two [B,n,H,W] fp32 log-probability maps (HBM): 453 MB per 1024x2048 image.
// Also fused here (src/eval_segmentation.py:124-126, 138-139; src/utils.py:219-229):
//   * flip test-time augmentation  code = (code(img) + code(img.flip(3)).flip(3)) / 2  — averaged on the low-res code
//     inside the preparation kernel (interpolation is linear, and the reference averages before it as well);
//   * UnsupervisedMetrics.update for both probes: the [pred][actual] confusion counts are accumulated per CTA in shared
//     memory from the argmax the kernel already has, one 64-bit atomic per non-zero cell per CTA.
#include "common.cuh"
#include "host_util.h"

namespace stego {

constexpr int EV_LD = 72;  // floats per low-res pixel: [0,32) linear logits, [32,64) centroid dots, 64.. Gram entries
constexpr int EV_SS = 64, EV_R = 65, EV_D = 66, EV_DR = 67, EV_DL = 68;

struct EvalPrepParams {
  const float* code;   // [B*h*w][ld] tokens-major
  const float* code_flip;  // code of the horizontally flipped image (same layout) or null: flip-TTA average
  long long ld;
  int B, h, w, C;
  const float* W;      // [n_lin][C]
  const float* bias;   // [n_lin]
  int n_lin;
  const float* clusters;  // [n_clu][C]
  int n_clu;
  float* lr;           // [B*h*w][EV_LD]
};

// one warp per low-res pixel; lanes = classes for the dot products, lanes = channels for the Gram entries
__global__ void __launch_bounds__(256)
eval_prep_kernel(EvalPrepParams p) {
  extern __shared__ float sm[];
  float* swT = sm;                 // [C][32] linear weights transposed
  float* scT = sm + p.C * 32;      // [C][32] normalised centroids transposed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.C * 32; i += blockDim.x) {
    const int c = i >> 5, k = i & 31;
    swT[i] = (k < p.n_lin) ? p.W[k * p.C + c] : 0.f;
    scT[i] = 0.f;
  }
  __syncthreads();
  for (int k = warp; k < p.n_clu; k += 8) {
    float ss = 0.f;
    for (int c = lane; c < p.C; c += 32) { const float v = p.clusters[k * p.C + c]; ss += v * v; }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < p.C; c += 32) scT[c * 32 + k] = p.clusters[k * p.C + c] * inv;
  }
  __syncthreads();
  const float bk = (lane < p.n_lin) ? p.bias[lane] : 0.f;
  const long long rows = 1ll * p.B * p.h * p.w;
  for (long long r = 1ll * blockIdx.x * 8 + warp; r < rows; r += 1ll * gridDim.x * 8) {
    const int x = static_cast<int>(r % p.w);
    const int y = static_cast<int>((r / p.w) % p.h);
    const float* cp = p.code + r * p.ld;
    // flipped image: its column w-1-x holds this pixel; moving right here is moving left there
    const float* fp = p.code_flip ? p.code_flip + (r - x + (p.w - 1 - x)) * p.ld : nullptr;
    float xr[3], nr[3], nd[3], ndr[3], ndl[3];
    const bool hr = x + 1 < p.w, hd = y + 1 < p.h, hl = x > 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int c = lane + 32 * k;
      const bool ok = c < p.C;
      xr[k] = ok ? cp[c] : 0.f;
      nr[k] = (ok && hr) ? cp[p.ld + c] : 0.f;
      nd[k] = (ok && hd) ? cp[p.w * p.ld + c] : 0.f;
      ndr[k] = (ok && hd && hr) ? cp[(p.w + 1) * p.ld + c] : 0.f;
      ndl[k] = (ok && hd && hl) ? cp[(p.w - 1) * p.ld + c] : 0.f;
      if (fp) {
        xr[k] = 0.5f * (xr[k] + (ok ? fp[c] : 0.f));
        nr[k] = 0.5f * (nr[k] + ((ok && hr) ? fp[c - p.ld] : 0.f));
        nd[k] = 0.5f * (nd[k] + ((ok && hd) ? fp[p.w * p.ld + c] : 0.f));
        ndr[k] = 0.5f * (ndr[k] + ((ok && hd && hr) ? fp[(p.w - 1) * p.ld + c] : 0.f));
        ndl[k] = 0.5f * (ndl[k] + ((ok && hd && hl) ? fp[(p.w + 1) * p.ld + c] : 0.f));
      }
    }
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      g0 = fmaf(xr[k], xr[k], g0);
      g1 = fmaf(xr[k], nr[k], g1);
      g2 = fmaf(xr[k], nd[k], g2);
      g3 = fmaf(xr[k], ndr[k], g3);
      g4 = fmaf(xr[k], ndl[k], g4);
    }
    g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2); g3 = warp_sum(g3); g4 = warp_sum(g4);
    float dl = bk, dc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const int c = 32 * k + j;
        if (c < p.C) {
          const float xc = __shfl_sync(0xffffffffu, xr[k], j);
          dl = fmaf(xc, swT[c * 32 + lane], dl);
          dc = fmaf(xc, scT[c * 32 + lane], dc);
        }
      }
    }
    float* o = p.lr + r * EV_LD;
    o[lane] = dl;
    o[32 + lane] = dc;
    if (lane == 0) { o[EV_SS] = g0; o[EV_R] = g1; o[EV_D] = g2; o[EV_DR] = g3; o[EV_DL] = g4; }
  }
}

struct EvalProbeParams {
  const float* lr;   // [B*h*w][EV_LD]
  int B, h, w, H, W, n_lin, n_clu;
  float alpha;
  float* lin_logp;   // [B][n_lin][H][W] or null
  float* clu_logp;   // [B][n_clu][H][W] or null
  unsigned char* lin_arg;  // [B][H][W] or null
  unsigned char* clu_arg;  // [B][H][W] or null
  int box_h, box_w;
  const void* label;       // [B][H][W] int64 / int32 / uint8 (label_bytes 8 / 4 / 1) or null
  int label_bytes;
  int n_cls;               // label classes: a pixel counts when 0 <= label < n_cls and pred < n_cls (utils.py:222)
  unsigned long long* lin_conf;  // [n_lin][n_cls] += counts of (pred, actual), or null
  unsigned long long* clu_conf;  // [n_clu][n_cls]
};

__device__ __forceinline__ void ev_src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (dst + 0.5f) - 0.5f;  // ATen area_pixel_compute_source_index, align_corners=False
  if (s < 0.f) s = 0.f;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - i0;
}

// Output tile of a CTA: 64 x 4 pixels, thread = pixel, x fastest (coalesced plane writes).  (Measured and rejected in
// round 2: 64 x 16 tiles in four passes under a 120-register cap, two CTAs per SM: 1.35 ms instead of 1.00 ms per 4 frames.)
constexpr int EVT_W = 64, EVT_ROWS = 4, EVT_PASSES = 1, EVT_H = EVT_ROWS * EVT_PASSES;

// Gram entry <code_p, code_q> for box-relative low-res pixels p, q that are equal or 8-neighbours
__device__ __forceinline__ float ev_gram(const float* slr, int bw, int py, int px, int qy, int qx) {
  int dy = qy - py, dx = qx - px;
  if (dy < 0 || (dy == 0 && dx < 0)) {  // look the pair up from the upper / left pixel
    const int ty = py, tx = px;
    py = qy; px = qx; qy = ty; qx = tx;
    dy = -dy; dx = -dx;
  }
  const float* e = slr + (py * bw + px) * EV_LD;
  if (dy == 0) return dx == 0 ? e[EV_SS] : e[EV_R];
  return dx == 0 ? e[EV_D] : (dx > 0 ? e[EV_DR] : e[EV_DL]);
}

__global__ void __launch_bounds__(EVT_W* EVT_ROWS)
eval_probe_kernel(EvalProbeParams p) {
  extern __shared__ float slr[];  // [box_h*box_w][EV_LD]
  __shared__ unsigned int hist[2][32 * 32];  // [probe][pred * 32 + actual]
  const bool want_conf = p.label != nullptr;
  if (want_conf)
    for (int i = threadIdx.x; i < 2 * 32 * 32; i += blockDim.x) (&hist[0][0])[i] = 0u;
  const int tiles_x = (p.W + EVT_W - 1) / EVT_W, tiles_y = (p.H + EVT_H - 1) / EVT_H;
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int X0 = tx * EVT_W, Y0 = ty * EVT_H;
  const int Xl = min(X0 + EVT_W - 1, p.W - 1), Yl = min(Y0 + EVT_H - 1, p.H - 1);
  const float sy = static_cast<float>(p.h) / p.H, sx = static_cast<float>(p.w) / p.W;
  int by0, by1, bx0, bx1, tmp;
  float ftmp;
  ev_src_index(Y0, sy, p.h, by0, tmp, ftmp);
  ev_src_index(Yl, sy, p.h, tmp, by1, ftmp);
  ev_src_index(X0, sx, p.w, bx0, tmp, ftmp);
  ev_src_index(Xl, sx, p.w, tmp, bx1, ftmp);
  const int bh = by1 - by0 + 1, bw = bx1 - bx0 + 1;
  const long long base = 1ll * b * p.h * p.w;
  for (int i = threadIdx.x; i < bh * bw * EV_LD; i += blockDim.x) {
    const int cell = i / EV_LD, k = i % EV_LD;
    const int r = cell / bw, c = cell % bw;
    slr[i] = p.lr[(base + 1ll * (by0 + r) * p.w + bx0 + c) * EV_LD + k];
  }
  __syncthreads();
  for (int pass = 0; pass < EVT_PASSES; ++pass) {
  const int X = X0 + (threadIdx.x % EVT_W), Y = Y0 + pass * EVT_ROWS + (threadIdx.x / EVT_W);
  const bool active = X < p.W && Y < p.H;
  int lin_pred = -1, clu_pred = -1;
  if (active) {
  int y0, y1, x0, x1;
  float ly, lx;
  ev_src_index(Y, sy, p.h, y0, y1, ly);
  ev_src_index(X, sx, p.w, x0, x1, lx);
  y0 -= by0; y1 -= by0; x0 -= bx0; x1 -= bx0;
  const float wa = (1.f - ly) * (1.f - lx), wb = (1.f - ly) * lx, wc = ly * (1.f - lx), wd = ly * lx;
  const float* ea = slr + (y0 * bw + x0) * EV_LD;
  const float* eb = slr + (y0 * bw + x1) * EV_LD;
  const float* ec = slr + (y1 * bw + x0) * EV_LD;
  const float* ed = slr + (y1 * bw + x1) * EV_LD;
  const long long plane = 1ll * p.H * p.W;
  const long long pix = 1ll * Y * p.W + X;
  // ---- linear probe: log_softmax of the interpolated logits
  if (p.lin_logp || p.lin_arg) {
    float z[32];
    float mx = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < p.n_lin) {
        z[k] = wa * ea[k] + wb * eb[k] + wc * ec[k] + wd * ed[k];
        if (z[k] > mx) { mx = z[k]; arg = k; }
      }
    }
    lin_pred = arg;
    if (p.lin_arg) p.lin_arg[b * plane + pix] = static_cast<unsigned char>(arg);
    if (p.lin_logp) {
      float se = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < p.n_lin) se += __expf(z[k] - mx);
      const float lse = mx + __logf(se);
      float* o = p.lin_logp + (1ll * b * p.n_lin) * plane + pix;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < p.n_lin) o[k * plane] = z[k] - lse;
    }
  }
  // ---- cluster probe: cosine similarity of the interpolated code with the centroids, log_softmax(alpha * .)
  if (p.clu_logp || p.clu_arg) {
    float n2 = wa * wa * ea[EV_SS] + wb * wb * eb[EV_SS] + wc * wc * ec[EV_SS] + wd * wd * ed[EV_SS];
    n2 += 2.f * (wa * wb * ev_gram(slr, bw, y0, x0, y0, x1) + wa * wc * ev_gram(slr, bw, y0, x0, y1, x0) +
                 wa * wd * ev_gram(slr, bw, y0, x0, y1, x1) + wb * wc * ev_gram(slr, bw, y0, x1, y1, x0) +
                 wb * wd * ev_gram(slr, bw, y0, x1, y1, x1) + wc * wd * ev_gram(slr, bw, y1, x0, y1, x1));
    const float inv = 1.0f / fmaxf(sqrtf(fmaxf(n2, 0.f)), 1e-12f);
    float z[32];
    float mx = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < p.n_clu) {
        z[k] = (wa * ea[32 + k] + wb * eb[32 + k] + wc * ec[32 + k] + wd * ed[32 + k]) * inv;
        if (z[k] > mx) { mx = z[k]; arg = k; }
      }
    }
    clu_pred = arg;
    if (p.clu_arg) p.clu_arg[b * plane + pix] = static_cast<unsigned char>(arg);
    if (p.clu_logp) {
      float m2 = -INFINITY;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < p.n_clu) m2 = fmaxf(m2, z[k] * p.alpha);
      float se = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < p.n_clu) se += __expf(z[k] * p.alpha - m2);
      const float lse = m2 + __logf(se);
      float* o = p.clu_logp + (1ll * b * p.n_clu) * plane + pix;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < p.n_clu) o[k * plane] = z[k] * p.alpha - lse;
    }
  }
  }  // active
  if (want_conf) {
    // UnsupervisedMetrics.update (src/utils.py:219-229): stats[pred][actual] += 1 over pixels with a valid label
    if (active) {
      const long long li = 1ll * b * p.H * p.W + 1ll * Y * p.W + X;
      long long lab;
      if (p.label_bytes == 8) lab = reinterpret_cast<const long long*>(p.label)[li];
      else if (p.label_bytes == 4) lab = reinterpret_cast<const int*>(p.label)[li];
      else lab = reinterpret_cast<const unsigned char*>(p.label)[li];
      if (lab >= 0 && lab < p.n_cls) {
        if (lin_pred >= 0 && lin_pred < p.n_cls) atomicAdd(&hist[0][lin_pred * 32 + static_cast<int>(lab)], 1u);
        if (clu_pred >= 0 && clu_pred < p.n_cls) atomicAdd(&hist[1][clu_pred * 32 + static_cast<int>(lab)], 1u);
      }
    }
  }
  }  // pass
  if (want_conf) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 32 * 32; i += blockDim.x) {
      const unsigned int cnt = (&hist[0][0])[i];
      if (cnt == 0u) continue;
      const int probe = i >> 10, pred = (i >> 5) & 31, act = i & 31;
      unsigned long long* dst = probe ? p.clu_conf : p.lin_conf;
      if (dst && act < p.n_cls && pred < (probe ? p.n_clu : p.n_lin)) atomicAdd(dst + pred * p.n_cls + act, (unsigned long long)cnt);
    }
  }
}


// ---- four pixels per thread (round 2) --------------------------------------------------------------------------------
// The pixel-per-thread kernel above is issue-bound, not HBM-bound (ncu: 70 % issue-active, 2900 warp instructions per
// 32 pixels, 20 % of the DRAM throughput): every class costs four scalar shared-memory loads, four FMAs and a strided
// 4-byte store per pixel.  When the horizontal upsampling factor is a multiple of 8, four x-adjacent output pixels
// (X = 4t .. 4t+3) always interpolate between the same two low-res columns, so a thread that owns all four
//   * loads the four corner rows once, as float4 (LDS.128, same address across most of the warp = broadcast),
//   * interpolates vertically once:  L[k] = a + ly (c - a),  R[k] = b + ly (d - b),  D[k] = R[k] - L[k],
//   * evaluates each pixel as z_j[k] = L[k] + lx_j D[k]  (one FMA per class), recomputing it in the three passes
//     (max/argmax, sum of exponentials, output) instead of keeping 4 x 27 logits in registers,
//   * writes one float4 per class plane (a warp covers 512 contiguous bytes of a plane row).
// Class counts are template parameters (27/27 = both shipped label sets); any other shape takes the generic kernel.
constexpr int EV4_TW = 64, EV4_ROWS = 4, EV4_W = 4 * EV4_TW;  // CTA tile: 256 x 4 pixels, 256 threads

template <int N, bool SCALED>
__device__ __forceinline__ void ev4_probe(const float* ea, const float* eb, const float* ec, const float* ed, float ly,
                                          const float (&lx)[4], const float (&scale)[4], float* out, long long plane,
                                          int (&pred)[4]) {
  constexpr int Q = (N + 3) / 4;
  float L[4 * Q], D[4 * Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const float4 a = reinterpret_cast<const float4*>(ea)[q], b = reinterpret_cast<const float4*>(eb)[q];
    const float4 c = reinterpret_cast<const float4*>(ec)[q], d = reinterpret_cast<const float4*>(ed)[q];
    const float l0 = fmaf(ly, c.x - a.x, a.x), l1 = fmaf(ly, c.y - a.y, a.y), l2 = fmaf(ly, c.z - a.z, a.z),
                l3 = fmaf(ly, c.w - a.w, a.w);
    L[4 * q] = l0; L[4 * q + 1] = l1; L[4 * q + 2] = l2; L[4 * q + 3] = l3;
    D[4 * q] = fmaf(ly, d.x - b.x, b.x) - l0;
    D[4 * q + 1] = fmaf(ly, d.y - b.y, b.y) - l1;
    D[4 * q + 2] = fmaf(ly, d.z - b.z, b.z) - l2;
    D[4 * q + 3] = fmaf(ly, d.w - b.w, b.w) - l3;
  }
  float lse[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float mx = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float z = fmaf(lx[j], D[k], L[k]);
      if (SCALED) z *= scale[j];
      if (z > mx) { mx = z; arg = k; }
    }
    pred[j] = arg;
    float se = 0.f;
    if (out) {
      const float nmx = -mx * 1.4426950408889634f;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        float z = fmaf(lx[j], D[k], L[k]);
        if (SCALED) z *= scale[j];
        se += ex2_approx(fmaf(z, 1.4426950408889634f, nmx));
      }
    }
    lse[j] = mx + __logf(se);
  }
  if (out) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float4 o;
      o.x = fmaf(lx[0], D[k], L[k]); o.y = fmaf(lx[1], D[k], L[k]); o.z = fmaf(lx[2], D[k], L[k]); o.w = fmaf(lx[3], D[k], L[k]);
      if (SCALED) { o.x *= scale[0]; o.y *= scale[1]; o.z *= scale[2]; o.w *= scale[3]; }
      o.x -= lse[0]; o.y -= lse[1]; o.z -= lse[2]; o.w -= lse[3];
      *reinterpret_cast<float4*>(out + k * plane) = o;
    }
  }
}

template <int NL, int NC>
__global__ void __launch_bounds__(EV4_TW* EV4_ROWS, 2)
eval_probe_vec4_kernel(EvalProbeParams p) {
  extern __shared__ __align__(16) float slr[];  // [box_h*box_w][EV_LD]
  __shared__ unsigned int hist[2][32 * 32];     // [probe][pred * 32 + actual]
  const bool want_conf = p.label != nullptr;
  if (want_conf)
    for (int i = threadIdx.x; i < 2 * 32 * 32; i += blockDim.x) (&hist[0][0])[i] = 0u;
  const int tiles_x = (p.W + EV4_W - 1) / EV4_W, tiles_y = (p.H + EV4_ROWS - 1) / EV4_ROWS;
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int X0 = tx * EV4_W, Y0 = ty * EV4_ROWS;
  const int Xl = min(X0 + EV4_W - 1, p.W - 1), Yl = min(Y0 + EV4_ROWS - 1, p.H - 1);
  const float sy = static_cast<float>(p.h) / p.H, sx = static_cast<float>(p.w) / p.W;
  int by0, by1, bx0, bx1, tmp;
  float ftmp;
  ev_src_index(Y0, sy, p.h, by0, tmp, ftmp);
  ev_src_index(Yl, sy, p.h, tmp, by1, ftmp);
  ev_src_index(X0, sx, p.w, bx0, tmp, ftmp);
  ev_src_index(Xl, sx, p.w, tmp, bx1, ftmp);
  const int bh = by1 - by0 + 1, bw = bx1 - bx0 + 1;
  const long long base = 1ll * b * p.h * p.w;
  {
    constexpr int LD4 = EV_LD / 4;
    float4* s4 = reinterpret_cast<float4*>(slr);
    const float4* g4 = reinterpret_cast<const float4*>(p.lr);
    for (int i = threadIdx.x; i < bh * bw * LD4; i += blockDim.x) {
      const int cell = i / LD4, k = i % LD4;
      const int r = cell / bw, c = cell % bw;
      s4[i] = g4[(base + 1ll * (by0 + r) * p.w + bx0 + c) * LD4 + k];
    }
  }
  __syncthreads();
  const int X = X0 + 4 * (threadIdx.x % EV4_TW), Y = Y0 + threadIdx.x / EV4_TW;
  const bool active = X < p.W && Y < p.H;  // W % 4 == 0: a group is entirely inside or outside
  if (active) {
    int y0, y1, x0, x1;
    float ly, lx[4];
    ev_src_index(Y, sy, p.h, y0, y1, ly);
    ev_src_index(X, sx, p.w, x0, x1, lx[0]);
#pragma unroll
    for (int j = 1; j < 4; ++j) {  // same two columns for the whole group (host checks the upsampling factor)
      int t0, t1;
      ev_src_index(X + j, sx, p.w, t0, t1, lx[j]);
    }
    y0 -= by0; y1 -= by0; x0 -= bx0; x1 -= bx0;
    const float* ea = slr + (y0 * bw + x0) * EV_LD;
    const float* eb = slr + (y0 * bw + x1) * EV_LD;
    const float* ec = slr + (y1 * bw + x0) * EV_LD;
    const float* ed = slr + (y1 * bw + x1) * EV_LD;
    const long long plane = 1ll * p.H * p.W;
    const long long pix = 1ll * Y * p.W + X;
    int lin_pred[4] = {-1, -1, -1, -1}, clu_pred[4] = {-1, -1, -1, -1};
    float one[4] = {1.f, 1.f, 1.f, 1.f};
    if (p.lin_logp || p.lin_arg) {
      ev4_probe<NL, false>(ea, eb, ec, ed, ly, lx, one, p.lin_logp ? p.lin_logp + (1ll * b * NL) * plane + pix : nullptr,
                           plane, lin_pred);
      if (p.lin_arg)
        *reinterpret_cast<uchar4*>(p.lin_arg + b * plane + pix) =
            make_uchar4((unsigned char)lin_pred[0], (unsigned char)lin_pred[1], (unsigned char)lin_pred[2], (unsigned char)lin_pred[3]);
    }
    if (p.clu_logp || p.clu_arg) {
      const float gaa = ea[EV_SS], gbb = eb[EV_SS], gcc = ec[EV_SS], gdd = ed[EV_SS];
      const float gab = ev_gram(slr, bw, y0, x0, y0, x1), gac = ev_gram(slr, bw, y0, x0, y1, x0);
      const float gad = ev_gram(slr, bw, y0, x0, y1, x1), gbc = ev_gram(slr, bw, y0, x1, y1, x0);
      const float gbd = ev_gram(slr, bw, y0, x1, y1, x1), gcd = ev_gram(slr, bw, y1, x0, y1, x1);
      float scale[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float wa = (1.f - ly) * (1.f - lx[j]), wb = (1.f - ly) * lx[j], wc = ly * (1.f - lx[j]), wd = ly * lx[j];
        float n2 = wa * wa * gaa + wb * wb * gbb + wc * wc * gcc + wd * wd * gdd;
        n2 += 2.f * (wa * wb * gab + wa * wc * gac + wa * wd * gad + wb * wc * gbc + wb * wd * gbd + wc * wd * gcd);
        scale[j] = p.alpha / fmaxf(sqrtf(fmaxf(n2, 0.f)), 1e-12f);
      }
      ev4_probe<NC, true>(ea + 32, eb + 32, ec + 32, ed + 32, ly, lx, scale,
                          p.clu_logp ? p.clu_logp + (1ll * b * NC) * plane + pix : nullptr, plane, clu_pred);
      if (p.clu_arg)
        *reinterpret_cast<uchar4*>(p.clu_arg + b * plane + pix) =
            make_uchar4((unsigned char)clu_pred[0], (unsigned char)clu_pred[1], (unsigned char)clu_pred[2], (unsigned char)clu_pred[3]);
    }
    if (want_conf) {  // UnsupervisedMetrics.update (src/utils.py:219-229)
      const long long li = 1ll * b * plane + pix;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        long long lab;
        if (p.label_bytes == 8) lab = reinterpret_cast<const long long*>(p.label)[li + j];
        else if (p.label_bytes == 4) lab = reinterpret_cast<const int*>(p.label)[li + j];
        else lab = reinterpret_cast<const unsigned char*>(p.label)[li + j];
        if (lab >= 0 && lab < p.n_cls) {
          if (lin_pred[j] >= 0 && lin_pred[j] < p.n_cls) atomicAdd(&hist[0][lin_pred[j] * 32 + static_cast<int>(lab)], 1u);
          if (clu_pred[j] >= 0 && clu_pred[j] < p.n_cls) atomicAdd(&hist[1][clu_pred[j] * 32 + static_cast<int>(lab)], 1u);
        }
      }
    }
  }
  if (want_conf) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 32 * 32; i += blockDim.x) {
      const unsigned int cnt = (&hist[0][0])[i];
      if (cnt == 0u) continue;
      const int probe = i >> 10, pred = (i >> 5) & 31, act = i & 31;
      unsigned long long* dst = probe ? p.clu_conf : p.lin_conf;
      if (dst && act < p.n_cls && pred < (probe ? NC : NL)) atomicAdd(dst + pred * p.n_cls + act, (unsigned long long)cnt);
    }
  }
}

}  // namespace stego

using namespace stego;

// code: tokens-major low-res code [B*h*w][ld_code] fp32 (what DinoFeaturizer produces); outputs at [H][W].
// code_flip: the code of the horizontally flipped images (flip-TTA) or null.  lr_scratch: [B*h*w][72] floats.
// label + confusion outputs (int64, accumulated): optional.  Any output pointer may be null.
extern "C" int stego_eval_probes(const float* code, const float* code_flip, long long ld_code, int C, int B, int h, int w,
                                 int H, int W, const float* lin_weight, const float* lin_bias, int n_lin,
                                 const float* clusters, int n_clu, float alpha, float* lr_scratch, float* lin_log_probs,
                                 float* clu_log_probs, unsigned char* lin_argmax, unsigned char* clu_argmax,
                                 const void* label, int label_bytes, int n_label_classes, long long* lin_confusion,
                                 long long* clu_confusion, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  STEGO_CHECK_ARG(code && lin_weight && lin_bias && clusters && lr_scratch, "stego_eval_probes: null pointer");
  STEGO_CHECK_ARG(C > 0 && C <= 96 && n_lin > 0 && n_lin <= 32 && n_clu > 0 && n_clu <= 32,
                  "stego_eval_probes: C=%d n_lin=%d n_clu=%d unsupported (C <= 96, classes <= 32)", C, n_lin, n_clu);
  STEGO_CHECK_ARG(B > 0 && h > 0 && w > 0 && H >= h && W >= w, "stego_eval_probes: bad sizes (upsampling only)");
  STEGO_CHECK_ARG(!label || ((label_bytes == 8 || label_bytes == 4 || label_bytes == 1) && n_label_classes > 0 &&
                             n_label_classes <= 32 && (lin_confusion || clu_confusion)),
                  "stego_eval_probes: confusion counts need label_bytes in {8,4,1}, n_label_classes <= 32 and an output");
  EvalPrepParams q;
  q.code = code; q.code_flip = code_flip; q.ld = ld_code; q.B = B; q.h = h; q.w = w; q.C = C; q.W = lin_weight; q.bias = lin_bias;
  q.n_lin = n_lin; q.clusters = clusters; q.n_clu = n_clu; q.lr = lr_scratch;
  const long long rows = 1ll * B * h * w;
  long long g = (rows + 7) / 8;
  const long long cap = 16ll * num_sms();
  eval_prep_kernel<<<(unsigned)(g < cap ? g : cap), 256, (size_t)C * 64 * sizeof(float), stream>>>(q);
  STEGO_CHECK_LAUNCH("eval_prep_kernel");
  EvalProbeParams p;
  p.lr = lr_scratch; p.B = B; p.h = h; p.w = w; p.H = H; p.W = W; p.n_lin = n_lin; p.n_clu = n_clu; p.alpha = alpha;
  p.lin_logp = lin_log_probs; p.clu_logp = clu_log_probs; p.lin_arg = lin_argmax; p.clu_arg = clu_argmax;
  p.label = label; p.label_bytes = label_bytes; p.n_cls = n_label_classes;
  p.lin_conf = reinterpret_cast<unsigned long long*>(lin_confusion);
  p.clu_conf = reinterpret_cast<unsigned long long*>(clu_confusion);
  // four pixels per thread when a group of four x-adjacent pixels always shares its two source columns: integer
  // horizontal factor that is a multiple of 8 (every patch-8 / patch-16 model evaluated at the label resolution)
  const bool vec4 = n_lin == 27 && n_clu == 27 && W % w == 0 && (W / w) % 8 == 0 &&
                    (!lin_log_probs || (reinterpret_cast<uintptr_t>(lin_log_probs) & 15) == 0) &&
                    (!clu_log_probs || (reinterpret_cast<uintptr_t>(clu_log_probs) & 15) == 0) &&
                    (!lin_argmax || (reinterpret_cast<uintptr_t>(lin_argmax) & 3) == 0) &&
                    (!clu_argmax || (reinterpret_cast<uintptr_t>(clu_argmax) & 3) == 0) &&
                    (reinterpret_cast<uintptr_t>(lr_scratch) & 15) == 0;
  const int tile_h = vec4 ? EV4_ROWS : EVT_H, tile_w = vec4 ? EV4_W : EVT_W;
  p.box_h = (int)((double)tile_h * h / H) + 3;
  p.box_w = (int)((double)tile_w * w / W) + 3;
  if (p.box_h > h) p.box_h = h;
  if (p.box_w > w) p.box_w = w;
  const size_t smem = (size_t)p.box_h * p.box_w * EV_LD * sizeof(float);
  STEGO_CHECK_ARG(smem <= 200 * 1024, "stego_eval_probes: upsample ratio needs %zu B of shared memory", smem);
  const long long tiles = 1ll * B * ((H + tile_h - 1) / tile_h) * ((W + tile_w - 1) / tile_w);
  STEGO_CHECK_ARG(tiles < (1ll << 31), "stego_eval_probes: too many tiles");
  if (vec4) {
    static size_t conf4 = 0;
    if (smem > 48 * 1024 && smem > conf4) {
      cudaError_t e = cudaFuncSetAttribute(eval_probe_vec4_kernel<27, 27>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(eval_probe_vec4)");
      conf4 = smem;
    }
    eval_probe_vec4_kernel<27, 27><<<(unsigned)tiles, EV4_TW * EV4_ROWS, smem, stream>>>(p);
    STEGO_CHECK_LAUNCH("eval_probe_vec4_kernel");
    return STEGO_OK;
  }
  static size_t conf = 0;
  if (smem > 48 * 1024 && smem > conf) {
    cudaError_t e = cudaFuncSetAttribute(eval_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(eval_probe)");
    conf = smem;
  }
  eval_probe_kernel<<<(unsigned)tiles, EVT_W * EVT_ROWS, smem, stream>>>(p);
  STEGO_CHECK_LAUNCH("eval_probe_kernel");
  return STEGO_OK;
}
