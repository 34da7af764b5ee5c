/*
 * stego_b200 — C-ABI of the B200-native STEGO correspondence-distillation hot path.
 *
 * This header is the drop-in boundary.  The reference (mhamilton723/STEGO) has no FFI layer: its
 * hot path is Python over torch ops in src/modules.py / src/dino/vision_transformer.py /
 * src/train_segmentation.py.  Each entry point below replaces the torch-op sequence cited next to
 * it (reference file:line) with hand-written sm_100a kernels; stego_b200/modules.py binds them with
 * ctypes behind the reference's own class / function names.
 *
 * Conventions
 *   - every pointer is a borrowed DEVICE pointer into caller-owned (PyTorch-owned) memory; the
 *     library allocates nothing and keeps no pointer after the call returns;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, -1 = bad argument, -2 = unsupported, -3 = CUDA error;
 *     stego_last_error() returns a thread-local message for the last non-zero status;
 *   - bf16 tensors are raw __nv_bfloat16 (uint16) storage; "tokens-major" means [.., HW, C] with the
 *     channel dimension contiguous (PyTorch channels_last for an NCHW view).
 */
#ifndef STEGO_B200_H_
#define STEGO_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define STEGO_API

/* ------------------------------------------------------------------------------------------------
 * Library
 * ---------------------------------------------------------------------------------------------- */
STEGO_API int stego_version(void);
STEGO_API const char* stego_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction (tcgen05 + TMA + TMEM):
 *     out[M,N] = act(A . B^T + bias[N]) + residual
 * replaces nn.Linear / 1x1 Conv2d calls of the path:
 *   src/dino/vision_transformer.py:80,88 (qkv, proj), :58-62 (fc1+GELU, fc2), :127-131 (patch embed),
 *   src/modules.py:73-81 (cluster1 / cluster2 heads) and their autograd backward (dgrad, wgrad).
 *   a_mn_major = 0: A is [M][lda] (K contiguous);   1: A is stored transposed, [K][lda] (M contiguous)
 *   b_mn_major = 0: B is [N][ldb] (K contiguous);   1: B is [K][ldb] (N contiguous)
 *   K must be a multiple of 64 (zero-pad operands); lda/ldb multiples of 8.
 *   act: 0 none, 1 GELU(erf) (nn.GELU default), 2 ReLU.
 *   residual: fp32 [M][ldr] added after the activation (may alias out for an in-place update).
 *   row_div > 0 (patch-embed mode): output row r goes to r + r/row_div + 1 (skips the cls slot of
 *     each image) and the residual row is r % row_div + 1 (positional embedding broadcast).
 *   splits > 1 requires atomic_out = 1: split-K partial sums are atomically added into fp32 `out`.
 * ---------------------------------------------------------------------------------------------- */
STEGO_API int stego_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major,
                              int M, int N, int K, void* out, int ldo, int out_bf16, const float* bias, int act,
                              const float* residual, int ldr, int row_div, int splits, int atomic_out,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STEGO_B200_H_ */
