"""Thin Python wrappers over the C-ABI (one function per extern "C" entry point).

These take torch CUDA tensors, pass raw device pointers + sizes + the current stream through
ctypes and return torch tensors.  They allocate outputs with torch (PyTorch owns all memory) and
never fall back to torch math.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
         a_mn: bool = False, b_mn: bool = False, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         residual: Optional[torch.Tensor] = None, row_div: int = 0, splits: int = 1,
         atomic: bool = False) -> torch.Tensor:
    """out[M,N] = act(A.B^T + bias) + residual on tcgen05 (see include/stego_b200.h: stego_gemm_bf16)."""
    _lib.require_cuda(a, b, out, bias, residual)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    assert out.dtype in (torch.bfloat16, torch.float32)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
    rc = _lib.load().stego_gemm_bf16(
        _lib.ptr(a), a.stride(0), int(a_mn), _lib.ptr(b), b.stride(0), int(b_mn), M, N, K,
        _lib.ptr(out), out.stride(0), int(out.dtype == torch.bfloat16), _lib.ptr(bias), act,
        _lib.ptr(residual), residual.stride(0) if residual is not None else 0, row_div, splits, int(atomic),
        _lib.stream())
    _lib.check(rc, "stego_gemm_bf16")
    return out
