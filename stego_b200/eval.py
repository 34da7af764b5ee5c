"""Fused evaluation probes: the inner loop of the reference's `eval_segmentation.py` (:124-141) without
materialising the upsampled code.

    code = F.interpolate(code, label.shape[-2:], mode='bilinear', align_corners=False)
    linear_probs  = torch.log_softmax(model.linear_probe(code), dim=1)
    cluster_probs = model.cluster_probe(code, 2, log_probs=True)

becomes `fused_probe_log_probs(code_lowres, model.linear_probe, model.cluster_probe, label.shape[-2:], 2)`.
At 1024x2048 the reference moves 587 MB (fp32 upsampled code) per image per probe before it even starts;
the fused kernel reads the 9 MB low-res code and writes only the outputs (stego_eval_probes, eval_probes.cu).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def fused_probe_log_probs(code: torch.Tensor, linear_probe: torch.nn.Module, cluster_probe: torch.nn.Module,
                          size: Sequence[int], alpha: float = 2.0, want_log_probs: bool = True,
                          want_argmax: bool = False):
    """code: low-res [B, C, h, w] (any strides, CUDA).  Returns (linear_log_probs, cluster_log_probs) [B,n,H,W]
    fp32, and with want_argmax also (linear_argmax, cluster_argmax) uint8 [B,H,W]."""
    _lib.require_cuda(code)
    if not code.is_cuda:
        raise RuntimeError("stego_b200.eval: CUDA tensors required (no CPU fallback)")
    B, C, h, w = code.shape
    H, W = int(size[0]), int(size[1])
    x = code.detach()
    if x.dtype != torch.float32 or x.stride(1) != 1 or x.stride(2) != w * x.stride(3):
        x = x.float().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)  # tokens-major
    ld = x.stride(3)
    wl = linear_probe.weight.detach().float().reshape(linear_probe.weight.shape[0], C).contiguous()
    bl = linear_probe.bias.detach().float().contiguous()
    cl = cluster_probe.clusters.detach().float().contiguous()
    n_lin, n_clu = wl.shape[0], cl.shape[0]
    dev = code.device
    scratch = torch.empty(B * h * w, 72, dtype=torch.float32, device=dev)
    lin = torch.empty(B, n_lin, H, W, dtype=torch.float32, device=dev) if want_log_probs else None
    clu = torch.empty(B, n_clu, H, W, dtype=torch.float32, device=dev) if want_log_probs else None
    la = torch.empty(B, H, W, dtype=torch.uint8, device=dev) if want_argmax else None
    ca = torch.empty(B, H, W, dtype=torch.uint8, device=dev) if want_argmax else None
    rc = _lib.load().stego_eval_probes(_lib.ptr(x), ld, C, B, h, w, H, W, _lib.ptr(wl), _lib.ptr(bl), n_lin, _lib.ptr(cl),
                                       n_clu, float(alpha), _lib.ptr(scratch), _lib.ptr(lin), _lib.ptr(clu), _lib.ptr(la),
                                       _lib.ptr(ca), _lib.stream())
    _lib.check(rc, "stego_eval_probes")
    if want_argmax:
        return lin, clu, la, ca
    return lin, clu
