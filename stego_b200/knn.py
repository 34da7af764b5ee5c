"""Fused kNN over image descriptors: the device-side replacement of the similarity + top-k
loop of the reference's `src/precompute_knns.py:83-96`, which produces the `nns_*.npz` files `img_pos` is drawn from."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def knn_topk(feats: torch.Tensor, k: int = 30, return_values: bool = False
             ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """feats: [n, E] fp32 CUDA (un-normalised descriptors, e.g. `model(img).mean([2, 3])`, precompute_knns.py:19).
    Returns int64 [n, k] neighbour indices by descending cosine similarity (each row contains itself, like the
    reference's `torch.topk(einsum("nf,mf->nm", ...), 30)[1]`), and the similarities if requested."""
    _lib.require_cuda(feats)
    if feats.dim() != 2 or feats.dtype != torch.float32:
        raise RuntimeError("stego_b200.knn_topk: feats must be a 2-D fp32 tensor")
    feats = feats.contiguous()
    n, E = feats.shape
    planes = torch.empty(2, n, E, dtype=torch.bfloat16, device=feats.device)
    idx = torch.empty(n, k, dtype=torch.long, device=feats.device)
    vals = torch.empty(n, k, dtype=torch.float32, device=feats.device) if return_values else None
    rc = _lib.load().stego_knn_topk(_lib.ptr(feats), n, E, k, _lib.ptr(planes), _lib.ptr(idx), _lib.ptr(vals), _lib.stream())
    _lib.check(rc, "stego_knn_topk")
    return idx, vals
