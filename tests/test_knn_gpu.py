"""Fused kNN kernel (stego_knn_topk) against the oracle restatement of
src/precompute_knns.py:83-92.  Similarities are fp32 sums of 384 products: the bf16 hi/lo tensor-core path differs from
the fp32 einsum by ~1e-6, so rankings are compared exactly where the oracle's neighbouring similarities are separated by
more than that, and through the similarity values everywhere."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,E,k", [(1000, 384, 30), (257, 768, 30), (40, 64, 5), (3000, 384, 30), (31, 64, 30)])
def test_knn_topk_matches_oracle(cuda_dev, n, E, k):
    import stego_oracle as O
    from stego_b200.knn import knn_topk
    g = torch.Generator().manual_seed(n + E)
    base = torch.randn(max(8, n // 20), E, generator=g)                       # clustered descriptors: realistic ties
    feats = base[torch.randint(0, base.shape[0], (n,), generator=g)] + 0.35 * torch.randn(n, E, generator=g)
    feats = feats * (0.5 + torch.rand(n, 1, generator=g))                    # un-normalised input
    normed = torch.nn.functional.normalize(feats, dim=1)
    want_idx, want_val = O.knn_indices(normed.double(), k + 1)               # fp64 reference similarities (+ the runner-up)
    next_gap = (want_val[:, k - 1] - want_val[:, k]).abs()                   # separation of the k-th from the (k+1)-th
    want_idx, want_val = want_idx[:, :k], want_val[:, :k]
    idx, val = knn_topk(feats.to(cuda_dev), k, return_values=True)
    idx, val = idx.cpu(), val.cpu().double()
    assert idx.shape == (n, k) and idx.dtype == torch.long
    assert (idx[:, 0] == torch.arange(n)).all()                              # a row is its own nearest neighbour
    assert (val[:, :-1] >= val[:, 1:]).all()                                 # sorted by descending similarity
    assert (val - want_val).abs().max().item() < 2e-5                        # the k best similarities, in order
    true_sims = (normed.double() @ normed.double().t()).gather(1, idx)       # returned indices really have those sims
    assert (true_sims - val).abs().max().item() < 2e-5
    gap = (want_val[:, :-1] - want_val[:, 1:]).abs()
    clear = torch.cat([gap > 1e-4, (next_gap > 1e-4).unsqueeze(1)], 1) & \
        torch.cat([torch.ones(n, 1, dtype=torch.bool), gap > 1e-4], 1)       # positions separated from both neighbours
    assert (idx[clear] == want_idx[clear]).all()


def test_knn_rejects_bad_shapes(cuda_dev):
    from stego_b200.knn import knn_topk
    with pytest.raises(RuntimeError, match="multiple of 64"):
        knn_topk(torch.zeros(100, 100, device=cuda_dev), 5)
    with pytest.raises(RuntimeError, match="k="):
        knn_topk(torch.zeros(100, 64, device=cuda_dev), 40)
