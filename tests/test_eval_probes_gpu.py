"""Fused eval probes (upsample + linear probe + ClusterLookup log-probs) vs the reference op sequence
(src/eval_segmentation.py:128-131) evaluated by the oracle on the CPU."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))


@pytest.mark.parametrize("B,h,w,H,W", [(2, 10, 10, 80, 80), (1, 7, 9, 50, 61), (1, 16, 32, 128, 256), (2, 6, 6, 6, 6)])
def test_fused_eval_probes_match_reference_sequence(cuda_dev, B, h, w, H, W):
    import stego_oracle as O
    from stego_b200.eval import fused_probe_log_probs
    from stego_b200.modules import ClusterLookup
    g = torch.Generator().manual_seed(h * 100 + W)
    code = torch.randn(B, 70, h, w, generator=g)
    lin = torch.nn.Conv2d(70, 27, (1, 1))
    with torch.no_grad():
        lin.weight.copy_(torch.randn(27, 70, 1, 1, generator=g) * 0.3)
        lin.bias.copy_(torch.randn(27, generator=g) * 0.1)
    clu = ClusterLookup(70, 27)
    with torch.no_grad():
        clu.clusters.copy_(torch.randn(27, 70, generator=g))
    # reference op sequence
    up = F.interpolate(code, (H, W), mode="bilinear", align_corners=False)
    want_lin = torch.log_softmax(F.conv2d(up, lin.weight, lin.bias), dim=1)
    want_clu = O.cluster_lookup(up, clu.clusters.detach(), 2.0, log_probs=True)
    got_lin, got_clu, la, ca = fused_probe_log_probs(code.to(cuda_dev), lin.to(cuda_dev), clu.to(cuda_dev), (H, W), 2.0,
                                                     want_argmax=True)
    assert got_lin.shape == want_lin.shape and got_clu.shape == want_clu.shape
    assert (got_lin.cpu() - want_lin).abs().max().item() < 2e-5
    assert (got_clu.cpu() - want_clu).abs().max().item() < 2e-5
    for got_arg, want in ((la, want_lin), (ca, want_clu)):
        top2 = want.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 1e-4
        assert torch.equal(got_arg.cpu().long()[safe], want.argmax(1)[safe])
        assert (~safe).float().mean() < 5e-3


def test_fused_eval_probes_full_frame_properties(cuda_dev):
    """configs[4] size (1024 x 2048 from a 128 x 256 code): log-probs normalise, argmax maps agree with them."""
    from stego_b200.eval import fused_probe_log_probs
    from stego_b200.modules import ClusterLookup
    torch.manual_seed(0)
    code = torch.randn(1, 70, 128, 256, device=cuda_dev)
    lin = torch.nn.Conv2d(70, 27, (1, 1)).to(cuda_dev)
    clu = ClusterLookup(70, 27).to(cuda_dev)
    l, c, la, ca = fused_probe_log_probs(code, lin, clu, (1024, 2048), 2.0, want_argmax=True)
    assert l.shape == (1, 27, 1024, 2048)
    assert (l.exp().sum(1) - 1).abs().max().item() < 1e-4
    assert (c.exp().sum(1) - 1).abs().max().item() < 1e-4
    # the kernel's argmax attains the maximum log-probability (indices may differ from torch.argmax only where two
    # classes round to the same fp32 log-prob)
    assert torch.equal(l.gather(1, la.long().unsqueeze(1)).squeeze(1), l.max(1).values)
    assert torch.equal(c.gather(1, ca.long().unsqueeze(1)).squeeze(1), c.max(1).values)
    assert (l.argmax(1) != la.long()).float().mean().item() < 1e-4
