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


@pytest.mark.parametrize("B,h,w,H,W,label_dtype", [(2, 10, 12, 80, 96, torch.int64), (1, 7, 9, 50, 61, torch.uint8),
                                                   (1, 16, 32, 128, 256, torch.int32)])
def test_flip_tta_and_confusion_match_reference_sequence(cuda_dev, B, h, w, H, W, label_dtype):
    """eval_segmentation.py:124-141 + utils.py:219-229: flip-TTA average of the two codes, upsample, both probes, argmax,
    UnsupervisedMetrics.update — one fused call against the reference op sequence on the CPU."""
    from stego_b200.eval import fused_probe_log_probs
    from stego_b200.modules import ClusterLookup
    import stego_oracle as O
    g = torch.Generator().manual_seed(h * 131 + W)
    code1 = torch.randn(B, 70, h, w, generator=g)
    code2 = torch.randn(B, 70, h, w, generator=g)  # code of the flipped image
    lin = torch.nn.Conv2d(70, 27, (1, 1))
    with torch.no_grad():
        lin.weight.copy_(torch.randn(27, 70, 1, 1, generator=g) * 0.3)
        lin.bias.copy_(torch.randn(27, generator=g) * 0.1)
    clu = ClusterLookup(70, 27)
    with torch.no_grad():
        clu.clusters.copy_(torch.randn(27, 70, generator=g))
    label = torch.randint(-1, 29, (B, H, W), generator=g)  # -1, 27, 28 are ignored
    # ---- reference sequence
    code = (code1 + code2.flip(dims=[3])) / 2
    up = F.interpolate(code, (H, W), mode="bilinear", align_corners=False)
    want_lin = torch.log_softmax(F.conv2d(up, lin.weight, lin.bias), dim=1)
    want_clu = O.cluster_lookup(up, clu.clusters.detach(), 2.0, log_probs=True)

    def update(preds, target, n=27):  # utils.py:219-229
        actual, preds = target.reshape(-1), preds.reshape(-1)
        mask = (actual >= 0) & (actual < n) & (preds >= 0) & (preds < n)
        return torch.bincount(n * actual[mask] + preds[mask], minlength=n * n).reshape(n, n).t()

    if label_dtype == torch.uint8:
        dev_label = label.clone()
        dev_label[(dev_label < 0) | (dev_label > 254)] = 255
        dev_label = dev_label.to(torch.uint8)
    else:
        dev_label = label.to(label_dtype)
    lc = torch.zeros(27, 27, dtype=torch.int64, device=cuda_dev)
    cc = torch.zeros(27, 27, dtype=torch.int64, device=cuda_dev)
    got_lin, got_clu, la, ca = fused_probe_log_probs(
        code1.to(cuda_dev), lin.to(cuda_dev), clu.to(cuda_dev), (H, W), 2.0, want_argmax=True,
        code_flipped=code2.to(cuda_dev), label=dev_label.to(cuda_dev), linear_confusion=lc, cluster_confusion=cc)
    assert (got_lin.cpu() - want_lin).abs().max().item() < 2e-5
    assert (got_clu.cpu() - want_clu).abs().max().item() < 2e-5
    # confusion counts are exact functions of the kernel's own argmax maps ...
    assert torch.equal(lc.cpu(), update(la.cpu().long(), label))
    assert torch.equal(cc.cpu(), update(ca.cpu().long(), label))
    # ... which equal the reference's argmax off fp32 near-ties
    for got_arg, want in ((la, want_lin), (ca, want_clu)):
        top2 = want.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 1e-4
        assert torch.equal(got_arg.cpu().long()[safe], want.argmax(1)[safe])
    assert (lc.cpu() - update(want_lin.argmax(1), label)).abs().sum().item() <= 2 * int((~safe).sum()) + 4
    # accumulation: a second call doubles the counts
    fused_probe_log_probs(code1.to(cuda_dev), lin, clu, (H, W), 2.0, want_log_probs=False, want_argmax=True,
                          code_flipped=code2.to(cuda_dev), label=dev_label.to(cuda_dev), linear_confusion=lc,
                          cluster_confusion=cc)
    assert torch.equal(lc.cpu(), 2 * update(la.cpu().long(), label))
