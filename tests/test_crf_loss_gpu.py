"""ContrastiveCRFLoss (src/modules.py:437-469) on the fused CUDA kernels (csrc/crf_loss.cu) against the
oracle restatement (pinned bit-exactly to the reference module in oracle/check_against_reference.py) and against the golden
fixture the reference itself produced.  fp32 FMA accumulation in a different order than torch's bmm: tolerance 2e-6 of the
output scale (forward), 1e-5 of the gradient scale (backward: atomics + recomputed exponentials)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
PARAMS = (.5, .15, .05, 10.0, 3.0, 0.00)  # train_config.yml:131-137 (alpha, beta, gamma, w1, w2, shift)


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,C,H,W,n,shift,layout", [(2, 70, 56, 56, 1000, 0.0, "nchw"), (3, 70, 20, 24, 203, 0.1, "nhwc"),
                                                     (1, 27, 9, 7, 64, 0.0, "nchw"), (2, 5, 12, 12, 500, 0.05, "nhwc")])
def test_crf_loss_matches_oracle(cuda_dev, B, C, H, W, n, shift, layout):
    import stego_oracle as O
    from stego_b200.modules import ContrastiveCRFLoss
    g = torch.Generator().manual_seed(B * 1000 + n)
    gd = torch.rand(B, 3, H, W, generator=g) * 4 - 2
    cl = torch.nn.functional.normalize(torch.randn(B, C, H, W, generator=g), dim=1)
    coords = torch.cat([torch.randint(0, H, size=[1, n], generator=g), torch.randint(0, W, size=[1, n], generator=g)], 0)
    up = torch.randn(B, n, n, generator=g)  # a non-uniform, non-symmetric upstream gradient
    p = PARAMS[:5] + (shift,)
    c_cpu = cl.clone().requires_grad_(True)
    want = O.contrastive_crf_loss(gd, c_cpu, coords, *p)
    gw, = torch.autograd.grad((want * up).sum(), c_cpu)
    c_dev = cl.to(cuda_dev)
    if layout == "nhwc":  # the kernels take element strides: channels-last views, no copy
        c_dev = c_dev.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    c_dev.requires_grad_(True)
    mod = ContrastiveCRFLoss(n, *p)
    got = mod.forward_with_coords(gd.to(cuda_dev), c_dev, coords.to(cuda_dev))
    assert got.shape == (B, n, n)
    gg, = torch.autograd.grad((got * up.to(cuda_dev)).sum(), c_dev)
    ef, eb = _rel(got.detach(), want.detach()), _rel(gg, gw)
    print(f"crf loss B={B} C={C} n={n}: fwd {ef:.2e} bwd {eb:.2e}")
    assert ef < 2e-6 and eb < 1e-5


def test_crf_loss_golden_and_rng(cuda_dev):
    """The reference's own output (tests/golden/contrastive_crf_loss.pt) through the CUDA path, and the RNG contract: forward()
    draws row indices then column indices with two torch.randint calls on the tensors' device."""
    from stego_b200.modules import ContrastiveCRFLoss
    g = torch.load(os.path.join(ROOT, "tests", "golden", "contrastive_crf_loss.pt"))
    torch.manual_seed(51)
    gd = torch.rand(2, 3, 56, 56) * 4 - 2
    cl = torch.nn.functional.normalize(torch.randn(2, 70, 56, 56), dim=1)
    torch.manual_seed(52)
    coords = torch.cat([torch.randint(0, 56, size=[1, 300]), torch.randint(0, 56, size=[1, 300])], 0)
    mod = ContrastiveCRFLoss(300, *PARAMS)
    c_dev = cl.to(cuda_dev).requires_grad_(True)
    out = mod.forward_with_coords(gd.to(cuda_dev), c_dev, coords.to(cuda_dev))
    grad, = torch.autograd.grad(out.mean(), c_dev)
    assert torch.allclose(out.detach().reshape(-1)[::97].cpu(), g["out_sub"], atol=2e-5)
    assert torch.allclose(grad.reshape(-1)[::53].cpu(), g["grad_sub"], atol=1e-9, rtol=1e-4)
    assert abs(grad.abs().sum().item() - g["grad_abs_sum"].item()) < 1e-4 * g["grad_abs_sum"].item()
    # RNG stream
    torch.manual_seed(77)
    want = torch.cat([torch.randint(0, 56, size=[1, 300], device=cuda_dev), torch.randint(0, 40, size=[1, 300], device=cuda_dev)], 0)
    torch.manual_seed(77)
    assert torch.equal(mod.draw_coords(56, 40, cuda_dev), want)
    # and the public forward is forward_with_coords on those draws
    torch.manual_seed(78)
    a = mod(gd.to(cuda_dev), c_dev)
    torch.manual_seed(78)
    b = mod.forward_with_coords(gd.to(cuda_dev), c_dev, mod.draw_coords(56, 56, cuda_dev))
    assert torch.equal(a, b)


def test_training_step_with_crf_term(cuda_dev):
    """train_segmentation.py:201-208: with crf_weight > 0 the step adds crf_weight * mean(ContrastiveCRFLoss(resize(img, 56),
    norm(resize(code, 56)))) and its gradient reaches the segmentation head."""
    from stego_b200.config import make_cfg
    from stego_b200.segmenter import LitUnsupervisedSegmenter
    losses = {}
    for wgt in (0.0, 0.5):
        torch.manual_seed(0)
        cfg = make_cfg(model_type="vit_small", res=64, batch_size=2, random_backbone_init=True, crf_weight=wgt, crf_samples=128)
        model = LitUnsupervisedSegmenter(27, cfg).to(cuda_dev).train()
        model.configure_optimizers()
        g = torch.Generator().manual_seed(5)
        batch = dict(img=torch.randn(2, 3, 64, 64, generator=g).to(cuda_dev), img_pos=torch.randn(2, 3, 64, 64, generator=g).to(cuda_dev),
                     label=torch.randint(-1, 27, (2, 64, 64), generator=g).to(cuda_dev))
        torch.manual_seed(9)
        losses[wgt] = float(model.training_step(batch, 0))
        model.flush()
        assert all(torch.isfinite(p).all() for p in model.parameters())
    assert torch.isfinite(torch.tensor(list(losses.values()))).all()
    assert abs(losses[0.5] - losses[0.0]) > 1e-6


@pytest.mark.parametrize("layout", ["nchw", "nhwc", "mixed"])
def test_pixel_cosine_matches_reference_ops(cuda_dev, layout):
    """train_segmentation.py:185,194-198: (norm(a) * norm(b)).sum(1) with F.normalize(eps=1e-10) and its gradients, incl. a
    zero vector (clamped norm) in each operand."""
    from stego_b200.modules import pixel_cosine
    g = torch.Generator().manual_seed(3)
    B, C, H, W = 3, 70, 9, 11
    a, b = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g) * 3
    a[0, :, 0, 0] = 0
    b[1, :, 2, 3] = 0
    up = torch.randn(B, H, W, generator=g)
    a_c, b_c = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    nrm = lambda t: torch.nn.functional.normalize(t, dim=1, eps=1e-10)
    want = (nrm(a_c) * nrm(b_c)).sum(1)
    wa, wb = torch.autograd.grad((want * up).sum(), (a_c, b_c))
    cl = lambda t: t.to(cuda_dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    a_d = (cl(a) if layout in ("nhwc", "mixed") else a.to(cuda_dev)).requires_grad_(True)
    b_d = (cl(b) if layout == "nhwc" else b.to(cuda_dev)).requires_grad_(True)
    got = pixel_cosine(a_d, b_d)
    ga, gb = torch.autograd.grad((got * up.to(cuda_dev)).sum(), (a_d, b_d))
    assert _rel(got.detach(), want.detach()) < 2e-6
    # the zero vectors get a 1/eps-scaled gradient (F.normalize clamps the norm): compare them on their own scale
    ga, gb = ga.cpu(), gb.cpu()
    assert _rel(ga[0, :, 0, 0], wa[0, :, 0, 0]) < 1e-5 and _rel(gb[1, :, 2, 3], wb[1, :, 2, 3]) < 1e-5
    for t in (ga, wa):
        t[0, :, 0, 0] = 0
    for t in (gb, wb):
        t[1, :, 2, 3] = 0
    assert _rel(ga, wa) < 1e-5 and _rel(gb, wb) < 1e-5
