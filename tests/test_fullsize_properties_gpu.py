"""Size-independent properties of the CUDA path at the FULL BASELINE.json sizes (c1: B=32, ViT-S/8 224^2 ->
28x28 maps, E=384; c2 map size 40x40, E=768), where running the CPU oracle would take too long:

  * intra-call code correlation of a normalised vector with itself is 1 on the diagonal;
  * the reported per-call statistics are consistent with the returned element tensors (mean of elements);
  * the analytic gradient matches a central finite difference of the loss along a random direction
    (directional derivative), which exercises the whole backward chain (einsum dgrads, norm, grid-sample scatter);
  * permutation equivariance: permuting the batch (with coords / perms permuted accordingly) permutes nothing in
    the scalar loss.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg():
    from stego_b200.config import make_cfg
    return make_cfg()


@pytest.mark.parametrize("B,E,h", [(32, 384, 28), (32, 768, 40)])
def test_full_size_loss_properties(cuda_dev, B, E, h):
    from stego_b200 import corr
    cfg = _cfg()
    spec = corr.LossSpec(cfg)
    g = torch.Generator(device=cuda_dev).manual_seed(B + h)
    rnd = lambda *s: torch.randn(*s, device=cuda_dev, generator=g)
    z = rnd(B, 16, h, h)
    bf, bc = rnd(16, E), rnd(16, 70)
    feats = (torch.einsum("bkhw,kc->bhwc", z, bf) + 0.1 * rnd(B, h, h, E)).bfloat16().permute(0, 3, 1, 2)
    zp = z + 0.3 * rnd(B, 16, h, h)
    feats_pos = (torch.einsum("bkhw,kc->bhwc", zp, bf) + 0.1 * rnd(B, h, h, E)).bfloat16().permute(0, 3, 1, 2)
    code = (torch.einsum("bkhw,kc->bhwc", z, bc) + 0.1 * rnd(B, h, h, 70)).permute(0, 3, 1, 2)
    code_pos = (torch.einsum("bkhw,kc->bhwc", zp, bc) + 0.1 * rnd(B, h, h, 70)).permute(0, 3, 1, 2)
    c1 = torch.rand(B, 11, 11, 2, device=cuda_dev, generator=g) * 2 - 1
    c2 = torch.rand(B, 11, 11, 2, device=cuda_dev, generator=g) * 2 - 1
    perms = torch.stack([torch.randperm(B, device=cuda_dev, generator=g) for _ in range(5)])
    w = torch.tensor([0.67, 0.25] + [0.63 / 5] * 5, device=cuda_dev)

    def total(cd_, cp_, elems=False):
        losses, cd_means, cd, el = corr.corr_loss(feats, feats_pos, cd_, cp_, c1, c2, perms, spec, want_elems=elems)
        return (losses * w).sum(), losses, cd_means, cd, el

    code_g, code_pos_g = code.clone().requires_grad_(True), code_pos.clone().requires_grad_(True)
    tot, losses, cd_means, cd, el = total(code_g, code_pos_g, elems=True)
    # diagonal of the intra-call code correlation
    diag = cd[0].diagonal(dim1=1, dim2=2)
    assert (diag - 1).abs().max().item() < 2e-5
    # statistics vs elements
    assert torch.allclose(cd.mean(dim=(1, 2, 3)), cd_means, atol=1e-6)
    assert torch.allclose(el.mean(dim=(1, 2, 3)), losses, atol=2e-6, rtol=1e-4)
    # directional derivative
    ga, gb = torch.autograd.grad(tot, [code_g, code_pos_g])
    da, db = rnd(*code.shape), rnd(*code_pos.shape)
    eps = 1e-2
    with torch.no_grad():
        lp = total(code + eps * da, code_pos + eps * db)[0].double()
        lm = total(code - eps * da, code_pos - eps * db)[0].double()
    fd = ((lp - lm) / (2 * eps)).item()
    an = ((ga.double() * da.double()).sum() + (gb.double() * db.double()).sum()).item()
    assert abs(fd - an) < 2e-2 * abs(an) + 1e-6, (fd, an)
    # batch-permutation equivariance of the scalar loss
    pi = torch.randperm(B, device=cuda_dev, generator=g)
    inv = torch.argsort(pi)
    perms_p = inv[perms[:, pi]]  # image pi[b] now sits at position b; its negative partner index is relabelled
    with torch.no_grad():
        losses_p, _, _, _ = corr.corr_loss(feats[pi], feats_pos[pi], code[pi], code_pos[pi], c1[pi], c2[pi], perms_p, spec)
    assert torch.allclose(losses_p, losses.detach(), rtol=2e-4, atol=2e-6)
