"""Generate tests/golden/*.pt from the REAL reference (build container only; /root/reference must exist).

    python oracle/make_golden.py

The reference ships no tests or golden vectors, so these fixtures are outputs of the
reference's own code (imported through oracle/reference_shim.py) on seeded inputs.  Large inputs are NOT
stored: they are regenerated from the recorded seeds with the CPU generator (same torch build on the GPU
box), only outputs / sub-sampled outputs are stored, so the fixtures stay small.
"""
from __future__ import annotations

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_shim  # noqa: E402
import stego_oracle as O  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def kat_inputs():
    """SURVEY.md §4.3 recipe (config c0 shapes)."""
    torch.manual_seed(1234)
    feats = torch.randn(2, 384, 28, 28)
    feats_pos = torch.randn(2, 384, 28, 28)
    code = torch.randn(2, 70, 28, 28)
    code_pos = torch.randn(2, 70, 28, 28)
    return feats, feats_pos, code, code_pos


def small_inputs():
    g = torch.Generator().manual_seed(4321)
    B, E, D, h = 3, 64, 70, 12
    basis_f, basis_c = torch.randn(8, E, generator=g), torch.randn(8, D, generator=g)
    z = torch.randn(B, 8, h, h, generator=g)
    zp = z + 0.3 * torch.randn(B, 8, h, h, generator=g)
    mk = lambda zz, bs, C: torch.einsum("bkhw,kc->bchw", zz, bs) + 0.1 * torch.randn(B, C, h, h, generator=g)
    return mk(z, basis_f, E), mk(zp, basis_f, E), mk(z, basis_c, D), mk(zp, basis_c, D)


def main():
    ref, vits = reference_shim.import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    cfg = O.LossCfg()
    ns = types.SimpleNamespace(**cfg.__dict__)

    # ---- 1. KAT (§4.3): scalars, grad statistics, sub-sampled grads --------------------------------
    feats, feats_pos, code, code_pos = kat_inputs()
    code.requires_grad_(True)
    code_pos.requires_grad_(True)
    torch.manual_seed(99)
    o = ref.ContrastiveCorrelationLoss(ns)(feats, feats_pos, None, None, code, code_pos)
    loss = .67 * o[0] + .25 * o[2] + .63 * o[4].mean()
    loss.backward()
    torch.manual_seed(99)
    c1, c2, perms = O.draw_loss_randomness(2, cfg)
    torch.save(dict(
        recipe="manual_seed(1234); feats, feats_pos = randn(2,384,28,28) x2; code, code_pos = randn(2,70,28,28) x2; "
               "manual_seed(99); ContrastiveCorrelationLoss(shipped cfg)",
        pos_intra_loss=o[0].detach(), pos_inter_loss=o[2].detach(), neg_inter_loss_mean=o[4].mean().detach(),
        cd_means=torch.stack([o[1].mean(), o[3].mean(), o[5].mean()]).detach(), total=loss.detach(),
        code_grad_norm=code.grad.norm(), code_grad_sum=code.grad.sum(),
        code_pos_grad_norm=code_pos.grad.norm(), code_pos_grad_sum=code_pos.grad.sum(),
        code_grad_sub=code.grad.reshape(-1)[::97].clone(), code_pos_grad_sub=code_pos.grad.reshape(-1)[::97].clone(),
        intra_cd_sub=o[1].detach().reshape(-1)[::211].clone(), neg_loss_sub=o[4].detach().reshape(-1)[::211].clone(),
        coords1=c1, coords2=c2, perms=torch.stack(perms)), os.path.join(OUT, "corr_kat_c0.pt"))

    # ---- 2. small correlated case with full inputs -------------------------------------------------
    f, fp, c, cp = small_inputs()
    c.requires_grad_(True)
    cp.requires_grad_(True)
    torch.manual_seed(5)
    o = ref.ContrastiveCorrelationLoss(ns)(f, fp, None, None, c, cp)
    loss = .67 * o[0] + .25 * o[2] + .63 * o[4].mean()
    loss.backward()
    torch.manual_seed(5)
    c1, c2, perms = O.draw_loss_randomness(3, cfg)
    torch.save(dict(feats=f, feats_pos=fp, code=c.detach(), code_pos=cp.detach(), coords1=c1, coords2=c2,
                    perms=torch.stack(perms), pos_intra_loss=o[0].detach(), pos_inter_loss=o[2].detach(),
                    neg_inter_loss_mean=o[4].mean().detach(), total=loss.detach(),
                    cd_means=torch.stack([o[1].mean(), o[3].mean(), o[5].mean()]).detach(),
                    inter_cd_sub=o[3].detach().reshape(-1)[::53].clone(), neg_loss_sub=o[4].detach().reshape(-1)[::53].clone(),
                    code_grad=c.grad.clone(), code_pos_grad=cp.grad.clone()), os.path.join(OUT, "corr_small.pt"))

    # ---- 3. ClusterLookup KAT ----------------------------------------------------------------------
    torch.manual_seed(7)
    cl = ref.ClusterLookup(70, 27)
    x = torch.randn(2, 70, 28, 28)
    l, p = cl(x, None)
    lp = cl(x, 2.0, log_probs=True)
    torch.save(dict(recipe="manual_seed(7); ClusterLookup(70,27); x = randn(2,70,28,28)",
                    clusters=cl.clusters.detach().clone(), cluster_loss=l.detach(), argmax=p.argmax(1).to(torch.int16),
                    log_probs_sum=lp.sum().detach(), log_probs_sub=lp.detach().reshape(-1)[::101].clone()),
               os.path.join(OUT, "cluster_lookup_kat.pt"))

    # ---- 4. ViT-S/8 tokens on a 32x32 image (pos-embed interpolation path) -------------------------
    sd = O.perturb_vit_state(O.vit_random_state("vit_small", 8, seed=3))
    model = vits.vit_small(patch_size=8, num_classes=0)
    model.load_state_dict(sd)
    model.eval()
    torch.manual_seed(11)
    img = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        feat, _, _ = model.get_intermediate_feat(img, n=1)
    torch.save(dict(recipe="sd = perturb_vit_state(vit_random_state('vit_small', 8, seed=3)); manual_seed(11); "
                           "img = randn(2,3,32,32); get_intermediate_feat(img)[0][0]",
                    tokens=feat[0].clone()), os.path.join(OUT, "vit_small8_32px.pt"))

    # ---- 5. super_perm draws -----------------------------------------------------------------------
    rows = []
    for size in (1, 2, 5, 16, 32):
        torch.manual_seed(1000 + size)
        rows.append(torch.stack([ref.super_perm(size, torch.device("cpu")) for _ in range(3)]))
    torch.save(dict(recipe="for size in (1,2,5,16,32): manual_seed(1000+size); 3 x super_perm(size)",
                    draws=rows), os.path.join(OUT, "super_perm.pt"))
    # ---- 6. ContrastiveCRFLoss (modules.py:437-469) at the training call's shapes (56 x 56, 70 channels), 300 samples
    torch.manual_seed(51)
    gd = torch.rand(2, 3, 56, 56) * 4 - 2
    cl = torch.nn.functional.normalize(torch.randn(2, 70, 56, 56), dim=1).requires_grad_(True)
    crf = ref.ContrastiveCRFLoss(300, .5, .15, .05, 10.0, 3.0, 0.00)
    torch.manual_seed(52)
    out = crf(gd, cl)
    g, = torch.autograd.grad(out.mean(), cl)
    torch.save(dict(recipe="manual_seed(51); guidance = rand(2,3,56,56)*4-2; clusters = normalize(randn(2,70,56,56), dim=1); "
                           "ContrastiveCRFLoss(300, .5, .15, .05, 10, 3, 0) under manual_seed(52) (coords = randint(56,[1,300]) x 2); "
                           "grad of out.mean()",
                    out_sub=out.detach().reshape(-1)[::97].clone(), out_mean=out.detach().mean(), out_abs_sum=out.detach().abs().sum(),
                    grad_sub=g.reshape(-1)[::53].clone(), grad_abs_sum=g.abs().sum()),
               os.path.join(OUT, "contrastive_crf_loss.pt"))
    for f_ in sorted(os.listdir(OUT)):
        print(f_, os.path.getsize(os.path.join(OUT, f_)))


if __name__ == "__main__":
    main()
