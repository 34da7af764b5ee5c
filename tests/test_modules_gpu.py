"""Module-level parity on the GPU: DinoFeaturizer (ViT + head), ClusterLookup, and one full training step
(head + correspondence loss + probes + backward + Adam) against the CPU oracle.

ViT tolerance: the CUDA backbone computes GEMM operands in bf16 with fp32 accumulation and an fp32 residual
stream; against the fp32 oracle the final features agree to ~3e-3 relative L2 (asserted < 1e-2).
Everything downstream of the backbone is compared on IDENTICAL (bf16-valued) features, masks, coords and
perms; losses and seg-head gradients are asserted within 1e-3 relative (the BASELINE.json bar), with the
oracle rounding GEMM operands to bf16 at the points a bf16 run of the reference would.
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))


def _rel(x, y):
    x, y = x.detach().double().cpu(), y.detach().double().cpu()
    return ((x - y).norm() / y.norm().clamp_min(1e-30)).item()


def _make_net(cfg, cuda_dev, seed=0):
    import stego_oracle as O
    from stego_b200.modules import DinoFeaturizer
    torch.manual_seed(seed)
    net = DinoFeaturizer(cfg.dim, cfg).to(cuda_dev)
    sd = O.perturb_vit_state(O.vit_random_state(cfg.model_type, 8, seed=3))
    net.model.load_state_dict(sd)
    return net, sd


@pytest.mark.parametrize("arch,res,B", [("vit_small", 224, 2), ("vit_small", 96, 3), ("vit_base", 64, 2)])
def test_vit_features_match_oracle(cuda_dev, arch, res, B):
    import stego_oracle as O
    from stego_b200.config import make_cfg
    cfg = make_cfg(model_type=arch, random_backbone_init=True)
    net, sd = _make_net(cfg, cuda_dev)
    net.eval()
    torch.manual_seed(5)
    img = torch.randn(B, 3, res, res)
    with torch.no_grad():
        want = O.vit_image_feat(sd, img, arch, 8)
        feat, code = net(img.to(cuda_dev))
    assert feat.shape == want.shape and code.shape == (B, 70, res // 8, res // 8)
    assert _rel(feat, want) < 1e-2
    # state-dict compatibility with the reference names
    assert set(sd.keys()) == set(net.model.state_dict().keys())


def test_featurizer_head_and_dropout_rng(cuda_dev):
    """Head parity (eval and train mode) and the Dropout2d RNG contract: our masks are the ones
    nn.Dropout2d would have drawn for the same generator state."""
    import stego_oracle as O
    from stego_b200.config import make_cfg
    cfg = make_cfg(random_backbone_init=True)
    net, sd = _make_net(cfg, cuda_dev)
    B, res = 2, 64
    torch.manual_seed(6)
    img = torch.randn(B, 3, res, res, device=cuda_dev)
    hp = {k: v.detach().cpu() for k, v in net.state_dict().items() if k.startswith("cluster")}
    net.eval()
    with torch.no_grad():
        feat, code = net(img)
    _, want_code = O.head_forward(feat.cpu(), hp, None, round_bf16=True)
    assert _rel(code, want_code) < 1e-3
    net.train()
    torch.manual_seed(123)
    feat_t, code_t = net(img)
    torch.manual_seed(123)
    drop = torch.nn.Dropout2d(p=.1)
    dummy = torch.ones(B, 384, 8, 8, device=cuda_dev)
    masks = [drop(dummy)[:, :, :1, :1].cpu() for _ in range(3)]
    want_feat, want_code = O.head_forward(feat.cpu(), hp, masks, round_bf16=True)
    assert _rel(feat_t, want_feat) < 1e-6
    assert _rel(code_t, want_code) < 1e-3


def test_cluster_lookup_matches_oracle(cuda_dev):
    import stego_oracle as O
    from stego_b200.modules import ClusterLookup
    torch.manual_seed(7)
    cl = ClusterLookup(70, 27).to(cuda_dev)
    x = torch.randn(2, 70, 28, 28)
    clusters = cl.clusters.detach().cpu()
    loss, probs = cl(x.to(cuda_dev), None)
    wl, wp = O.cluster_lookup(x, clusters, None)
    ip = O.cluster_lookup(x, clusters, 1.0, log_probs=True)  # margins via the soft scores
    top2 = ip.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-5
    assert torch.equal(probs.argmax(1).cpu()[safe], wp.argmax(1)[safe])  # bit-exact assignments (off ties)
    assert (~safe).float().mean() < 1e-3
    assert abs(loss.item() - wl.item()) < 1e-5
    # channels-last input, softmax and log-prob modes
    xc = x.to(cuda_dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    l2, p2 = cl(xc, 3.0)
    wl2, wp2 = O.cluster_lookup(x, clusters, 3.0)
    assert _rel(p2, wp2) < 1e-5 and abs(l2.item() - wl2.item()) < 1e-5
    lp = cl(xc, 2.0, log_probs=True)
    assert _rel(lp, O.cluster_lookup(x, clusters, 2.0, log_probs=True)) < 1e-5
    # gradient wrt the centroids (both modes)
    for alpha in (None, 2.0):
        c_ref = clusters.clone().requires_grad_(True)
        O.cluster_lookup(x, c_ref, alpha)[0].backward()
        cl.clusters.grad = None
        cl(x.to(cuda_dev), alpha)[0].backward()
        assert _rel(cl.clusters.grad, c_ref.grad) < 1e-4


@pytest.mark.parametrize("fused", [True, False], ids=["handscheduled", "autograd"])
@pytest.mark.parametrize("arch,res,B", [("vit_small", 224, 2), ("vit_small", 64, 4)])
def test_training_step_matches_oracle(cuda_dev, arch, res, B, fused):
    import stego_oracle as O
    from stego_b200.config import make_cfg
    from stego_b200.modules import super_perm
    from stego_b200.segmenter import LitUnsupervisedSegmenter
    cfg = make_cfg(model_type=arch, random_backbone_init=True, fused_step=fused)
    ocfg = O.LossCfg()
    torch.manual_seed(0)
    model = LitUnsupervisedSegmenter(27, cfg).to(cuda_dev)
    model.net.model.load_state_dict(O.perturb_vit_state(O.vit_random_state(arch, 8, seed=3)))
    model.train()
    model.configure_optimizers()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(B, 3, res, res, generator=g)
    img_pos = img + 0.3 * torch.randn(B, 3, res, res, generator=g)
    label = torch.randint(-1, 27, (B, res, res), generator=g)
    batch = dict(img=img.to(cuda_dev), img_pos=img_pos.to(cuda_dev), label=label.to(cuda_dev))
    names = ["net.cluster1.0.weight", "net.cluster1.0.bias", "net.cluster2.0.weight", "net.cluster2.0.bias",
             "net.cluster2.2.weight", "net.cluster2.2.bias", "linear_probe.weight", "linear_probe.bias",
             "cluster_probe.clusters"]
    params0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if k in names}

    # replay the RNG stream of the step to learn the draws it will make
    torch.manual_seed(777)
    m = model.net.draw_masks(B, cuda_dev)
    mp = model.net.draw_masks(B, cuda_dev)
    c1, c2 = model.contrastive_corr_loss_fn.draw_coords(batch["img"], None, None)
    perms = [super_perm(B, cuda_dev) for _ in range(5)]
    torch.manual_seed(777)
    loss = model.training_step(batch, 0)
    assert (model._fused is not None and model._fused.step_idx == 1) == fused
    grads = {k: dict(model.named_parameters())[k].grad.detach().cpu().clone() for k in names}
    params1 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if k in names}

    # oracle on the SAME backbone features (bf16-valued) and the same draws
    with torch.no_grad():
        tok = model.net.backbone_tokens(torch.cat([batch["img"], batch["img_pos"]], 0)).float().cpu()
    h = res // 8
    f_all = tok.view(2 * B, h, h, -1).permute(0, 3, 1, 2)
    hp = {k[len("net."):]: v.clone().requires_grad_(True) for k, v in params0.items() if k.startswith("net.")}
    probes = {k: v.clone().requires_grad_(True) for k, v in params0.items() if not k.startswith("net.")}
    to4 = lambda t: t.cpu().view(B, -1, 1, 1)
    out = O.training_losses(f_all[:B], f_all[B:], hp, probes, label, [to4(x) for x in m], [to4(x) for x in mp],
                            c1.cpu(), c2.cpu(), [p.cpu() for p in perms], ocfg, 27, round_bf16=True)
    out["total"].backward()
    logged = {k: float(v) for k, v in model.logged.items()}
    elem_scale = 0.05  # typical |loss element|; call means cancel to ~1e-3 on random features
    assert abs(logged["loss/linear"] - out["linear"].item()) < 1e-4 * abs(out["linear"].item()) + 1e-6
    assert abs(logged["loss/cluster"] - out["cluster"].item()) < 2e-4 * abs(out["cluster"].item()) + 1e-6
    for k_log, k_or in [("loss/pos_intra", "pos_intra"), ("loss/pos_inter", "pos_inter"), ("loss/neg_inter", "neg_inter")]:
        assert abs(logged[k_log] - out[k_or].item()) < 1e-3 * abs(out[k_or].item()) + 1e-3 * elem_scale
    assert abs(float(loss) - out["total"].item()) < 1e-3 * abs(out["total"].item())
    want_g = {("net." + k): v.grad for k, v in hp.items()}
    want_g.update({k: v.grad for k, v in probes.items()})
    for k in names:
        assert _rel(grads[k], want_g[k]) < 1e-3, (k, _rel(grads[k], want_g[k]))
    # one Adam step (lr 5e-4 head, 5e-3 probes): the fused kernel must reproduce torch.optim.Adam arithmetic
    # on the gradients it was given (checked against the oracle's restatement of Adam)
    for k in names:
        p = params0[k].clone()
        O.adam_step(p, grads[k], torch.zeros_like(p), torch.zeros_like(p), 1, 5e-4 if k.startswith("net.") else 5e-3)
        assert _rel(params1[k] - params0[k], p - params0[k]) < 1e-4, k


@pytest.mark.parametrize("B,h,w,H,W", [(2, 28, 28, 224, 224), (3, 7, 9, 50, 61), (1, 12, 12, 12, 12), (2, 40, 40, 320, 320)])
def test_linear_probe_ce_matches_oracle(cuda_dev, B, h, w, H, W):
    import stego_oracle as O
    from stego_b200.segmenter import linear_probe_ce
    g = torch.Generator().manual_seed(B * 1000 + H)
    code = torch.randn(B, 70, h, w, generator=g)
    weight = (torch.randn(27, 70, 1, 1, generator=g) * 0.3).requires_grad_(True)
    bias = (torch.randn(27, generator=g) * 0.1).requires_grad_(True)
    label = torch.randint(-1, 29, (B, H, W), generator=g)  # includes ignored labels (-1, 27, 28)
    want = O.linear_probe_loss(code, weight, bias, label, 27)
    gw, gb = torch.autograd.grad(want, [weight, bias])
    wg = weight.detach().to(cuda_dev).requires_grad_(True)
    bg = bias.detach().to(cuda_dev).requires_grad_(True)
    cg = code.to(cuda_dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    got = linear_probe_ce(cg, wg, bg, label.to(cuda_dev))
    ggw, ggb = torch.autograd.grad(got * 2.0, [wg, bg])
    assert abs(got.item() - want.item()) < 1e-5 * abs(want.item()) + 1e-6
    assert _rel(ggw, 2 * gw) < 1e-4 and _rel(ggb, 2 * gb) < 1e-4
