"""The boundary symbols `from modules import *` hands to the reference scripts, exercised on the GPU
the way the reference calls them:

  * ContrastiveCorrelationLoss.forward — the 6-tuple entry (src/modules.py:349-398) incl. autograd into the codes;
  * tensor_correlation (src/modules.py:283-284);
  * DinoFeaturizer.forward with dino_feat_type "KK" and with return_class_feat (src/modules.py:98-106);
  * ClusterLookup argmax assignments: EXACT flip counts against the fp32 oracle and against an fp64 evaluation;
  * the reference's own `LitUnsupervisedSegmenter.training_step` TEXT (src/train_segmentation.py:112-245) executed
    over stego_b200.modules through a stub Lightning base (oracle/lightning_harness.py), compared with the same text
    over the reference's own modules.py in PyTorch eager on the same GPU.
"""
import os
import sys

import pytest
import torch

from _parity_util import fp32_strict, record, rel

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))


def _correlated(B, C, h, w, g, rank=16):
    """Low-rank + noise maps: correlations span [-0.2, 0.9] so that clamp / shift branches fire."""
    basis = torch.randn(rank, C, generator=g)
    mix = torch.randn(B, h, w, rank, generator=g)
    return (mix @ basis + 0.1 * torch.randn(B, h, w, C, generator=g)).permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("kind", ["iid", "correlated"])
def test_contrastive_correlation_loss_forward_6tuple(cuda_dev, kind):
    import stego_oracle as O
    from stego_b200.config import make_cfg
    from stego_b200.modules import ContrastiveCorrelationLoss
    cfg = make_cfg()
    B, E, D, h = 3, 384, 70, 28
    g = torch.Generator().manual_seed(5)
    mk = (lambda c: torch.randn(B, c, h, h, generator=g)) if kind == "iid" else (lambda c: _correlated(B, c, h, h, g))
    feats, feats_pos, code, code_pos = mk(E), mk(E), mk(D), mk(D)
    lossfn = ContrastiveCorrelationLoss(cfg)
    dc = code.to(cuda_dev).requires_grad_(True)
    dcp = code_pos.to(cuda_dev).requires_grad_(True)
    torch.manual_seed(99)
    c1, c2, perms = O.draw_loss_randomness(B, O.LossCfg(), device=cuda_dev)  # the draws forward() is about to make
    torch.manual_seed(99)
    out = lossfn(feats.to(cuda_dev), feats_pos.to(cuda_dev), None, None, dc, dcp)
    assert len(out) == 6
    five = (11, 11, 11, 11)
    assert out[0].shape == () and out[2].shape == ()
    assert out[1].shape == (B, *five) and out[3].shape == (B, *five)
    assert out[4].shape == (5 * B, *five) and out[5].shape == (5 * B, *five)
    oc = code.clone().requires_grad_(True)
    ocp = code_pos.clone().requires_grad_(True)
    want = O.correlation_loss(feats, feats_pos, oc, ocp, c1.cpu(), c2.cpu(), [p.cpu() for p in perms], O.LossCfg())
    scale = max(want[4].abs().max().item(), 1e-3)
    for i in (1, 3, 5):  # cd tensors
        assert (out[i].cpu() - want[i]).abs().max().item() < 5e-5, i
    assert (out[4].cpu() - want[4]).abs().max().item() < 1e-4 * scale + 1e-5
    for i in (0, 2):
        assert abs(out[i].item() - want[i].item()) < 1e-3 * abs(want[i].item()) + 1e-4 * scale
    # the weighting of train_segmentation.py:169-181, backward into both codes (what manual_backward drives)
    (0.67 * out[0] + 0.25 * out[2] + 0.63 * out[4].mean()).backward()
    (0.67 * want[0] + 0.25 * want[2] + 0.63 * want[4].mean()).backward()
    # clamp(cd, 0) makes dL/dcd jump at cd = 0: an element whose |cd| is below the two implementations' fp32 difference
    # (~1e-6) can land on either side, and ONE such element moves the per-pixel code gradient by ~1/sqrt(active
    # elements).  The bound therefore counts those edge elements instead of depending on a lucky seed.
    def tol(cds):
        edge = sum(int((c.abs() < 5e-6).sum()) for c in cds)
        act = sum(int((c >= 0).sum()) for c in cds)
        return 1e-3 + 3.0 * (edge / max(act, 1)) ** 0.5
    assert rel(dc.grad, oc.grad) < tol([want[1], want[3], want[5]])
    assert rel(dcp.grad, ocp.grad) < tol([want[3]])


@pytest.mark.parametrize("n,c,hw,ij", [(2, 384, (11, 11), (11, 11)), (1, 70, (28, 28), (5, 7)), (3, 768, (3, 4), (40, 40))])
def test_tensor_correlation(cuda_dev, n, c, hw, ij):
    from stego_b200.modules import norm, tensor_correlation
    fp32_strict()
    g = torch.Generator().manual_seed(c)
    a = torch.randn(n, c, *hw, generator=g).to(cuda_dev)
    b = torch.randn(n, c, *ij, generator=g).to(cuda_dev)
    got = tensor_correlation(norm(a), norm(b))
    want = torch.einsum("nchw,ncij->nhwij", norm(a).double(), norm(b).double())
    assert got.shape == want.shape
    assert (got.double() - want).abs().max().item() < 2e-5


def test_featurizer_kk_and_class_feat(cuda_dev):
    """feat_type "KK" (keys of the last block, heads concatenated) and return_class_feat against the reference's own
    DinoFeaturizer run in PyTorch eager (fp32) on the same weights."""
    import lightning_harness as H
    if not H.available():
        pytest.skip("reference sources (baseline/_ref) not present")
    import tempfile
    from stego_b200.config import make_cfg
    from stego_b200.modules import DinoFeaturizer
    fp32_strict()
    ts = H.load_reference_segmenter("reference")
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "dino.pth")
        H.write_random_dino_checkpoint(ck, "vit_small")
        torch.manual_seed(6)
        img = torch.randn(2, 3, 64, 96, device=cuda_dev)
        for feat_type in ("KK", "feat"):
            cfg = make_cfg(dino_feat_type=feat_type, pretrained_weights=ck)
            torch.manual_seed(0)
            ref = ts._modules.DinoFeaturizer(70, cfg).to(cuda_dev).eval()
            torch.manual_seed(0)
            ours = DinoFeaturizer(70, cfg).to(cuda_dev).eval()
            ours.load_state_dict(ref.state_dict())
            with torch.no_grad():
                rf, rc = ref(img)
                of, oc = ours(img)
                assert of.shape == rf.shape and oc.shape == rc.shape
                assert rel(of, rf) < 1e-2, (feat_type, rel(of, rf))
                assert rel(oc, rc) < 2e-2, (feat_type, rel(oc, rc))
                if feat_type == "feat":
                    rcls = ref(img, return_class_feat=True)
                    ocls = ours(img, return_class_feat=True)
                    assert ocls.shape == rcls.shape == (2, 384, 1, 1)
                    assert rel(ocls, rcls) < 1e-2


@pytest.mark.parametrize("B,h,w", [(2, 28, 28), (2, 40, 40), (1, 56, 56), (1, 128, 256)])
def test_cluster_lookup_assignment_flip_count(cuda_dev, B, h, w):
    """north_star: ClusterLookup assignments bit-exact.  The argmax is over 27 fp32 inner products whose summation order
    differs between any two implementations (MKL, cuBLAS, this kernel), so exactness is stated against an fp64
    evaluation: EVERY pixel whose fp64 top-2 margin exceeds 1e-6 must get the fp64 argmax, and the number of pixels that
    differ from the fp32 CPU oracle / the fp32 GPU oracle is counted and reported, not hidden behind a carve-out."""
    import stego_oracle as O
    from stego_b200.modules import ClusterLookup
    fp32_strict()
    torch.manual_seed(7)
    cl = ClusterLookup(70, 27).to(cuda_dev)
    g = torch.Generator().manual_seed(h * w)
    x = torch.randn(B, 70, h, w, generator=g)
    clusters = cl.clusters.detach().cpu()
    _, probs = cl(x.to(cuda_dev), None)
    got = probs.argmax(1).cpu()
    assert torch.equal(probs.sum(1).cpu(), torch.ones(B, h, w))  # one-hot
    nc = clusters.double() / clusters.double().norm(dim=1, keepdim=True)
    nx = x.double() / x.double().norm(dim=1, keepdim=True)
    sim64 = torch.einsum("bchw,nc->bnhw", nx, nc)
    top2 = sim64.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    want64 = sim64.argmax(1)
    want32_cpu = O.cluster_lookup(x, clusters, None)[1].argmax(1)
    want32_gpu = O.cluster_lookup(x.to(cuda_dev), clusters.to(cuda_dev), None)[1].argmax(1).cpu()
    flips64 = int((got != want64).sum())
    flips_cpu = int((got != want32_cpu).sum())
    flips_gpu = int((got != want32_gpu).sum())
    oracle_disagree = int((want32_cpu != want32_gpu).sum())
    npix = B * h * w
    worst_margin = float(margin[got != want64].max()) if flips64 else 0.0
    record(f"cluster_lookup_flips_{h}x{w}", dict(pixels=npix, flips_vs_fp64=flips64, flips_vs_fp32_cpu_oracle=flips_cpu,
                                                 flips_vs_fp32_gpu_oracle=flips_gpu,
                                                 cpu_vs_gpu_oracle_disagreements=oracle_disagree,
                                                 largest_fp64_margin_among_flips=worst_margin,
                                                 pixels_with_margin_below_1e6=int((margin <= 1e-6).sum())))
    print(f"ClusterLookup {B}x{h}x{w}: {flips64}/{npix} differ from fp64, {flips_cpu} from the fp32 CPU oracle, "
          f"{flips_gpu} from the fp32 GPU oracle (the two oracles differ on {oracle_disagree}); worst margin {worst_margin:.1e}")
    safe = margin > 1e-6
    assert torch.equal(got[safe], want64[safe])
    assert flips_cpu <= max(2, oracle_disagree + 2)  # no more flips than two fp32 library paths have between themselves (+2)


def test_eval_frame_assignments_vs_oracle(cuda_dev):
    """configs[4] frame (1024 x 2048 from a 128 x 256 code): fused upsample + ClusterLookup argmax against the oracle's
    reference op sequence (F.interpolate -> ClusterLookup) run in fp32 on the GPU; flips counted."""
    import torch.nn.functional as F
    import stego_oracle as O
    from stego_b200.eval import fused_probe_log_probs
    from stego_b200.modules import ClusterLookup
    fp32_strict()
    torch.manual_seed(0)
    code = torch.randn(1, 70, 128, 256, device=cuda_dev)
    lin = torch.nn.Conv2d(70, 27, (1, 1)).to(cuda_dev)
    clu = ClusterLookup(70, 27).to(cuda_dev)
    _, c, la, ca = fused_probe_log_probs(code, lin, clu, (1024, 2048), 2.0, want_argmax=True)
    with torch.no_grad():
        up = F.interpolate(code, (1024, 2048), mode="bilinear", align_corners=False)
        want = O.cluster_lookup(up, clu.clusters.detach(), 2.0, log_probs=True)
        want_lin = torch.log_softmax(F.conv2d(up, lin.weight, lin.bias), dim=1)
    assert (c - want).abs().max().item() < 5e-5
    for name, got_arg, w in (("cluster", ca, want), ("linear", la, want_lin)):
        top2 = w.topk(2, dim=1).values
        margin = top2[:, 0] - top2[:, 1]
        diff = got_arg.long() != w.argmax(1)
        flips = int(diff.sum())
        worst = float(margin[diff].max()) if flips else 0.0
        record(f"eval_frame_flips_{name}", dict(pixels=1024 * 2048, flips=flips, largest_margin_among_flips=worst))
        print(f"eval frame {name}: {flips} / {1024 * 2048} argmax differences, worst log-prob margin {worst:.1e}")
        assert worst < 2e-5  # differences only where the two top log-probs are within fp32 rounding of each other
        assert flips < 200


def test_reference_training_step_text_over_stego_modules(cuda_dev):
    """the reference `training_step` source runs unchanged over stego_b200.modules."""
    import lightning_harness as H
    if not H.available():
        pytest.skip("reference sources (baseline/_ref) not present")
    import tempfile
    from stego_b200.config import make_cfg
    fp32_strict()
    B, res = 4, 64
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "dino.pth")
        H.write_random_dino_checkpoint(ck, "vit_small")
        cfg = make_cfg(pretrained_weights=ck)
        batch = H.make_batch(B, res, cuda_dev)
        runs = {}
        for impl in ("stego_b200", "reference"):
            ts = H.load_reference_segmenter(impl)
            torch.manual_seed(0)
            m = ts.LitUnsupervisedSegmenter(27, cfg).to(cuda_dev)
            m.train()
            torch.manual_seed(777)
            losses = []
            for s in range(2):
                losses.append(float(m.training_step(batch, s).detach()))
                m.global_step += 1
            names = [n for n, p in m.named_parameters() if p.requires_grad]
            runs[impl] = dict(losses=losses, logged={k: float(v) for k, v in m.logged.items()},
                              grads={n: dict(m.named_parameters())[n].grad.detach().clone() for n in names
                                     if dict(m.named_parameters())[n].grad is not None},
                              params={n: dict(m.named_parameters())[n].detach().clone() for n in names})
            if impl == "stego_b200":
                assert type(m.net).__module__ == "stego_b200.modules"  # the class the reference text instantiated
    ours, ref = runs["stego_b200"], runs["reference"]
    trained = [n for n in ref["grads"] if n.startswith(("net.cluster", "linear_probe", "cluster_probe"))]
    assert set(trained) <= set(ours["grads"])
    errs = {n: rel(ours["grads"][n], ref["grads"][n]) for n in trained}
    record("dropin_reference_training_step", dict(losses_ours=ours["losses"], losses_reference=ref["losses"],
                                                   logged_ours=ours["logged"], logged_reference=ref["logged"],
                                                   grad_rel=errs))
    print("reference training_step text: ours", ours["losses"], "reference-eager", ref["losses"], "grad rel", errs)
    for a, b in zip(ours["losses"], ref["losses"]):
        assert abs(a - b) < 5e-3 * abs(b), (a, b)  # includes the bf16-operand backbone vs the fp32 eager backbone
    for k in ("loss/linear", "loss/cluster"):
        assert abs(ours["logged"][k] - ref["logged"][k]) < 5e-3 * abs(ref["logged"][k]) + 1e-4
    for n in ("linear_probe.weight", "linear_probe.bias", "cluster_probe.clusters"):
        assert errs[n] < 5e-2, (n, errs[n])
