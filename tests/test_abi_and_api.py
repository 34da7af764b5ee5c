"""CPU: the C-ABI library loads and exports every symbol include/stego_b200.h declares; argument
validation works without a GPU; the Python API mirrors the reference's `modules.py` surface and fails
loudly (no CPU fallback) when asked to compute off-device."""
import ctypes
import inspect
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from stego_b200 import _lib
    protos = _lib.header_prototypes()
    assert len(protos) >= 19
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/stego_b200.h but not exported"
    assert _lib.load().stego_version() >= 100


def test_bad_arguments_are_rejected_before_any_cuda_call():
    from stego_b200 import _lib
    lib = _lib.load()
    rc = lib.stego_gemm_bf16(0, 8, 0, 0, 8, 0, 128, 128, 64, 0, 128, 0, 0, 0, 0, 0, 0, 1, 0, 0)
    assert rc == -1 and "null pointer" in _lib.last_error()
    rc = lib.stego_attention_fwd(16, 16, 1, 10, 100, 2, 0)
    assert rc == -1 and "head_dim" in _lib.last_error()
    rc = lib.stego_layernorm_bf16(16, 16, 16, 16, 4, 100, 1e-6, 0, 0)
    assert rc == -2 and "unsupported" in _lib.last_error()


def test_modules_surface_matches_reference():
    from stego_b200 import modules as M
    for name in ["LambdaLayer", "DinoFeaturizer", "ResizeAndClassify", "ClusterLookup", "FeaturePyramidNet",
                 "DoubleConv", "norm", "average_norm", "tensor_correlation", "sample", "super_perm",
                 "sample_nonzero_locations", "ContrastiveCorrelationLoss", "Decoder", "NetWithActivations",
                 "ContrastiveCRFLoss"]:
        assert hasattr(M, name), name
    assert list(inspect.signature(M.DinoFeaturizer.forward).parameters) == ["self", "img", "n", "return_class_feat"]
    assert list(inspect.signature(M.ContrastiveCorrelationLoss.forward).parameters) == [
        "self", "orig_feats", "orig_feats_pos", "orig_salience", "orig_salience_pos", "orig_code", "orig_code_pos"]
    assert list(inspect.signature(M.ClusterLookup.forward).parameters) == ["self", "x", "alpha", "log_probs"]
    assert list(inspect.signature(M.FeaturePyramidNet.__init__).parameters) == [
        "self", "granularity", "cut_model", "dim", "continuous"]


def test_state_dict_keys_match_reference_checkpoints():
    from stego_b200.config import make_cfg
    from stego_b200.segmenter import LitUnsupervisedSegmenter
    torch.manual_seed(0)
    m = LitUnsupervisedSegmenter(27, make_cfg(random_backbone_init=True))
    keys = set(m.state_dict().keys())
    vit = {k for k in keys if k.startswith("net.model.")}
    assert len(vit) == 150  # 150 ViT keys
    for k in ["net.model.cls_token", "net.model.pos_embed", "net.model.patch_embed.proj.weight",
              "net.model.blocks.11.attn.qkv.bias", "net.model.blocks.0.mlp.fc2.weight", "net.model.norm.bias",
              "net.cluster1.0.weight", "net.cluster1.0.bias", "net.cluster2.0.weight", "net.cluster2.2.bias",
              "train_cluster_probe.clusters", "cluster_probe.clusters", "linear_probe.weight", "linear_probe.bias",
              "decoder.weight", "decoder.bias"]:
        assert k in keys, k
    assert m.net.n_feats == 384 and m.net.model.pos_embed.shape == (1, 785, 384)
    n_train = sum(p.numel() for n, p in m.named_parameters() if p.requires_grad and n.startswith("net."))
    assert n_train == 201740  # cluster1 + cluster2 for ViT-S
    assert sum(p.numel() for p in m.net.model.parameters()) == 21670272


def test_hot_path_refuses_cpu_tensors():
    from stego_b200 import modules as M
    from stego_b200.config import make_cfg
    cfg = make_cfg(random_backbone_init=True)
    loss = M.ContrastiveCorrelationLoss(cfg)
    f, c = torch.randn(2, 384, 4, 4), torch.randn(2, 70, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        loss(f, f, None, None, c, c)
    with pytest.raises(RuntimeError, match="CUDA"):
        M.ClusterLookup(70, 27)(c, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        M.tensor_correlation(f, f)
    with pytest.raises(ValueError):
        M.DinoFeaturizer(70, make_cfg(model_type="vit_huge"))


def test_rng_helpers_follow_reference_stream():
    """super_perm / coordinate draws consume the torch generator exactly like the reference code."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import stego_oracle as O
    from stego_b200 import modules as M
    from stego_b200.config import make_cfg
    cfg = make_cfg()
    lossfn = M.ContrastiveCorrelationLoss(cfg)
    torch.manual_seed(99)
    c1, c2 = lossfn.draw_coords(torch.zeros(2, 384, 28, 28), None, None)
    perms = [M.super_perm(2, torch.device("cpu")) for _ in range(5)]
    torch.manual_seed(99)
    w1, w2, wp = O.draw_loss_randomness(2, O.LossCfg())
    assert torch.equal(c1, w1) and torch.equal(c2, w2) and all(torch.equal(a, b) for a, b in zip(perms, wp))
    g = torch.load(os.path.join(ROOT, "tests", "golden", "super_perm.pt"))
    for size, want in zip((1, 2, 5, 16, 32), g["draws"]):
        torch.manual_seed(1000 + size)
        assert torch.equal(torch.stack([M.super_perm(size, torch.device("cpu")) for _ in range(3)]), want)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from stego_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.load()


def test_unsupervised_metrics_match_reference_arithmetic():
    """stego_b200.eval.UnsupervisedMetrics vs a direct restatement of src/utils.py:219-274 (bincount confusion, Hungarian,
    mIoU / accuracy) on random predictions, with and without extra clusters."""
    import numpy as np
    import torch
    from scipy.optimize import linear_sum_assignment
    from stego_b200.eval import UnsupervisedMetrics
    g = torch.Generator().manual_seed(0)
    n = 6
    target = torch.randint(-1, n + 1, (4, 17, 19), generator=g)
    preds = (target.clamp(0, n - 1) + (torch.rand(4, 17, 19, generator=g) < 0.3).long() * 2) % n
    perm = torch.randperm(n, generator=g)
    m = UnsupervisedMetrics("test/", n, 0, True)
    m.update(perm[preds], target)
    m.update(perm[preds], target)
    out = m.compute()
    mask = (target >= 0) & (target < n)
    stats = torch.zeros(n, n, dtype=torch.int64)
    for p_, a_ in zip(perm[preds][mask].tolist(), target[mask].tolist()):
        stats[p_, a_] += 2
    assert torch.equal(m.stats, stats)
    rows, cols = linear_sum_assignment(stats, maximize=True)
    hist = stats[np.argsort(cols), :].double()
    tp = torch.diag(hist)
    iou = tp / (hist.sum(0) + hist.sum(1) - tp)
    assert abs(out["test/mIoU"] - 100 * iou[~torch.isnan(iou)].mean().item()) < 1e-9
    assert abs(out["test/Accuracy"] - 100 * (tp.sum() / hist.sum()).item()) < 1e-9
    lin = UnsupervisedMetrics("lin/", n, 0, False)
    lin.update(preds, target)
    assert 0 < lin.compute()["lin/Accuracy"] <= 100
    ex = UnsupervisedMetrics("ex/", n, 2, True)
    ex.update(preds, target)
    assert ex.stats.shape == (n + 2, n) and "ex/mIoU" in ex.compute()


def test_knn_file_helpers_round_trip(tmp_path):
    """src/precompute_knns.py:66-67,94 / src/data.py:503-511: file name pattern and the `nns` key of the .npz."""
    import numpy as np
    from stego_b200.knn import nns_file_name, save_nns
    name = nns_file_name("vit_small", "cocostuff27", "train", "five", 224)
    assert name == "nns_vit_small_cocostuff27_train_five_224.npz"
    nns = torch.arange(60, dtype=torch.int32).reshape(2, 30)
    save_nns(str(tmp_path / name), nns)
    back = np.load(tmp_path / name)
    assert list(back.keys()) == ["nns"] and back["nns"].dtype == np.int64
    assert (back["nns"] == nns.numpy()).all()


def test_crf_host_helpers():
    """Host-side pieces of the dense-CRF drop-in that need no GPU: lattice key packing (bits = 60 // d per coordinate,
    negative coordinates included) and the image preparation of src/crf.py:23 against the oracle's."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import crf_oracle as CO
    from stego_b200 import crf
    g = torch.Generator().manual_seed(0)
    for d in (2, 5):
        bits = 60 // d
        lim = 1 << (bits - 2)
        coords = torch.randint(-lim, lim, (1000, d), generator=g)
        keys = crf._pack(coords, d, bits)
        assert torch.equal(crf._unpack(keys, d, bits), coords)
        assert keys.unique().numel() == torch.unique(coords, dim=0).shape[0]
        # the packed order is the lexicographic order of the coordinates (what searchsorted relies on)
        order = torch.argsort(keys)
        srt = coords[order]
        as_tuples = [tuple(r.tolist()) for r in srt]
        assert as_tuples == sorted(as_tuples)
    img = torch.randn(3, 17, 23, generator=g)
    got = crf.prepare_image(img)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (17, 23, 3)
    assert (got.numpy() == CO.prepare_image(img)).all()


def test_refused_without_cuda_for_new_entry_points():
    """The round-2 entry points keep the no-CPU-fallback rule."""
    from stego_b200 import crf
    from stego_b200.modules import ContrastiveCRFLoss, pixel_cosine
    with pytest.raises(RuntimeError):
        crf.dense_crf(torch.randn(3, 8, 8), torch.randn(4, 8, 8))
    with pytest.raises(RuntimeError):
        ContrastiveCRFLoss(16, .5, .15, .05, 10.0, 3.0, 0.0)(torch.rand(1, 3, 8, 8), torch.rand(1, 5, 8, 8))
    with pytest.raises(RuntimeError):
        pixel_cosine(torch.randn(1, 4, 3, 3), torch.randn(1, 4, 3, 3))
    # caller-supplied sample positions are validated before they reach the kernels
    crf_loss = ContrastiveCRFLoss(4, .5, .15, .05, 10.0, 3.0, 0.0)
    with pytest.raises(ValueError):
        crf_loss.forward_with_coords(torch.rand(1, 3, 8, 8), torch.rand(1, 5, 8, 8), torch.tensor([[0, 1, 2, 8], [0, 1, 2, 3]]))
    with pytest.raises(ValueError):
        crf_loss.forward_with_coords(torch.rand(1, 3, 8, 8), torch.rand(1, 5, 8, 8), torch.zeros(2, 4, dtype=torch.int32))


def test_segmenter_carries_reference_metric_objects():
    """train_segmentation.py:80-88: cluster / linear metric objects (validation and final) exist under the reference's names,
    add nothing to the state dict (torchmetrics states are non-persistent) and follow the module across `.to()`."""
    from stego_b200.config import make_cfg
    from stego_b200.segmenter import LitUnsupervisedSegmenter
    m = LitUnsupervisedSegmenter(27, make_cfg(random_backbone_init=True, extra_clusters=3))
    assert m.cluster_metrics.prefix == "test/cluster/" and m.test_linear_metrics.prefix == "final/linear/"
    assert tuple(m.test_cluster_metrics.stats.shape) == (30, 27) and tuple(m.linear_metrics.stats.shape) == (27, 27)
    assert not [k for k in m.state_dict() if "metrics" in k]
    m = m.to("meta")
    assert m.test_cluster_metrics.stats.device.type == "meta" and m.test_cluster_metrics.stats.dtype == torch.int64
