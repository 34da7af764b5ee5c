"""Launch the step's main kernels once each at the c1 bench shapes (for `ncu --set full`).
    ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attention_fwd|corr_kernel' \
        -o gpurun_out/prof python profiles/prof_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from stego_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1:] or ["qkv", "proj", "fc1", "fc2", "attn"]
E, heads, N, B2 = 384, 6, 785, 64
M = B2 * N


def gemm(Nn, K, act=0, residual=False):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(Nn, K, device=dev) * K ** -0.5).bfloat16()
    bias = torch.randn(Nn, device=dev)
    o = torch.zeros(M, Nn, device=dev, dtype=torch.float32 if residual else torch.bfloat16)
    for _ in range(2):
        ops.gemm(a, w, o, M=M, N=Nn, K=K, bias=bias, act=act, residual=o if residual else None)


if "qkv" in which:
    gemm(3 * E, E)
if "proj" in which:
    gemm(E, E, residual=True)
if "fc1" in which:
    gemm(4 * E, E, act=1)
if "fc2" in which:
    gemm(E, 4 * E, residual=True)
if "attn" in which:
    qkv = torch.randn(M, 3 * E, device=dev).bfloat16()
    ao = torch.empty(M, E, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.attention(qkv, ao, B2, N, E, heads)
if "corr" in which:
    from stego_b200 import corr
    from stego_b200.config import make_cfg
    spec = corr.LossSpec(make_cfg())
    Bq, h = 32, 28
    feats = torch.randn(Bq, h, h, E, device=dev).bfloat16().permute(0, 3, 1, 2)
    feats_pos = torch.randn(Bq, h, h, E, device=dev).bfloat16().permute(0, 3, 1, 2)
    code = torch.randn(Bq, h, h, 72, device=dev)[..., :70].permute(0, 3, 1, 2).requires_grad_(True)
    code_pos = torch.randn(Bq, h, h, 72, device=dev)[..., :70].permute(0, 3, 1, 2).requires_grad_(True)
    c1 = torch.rand(Bq, 11, 11, 2, device=dev) * 2 - 1
    c2 = torch.rand(Bq, 11, 11, 2, device=dev) * 2 - 1
    perms = torch.stack([torch.randperm(Bq, device=dev) for _ in range(5)])
    for _ in range(2):
        losses, _, _, _ = corr.corr_loss(feats, feats_pos, code, code_pos, c1, c2, perms, spec)
        losses.sum().backward()
if "ce" in which:
    from stego_b200.segmenter import linear_probe_ce
    code = torch.randn(32, 28, 28, 72, device=dev)[..., :70].permute(0, 3, 1, 2)
    w = (torch.randn(27, 70, 1, 1, device=dev) * 0.3).requires_grad_(True)
    b = torch.zeros(27, device=dev, requires_grad=True)
    label = torch.randint(-1, 27, (32, 224, 224), device=dev)
    for _ in range(2):
        linear_probe_ce(code, w, b, label).backward()
if "ln" in which:
    x = torch.randn(M, E, device=dev)
    g = torch.ones(E, device=dev)
    y = torch.empty(M, E, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.layernorm(x, g, g, y)
if "eval" in which:
    from stego_b200.eval import fused_probe_log_probs
    from stego_b200.modules import ClusterLookup
    code = torch.randn(1, 128, 256, 70, device=dev).permute(0, 3, 1, 2)
    lin = torch.nn.Conv2d(70, 27, (1, 1)).to(dev)
    clu = ClusterLookup(70, 27).to(dev)
    for _ in range(2):
        fused_probe_log_probs(code, lin, clu, (1024, 2048), 2.0)
if "knn" in which:
    from stego_b200.knn import knn_topk
    feats = torch.randn(20000, 384, device=dev)
    for _ in range(2):
        knn_topk(feats, 30)
if "crf" in which:
    from stego_b200 import crf
    img = torch.rand(3, 512, 1024, device=dev)
    logp = torch.log_softmax(torch.randn(27, 512, 1024, device=dev), 0)
    crf.mean_field(logp, crf.prepare_image((img - 0.45) / 0.225), 2)
if "corrdense" in which:
    f = torch.nn.functional.normalize(torch.randn(32, 1600, 768, device=dev), dim=2).bfloat16()
    out = torch.empty(32, 1600, 1600, device=dev)
    for _ in range(2):
        ops.gemm_batched(f, f, out)
torch.cuda.synchronize()
print("done")
