"""Small launches of every mbarrier / shared-memory pipelined kernel for compute-sanitizer (racecheck / memcheck):
    compute-sanitizer --tool racecheck python profiles/sanitizer_smoke.py
Shapes are tiny on purpose (the tools slow kernels down by 100x or more)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stego_b200 import corr, ops
from stego_b200.config import make_cfg
dev = torch.device("cuda:0")
torch.manual_seed(0)
a = torch.randn(300, 192, device=dev).bfloat16()
w = torch.randn(160, 192, device=dev).bfloat16()
o = torch.empty(300, 160, device=dev)
ops.gemm(a, w, o, M=300, N=160, K=192, bias=torch.zeros(160, device=dev))
ob = torch.empty(300, 160, device=dev, dtype=torch.bfloat16)
ops.gemm(a, w, ob, M=300, N=160, K=192, act=ops.ACT_GELU)
ops.gemm_batched(a.view(3, 100, 192), w.view(2, 80, 192)[:1].expand(3, 80, 192).contiguous(), torch.empty(3, 100, 80, device=dev))
qkv = torch.randn(2 * 150, 3 * 128, device=dev).bfloat16()
ao = torch.empty(2 * 150, 128, device=dev, dtype=torch.bfloat16)
ops.attention(qkv, ao, 2, 150, 128, 2)
spec = corr.LossSpec(make_cfg())
B, h, E = 2, 8, 128
f = torch.randn(B, E, h, h, device=dev)
c = torch.randn(B, 70, h, h, device=dev, requires_grad=True)
c1 = torch.rand(B, 11, 11, 2, device=dev) * 2 - 1
perms = [torch.tensor([1, 0], device=dev) for _ in range(5)]
losses, _, _, _ = corr.corr_loss(f, f.flip(0), c, c.detach().flip(0).requires_grad_(True), c1, c1.flip(0), perms, spec)
losses.sum().backward()
torch.cuda.synchronize()
print("sanitizer smoke done")
