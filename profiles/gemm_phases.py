"""Phase isolation of the tcgen05 GEMM at the ViT-S shapes: STEGO_GEMM_DIAG makes the kernel skip the epilogue work (1),
the MMAs (2) or the TMA loads (4) — results are garbage, only the time matters.  Shows which pipeline bounds each GEMM.
    python profiles/gemm_phases.py > gpurun_out/gemm_phases.md
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from stego_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
E, N, B2 = 384, 785, 64
M = B2 * N
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


cases = [("qkv      N=1152 K=384  bf16 out", 3 * E, E, dict()),
         ("fc1+gelu N=1536 K=384  bf16 out", 4 * E, E, dict(act=1)),
         ("proj     N=384  K=384  fp32 +=", E, E, dict(residual=True)),
         ("fc2      N=384  K=1536 fp32 +=", E, 4 * E, dict(residual=True))]
modes = [("full", 0), ("no epilogue work (TMA+MMA)", 1), ("MMA only", 5), ("TMA only", 3), ("epilogue only", 6),
         ("TMA + epilogue (no MMA)", 2), ("MMA + epilogue (no TMA)", 4)]
print("| GEMM | " + " | ".join(m for m, _ in modes) + " |")
print("|---|" + "---|" * len(modes))
for name, Nn, K, kw in cases:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(Nn, K, device=dev) * K ** -0.5).bfloat16()
    bias = torch.randn(Nn, device=dev)
    res = kw.pop("residual", False)
    o = torch.zeros(M, Nn, device=dev, dtype=torch.float32 if res else torch.bfloat16)
    row = []
    for _, d in modes:
        os.environ["STEGO_GEMM_DIAG"] = str(d)
        row.append(timeit(lambda: ops.gemm(a, w, o, M=M, N=Nn, K=K, bias=bias, residual=o if res else None, **kw)))
    os.environ["STEGO_GEMM_DIAG"] = "0"
    print(f"| {name} | " + " | ".join(f"{t:.1f}" for t in row) + " |")
