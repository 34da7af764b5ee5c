"""Phase isolation of the fused attention kernel (ViT-S/8 224^2, 64 images): STEGO_ATT_DIAG makes it skip the softmax
math (1), the MMAs (2) or the TMA loads (4); results are garbage, only the time matters.
    python profiles/attn_phases.py > gpurun_out/attn_phases.md
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from stego_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
E, heads, N, B2 = 384, 6, 785, 64
M = B2 * N
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
qkv = torch.randn(M, 3 * E, device=dev).bfloat16()
ao = torch.empty(M, E, device=dev, dtype=torch.bfloat16)


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


modes = [("full", 0), ("no KV tiles: launch + prologue + merge + output store", 8 | 4),
         ("no KV tiles, no output store", 8 | 4 | 16), ("no softmax math (TMA + MMA + barrier protocol)", 1),
         ("barrier protocol only", 7),
         ("softmax + barriers (no TMA, no MMA)", 6), ("softmax + MMA (no TMA)", 4), ("softmax + TMA (no MMA)", 2)]
print("| mode | us |\n|---|---|")
for name, d in modes:
    os.environ["STEGO_ATT_DIAG"] = str(d)
    print(f"| {name} | {timeit(lambda: ops.attention(qkv, ao, B2, N, E, heads)):.1f} |")
os.environ["STEGO_ATT_DIAG"] = "0"
