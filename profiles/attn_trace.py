"""Per-CTA pipeline timeline of the fused attention kernel (diagnostic build only).

    STEGO_NVCC_DEFS=-DSTEGO_ATT_TRACE python stego_b200/build.py -f      # build libstego_b200.so with the trace points
    python profiles/attn_trace.py > gpurun_out/attn_trace.md
    python stego_b200/build.py -f                                        # back to the shipped build

Lane 0 of every warp of 8 CTAs (every `EVERY`-th of the grid) stamps (globaltimer, event) pairs:
  1 CTA start, 2 prologue done, 140+j K/V tile j requested (TMA warp), 10+j S_j issued, 40+j P_j V_j issued (MMA warp),
  70+j S_j visible to a softmax warp, 100+j P_j stored (p_full signalled), 130 KV loop done, 131 role finished.
The output lists, per traced CTA, the events of all ten warps in time order (ns since the CTA's first event) — enough to see which hand-off a KV tile spends its time in.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from stego_b200 import _lib, ops  # noqa: E402

EVENTS, SLOTS, WARPS, EVERY = 256, 8, 10, 331
dev = torch.device("cuda:0")
lib = _lib.load()
if not hasattr(lib, "stego_attention_set_trace"):
    sys.exit("libstego_b200.so was not built with -DSTEGO_ATT_TRACE (see the docstring)")
E, heads, N, B2 = 384, 6, 785, 64
M = B2 * N
qkv = torch.randn(M, 3 * E, device=dev).bfloat16()
ao = torch.empty(M, E, device=dev, dtype=torch.bfloat16)
ops.attention(qkv, ao, B2, N, E, heads)  # warm-up without tracing
buf = torch.zeros(SLOTS * WARPS * EVENTS * 2, dtype=torch.int64, device=dev)
lib.stego_attention_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.stego_attention_set_trace(ctypes.c_void_p(buf.data_ptr()), EVERY)
ops.attention(qkv, ao, B2, N, E, heads)
torch.cuda.synchronize()
lib.stego_attention_set_trace(None, 0)
t = buf.cpu().view(SLOTS, WARPS, EVENTS, 2)


def name(ev):
    if ev in (1, 2, 130, 131):
        return {1: "cta start", 2: "prologue done", 130: "kv loop done", 131: "role finished"}[ev]
    for base, what in ((140, "kv req"), (100, "P stored"), (70, "S seen"), (40, "PV issued"), (10, "S issued")):
        if ev >= base:
            return f"{what} {ev - base}"
    return str(ev)


for slot in range(SLOTS):
    rows = []
    for w in range(WARPS):  # 0 TMA warp, 1 MMA warp, 2..5 softmax warpgroup 0, 6..9 softmax warpgroup 1
        for i in range(EVENTS):
            ts, word = int(t[slot, w, i, 0]), int(t[slot, w, i, 1])
            if ts == 0:
                break
            rows.append((ts, w, word & 0xFFFFFFFF, word >> 32))
    if not rows:
        continue
    rows.sort()
    t0 = rows[0][0]
    print(f"\n## CTA {rows[0][3]} (slot {slot}): {rows[-1][0] - t0} ns\n\n| ns | warp | event |\n|---|---|---|")
    for ts, w, ev, _ in rows:
        print(f"| {ts - t0} | {w} | {name(ev)} |")
