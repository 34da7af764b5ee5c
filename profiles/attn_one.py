"""One launch of the attention kernel at a bench shape (for ncu): python profiles/attn_one.py [c1|c2|c3]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stego_b200 import ops
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
B, N, heads = {"c1": (64, 785, 6), "c2": (64, 1601, 12), "c3": (32, 3137, 12)}[name]
E = heads * 64
dev = torch.device("cuda:0")
qkv = torch.randn(B * N, 3 * E, device=dev).bfloat16()
out = torch.empty(B * N, E, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, out, B, N, E, heads)
torch.cuda.synchronize()
