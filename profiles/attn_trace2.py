"""Synchronisation-point timeline of the attention kernel (diagnostic build):
    STEGO_NVCC_DEFS=-DATT_TRACE python -m stego_b200.build -f && python profiles/attn_trace2.py
clock64() stamps of CTA 0 during its first work item, per KV tile j and query tile t:
  MMA warp:      [0+t] before waiting for P^t_j   [2+t] P^t_j seen   [4+t] P V + next S issued
  softmax warp:  [6+t] before waiting for S^t_j   [8+t] S^t_j seen   [14+t] scores in registers   [16+t] max / rescale done
                 [18+t] exponentials + P stores issued   [10+t] P visible (wait::st + fence)   [12+t] arrived on p_full"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stego_b200 import _lib, ops
dev = torch.device("cuda:0")
B, N, heads = 64, 785, 6
E = heads * 64
qkv = torch.randn(B * N, 3 * E, device=dev).bfloat16()
out = torch.empty(B * N, E, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv, out, B, N, E, heads)
buf = torch.zeros(24 * 64, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.stego_attention_set_trace.argtypes = [ctypes.c_void_p]
lib.stego_attention_set_trace(buf.data_ptr())
ops.attention(qkv, out, B, N, E, heads)
torch.cuda.synchronize()
lib.stego_attention_set_trace(None)
t = buf.cpu().view(24, 64)
t0 = int(t[t > 0].min())
names = ["mma wait P0", "mma wait P1", "mma got P0", "mma got P1", "mma issued 0", "mma issued 1", "sm0 wait S", "sm1 wait S",
         "sm0 got S", "sm1 got S", "sm0 P fenced", "sm1 P fenced", "sm0 arrived", "sm1 arrived", "sm0 S in regs", "sm1 S in regs",
         "sm0 max done", "sm1 max done", "sm0 exp done", "sm1 exp done"]
nkv = (N + 95) // 96
print("clock cycles since the first stamp; one column per KV tile j = 0..%d" % (nkv - 1))
for i, n in enumerate(names):
    print(f"{n:14s}", " ".join(f"{int(t[i, j]) - t0:6d}" if t[i, j] > 0 else "     -" for j in range(nkv)))
