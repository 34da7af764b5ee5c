"""LayerNorm / patchify / fused attention kernels against plain PyTorch fp32 references."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(x, y):
    return ((x.float() - y.float()).norm() / y.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("E", [384, 768])
def test_layernorm(cuda_dev, E):
    from stego_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(1000, E, device=cuda_dev) * 3 + 0.5
    g, b = torch.randn(E, device=cuda_dev), torch.randn(E, device=cuda_dev)
    out = torch.empty(1000, E, device=cuda_dev, dtype=torch.bfloat16)
    ops.layernorm(x, g, b, out)
    want = torch.nn.functional.layer_norm(x, (E,), g, b, eps=1e-6)
    assert ((out.float() - want).abs() / (want.abs() + 1.0)).max().item() < 8e-3  # bf16 output rounding
    assert _rel(out, want) < 4e-3


def test_layernorm_drop_cls(cuda_dev):
    from stego_b200 import ops
    torch.manual_seed(1)
    B, ntok, E = 3, 50, 384
    x = torch.randn(B * ntok, E, device=cuda_dev)
    g, b = torch.randn(E, device=cuda_dev), torch.randn(E, device=cuda_dev)
    out = torch.zeros(B * (ntok - 1), E, device=cuda_dev, dtype=torch.bfloat16)
    ops.layernorm(x, g, b, out, drop_cls_ntok=ntok)
    want = torch.nn.functional.layer_norm(x, (E,), g, b, eps=1e-6).view(B, ntok, E)[:, 1:].reshape(-1, E)
    assert _rel(out, want) < 4e-3


def test_patchify_matches_conv(cuda_dev):
    from stego_b200 import ops
    torch.manual_seed(2)
    img = torch.randn(2, 3, 32, 48, device=cuda_dev)
    w = torch.randn(16, 3, 8, 8, device=cuda_dev)
    rows = ops.patchify(img, 8)
    torch.backends.cuda.matmul.allow_tf32 = False
    got = rows.float() @ w.reshape(16, -1).t()
    torch.backends.cudnn.allow_tf32 = False
    want = torch.nn.functional.conv2d(img.bfloat16().float(), w, stride=8).flatten(2).transpose(1, 2).reshape(-1, 16)
    assert _rel(got, want) < 1e-5


@pytest.mark.parametrize("B,N,heads", [(1, 128, 1), (2, 785, 6), (1, 1601, 12), (3, 50, 6), (1, 257, 2)])
def test_attention(cuda_dev, B, N, heads):
    from stego_b200 import ops
    torch.manual_seed(3)
    E = heads * 64
    qkv = (torch.randn(B * N, 3 * E, device=cuda_dev) * 1.5).bfloat16()
    out = torch.full((B * N, E), float("nan"), device=cuda_dev, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, N, E, heads)
    q, k, v = qkv.float().view(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    attn = ((q @ k.transpose(-2, -1)) * 0.125).softmax(-1)
    want = (attn @ v).transpose(1, 2).reshape(B * N, E)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, want) < 1e-2
