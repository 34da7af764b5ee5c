"""CPU, world_size 2, gloo: the data-parallel plumbing — ONE all-reduce of the flat gradient buffer,
per-rank (local shard) loss semantics, averaged gradients — checked against a single-process
shard-and-average emulation with the oracle."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _shard_grads(rank):
    """Oracle correspondence-loss gradients wrt a tiny 'head' on this rank's shard."""
    import stego_oracle as O
    cfg = O.LossCfg(feature_samples=5, neg_samples=2)
    g = torch.Generator().manual_seed(100 + rank)  # data differs per rank
    B, E, D, h = 2, 16, 8, 6
    feats, feats_pos = torch.randn(B, E, h, h, generator=g), torch.randn(B, E, h, h, generator=g)
    torch.manual_seed(0)  # seed_everything(0) on every rank: identical parameters and RNG streams
    w = torch.nn.Parameter(torch.randn(D, E, 1, 1) * 0.1)
    b = torch.nn.Parameter(torch.zeros(D))
    c1, c2, perms = O.draw_loss_randomness(B, cfg)
    code = torch.nn.functional.conv2d(feats, w, b)
    code_pos = torch.nn.functional.conv2d(feats_pos, w, b)
    out = O.correlation_loss(feats, feats_pos, code, code_pos, c1, c2, perms, cfg)
    O.weighted_correspondence_loss(out, cfg).backward()
    return w, b


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stego_b200.segmenter import FlatParams, allreduce_gradients
    w, b = _shard_grads(rank)
    gw, gb = w.grad.clone(), b.grad.clone()
    flat = FlatParams([[w], [b]], [5e-4, 5e-3])
    flat.grad[:w.numel()].copy_(gw.reshape(-1))
    flat.grad[w.numel():].copy_(gb.reshape(-1))
    assert w.grad.data_ptr() == flat.grad.data_ptr()  # .grad lives inside the flat buffer
    allreduce_gradients(flat)
    ret[rank] = (flat.grad.clone() * flat.grad_scale, flat.grad_scale)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    want = None
    for r in range(world):
        w, b = _shard_grads(r)
        v = torch.cat([w.grad.reshape(-1), b.grad.reshape(-1)])
        want = v if want is None else want + v
    want = want / world
    for r in range(world):
        got, scale = ret[r]
        assert scale == 0.5
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-9)


def test_single_process_is_identity():
    from stego_b200.segmenter import FlatParams, allreduce_gradients
    w = torch.nn.Parameter(torch.randn(3, 4))
    flat = FlatParams([[w]], [1e-3])
    flat.grad.fill_(2.0)
    allreduce_gradients(flat)
    assert flat.grad_scale == 1.0 and torch.all(w.grad == 2.0)
    assert w.data.data_ptr() == flat.param.data_ptr()
