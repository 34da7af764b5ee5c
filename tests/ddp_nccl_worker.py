"""Worker of tests/test_ddp_nccl_gpu.py (launched by torchrun, one rank per GPU, NCCL).

Checks the data-parallel CUDA step against shard-and-average (reference semantics:
Lightning-DDP, src/train_segmentation.py:476 — per-rank local statistics, mean of the per-rank gradients):

  1. BEFORE the process group exists, every rank runs the single-GPU hand-scheduled step once per shard (all W shards,
     sequentially, fresh model each, identical seeds) and keeps each shard's gradients  -> mean over shards.
  2. Then the process group is initialised and the SAME step runs data-parallel: rank r gets shard r, the exchange of the
     flat gradient buffer happens inside training_step (side stream) — by default the fused peer-memory all-reduce + Adam
     kernel (csrc/p2p_update.cu), with STEGO_TEST_P2P=0 one NCCL all-reduce + three Adam launches — Adam consumes grad/W.
  3. Compared on every rank: all-reduced gradient / W == mean of the shard gradients; parameters after the update ==
     torch-Adam arithmetic on that mean; parameters identical across ranks (all-gather); two more steps keep the ranks
     bit-identical (the replicas never drift).
Prints one JSON line per rank; exit code 0 only if every check passed.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    from _parity_util import NAMES, grads_of, lr_of, make_batch, make_model, params_of, rel
    import stego_oracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    arch, res, B = "vit_small", 96, 4
    shards = [make_batch(B, res, dev, seed=100 + r) for r in range(world)]

    # ---- 1. shard-and-average emulation, no process group
    shard_grads, p0 = [], None
    for r in range(world):
        model, _ = make_model(arch, dev, fused=True, seed=0)
        p0 = params_of(model)
        torch.manual_seed(777)  # seed_everything on every rank: identical RNG streams (train_segmentation.py:403)
        model.training_step(shards[r], 0)
        shard_grads.append(grads_of(model))
        del model
    mean_g = {k: sum(g[k] for g in shard_grads) / world for k in NAMES}

    # ---- 2. the data-parallel step
    dist.init_process_group("nccl", device_id=dev)
    want_p2p = os.environ.get("STEGO_TEST_P2P", "1") == "1"
    model, _ = make_model(arch, dev, fused=True, seed=0, p2p_update=want_p2p)
    # the exchange under test: the fused peer-memory all-reduce + Adam kernel (default) or the NCCL fallback
    exchange = "p2p" if getattr(model, "_peer", None) is not None else "nccl"
    torch.manual_seed(777)
    loss = model.training_step(shards[rank], 0)
    g = grads_of(model)  # flat buffer after the SUM all-reduce
    p1 = params_of(model)
    res_ = dict(rank=rank, world=world, exchange=exchange, loss=float(loss), grad_scale=model._flat.grad_scale)
    ok = abs(model._flat.grad_scale - 1.0 / world) < 1e-12
    ok &= exchange == ("p2p" if want_p2p else "nccl")
    worst_g, worst_p = 0.0, 0.0
    for k in NAMES:
        e = rel(g[k] / world, mean_g[k])
        worst_g = max(worst_g, e)
        ok &= e < 1e-5
        # Adam's first step is sign-like (g / (|g| + eps)): elements with |g| ~ 1e-8 amplify the 2e-6 difference between
        # the all-reduced gradient and the emulated mean, so the update is checked exactly against the all-reduced
        # gradient (arithmetic + 1/world scale) and loosely against the emulated mean
        want = p0[k].clone()
        O.adam_step(want, g[k] / world, torch.zeros_like(want), torch.zeros_like(want), 1, lr_of(k))
        ok &= rel(p1[k] - p0[k], want - p0[k]) < 1e-4  # fp32 Adam arithmetic: rsqrt / division order differs from the oracle
        want = p0[k].clone()
        O.adam_step(want, mean_g[k], torch.zeros_like(want), torch.zeros_like(want), 1, lr_of(k))
        e = rel(p1[k] - p0[k], want - p0[k])
        worst_p = max(worst_p, e)
        ok &= e < 2e-3
    # the shards really differ (otherwise the test proves nothing)
    ok &= rel(shard_grads[0]["linear_probe.weight"], shard_grads[-1]["linear_probe.weight"]) > 1e-2
    # ---- 3. replicas stay identical
    for s in range(1, 3):
        model.training_step(shards[rank], s)
    flat = torch.cat([v.reshape(-1) for v in params_of(model).values()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(gathered[0], t) for t in gathered)
    ok &= identical
    res_.update(grad_rel_vs_shard_mean=worst_g, param_delta_rel_vs_adam_on_mean=worst_p,
                replicas_bit_identical_after_3_steps=identical, ok=bool(ok))
    model.check_update_health()  # raises if a rank missed the peer-memory rendezvous
    print("DDP_NCCL_RESULT " + json.dumps(res_), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
