"""Host side of the peer-memory parameter update (csrc/p2p_update.cu): the gradient all-reduce of the data-parallel step
(Lightning-DDP's all-reduce behind `manual_backward`, src/train_segmentation.py:227,476) fused into the Adam kernel, with
the gradients read straight from the other ranks' HBM over NVLink.

One process per GPU, all on one node.  `torch.distributed` is used once, at construction, to exchange the 64-byte CUDA IPC
handles of the ranks' peer-visible blocks (and to agree whether every rank could map every block); the per-step exchange is
three launches of this library's kernels on the update stream and involves neither NCCL nor the host.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib

# A peer that does not publish its gradient within this time sets the status flag (RuntimeError at the next health check)
# instead of hanging the GPU.  Generous on purpose: the first steps of a run (graph capture, lazy module loads) can skew the
# ranks by seconds, and the waiting CTA is 32 sleeping threads.
TIMEOUT_MS = 60_000


class PeerUpdate:
    """Owns this rank's export block and the mappings of the other ranks' blocks."""

    def __init__(self, flat, group=None):
        import torch.distributed as dist
        self.flat = flat
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("stego_b200.p2p: at most 16 ranks (one node)")
        lib = _lib.load()
        n = flat.grad.numel()
        self.n = n
        self.n_pad = (n + 3) // 4 * 4
        flags_off = 2 * self.n_pad * 4
        total = flags_off + 256
        ptr = torch.zeros(1, dtype=torch.int64)
        handle = torch.zeros(64, dtype=torch.uint8)
        self.local_ptr = 0
        self._opened: List[int] = []
        ok, err = 1, ""
        # every rank takes part in both collectives below whatever happens locally (a rank that raised early would leave
        # the others hanging in them)
        if lib.stego_p2p_alloc(total, ptr.data_ptr(), handle.data_ptr()) != 0:
            ok, err = 0, _lib.last_error()
        else:
            self.local_ptr = int(ptr[0])
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.numpy().tobytes()) if ok else b"", group=group)
        base = [0] * self.world
        if ok and all(len(h) == 64 for h in handles):
            for r in range(self.world):
                if r == self.rank:
                    base[r] = self.local_ptr
                    continue
                h = torch.frombuffer(bytearray(handles[r]), dtype=torch.uint8)
                if lib.stego_p2p_open(h.data_ptr(), ptr.data_ptr()) != 0:
                    ok, err = 0, _lib.last_error()
                    break
                base[r] = int(ptr[0])
                self._opened.append(base[r])
        else:
            ok = 0
        agree = torch.tensor([ok], dtype=torch.int32, device=flat.grad.device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)
        if int(agree.item()) != 1:
            self.close()
            raise RuntimeError(f"stego_b200.p2p: peer mapping failed on some rank ({err or 'another rank'})")
        self.export = [[b + s * self.n_pad * 4 for b in base] for s in (0, 1)]     # [slot][rank] addresses
        self.export_addr = [torch.tensor(e, dtype=torch.int64) for e in self.export]
        self.flags_addr = torch.tensor([b + flags_off for b in base], dtype=torch.int64)
        self.status = torch.zeros(1, dtype=torch.int32, device=flat.grad.device)
        self.epoch = 0
        self._desc = torch.zeros(4 * 7, dtype=torch.float64)

    def step(self, optimizers: Sequence) -> None:
        """All-reduce (sum) of the flat gradient + one Adam step of every optimiser group, on the current stream."""
        lib = _lib.load()
        flat = self.flat
        self.epoch += 1
        slot = self.epoch & 1
        _lib.check(lib.stego_p2p_publish(_lib.ptr(flat.grad), self.n, self.export[slot][self.rank], self.flags_addr.data_ptr(),
                                         self.rank, self.world, self.epoch, _lib.ptr(self.status), TIMEOUT_MS, _lib.stream()),
                   "stego_p2p_publish")
        for k, opt in enumerate(optimizers):
            opt.steps += 1
            g, pg = opt.group, opt.param_groups[0]
            self._desc[7 * k:7 * k + 7] = torch.tensor([g.start, g.numel, pg["lr"], pg["betas"][0], pg["betas"][1], pg["eps"],
                                                        opt.steps], dtype=torch.float64)
        flat.grad_scale = 1.0 / self.world
        _lib.check(lib.stego_p2p_adam(self.export_addr[slot].data_ptr(), self.world, _lib.ptr(flat.param), _lib.ptr(flat.grad),
                                      _lib.ptr(flat.exp_avg), _lib.ptr(flat.exp_avg_sq), self.n, self._desc.data_ptr(),
                                      len(optimizers), flat.grad_scale, _lib.stream()), "stego_p2p_adam")

    def check(self) -> None:
        """Host-side check of the time-out flag (synchronises: called from flush(), not from the step)."""
        if int(self.status.item()) != 0:
            raise RuntimeError("stego_b200.p2p: a rank did not publish its gradient in time; the replicas have diverged")

    def close(self) -> None:
        lib = _lib.load()
        for p in self._opened:
            lib.stego_p2p_close(p)
        self._opened = []
        if getattr(self, "local_ptr", 0):
            torch.cuda.synchronize()
            lib.stego_p2p_free(self.local_ptr)
            self.local_ptr = 0
