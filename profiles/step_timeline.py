"""In-stream timeline of one training step (CUPTI via torch.profiler): every kernel of the step in launch order with
its duration and the idle gap before it.  Unlike the ncu launch list this is warm and NOT serialised, so absolute
times and gaps are meaningful (profiler overhead: a few % on a 6 ms step).
    python profiles/step_timeline.py [c1|c2] > gpurun_out/timeline.md
Under torchrun (WORLD_SIZE > 1) every rank runs the data-parallel step and rank 0 prints; the report then also says how
much of the update (exchange + Adam, side stream) is EXPOSED, i.e. not overlapped by any other kernel of the next step's
frozen backbone.  CAVEAT (measured at 8 ranks, round 2): the profiler start-up skews the ranks by tens of milliseconds, a
collective kernel that waits for its peers inside the kernel (NCCL) then shows that skew as kernel time, and the EXPOSED
figure is meaningless — use it at N = 1 / 2 only, and judge N = 8 by the step times of bench.py with both exchanges.
"""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from stego_b200.config import make_cfg  # noqa: E402
from stego_b200.segmenter import LitUnsupervisedSegmenter  # noqa: E402

cfgs = {"c1": ("vit_small", 224, 32), "c2": ("vit_base", 320, 32)}
model_type, res, B = cfgs[sys.argv[1] if len(sys.argv) > 1 else "c1"]
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = LitUnsupervisedSegmenter(27, make_cfg(model_type=model_type, res=res, batch_size=B, random_backbone_init=True)).to(dev)
model.train()
model.configure_optimizers()
batch = dict(img=torch.randn(B, 3, res, res, device=dev), img_pos=torch.randn(B, 3, res, res, device=dev),
             label=torch.randint(-1, 27, (B, res, res), device=dev))
for i in range(5):
    model.training_step(batch, i)
torch.cuda.synchronize()
NSTEP = 4
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(NSTEP):
        model.training_step(batch, i)
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank != 0:
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into steps at the patchify kernel (first kernel of the ViT graph)
starts = [i for i, e in enumerate(evs) if "patchify" in e.name]
assert len(starts) >= NSTEP, (len(starts), len(evs))
lo, hi = starts[-2], starts[-1]  # the second-to-last step, complete
step = evs[lo:hi]
t0 = step[0].time_range.start
span = evs[hi].time_range.start - t0
busy = sum(e.time_range.end - e.time_range.start for e in step)
print(f"# step timeline ({model_type} {res}^2 B={B}): {len(step)} device activities, span {span:.1f} us, "
      f"busy {busy:.1f} us, idle {span - busy:.1f} us")
agg = OrderedDict()
prev_end = t0
rows = []
for e in step:
    d = e.time_range.end - e.time_range.start
    gap = e.time_range.start - prev_end
    prev_end = max(prev_end, e.time_range.end)
    nm = e.name.replace("void ", "")[:70]
    a = agg.setdefault(nm, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d
    a[2] += max(gap, 0.0)
    rows.append((e.time_range.start - t0, d, gap, nm))
print("\n## by kernel (sum over the step)\n\n| kernel | launches | busy us | gap-before us |\n|---|---|---|---|")
for nm, (n, d, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{nm}` | {n} | {d:.1f} | {g:.1f} |")
# exposed update time: intervals of the NCCL / Adam kernels not covered by any other kernel
upd = [(e.time_range.start, e.time_range.end) for e in evs if ("nccl" in e.name.lower() or "adam_kernel" in e.name)]
oth = sorted((e.time_range.start, e.time_range.end) for e in evs if not ("nccl" in e.name.lower() or "adam_kernel" in e.name))
merged = []
for a, b in oth:
    if merged and a <= merged[-1][1]:
        merged[-1][1] = max(merged[-1][1], b)
    else:
        merged.append([a, b])
def covered(a, b):
    c = 0.0
    for x, y in merged:
        lo, hi = max(a, x), min(b, y)
        if hi > lo:
            c += hi - lo
    return c
tot_upd = sum(b - a for a, b in upd)
exp_upd = sum((b - a) - covered(a, b) for a, b in upd)
nsteps_seen = max(1, len(starts) - 1)
print(f"\n## update (all-reduce + Adam) on the side stream, {world} GPU(s): {tot_upd / nsteps_seen:.1f} us of kernel time per step, "
      f"{exp_upd / nsteps_seen:.1f} us of it EXPOSED (no other kernel running); the rest runs under the next step's backbone")
for e in evs:
    if "nccl" in e.name.lower():
        print(f"  NCCL kernel: {e.name[:60]}  {e.time_range.end - e.time_range.start:.1f} us")
        break
print("\n## tail of the step in launch order (after the frozen ViT)\n\n| t us | dur us | gap us | kernel |\n|---|---|---|---|")
last_vit = max(i for i, r in enumerate(rows) if "attention_fwd" in r[3])
for t, d, g, nm in rows[last_vit:]:
    print(f"| {t:.1f} | {d:.1f} | {g:.1f} | `{nm}` |")
