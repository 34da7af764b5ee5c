"""In-stream timeline of one training step (CUPTI via torch.profiler): every kernel of the step in launch order with
its duration and the idle gap before it.  Unlike the ncu launch list this is warm and NOT serialised, so absolute
times and gaps are meaningful (profiler overhead: a few % on a 6 ms step).
    python profiles/step_timeline.py [c1|c2] > gpurun_out/timeline.md
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
dev = torch.device("cuda:0")
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
print("\n## tail of the step in launch order (after the frozen ViT)\n\n| t us | dur us | gap us | kernel |\n|---|---|---|---|")
last_vit = max(i for i, r in enumerate(rows) if "attention_fwd" in r[3])
for t, d, g, nm in rows[last_vit:]:
    print(f"| {t:.1f} | {d:.1f} | {g:.1f} | `{nm}` |")
