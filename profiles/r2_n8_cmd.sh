#!/bin/bash
# 8-GPU runs (gpurun --gpus 8): the bench at c1 / c2 / c3 and the in-stream timeline of the data-parallel step
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29500
for c in c1 c2 c3; do
  port=$((port+1))
  timeout 400 $TR --master-port $port bench.py --gpus 8 --config $c --no-cpu-baseline --no-kernel-rooflines --sustain-seconds 2 > gpurun_out/bench_r2_${c}_n8.json 2> gpurun_out/bench_${c}_n8.err
  tail -c 900 gpurun_out/bench_r2_${c}_n8.json; echo
done
timeout 300 $TR --master-port 29511 profiles/step_timeline.py c1 > gpurun_out/r2_timeline_n8.md 2> gpurun_out/timeline_n8.err
grep -A3 "## update" gpurun_out/r2_timeline_n8.md
timeout 300 $TR --master-port 29512 profiles/step_timeline.py c2 > gpurun_out/r2_timeline_n8_c2.md 2> gpurun_out/timeline_n8_c2.err
grep -A3 "## update" gpurun_out/r2_timeline_n8_c2.md
