#!/bin/bash
# 2-GPU call: data-parallel parity (peer-memory exchange and NCCL fallback), the new loss kernels, bench with both exchanges
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py tests/test_crf_loss_gpu.py -x -q -m gpu 2>&1 | tail -15
grep -h DDP_NCCL_RESULT gpurun_out/ddp_p2p_parity.log gpurun_out/ddp_nccl_parity.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for ex in p2p nccl p2p nccl; do
  port=$((29700 + RANDOM % 200))
  timeout 300 $TR --master-port $port bench.py --gpus 2 --config c1 --exchange $ex --no-cpu-baseline --no-kernel-rooflines --sustain-seconds 2 > gpurun_out/bench_r2_c1_n2_$ex.json 2> gpurun_out/bench_n2_$ex.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_c1_n2_$ex.json").read().strip().splitlines()[-1])
    print("$ex", round(d["value"], 1), "img/s", round(d["ms_per_step"], 4), "ms/step; e2e", round(d["e2e"]["value"], 1), d["config"]["exchange"][:40])
except Exception as e:
    print("$ex FAILED", e); print(open("gpurun_out/bench_n2_$ex.err").read()[-1500:])
PY
done
