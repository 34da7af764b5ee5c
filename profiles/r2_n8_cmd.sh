#!/bin/bash
# 8-GPU call (gpurun --gpus 8): the bench at c1 with both exchanges back to back, c2 / c3 with the peer-memory exchange, and
# the shard-and-average parity worker at world = 8
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=$((29500 + RANDOM % 100))
run() {  # config exchange
  port=$((port+1))
  timeout 300 $TR --master-port $port bench.py --gpus 8 --config $1 --exchange $2 --no-cpu-baseline --no-kernel-rooflines --sustain-seconds 2 > gpurun_out/bench_r2_$1_n8_$2.json 2> gpurun_out/bench_$1_n8_$2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_$1_n8_$2.json").read().strip().splitlines()[-1])
    print("$1 $2", round(d["value"], 1), "img/s", round(d["ms_per_step"], 4), "ms/step; e2e", round(d["e2e"]["value"], 1), "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
except Exception as e:
    print("$1 $2 FAILED", e); print(open("gpurun_out/bench_$1_n8_$2.err").read()[-1500:])
PY
}
run c1 p2p
run c1 nccl
run c2 p2p
run c3 p2p
port=$((port+1))
STEGO_TEST_P2P=1 timeout 300 $TR --master-port $port tests/ddp_nccl_worker.py > gpurun_out/ddp_p2p_parity_n8.log 2>&1
grep -h DDP_NCCL_RESULT gpurun_out/ddp_p2p_parity_n8.log | cut -c1-330 | head -3
grep -c '"ok": true' gpurun_out/ddp_p2p_parity_n8.log
