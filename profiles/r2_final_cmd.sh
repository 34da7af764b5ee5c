#!/bin/bash
# final single-GPU call: the whole GPU test suite, smoke(), and the bench lines of the final build
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r2_c1.json 2> gpurun_out/bench_c1.err; tail -c 600 gpurun_out/bench_r2_c1.json; echo
for c in c2 c3 c4; do
  timeout 400 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_$c.json").read().strip().splitlines()[-1])
    print("$c", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 4), "ms/step; e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 3), d.get("eval_pipeline", {}).get("ms_per_frame"))
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/bench_$c.err").read()[-1500:])
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference_arm_c1.json 2> gpurun_out/bench_ref.err; tail -c 700 gpurun_out/bench_r2_reference_arm_c1.json
