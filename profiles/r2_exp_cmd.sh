#!/bin/bash
# one GPU call: eval-kernel tests + c4 bench, then the attention variants back to back
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eval_probes_gpu.py tests/test_boundary_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --config c4 --no-cpu-baseline > gpurun_out/bench_r2_c4_vec4.json 2> gpurun_out/bench_c4.err; tail -c 1500 gpurun_out/bench_r2_c4_vec4.json
for v in 0_0 1_0 1_1 1_2; do
  echo "== variant $v"
  STEGO_PROFILE_LIB=profiles/_variants/libstego_att_$v.so timeout 300 python profiles/attn_bench.py 2>&1 | grep -E "time|ALL OK|MISMATCH|rel [0-9.e-]+ " | tr '\n' ';'
  echo
done 2>&1 | tee gpurun_out/attn_variants_r2.log
