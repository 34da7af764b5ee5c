#!/bin/bash
# last single-GPU call of the round (5 GPU-minutes left): the default bench line of the final build first, then smoke(), then
# as much of the kernel / step parity tests as fits
cd /root/repo
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_r2_c1.json 2> gpurun_out/bench_c1.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/bench_r2_c1.json; echo; tail -3 gpurun_out/bench_c1.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 | tee gpurun_out/r2_smoke_final.log
timeout 600 python -m pytest tests/test_vit_kernels_gpu.py tests/test_step_parity_gpu.py tests/test_modules_gpu.py -q -m gpu -x -k "not fullsize" 2>&1 | tail -4 | tee gpurun_out/r2_tests_final.log
