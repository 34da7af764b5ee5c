# Round-2 evidence captures (run on the GPU box through gpurun; summaries are made on the box, the .ncu-rep is too big to
# bring back with sources: 98 MB > gpurun's 64 MiB limit)
set -x
ncu --set full --clock-control none -k regex:'gemm_bf16|attention_fwd|corr_kernel|sample_norm|eval_probe|knn_topk|layernorm_kernel|crf_splat|crf_update|crf_blur' -c 44 -o /tmp/r2_full --force-overwrite python profiles/prof_kernels.py qkv proj fc1 fc2 attn corr ln eval knn crf > gpurun_out/r2_full.log 2>&1
tail -2 gpurun_out/r2_full.log
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
python profiles/summarize.py full /tmp/r2_full.ncu-rep > gpurun_out/r2_ncu_full.md
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-rooflines --sustain-seconds 0 > gpurun_out/r2_launches.log 2>&1
python profiles/summarize.py launches gpurun_out/r2_launches.csv > gpurun_out/r2_launches.md
compute-sanitizer --tool racecheck --print-limit 10 python profiles/sanitizer_smoke.py > gpurun_out/r2_racecheck.log 2>&1; tail -3 gpurun_out/r2_racecheck.log
compute-sanitizer --tool memcheck --print-limit 10 python profiles/sanitizer_smoke.py > gpurun_out/r2_memcheck.log 2>&1; tail -3 gpurun_out/r2_memcheck.log
ls -la gpurun_out/ | head -20; du -sh gpurun_out
