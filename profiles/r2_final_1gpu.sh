#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, bench line, launch lists (evaluate + training), ncu --set full of the
# training attention kernels and the TF32 GEMM.  Every profiled command first runs to completion without ncu.
set -x
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -6 gpurun_out/r2_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; tail -3 gpurun_out/r2_pytest_final.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err
python profiles/bench_brief.py gpurun_out/r2_bench_1gpu_final.json
timeout 300 python profiles/eval_ncu.py > gpurun_out/eval_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_eval_final.csv python profiles/eval_ncu.py > /dev/null 2>&1
timeout 300 python profiles/train_step_probe.py 4 > gpurun_out/train_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_train_final.csv python profiles/train_step_probe.py 4 > /dev/null 2>&1
if [ -z "$SKIP_FULL" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_kernel|attn_bwd_kernel" -c 2 -o gpurun_out/r2_prof_attn_mma -f python profiles/train_step_probe.py 1 > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"tc_gemm_nt_kernel" -c 12 -o gpurun_out/r2_prof_train_gemm -f python profiles/train_step_probe.py 1 > gpurun_out/ncu_gemm.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
