#!/bin/bash
# third measurement pass: persistent stage-1 kernel (A/B against one-tile-per-CTA and a 4-deep ring), DF profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_df.py tests/test_short_range.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r01c_pytest_df.log
for v in "1 8" "0 8" "1 4"; do set -- $v
  B200JK_AR_PERSIST=$1 B200JK_AR_NSA=$2 timeout 300 python bench.py --workload c60-def2svp-df --no-cpu > gpurun_out/r01c_df_p$1_a$2.json 2> gpurun_out/r01c_df_p$1_a$2.err
done
B200JK_DF_PROFILE=1 timeout 300 python tools/gpu_dfprof.py > gpurun_out/r01c_dfprof.log 2>&1
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/r01c_df_p*.json; tail -n 5 gpurun_out/r01c_dfprof.log; cat gpurun_out/r01c_*.err | tail -5
