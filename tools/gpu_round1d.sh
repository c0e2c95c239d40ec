#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_df.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r01d_pytest_df.log
timeout 300 python bench.py --workload c60-def2svp-df --no-cpu > gpurun_out/r01d_df.json 2> gpurun_out/r01d_df.err
B200JK_DF_PROFILE=1 timeout 300 python tools/gpu_dfprof.py > gpurun_out/r01d_dfprof.log 2>&1
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/r01d_df.json; tail -n 4 gpurun_out/r01d_dfprof.log; cat gpurun_out/r01d_*.err | tail -5
