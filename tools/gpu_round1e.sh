#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_df.py tests/test_gpu_direct.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r01e_pytest.log
B200JK_I8_DEBUG=1 timeout 300 python bench.py --workload c60-def2svp-df --no-cpu --steps 5 > gpurun_out/r01e_df.json 2> gpurun_out/r01e_df.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/r01e_df.json; grep i8gemm_ar gpurun_out/r01e_df.err | head -4
B200JK_DF_PROFILE=1 timeout 300 python tools/gpu_dfprof.py 2>&1 | tail -2
timeout 300 python bench.py --no-cpu > gpurun_out/r01e_direct.json 2> gpurun_out/r01e_direct.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/r01e_direct.json
