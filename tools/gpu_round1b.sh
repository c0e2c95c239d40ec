#!/bin/bash
# second measurement pass of round 1: GPU tests, DF-K stage 1 with/without stacked-N MMAs, direct bench, one ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r01b_pytest_gpu.log
python bench.py --workload c60-def2svp-df --no-cpu > gpurun_out/r01b_df_stack1.json 2> gpurun_out/r01b_df_stack1.err
B200JK_AR_STACK=0 python bench.py --workload c60-def2svp-df --no-cpu > gpurun_out/r01b_df_stack0.json 2> gpurun_out/r01b_df_stack0.err
python bench.py --no-cpu > gpurun_out/r01b_direct.json 2> gpurun_out/r01b_direct.err
tail -n 3 gpurun_out/r01b_*.json gpurun_out/r01b_*.err
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:i8gemm_ar_kernel" -s 2 -c 1 -o gpurun_out/r01b_i8ar -f \
    python bench.py --workload c60-def2svp-df --steps 1 --warmup 3 --no-cpu > gpurun_out/r01b_i8ar.log 2>&1
ncu -i gpurun_out/r01b_i8ar.ncu-rep --page raw --csv > gpurun_out/r01b_i8ar.raw.csv 2>/dev/null
ls -la gpurun_out | tail
