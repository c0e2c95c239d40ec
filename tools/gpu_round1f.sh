#!/bin/bash
# Round-1 closing pass: full GPU test tier, both bench lines, smoke, then the DF launch list and one full capture of
# the persistent stage-1 tcgen05 kernel (the kernel that changed since the r01 profiles).
mkdir -p gpurun_out
R=r01f
(timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${R}_pytest.log; cat gpurun_out/${R}_pytest.log
timeout 150 python bench.py > gpurun_out/${R}_direct.json 2> gpurun_out/${R}_direct.err
timeout 200 python bench.py --workload c60-def2svp-df --steps 5 > gpurun_out/${R}_df.json 2> gpurun_out/${R}_df.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/${R}_direct.json gpurun_out/${R}_df.json
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/${R}_smoke.log; cat gpurun_out/${R}_smoke.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/${R}_launches_df.csv \
    python bench.py --workload c60-def2svp-df --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_bench_df_under_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:i8gemm_ar" -s 2 -c 1 \
    -o gpurun_out/${R}_i8ar -f python bench.py --workload c60-def2svp-df --steps 1 --warmup 3 --no-cpu > gpurun_out/${R}_i8ar.log 2>&1
ncu -i gpurun_out/${R}_i8ar.ncu-rep --page raw --csv > gpurun_out/${R}_i8ar.raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
