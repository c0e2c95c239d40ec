#!/bin/bash
# Closing pass of a round (usage: bash tools/gpu_closing_pass.sh r02): GPU test tier, both bench lines, smoke, per-class times,
# launch list of the direct bench and one full capture of the (fd|dp) kernel.
mkdir -p gpurun_out
R=${1:-r02}
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${R}_pytest.log; cat gpurun_out/${R}_pytest.log
timeout 150 python bench.py > gpurun_out/${R}_direct.json 2> gpurun_out/${R}_direct.err
timeout 200 python bench.py --workload c60-def2svp-df --steps 5 > gpurun_out/${R}_df.json 2> gpurun_out/${R}_df.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/${R}_direct.json gpurun_out/${R}_df.json; tail -2 gpurun_out/${R}_direct.err gpurun_out/${R}_df.err
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/${R}_smoke.log; cat gpurun_out/${R}_smoke.log
timeout 60 python tools/ab_direct.py pyscf_b200/libb200jk.so 2>&1 | tail -1; cp gpurun_out/ab_direct_benzene_cc-pvtz.json gpurun_out/${R}_class_times.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/${R}_launches_direct.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_bench_direct_under_ncu.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:jk_class_kernel_2cta<b200jk::QClass<\(int\)3, \(int\)2, \(int\)2, \(int\)1," -c 1 \
    -o gpurun_out/${R}_fddp -f python tools/profile_classes.py > gpurun_out/${R}_fddp.log 2>&1
ncu -i gpurun_out/${R}_fddp.ncu-rep --page raw --csv > gpurun_out/${R}_fddp.raw.csv 2>/dev/null
ls -la gpurun_out | tail -14
