#!/bin/bash
# Round profiles: launch lists of the bench commands + one `--set full` capture of the dominant kernels.
# Run under gpurun; outputs land in gpurun_out/ and are summarised into profiles/ by tools/ncu_summary.py.
mkdir -p gpurun_out
R=${1:-r01}
# (1) launch lists (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/${R}_launches_direct.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_bench_direct_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/${R}_launches_df.csv \
    python bench.py --workload c60-def2svp-df --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_bench_df_under_ncu.log 2>&1
# (2) full captures
cap() { # name regex cmd...
  n=$1; re=$2; shift 2
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -s 2 -c 1 -o gpurun_out/${R}_$n -f "$@" > gpurun_out/${R}_$n.log 2>&1
  ncu -i gpurun_out/${R}_$n.ncu-rep --page raw --csv > gpurun_out/${R}_$n.raw.csv 2>/dev/null
}
cap fsdp 'jk_class_kernel<b200jk::QClass<\(int\)3, \(int\)0, \(int\)2, \(int\)1,' python tools/profile_classes.py
cap psss 'jk_tpq_kernel<b200jk::QClass<\(int\)1, \(int\)0, \(int\)0, \(int\)0,' python tools/profile_classes.py
cap i8ar 'i8gemm_ar_kernel' python bench.py --workload c60-def2svp-df --steps 1 --warmup 3 --no-cpu
cap i8g2 'i8gemm_kernel' python bench.py --workload c60-def2svp-df --steps 1 --warmup 3 --no-cpu
cap dfjacc 'dfj_acc_kernel' python bench.py --workload c60-def2svp-df --steps 1 --warmup 3 --no-cpu
ls -la gpurun_out | tail -30
