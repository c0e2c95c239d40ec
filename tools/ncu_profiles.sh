#!/bin/bash
# Round profiles: launch lists of the bench commands + one `--set full` capture of the dominant kernels.
# Run under gpurun (ONE GPU); outputs land in gpurun_out/ and are summarised into profiles/ by tools/ncu_summary.py.
# usage: tools/ncu_profiles.sh <round tag, e.g. r02>
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
R=${1:-r02}
# (1) launch lists (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/${R}_launches_direct.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-df > gpurun_out/${R}_bench_direct_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file gpurun_out/${R}_launches_df.csv \
    python bench.py --workload c60-def2svp-df --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_bench_df_under_ncu.log 2>&1
# (2) full captures
cap() { # name regex skip cmd...
  n=$1; re=$2; sk=$3; shift 3
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -s $sk -c 1 -o gpurun_out/${R}_$n -f "$@" > gpurun_out/${R}_$n.log 2>&1
  ncu -i gpurun_out/${R}_$n.ncu-rep --page raw --csv > gpurun_out/${R}_$n.raw.csv 2>/dev/null
  rm -f gpurun_out/${R}_$n.ncu-rep     # the csv export is what is read; the report itself is tens of MB
}
DF="python bench.py --workload c60-def2svp-df --steps 1 --warmup 3 --no-cpu"
cap i8ar 'i8gemm_ar_kernel' 14 $DF
cap i8g2 'i8gemm_kernel' 14 $DF
cap dfjrho 'dfj_rho_kernel' 2 $DF
cap dfjacc 'dfj_acc_kernel' 2 $DF
cap splitpacked 'split_packed_kernel' 0 $DF
cap dpps 'jk_class_kernel.*QClass<\(int\)2, \(int\)1, \(int\)1, \(int\)0,' 2 python tools/profile_classes.py
cap psss 'jk_tpq_kernel<b200jk::QClass<\(int\)1, \(int\)0, \(int\)0, \(int\)0,' 2 python tools/profile_classes.py
python tools/ncu_summary.py gpurun_out/${R}_*.raw.csv > gpurun_out/${R}_ncu_summary.txt 2>&1
python tools/launch_summary.py gpurun_out/${R}_launches_direct.csv > gpurun_out/${R}_launches_direct_summary.txt 2>&1
python tools/launch_summary.py gpurun_out/${R}_launches_df.csv > gpurun_out/${R}_launches_df_summary.txt 2>&1
ls -la gpurun_out | tail -30
