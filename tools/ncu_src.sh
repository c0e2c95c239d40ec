#!/bin/bash
# usage: tools/ncu_src.sh <name> <kernel> LI LJ LK LL : full capture + per-source-line sample export
mkdir -p gpurun_out
RE="$2<b200jk::QClass<\(int\)$3, \(int\)$4, \(int\)$5, \(int\)$6,"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$RE" -c 1 \
    -o gpurun_out/$1 -f python tools/profile_classes.py > gpurun_out/$1.log 2>&1 || tail -5 gpurun_out/$1.log
ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
ncu -i gpurun_out/$1.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/$1.src.csv 2>/dev/null
