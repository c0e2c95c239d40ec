#!/bin/bash
# A/B pass over the tuning variants of the direct J/K kernels (tools/build_variant.sh), benzene/cc-pVTZ
mkdir -p gpurun_out
timeout 400 python tools/ab_direct.py pyscf_b200/libb200jk_base.so pyscf_b200/libb200jk.so $(ls pyscf_b200/libb200jk_*.so | grep -v _base) 2>&1 | tee gpurun_out/r01g_ab.log
cp gpurun_out/ab_direct_benzene_cc-pvtz.json gpurun_out/r01g_ab_$(date +%H%M).json
