#!/bin/bash
# One GPU pass (gpurun): usage tools/gpu_pass.sh <tag> <step> [<step> ...]; steps: tests ab bench ref smoke ncu_direct ncu_df
# Outputs under gpurun_out/<tag>_*.  Every step runs under its own timeout so that a hang cannot eat the GPU budget.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader | head -8
for step in "$@"; do
  echo "== step $step"
  case $step in
    tests)  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_tests.txt ;;
    ab)     timeout 600 python tools/ab_direct.py $(ls pyscf_b200/libb200jk*.so) 2>&1 | tee gpurun_out/${TAG}_ab.txt
            cp gpurun_out/ab_direct_benzene_cc-pvtz.json gpurun_out/${TAG}_ab.json 2>/dev/null ;;
    bench)  timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
            tail -c 1500 gpurun_out/${TAG}_bench.err; head -c 6000 gpurun_out/${TAG}_bench.json; cp /tmp/b200jk_bench_*.log gpurun_out/ 2>/dev/null ;;
    benchnodf) timeout 600 python bench.py --steps 10 --warmup 3 --no-df > gpurun_out/${TAG}_bench_nodf.json 2> gpurun_out/${TAG}_bench_nodf.err
            head -c 3000 gpurun_out/${TAG}_bench_nodf.json ;;
    ref)    timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
            head -c 3000 gpurun_out/${TAG}_ref.json ;;
    pytest:*) timeout 1500 python -m pytest tests -m gpu -x -q -s -k "${step#pytest:}" 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest.txt ;;
    bench:*) W=${step#bench:}; timeout 900 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
            tail -c 800 gpurun_out/${TAG}_bench_$W.err; python tools/bench_brief.py gpurun_out/${TAG}_bench_$W.json ;;
    smoke)  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/${TAG}_smoke.txt ;;
    *)      echo "unknown step $step" ;;
  esac
done
