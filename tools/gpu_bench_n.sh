#!/bin/bash
# Multi-GPU bench exactly as the driver launches it: tools/gpu_bench_n.sh <tag> <N> [extra bench.py args]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8
if [ "$N" = 1 ]; then
  timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" \
    > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
fi
echo "exit $?"
tail -c 1500 gpurun_out/${TAG}_bench_n$N.err
python tools/bench_brief.py gpurun_out/${TAG}_bench_n$N.json
cp /tmp/b200jk_bench_*_r0.log gpurun_out/ 2>/dev/null
