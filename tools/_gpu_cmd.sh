cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-160
echo "== bench (default, as the driver runs it)"; timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err; tail -c 400 gpurun_out/r02p_bench_n1.err; python tools/bench_brief.py gpurun_out/r02p_bench_n1.json | cut -c1-260
echo "== reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02p_ref.json 2> gpurun_out/r02p_ref.err; python tools/bench_brief.py gpurun_out/r02p_ref.json | cut -c1-400
