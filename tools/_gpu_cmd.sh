cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
echo "== df tests"; timeout 1400 python -m pytest tests -m gpu -x -q -k "test_df" 2>&1 | tail -3
echo "== taxol"; timeout 600 python bench.py --workload taxol-def2tzvp-df --steps 4 --warmup 3 --no-cpu > gpurun_out/r02n_taxol.json 2> gpurun_out/r02n_taxol.err; tail -c 600 gpurun_out/r02n_taxol.err; python tools/bench_brief.py gpurun_out/r02n_taxol.json | cut -c1-250
nvidia-smi --query-gpu=memory.used --format=csv,noheader
