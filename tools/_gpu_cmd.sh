cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
echo "== df tests"; timeout 1400 python -m pytest tests -m gpu -x -q -k "test_df or i8gemm" 2>&1 | tail -12
echo "== c60"; timeout 600 python bench.py --workload c60-def2svp-df --steps 5 --warmup 3 --no-cpu > gpurun_out/r02q_c60.json 2> gpurun_out/r02q_c60.err; tail -c 300 gpurun_out/r02q_c60.err; python tools/bench_brief.py gpurun_out/r02q_c60.json | cut -c1-900
echo "== taxol"; timeout 600 python bench.py --workload taxol-def2tzvp-df --steps 4 --warmup 3 --no-cpu > gpurun_out/r02q_taxol.json 2> gpurun_out/r02q_taxol.err; python tools/bench_brief.py gpurun_out/r02q_taxol.json | cut -c1-900
