cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== df tests"; timeout 1200 python -m pytest tests -m gpu -x -q -k "test_df_gpu or orbital or parity_at_size" 2>&1 | tail -3
echo "== c60"; B200JK_I8_DEBUG=1 timeout 600 python bench.py --workload c60-def2svp-df --steps 5 --warmup 3 --no-cpu > gpurun_out/r02g_c60.json 2> gpurun_out/r02g_c60.err; grep -E "i8gemm_ar CTA0" gpurun_out/r02g_c60.err | head -4; python tools/bench_brief.py gpurun_out/r02g_c60.json | cut -c1-200
