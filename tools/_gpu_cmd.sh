cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
echo "== tests (df + direct + size)"; timeout 1400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== c60 new g2"; timeout 600 python bench.py --workload c60-def2svp-df --steps 5 --warmup 3 --no-cpu > gpurun_out/r02i_c60.json 2> gpurun_out/r02i_c60.err; python tools/bench_brief.py gpurun_out/r02i_c60.json | cut -c1-250
echo "== c60 old g2"; B200JK_G2_OLD=1 timeout 600 python bench.py --workload c60-def2svp-df --steps 5 --warmup 3 --no-cpu > gpurun_out/r02i_c60old.json 2> gpurun_out/r02i_c60old.err; python tools/bench_brief.py gpurun_out/r02i_c60old.json | grep -E "ms/step|k_gemm2" | cut -c1-120
echo "== taxol"; timeout 600 python bench.py --workload taxol-def2tzvp-df --steps 4 --warmup 3 --no-cpu > gpurun_out/r02i_taxol.json 2> gpurun_out/r02i_taxol.err; python tools/bench_brief.py gpurun_out/r02i_taxol.json | cut -c1-250
echo "== direct"; timeout 600 python bench.py --steps 10 --warmup 3 --no-df --no-cpu > gpurun_out/r02i_direct.json 2> gpurun_out/r02i_direct.err; python tools/bench_brief.py gpurun_out/r02i_direct.json | cut -c1-250
