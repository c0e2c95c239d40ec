cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
run() { tag=$1; shift; echo "== $tag"; env "$@" timeout 300 python bench.py --workload c60-def2svp-df --steps 8 --warmup 3 --no-cpu > gpurun_out/r02x_$tag.json 2> gpurun_out/r02x_$tag.err; python tools/bench_brief.py gpurun_out/r02x_$tag.json | grep -E "ms/step|k_gemm" | cut -c1-90; }
run auto X=1
run ks5 B200JK_G2_KS=5
run ks8 B200JK_G2_KS=8
run ks21 B200JK_G2_KS=21
run ks3 B200JK_G2_KS=3
