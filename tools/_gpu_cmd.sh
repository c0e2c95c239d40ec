cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
run() { tag=$1; shift; echo "== $tag"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-df --no-cpu > gpurun_out/r02w_$tag.json 2> gpurun_out/r02w_$tag.err; python tools/bench_brief.py gpurun_out/r02w_$tag.json | cut -c1-100; }
run carve50 X=1
run carve25 B200JK_CARVEOUT=25
run carve75 B200JK_CARVEOUT=75
run carve100 B200JK_CARVEOUT=100
run carve0 B200JK_CARVEOUT=0
