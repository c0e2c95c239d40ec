cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
echo "== direct tests"; timeout 900 python -m pytest tests -m gpu -x -q -k "direct or short_range or edge or screening or cart or grad or incore" 2>&1 | tail -3
run() { tag=$1; shift; echo "== $tag"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-df --no-cpu > gpurun_out/r02v_$tag.json 2> gpurun_out/r02v_$tag.err; python tools/bench_brief.py gpurun_out/r02v_$tag.json | cut -c1-100; }
run default X=1
run cap512 B200JK_KETS_CAP=1
run pslice16 B200JK_TPQ_PSLICE=16
run pslice4 B200JK_TPQ_PSLICE=4
run want1 B200JK_WANT_CTAS=1
run want3 B200JK_WANT_CTAS=3
