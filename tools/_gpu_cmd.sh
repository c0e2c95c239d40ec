cd "${GRAFT_REPO_ROOT}"; mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-200
