set -x
mkdir -p gpurun_out/r2x
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2x/gpu_all.log 2>&1
tail -3 gpurun_out/r2x/gpu_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
