set -x
mkdir -p gpurun_out/r2v
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2v/gpu_all.log 2>&1
tail -3 gpurun_out/r2v/gpu_all.log
python bench.py > gpurun_out/r2v/bench_default.json 2> gpurun_out/r2v/bench_default.err; tail -c 200 gpurun_out/r2v/bench_default.err; cut -c1-200 gpurun_out/r2v/bench_default.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
