set -x
mkdir -p gpurun_out/r2q
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2q/gpu_all.log 2>&1
tail -4 gpurun_out/r2q/gpu_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2q/smoke.log 2>&1; tail -2 gpurun_out/r2q/smoke.log
python bench.py > gpurun_out/r2q/bench_default.json 2> gpurun_out/r2q/bench_default.err; tail -c 300 gpurun_out/r2q/bench_default.err; cut -c1-330 gpurun_out/r2q/bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2q/bench_reference.json 2> gpurun_out/r2q/bench_reference.err; cut -c1-400 gpurun_out/r2q/bench_reference.json
python tools/trace_time.py 2>&1 | tail -3
