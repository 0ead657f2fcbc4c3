set -x
mkdir -p gpurun_out/r2b
nproc
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2b/gpu_all.log 2>&1
tail -6 gpurun_out/r2b/gpu_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2b/smoke.log 2>&1; tail -2 gpurun_out/r2b/smoke.log
python bench.py --workload loop --loci 2000 --steps 3 --warmup 3 > gpurun_out/r2b/loop_n1.json 2> gpurun_out/r2b/loop_n1.err; tail -c 400 gpurun_out/r2b/loop_n1.err; cat gpurun_out/r2b/loop_n1.json
python bench.py > gpurun_out/r2b/bench_default.json 2> gpurun_out/r2b/bench_default.err; tail -c 300 gpurun_out/r2b/bench_default.err; cat gpurun_out/r2b/bench_default.json
