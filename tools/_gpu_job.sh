set -x
mkdir -p gpurun_out/r2u
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2u/gpu_all.log 2>&1
tail -3 gpurun_out/r2u/gpu_all.log
python tools/trace_time.py 2>&1 | tail -3 > gpurun_out/r2u/trace_time.log; cat gpurun_out/r2u/trace_time.log
python bench.py --workload loop --loci 2000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2u/loop_n1.json 2> gpurun_out/r2u/loop_n1.err; tail -c 200 gpurun_out/r2u/loop_n1.err; cut -c1-200 gpurun_out/r2u/loop_n1.json
