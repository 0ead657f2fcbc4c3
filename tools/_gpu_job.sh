set -x
mkdir -p gpurun_out/r2l
python bench.py --workload loop --loci 4000 --steps 1 --warmup 3 > gpurun_out/r2l/loop_n1_4000.json 2> gpurun_out/r2l/loop_n1_4000.err; tail -c 500 gpurun_out/r2l/loop_n1_4000.err; cut -c1-300 gpurun_out/r2l/loop_n1_4000.json
python tools/trace_time.py 2>&1 | tail -7 > gpurun_out/r2l/trace_time.log; cat gpurun_out/r2l/trace_time.log
