set -x
mkdir -p gpurun_out/r2w
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2w/gpu_all.log 2>&1
tail -3 gpurun_out/r2w/gpu_all.log
python bench.py --workload loop --loci 2000 --steps 1 --warmup 3 > gpurun_out/r2w/loop_n1.json 2> gpurun_out/r2w/loop_n1.err; tail -c 200 gpurun_out/r2w/loop_n1.err; cut -c1-200 gpurun_out/r2w/loop_n1.json
