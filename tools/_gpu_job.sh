set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2/gpu_all1.log 2>&1
tail -8 gpurun_out/r2/gpu_all1.log
python tools/loop_time.py 200 8 1 0.05 1 > gpurun_out/r2/loop_v2.log 2>&1
tail -7 gpurun_out/r2/loop_v2.log
