set -x
mkdir -p gpurun_out/r2e
python tools/quick_time.py 200 8 4 > gpurun_out/r2e/quick.log 2>&1; tail -6 gpurun_out/r2e/quick.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2e/gpu_all.log 2>&1
tail -4 gpurun_out/r2e/gpu_all.log
