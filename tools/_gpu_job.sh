set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2/parity5.log 2>&1
tail -3 gpurun_out/r2/parity5.log
python tools/quick_time.py 200 8 3 > gpurun_out/r2/quick_v2e.log 2>&1
tail -4 gpurun_out/r2/quick_v2e.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_v2e.csv python tools/quick_time.py 60 8 2 > gpurun_out/r2/ncu_l.log 2>&1
