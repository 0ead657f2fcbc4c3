set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2/parity4.log 2>&1
tail -3 gpurun_out/r2/parity4.log
python tools/quick_time.py 200 8 3 > gpurun_out/r2/quick_v2d.log 2>&1
tail -5 gpurun_out/r2/quick_v2d.log
HIPSTR_ALIGN_WARPS=20 python tools/quick_time.py 200 8 3 > gpurun_out/r2/quick_v2d20.log 2>&1
tail -5 gpurun_out/r2/quick_v2d20.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_v2d.csv python tools/quick_time.py 60 8 2 > gpurun_out/r2/ncu_l.log 2>&1
HIPSTR_ALIGN_WARPS=20 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_v2d20.csv python tools/quick_time.py 60 8 2 > gpurun_out/r2/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_align -s 4 -c 1 -o gpurun_out/r2/k1b_v2d python tools/quick_time.py 60 8 2 > gpurun_out/r2/ncu_k1b.log 2>&1
