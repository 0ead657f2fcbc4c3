set -x
mkdir -p gpurun_out/r2j
timeout 900 python -m pytest tests/test_gpu_fuzz.py -x -q -m gpu -s 2>&1 | tail -15 > gpurun_out/r2j/fuzz.log; cat gpurun_out/r2j/fuzz.log
for mb in 8192 4096 16384; do
  echo "== budget $mb" >> gpurun_out/r2j/budget.log
  HIPSTR_T_BUDGET_MB=$mb python tools/quick_time.py 200 8 4 2>&1 | grep "run 3\|checksum" >> gpurun_out/r2j/budget.log
done
cat gpurun_out/r2j/budget.log
python tools/trace_time.py 2>&1 | tail -4
