set -x
mkdir -p gpurun_out/r2k
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2k/gpu_all.log 2>&1
tail -4 gpurun_out/r2k/gpu_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2k/smoke.log 2>&1; tail -2 gpurun_out/r2k/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2k/bench_default.json 2> gpurun_out/r2k/bench_default.err; tail -c 300 gpurun_out/r2k/bench_default.err; cut -c1-400 gpurun_out/r2k/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --loci 200 --no-cpu-baseline --no-full-loop --no-e2e > gpurun_out/r2k/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stutter -s 1 -c 1 -o gpurun_out/r2k/k1a_final python tools/quick_time.py 60 8 2 > gpurun_out/r2k/ncu_k1a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_align -s 4 -c 1 -o gpurun_out/r2k/k1b_final python tools/quick_time.py 60 8 2 > gpurun_out/r2k/ncu_k1b.log 2>&1
ls -la gpurun_out/r2k
