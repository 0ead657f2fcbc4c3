set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2/gpu_all2.log 2>&1
tail -6 gpurun_out/r2/gpu_all2.log
python bench.py --workload cfg4_em --loci 1000 --steps 3 --warmup 3 > gpurun_out/r2/bench_em.json 2> gpurun_out/r2/bench_em.err; tail -c 300 gpurun_out/r2/bench_em.err
python bench.py --workload sweep --steps 3 > gpurun_out/r2/bench_sweep.json 2> gpurun_out/r2/bench_sweep.err; tail -c 300 gpurun_out/r2/bench_sweep.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --loci 200 --no-cpu-baseline --no-full-loop --no-e2e > gpurun_out/r2/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stutter -s 1 -c 1 -o gpurun_out/r2/k1a_final python tools/quick_time.py 60 8 2 > gpurun_out/r2/ncu_k1a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_align -s 4 -c 1 -o gpurun_out/r2/k1b_final python tools/quick_time.py 60 8 2 > gpurun_out/r2/ncu_k1b.log 2>&1
