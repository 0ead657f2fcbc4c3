set -x
mkdir -p gpurun_out/r2
python bench.py --steps 5 --warmup 3 > gpurun_out/r2/bench_n1_a.json 2> gpurun_out/r2/bench_n1_a.err; tail -c 600 gpurun_out/r2/bench_n1_a.err
python bench.py --workload cfg4_em --loci 500 --steps 3 --warmup 3 > gpurun_out/r2/bench_em.json 2> gpurun_out/r2/bench_em.err; tail -c 400 gpurun_out/r2/bench_em.err
