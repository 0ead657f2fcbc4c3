set -x
mkdir -p gpurun_out/r2t
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2t/gpu_all.log 2>&1
tail -5 gpurun_out/r2t/gpu_all.log
python bench.py > gpurun_out/r2t/bench_default.json 2> gpurun_out/r2t/bench_default.err; tail -c 300 gpurun_out/r2t/bench_default.err; cut -c1-200 gpurun_out/r2t/bench_default.json
python bench.py --workload cfg4_em --loci 1000 --steps 3 --warmup 3 > gpurun_out/r2t/bench_em.json 2> gpurun_out/r2t/bench_em.err; tail -c 300 gpurun_out/r2t/bench_em.err; cut -c1-300 gpurun_out/r2t/bench_em.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
