set -x
mkdir -p gpurun_out/r2
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2/bench_n2_a.json 2> gpurun_out/r2/bench_n2_a.err; tail -c 300 gpurun_out/r2/bench_n2_a.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload loop --loci 2000 > gpurun_out/r2/bench_loop_n2.json 2> gpurun_out/r2/bench_loop_n2.err; tail -c 300 gpurun_out/r2/bench_loop_n2.err
python bench.py --workload loop --loci 2000 > gpurun_out/r2/bench_loop_n1_2000.json 2> gpurun_out/r2/bench_loop_n1_2000.err; tail -c 300 gpurun_out/r2/bench_loop_n1_2000.err
