set -x
mkdir -p gpurun_out/r2n8
nproc; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload loop --loci 10000 --alleles 16 --steps 1 --warmup 3 > gpurun_out/r2n8/loop_n8.json 2> gpurun_out/r2n8/loop_n8.err; tail -c 600 gpurun_out/r2n8/loop_n8.err; cat gpurun_out/r2n8/loop_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n8/bench_n8.json 2> gpurun_out/r2n8/bench_n8.err; tail -c 600 gpurun_out/r2n8/bench_n8.err; cat gpurun_out/r2n8/bench_n8.json
