set -x
mkdir -p gpurun_out/r2n8b
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n8b/bench_n8.json 2> gpurun_out/r2n8b/bench_n8.err; tail -c 600 gpurun_out/r2n8b/bench_n8.err; cut -c1-300 gpurun_out/r2n8b/bench_n8.json
