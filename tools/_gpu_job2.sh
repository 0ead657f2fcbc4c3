set -x
mkdir -p gpurun_out/r2p
nproc
run() { # name, env, pipelines
  env $2 taskset -c 0-7 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 2 --workload loop --loci 2000 --steps 1 --warmup 3 --pipelines $3 --no-cpu-baseline > gpurun_out/r2p/$1.json 2> gpurun_out/r2p/$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2p/$1.json").read().strip().splitlines()[-1]); fl = d["full_loop"]
    print("$1", round(fl["loci_per_s"], 1), "loci/s; host_threads", fl["host_threads"], "windows", fl["windows_per_worker_this_rank"], {k: v for k, v in fl["stage_seconds_summed_over_windows_this_rank"].items() if k in ("construct", "decide", "align", "trace_device", "vcf")})
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r2p/$1.err").read()[-800:])
PY
}
run16() { # name, env, pipelines
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 2 --workload loop --loci 2000 --steps 1 --warmup 3 --pipelines $3 --no-cpu-baseline > gpurun_out/r2p/$1.json 2> gpurun_out/r2p/$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2p/$1.json").read().strip().splitlines()[-1]); fl = d["full_loop"]
    print("$1", round(fl["loci_per_s"], 1), "loci/s; host_threads", fl["host_threads"], "windows", fl["windows_per_worker_this_rank"], {k: v for k, v in fl["stage_seconds_summed_over_windows_this_rank"].items() if k in ("construct", "decide", "align", "trace_device", "vcf")})
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r2p/$1.err").read()[-800:])
PY
}
run sleep_t4_p4 "HIPSTR_HOST_THREADS=4" 4 29601
run spin_t4_p4 "HIPSTR_HOST_THREADS=4 HIPSTR_SPIN_WAITS=1" 4 29602
run sleep_t4_p6 "HIPSTR_HOST_THREADS=4" 6 29603
run spin_t4_p6 "HIPSTR_HOST_THREADS=4 HIPSTR_SPIN_WAITS=1" 6 29605
run sleep_t4_p8 "HIPSTR_HOST_THREADS=4" 8 29604
