#!/bin/bash
# round 2, call U (8 GPUs): the driver's scaling run shape at N = 8 (default config, replicas) with ranks pinned to their GPU's local CPUs
mkdir -p gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2u_bench_n8.json 2> gpurun_out/r2u_bench_n8.err; cut -c1-300 gpurun_out/r2u_bench_n8.json; tail -2 gpurun_out/r2u_bench_n8.err
timeout -k 5 300 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline --no-fast > gpurun_out/r2u_bench_n1.json 2> /dev/null
python - <<'PY'
import json
try:
    a = json.load(open("gpurun_out/r2u_bench_n1.json")); b = json.load(open("gpurun_out/r2u_bench_n8.json"))
    print("N=1 e2e", a["e2e"]["value"], "ms", a["ms_per_step"], "| N=8 e2e", b["e2e"]["value"], "ms", b["ms_per_step"], "value", b["value"], "| e2e efficiency", round(b["e2e"]["value"] / (8 * a["e2e"]["value"]), 4), "| parallelism:", b["config"]["parallelism"])
except Exception as e:
    print("failed:", e)
PY
