#!/bin/bash
# round 2, call I: cluster decode kernel (one 16-CTA cluster, DSMEM exchanges): parity + speed
mkdir -p gpurun_out
(BARK_B200_DECODE=cluster timeout -k 5 400 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "not experimental_quant and not one_host and not packed" 2>&1 | tail -25) > gpurun_out/r2i_pytest_cluster.log; tail -25 gpurun_out/r2i_pytest_cluster.log
BARK_B200_DECODE=cluster timeout -k 5 120 python tools/decode_bench.py --n-past 300,700 40:500:2000 > gpurun_out/r2i_decode_cluster.txt 2>&1; tail -3 gpurun_out/r2i_decode_cluster.txt
timeout -k 5 120 python tools/decode_bench.py --n-past 300,700 40:500:2000 > gpurun_out/r2i_decode_default.txt 2>&1; tail -3 gpurun_out/r2i_decode_default.txt
BARK_B200_DECODE=cluster timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-fast > gpurun_out/r2i_bench_cluster.json 2> gpurun_out/r2i_bench_cluster.err; tail -2 gpurun_out/r2i_bench_cluster.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2i_bench_cluster.json"))
    print("cluster e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "parity", d.get("parity", {}).get("ok"))
    print("roofline", d["roofline"]["kernel"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"])
except Exception as e:
    print("bench failed:", e)
PY
