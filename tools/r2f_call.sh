#!/bin/bash
# round 2, call F: per-SM streaming bandwidth (decides the 16-CTA-cluster decode design), flash v2 correctness + speed, bench
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/microbench/stream_bw.cu -o /tmp/stream_bw && timeout -k 5 120 /tmp/stream_bw > gpurun_out/r2f_stream_bw.txt 2>&1; cat gpurun_out/r2f_stream_bw.txt
(timeout -k 5 300 python -m pytest tests/test_fast_mode.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2f_pytest_fast.log; tail -3 gpurun_out/r2f_pytest_fast.log
timeout -k 5 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -2 gpurun_out/r2f_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2f_bench.json"))
    print("e2e", d["e2e"]["value"], "value", d["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()})
    print("roofline", d["roofline"]["kernel"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"]); print("parity", d.get("parity", {}).get("ok")); print("cpu", d.get("cpu_baseline", {}).get("value"))
    f = d.get("fast_mode", {}); print("fast", f.get("e2e"), f.get("fine_pass_ms"), f.get("fine_ids_equal_to_parity"), json.dumps(f.get("tensor_kernels")), f.get("roofline", {}).get("frac"))
except Exception as e:
    print("bench failed:", e)
PY
BARK_B200_FLASH=v1 timeout -k 5 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_flashv1.json 2> /dev/null
python - <<'PY'
import json
try:
    f = json.load(open("gpurun_out/r2f_bench_flashv1.json")).get("fast_mode", {}); print("flash v1:", json.dumps(f.get("tensor_kernels")), f.get("fine_pass_ms"))
except Exception as e:
    print("failed:", e)
PY
