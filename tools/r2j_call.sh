#!/bin/bash
# round 2, call J: q4_0 per-op decode path after the tiled-attention / load-unroll changes: parity + configs[3] bench line
mkdir -p gpurun_out
(timeout -k 5 600 python -m pytest tests/test_parity_gpu.py tests/test_true_size_gpu.py -m gpu -q -x -k "q4 or quant or variants or tokens_bit_exact" 2>&1 | tail -6) > gpurun_out/r2j_pytest_q4.log; tail -4 gpurun_out/r2j_pytest_q4.log
timeout -k 5 400 python bench.py --config small_q4_0 --steps 2 --warmup 3 --no-fast > gpurun_out/r2j_bench_q4.json 2> gpurun_out/r2j_bench_q4.err; tail -2 gpurun_out/r2j_bench_q4.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2j_bench_q4.json"))
    print("q4 e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "parity", d.get("parity", {}).get("ok"), "cpu", d.get("cpu_baseline", {}).get("value"))
    print({k: v for k, v in list(d["kernels"].items())[:8]})
except Exception as e:
    print("bench failed:", e)
PY
