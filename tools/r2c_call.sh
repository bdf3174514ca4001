#!/bin/bash
# round 2, call C: fixed exchange microbench; fast-mode (tcgen05) kernel tests in their own process; full GPU suite; bench
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/microbench/exchange_rounds.cu -o /tmp/exchange_rounds && timeout -k 5 120 /tmp/exchange_rounds > gpurun_out/r2c_exchange_rounds.txt 2>&1
grep -E "W  768.*poll  40" gpurun_out/r2c_exchange_rounds.txt | head -12
(timeout -k 5 300 python -m pytest tests/test_fast_mode.py -m gpu -q -x -s 2>&1 | tail -40) > gpurun_out/r2c_pytest_fast.log; tail -25 gpurun_out/r2c_pytest_fast.log
(timeout -k 5 900 python -m pytest tests -m gpu -q --deselect tests/test_fast_mode.py --durations=8 2>&1 | tail -40) > gpurun_out/r2c_pytest.log; tail -30 gpurun_out/r2c_pytest.log
timeout -k 5 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; cut -c1-400 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c_bench.json"))
    print("e2e", d["e2e"]["value"], "value", d["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()})
    print("roofline", d["roofline"]["kernel"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"]); print("parity", d.get("parity")); print("cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("bench failed:", e)
PY
