#!/bin/bash
# round 2, call L: q4_0 inside the persistent decode kernel (parity + configs[3] bench line); fast mode with PDL
mkdir -p gpurun_out
(timeout -k 5 600 python -m pytest tests/test_parity_gpu.py tests/test_true_size_gpu.py -m gpu -q -x -k "q4" 2>&1 | tail -12) > gpurun_out/r2l_pytest_q4.log; tail -8 gpurun_out/r2l_pytest_q4.log
timeout -k 5 400 python bench.py --config small_q4_0 --steps 3 --warmup 3 --no-fast > gpurun_out/r2l_bench_q4.json 2> gpurun_out/r2l_bench_q4.err; tail -2 gpurun_out/r2l_bench_q4.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2l_bench_q4.json"))
    print("q4 e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "parity", d.get("parity", {}).get("ok"), "cpu", d.get("cpu_baseline", {}).get("value"))
    print("roofline", d["roofline"]["kernel"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d["roofline"]["algorithmic_bytes_per_launch"])
    print({k: v for k, v in list(d["kernels"].items())[:6]})
except Exception as e:
    print("bench failed:", e)
PY
(timeout -k 5 300 python -m pytest tests/test_fast_mode.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2l_pytest_fast.log; tail -3 gpurun_out/r2l_pytest_fast.log
timeout -k 5 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -2 gpurun_out/r2l_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2l_bench.json"))
    print("e2e", d["e2e"]["value"], "value", d["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "parity", d.get("parity", {}).get("ok"))
    f = d.get("fast_mode", {}); print("fast", f.get("e2e"), f.get("ms_per_step"), f.get("fine_pass_ms"), f.get("fine_ids_equal_to_parity"), json.dumps(f.get("tensor_kernels")), f.get("roofline", {}).get("frac"))
except Exception as e:
    print("bench failed:", e)
PY
