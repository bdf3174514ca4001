#!/bin/bash
# round 2, call E: fast GEMM fixes, adaptive head starts A/B + timeline, parity, bench, large bench
mkdir -p gpurun_out
(timeout -k 5 300 python -m pytest tests/test_fast_mode.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2e_pytest_fast.log; tail -3 gpurun_out/r2e_pytest_fast.log
(timeout -k 5 600 python -m pytest tests/test_parity_gpu.py tests/test_true_size_gpu.py -m gpu -q -x -k "not experimental_quant and not large" 2>&1 | tail -8) > gpurun_out/r2e_pytest.log; tail -4 gpurun_out/r2e_pytest.log
BARK_B200_ADAPT=0 timeout -k 5 120 python tools/decode_bench.py --n-past 300,700 40:500:2000 > gpurun_out/r2e_knobs_fixed.txt 2>&1; tail -2 gpurun_out/r2e_knobs_fixed.txt
timeout -k 5 120 python tools/decode_bench.py --n-past 300,700 40:500:2000 40:0:0 > gpurun_out/r2e_knobs_adapt.txt 2>&1; tail -4 gpurun_out/r2e_knobs_adapt.txt
timeout -k 5 120 python tools/decode_timing.py --sweep 480:40:500 300 > gpurun_out/r2e_timeline.txt 2>&1; head -40 gpurun_out/r2e_timeline.txt
timeout -k 5 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -2 gpurun_out/r2e_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2e_bench.json"))
    print("e2e", d["e2e"]["value"], "value", d["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()})
    print("roofline", d["roofline"]["kernel"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"]); print("parity", d.get("parity", {}).get("ok")); print("cpu", d.get("cpu_baseline", {}).get("value"))
    f = d.get("fast_mode", {}); print("fast", f.get("e2e"), f.get("fine_pass_ms"), f.get("fine_ids_equal_to_parity"), json.dumps(f.get("tensor_kernels")), f.get("roofline", {}).get("frac"))
except Exception as e:
    print("bench failed:", e)
PY
timeout -k 5 400 python bench.py --config large --steps 2 --warmup 3 --no-cpu-baseline --no-fast > gpurun_out/r2e_bench_large.json 2> gpurun_out/r2e_bench_large.err; cut -c1-260 gpurun_out/r2e_bench_large.json; tail -2 gpurun_out/r2e_bench_large.err
