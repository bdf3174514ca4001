#!/bin/bash
# round 2, call D: fast-mode tests (fixed), new parity tests, decode knob sweep with the replicated exchange, bench (+ fast leg), ncu captures, configs[2]/[3] bench lines
mkdir -p gpurun_out
(timeout -k 5 300 python -m pytest tests/test_fast_mode.py -m gpu -q -x -s 2>&1 | tail -15) > gpurun_out/r2d_pytest_fast.log; tail -6 gpurun_out/r2d_pytest_fast.log
(timeout -k 5 600 python -m pytest tests -m gpu -q -k "large_full_depth_f16 or one_host_thread or epochs_survive or bench_clip" 2>&1 | tail -15) > gpurun_out/r2d_pytest_new.log; tail -4 gpurun_out/r2d_pytest_new.log
timeout -k 5 240 python tools/decode_bench.py --n-past 300,700 40:500:2000 40:0:2000 0:0:0 40:200:1000 100:0:0 40:500:0 > gpurun_out/r2d_knobs.txt 2>&1; cat gpurun_out/r2d_knobs.txt | tail -14
timeout -k 5 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -2 gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2d_bench.json"))
    print("e2e", d["e2e"]["value"], "value", d["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()})
    print("roofline", d["roofline"]["kernel"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"]); print("parity", d.get("parity", {}).get("ok")); print("cpu", d.get("cpu_baseline", {}).get("value"))
    print("fast", json.dumps(d.get("fast_mode"))[:1500])
    print({k: v for k, v in list(d["kernels"].items())[:8]})
except Exception as e:
    print("bench failed:", e)
PY
tools/ncu_one.sh gpt_decode_step 300 1 r2d_decode 2>&1 | tail -2
BARK_B200_MODE=fast ncu --set full --clock-control none --import-source on -k "regex:umma_gemm|flash_attn" -s 30 -c 8 -f -o gpurun_out/r2d_fast python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fast > gpurun_out/r2d_fast.log 2>&1; tail -2 gpurun_out/r2d_fast.log
timeout -k 5 400 python bench.py --config small_q4_0 --steps 2 --warmup 3 --no-fast > gpurun_out/r2d_bench_q4.json 2> gpurun_out/r2d_bench_q4.err; cut -c1-300 gpurun_out/r2d_bench_q4.json; tail -2 gpurun_out/r2d_bench_q4.err
timeout -k 5 400 python bench.py --config large --steps 2 --warmup 3 --no-cpu-baseline --no-fast > gpurun_out/r2d_bench_large.json 2> gpurun_out/r2d_bench_large.err; cut -c1-300 gpurun_out/r2d_bench_large.json; tail -2 gpurun_out/r2d_bench_large.err
ls -la gpurun_out | tail -20
