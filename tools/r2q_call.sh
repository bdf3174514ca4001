#!/bin/bash
# round 2, call Q: evidence for profiles/: launch list, ncu full of the decode step (+ DRAM traffic) and of the fast kernels, decode timeline, bench lines of the other configs
mkdir -p gpurun_out
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fast > gpurun_out/r2q_launches.log 2>&1; tail -1 gpurun_out/r2q_launches.log | cut -c1-120
tools/ncu_one.sh gpt_decode_step 300 1 r2q_decode 2>&1 | tail -1
BARK_B200_MODE=fast timeout -k 5 600 ncu --set full --clock-control none --import-source on -k "regex:umma_gemm|flash_attn|ln_rows" -s 36 -c 10 -f -o gpurun_out/r2q_fast python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fast > gpurun_out/r2q_fast.log 2>&1; tail -1 gpurun_out/r2q_fast.log | cut -c1-120
timeout -k 5 120 python tools/decode_timing.py --sweep 480:40:500 300 > gpurun_out/r2q_timeline.txt 2>&1; head -34 gpurun_out/r2q_timeline.txt
timeout -k 5 400 python bench.py --config large --steps 2 --warmup 3 --no-cpu-baseline --no-fast > gpurun_out/r2q_bench_large.json 2> gpurun_out/r2q_bench_large.err; tail -1 gpurun_out/r2q_bench_large.err
timeout -k 5 400 python bench.py --config small_q4_0 --steps 3 --warmup 3 --no-fast > gpurun_out/r2q_bench_q4.json 2> gpurun_out/r2q_bench_q4.err; tail -1 gpurun_out/r2q_bench_q4.err
python - <<'PY'
import json
for n in ("r2q_bench_large", "r2q_bench_q4"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "decode us", d["roofline"]["avg_launch_us"], d["roofline"]["frac"], "parity", d.get("parity", {}).get("ok"), "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(n, "failed:", e)
PY
