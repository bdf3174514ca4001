#!/bin/bash
# round 2, call O: flash attention with 8 soft_max warps (correctness + A-B), default decode CTAs = 128
mkdir -p gpurun_out
(timeout -k 5 300 python -m pytest tests/test_fast_mode.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2o_pytest_fast.log; tail -3 gpurun_out/r2o_pytest_fast.log
timeout -k 5 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -1 gpurun_out/r2o_bench.err
BARK_B200_FLASH=v2 timeout -k 5 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_flashv2.json 2> /dev/null
python - <<'PY'
import json
for n in ("r2o_bench", "r2o_bench_flashv2"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "decode us", d["roofline"]["avg_launch_us"])
        f = d.get("fast_mode", {}); print("   fast", f.get("e2e"), f.get("ms_per_step"), f.get("fine_pass_ms"), f.get("fine_ids_equal_to_parity"), json.dumps(f.get("tensor_kernels")), f.get("roofline", {}).get("frac"))
    except Exception as e:
        print(n, "failed:", e)
PY
