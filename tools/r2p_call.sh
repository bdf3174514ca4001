#!/bin/bash
# round 2, call P: full GPU suite after the fused sampler / 128-CTA default / flash v3; bench A-B fused vs separate sampler
mkdir -p gpurun_out
(timeout -k 5 1200 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -25) > gpurun_out/r2p_pytest.log; tail -14 gpurun_out/r2p_pytest.log
timeout -k 5 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -1 gpurun_out/r2p_bench.err
BARK_B200_FUSE_SAMPLER=0 timeout -k 5 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fast > gpurun_out/r2p_bench_nofuse.json 2> /dev/null
python - <<'PY'
import json
for n in ("r2p_bench", "r2p_bench_nofuse"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, "e2e", d["e2e"]["value"], "value", d["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "decode us", d["roofline"]["avg_launch_us"], d["roofline"]["frac"], "launches", d["gpu_launches"], "parity", d.get("parity", {}).get("ok"))
        f = d.get("fast_mode") or {}; print("   fast", f.get("e2e"), f.get("ms_per_step"), f.get("fine_pass_ms"))
    except Exception as e:
        print(n, "failed:", e)
PY
