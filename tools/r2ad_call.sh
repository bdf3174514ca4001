#!/bin/bash
# round 2, call AD: full GPU suite + bench on the final code
mkdir -p gpurun_out
(timeout -k 5 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2ad_pytest.log; tail -3 gpurun_out/r2ad_pytest.log
timeout -k 5 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; tail -1 gpurun_out/r2ad_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2ad_bench.json"))
    print("e2e", d["e2e"]["value"], "value", d["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "decode us", d["roofline"]["avg_launch_us"], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "parity", d.get("parity", {}).get("ok"), "cpu", d.get("cpu_baseline", {}).get("value"))
    f = d.get("fast_mode") or {}; print("   fast", f.get("e2e"), f.get("ms_per_step"), f.get("fine_pass_ms"), f.get("fine_ids_equal_to_parity"))
except Exception as e:
    print("bench failed:", e)
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
