#!/bin/bash
# round 2, call S: parity after the PV unroll / single-read LayerNorm kernels; bench; fast LN
mkdir -p gpurun_out
(timeout -k 5 900 python -m pytest tests/test_parity_gpu.py tests/test_true_size_gpu.py tests/test_fast_mode.py tests/test_baseline_config0.py -m gpu -q -x -k "not experimental_quant and not large" 2>&1 | tail -5) > gpurun_out/r2s_pytest.log; tail -3 gpurun_out/r2s_pytest.log
timeout -k 5 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -1 gpurun_out/r2s_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2s_bench.json"))
    print("e2e", d["e2e"]["value"], "value", d["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "decode us", d["roofline"]["avg_launch_us"], d["roofline"]["frac"])
    f = d.get("fast_mode") or {}; print("   fast", f.get("e2e"), f.get("ms_per_step"), f.get("fine_pass_ms"), f.get("fine_ids_equal_to_parity"))
    print({k: v["ms"] for k, v in list(d["kernels"].items())[:8]})
except Exception as e:
    print("bench failed:", e)
PY
