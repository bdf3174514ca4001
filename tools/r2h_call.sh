#!/bin/bash
# round 2, call H: f16-values-in-f32-containers option of the parity tiled GEMM: bit-exactness + speed A-B
mkdir -p gpurun_out
(BARK_B200_GEMM_F32C=1 timeout -k 5 600 python -m pytest tests/test_parity_gpu.py tests/test_true_size_gpu.py -m gpu -q -x -k "not experimental_quant and not large and not q4 and not one_host" 2>&1 | tail -6) > gpurun_out/r2h_pytest_f32c.log; tail -3 gpurun_out/r2h_pytest_f32c.log
BARK_B200_GEMM_F32C=1 timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-fast > gpurun_out/r2h_bench_f32c.json 2> gpurun_out/r2h_bench_f32c.err; tail -2 gpurun_out/r2h_bench_f32c.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2h_bench_f32c.json"))
    print("F32C e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()}, "parity", d.get("parity", {}).get("ok"))
    print({k: v for k, v in list(d["kernels"].items())[:7]})
except Exception as e:
    print("bench failed:", e)
PY
