#!/bin/bash
# First GPU-box call of round 2 (≈5 min of run time): everything round 1 prepared but could not run.
#   usage: gpurun --timeout 900 -- 'tools/round2_first_call.sh'
# 1. micro-benchmarks of the decode step's primitives (exchange latency vs polling, grid.sync, cluster barrier / DSMEM, nanosleep, TMA issue)
# 2. the default parity suite (must stay green)
# 3. the opt-in suites: packed-FMA (FFMA2) variants and the experimental quantised types
# 4. bench with and without BARK_B200_FFMA2=1
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/microbench/exchange_bench.cu -o gpurun_out/exchange_bench 2> gpurun_out/r2a_microbench_build.log \
  && timeout -k 5 60 gpurun_out/exchange_bench > gpurun_out/r2a_microbench.txt 2>&1
tail -60 gpurun_out/r2a_microbench.txt
(timeout -k 5 300 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r2a_pytest.log; tail -2 gpurun_out/r2a_pytest.log
(BARK_B200_TEST_EXPERIMENTAL=1 timeout -k 5 300 python -m pytest tests -m gpu -q -k "packed_fma or experimental_quant" 2>&1 | tail -15) > gpurun_out/r2a_pytest_experimental.log; tail -6 gpurun_out/r2a_pytest_experimental.log
timeout -k 5 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; cut -c1-200 gpurun_out/r2a_bench_default.json
BARK_B200_FFMA2=1 timeout -k 5 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_ffma2.json 2> gpurun_out/r2a_bench_ffma2.err; cut -c1-200 gpurun_out/r2a_bench_ffma2.json
python - <<'PY'
import json
for n in ("default", "ffma2"):
    try:
        d = json.load(open(f"gpurun_out/r2a_bench_{n}.json"))
        k = d["kernels"]
        print(n, "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "fine ms", d["stages"]["fine"]["ms"],
              "gemm ms", k.get("lane_gemm_tiled_kernel<__half>", k.get("lane_gemm_tiled_kernel<__half, true>", {})).get("ms"))
    except Exception as e:
        print(n, "failed:", e)
PY
