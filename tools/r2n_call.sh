#!/bin/bash
# round 2, call N: decode CTA count x ff head start; q4_0 decode after the ld.shared change; parity at 128 CTAs
mkdir -p gpurun_out
: > gpurun_out/r2n_sweep.txt
for c in 148 128 120; do for hs in 0:2000:500:400:500:0 0:2000:500:600:500:0 0:2000:700:400:700:0; do
  echo "== CTAS $c HEADSTART $hs" >> gpurun_out/r2n_sweep.txt
  BARK_B200_DECODE_CTAS=$c BARK_B200_HEADSTART=$hs timeout -k 5 100 python tools/decode_bench.py --n-past 300,700 40:500:2000 2>&1 | tail -2 >> gpurun_out/r2n_sweep.txt
done; done
cat gpurun_out/r2n_sweep.txt
(BARK_B200_DECODE_CTAS=128 timeout -k 5 400 python -m pytest tests/test_parity_gpu.py tests/test_true_size_gpu.py -m gpu -q -x -k "q4 or bench_clip or teacher or tokens_bit_exact or full_size" 2>&1 | tail -5) > gpurun_out/r2n_pytest.log; tail -3 gpurun_out/r2n_pytest.log
timeout -k 5 400 python bench.py --config small_q4_0 --steps 3 --warmup 3 --no-fast --no-cpu-baseline > gpurun_out/r2n_bench_q4.json 2> gpurun_out/r2n_bench_q4.err; tail -1 gpurun_out/r2n_bench_q4.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2n_bench_q4.json"))
    print("q4 e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "stages", {k: v["ms"] for k, v in d["stages"].items()})
    print("roofline", d["roofline"]["kernel"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"])
except Exception as e:
    print("bench failed:", e)
PY
