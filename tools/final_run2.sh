#!/bin/bash
# Last GPU-box call of the round: full parity suite on the final code, exchange-knob sweep on the production kernel, bench line.
tag=${1:-r1l}
mkdir -p gpurun_out
(timeout -k 5 260 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout -k 5 100 python tools/decode_bench.py --n-past 300 40:500:0 40:500:1500 40:500:3000 40:0:0 40:800:3000 40:300:1500 > gpurun_out/${tag}_decode_bench.txt 2>&1; cat gpurun_out/${tag}_decode_bench.txt | grep "^poll"
python - <<'PY' > gpurun_out/best_knobs2.env
import json
try:
    d = json.load(open("gpurun_out/decode_bench.json"))
    b = min(d, key=lambda e: e["us_per_token"])
    print(f"export BARK_B200_POLL_NS={b['poll_ns']} BARK_B200_POLL_FIRST_NS={b['first_ns']} BARK_B200_POLL_ATT_NS={b['att_ns']}")
except Exception as ex:
    print("# sweep failed:", ex)
PY
cat gpurun_out/best_knobs2.env; source gpurun_out/best_knobs2.env
timeout -k 5 150 python bench.py --steps 3 --warmup 3 --cpu-budget 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-260 gpurun_out/${tag}_bench.json
