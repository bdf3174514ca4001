#!/bin/bash
# One GPU-box call: decode-kernel knob sweep -> pick the fastest -> sanity tests, bench, ncu launch list, ncu full capture of the decode step.
# usage (GPU box): tools/final_run.sh <tag>
tag=${1:-r1k}
mkdir -p gpurun_out
timeout -k 5 150 python tools/decode_timing.py --sweep 0:40:0,480:40:0,480:150:0,480:400:0,480:1000:0,480:40:500,480:40:1000,480:150:1000,480:40:1500 300 > gpurun_out/${tag}_sweep.txt 2>&1
grep "^==" gpurun_out/${tag}_sweep.txt
python - <<'PY' > gpurun_out/best_knobs.env
import json
try:
    d = [e for e in json.load(open("gpurun_out/decode_sweep.json")) if e["tid"] == 480]
    b = min(d, key=lambda e: e["us_per_layer"])
    print(f"export BARK_B200_POLL_NS={b['poll_ns']} BARK_B200_POLL_FIRST_NS={b['first_ns']}")
except Exception as ex:
    print("# sweep failed:", ex)
PY
cat gpurun_out/best_knobs.env
source gpurun_out/best_knobs.env
(timeout -k 5 200 python -m pytest tests -m gpu -q -x -k "teacher_forced or tokens_bit_exact or prefix_reuse" 2>&1 | tail -5) > gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_pytest.log
timeout -k 5 200 python bench.py --steps 3 --warmup 3 --cpu-budget 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-260 gpurun_out/${tag}_bench.json
timeout -k 5 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; wc -l gpurun_out/launches_${tag}.csv
timeout -k 5 200 tools/ncu_one.sh gpt_decode_step 300 1 decode_${tag}
