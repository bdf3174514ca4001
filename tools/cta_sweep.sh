#!/bin/bash
# GPU box: decode-step time vs number of CTAs (BARK_B200_DECODE_CTAS), then parity tests at the fastest count if it is not the default.
mkdir -p gpurun_out
: > gpurun_out/cta_sweep.txt
for n in 148 136 128 112 96; do
  echo "== CTAS $n" >> gpurun_out/cta_sweep.txt
  BARK_B200_DECODE_CTAS=$n timeout -k 5 60 python tools/decode_bench.py --n-past 300,700 40:500:2000 2>/dev/null | grep "^poll" >> gpurun_out/cta_sweep.txt
done
cat gpurun_out/cta_sweep.txt
best=$(python - <<'PY'
import re
cur=None; tot={}
for l in open("gpurun_out/cta_sweep.txt"):
    m=re.match(r"== CTAS (\d+)", l)
    if m: cur=int(m.group(1)); tot[cur]=0.0; continue
    m=re.search(r":\s+([\d.]+) us per decode step", l)
    if m and cur is not None: tot[cur]+=float(m.group(1))
tot={k:v for k,v in tot.items() if v>0}
print(min(tot, key=tot.get) if tot else 148)
PY
)
echo "best CTAS $best"
if [ "$best" != "148" ]; then
  (BARK_B200_DECODE_CTAS=$best timeout -k 5 150 python -m pytest tests -m gpu -q -x -k "teacher_forced or full_size or bark_large" 2>&1 | tail -3) > gpurun_out/cta_pytest.log; cat gpurun_out/cta_pytest.log
fi
