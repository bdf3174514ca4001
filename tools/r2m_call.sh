#!/bin/bash
# round 2, call M: decode head-start sweep per exchange type and CTA-count sweep (timing only; results unchanged)
mkdir -p gpurun_out
: > gpurun_out/r2m_sweep.txt
for hs in 0:2000:500:0:500:0 300:2000:500:0:500:0 0:2000:500:400:500:0 0:2000:500:0:500:1500 300:2000:500:400:500:1500 200:2500:700:300:700:1000 0:1500:300:200:300:800; do
  echo "== HEADSTART $hs" >> gpurun_out/r2m_sweep.txt
  BARK_B200_HEADSTART=$hs timeout -k 5 100 python tools/decode_bench.py --n-past 300,700 40:500:2000 2>&1 | tail -2 >> gpurun_out/r2m_sweep.txt
done
for c in 136 128 112; do
  echo "== CTAS $c" >> gpurun_out/r2m_sweep.txt
  BARK_B200_DECODE_CTAS=$c timeout -k 5 100 python tools/decode_bench.py --n-past 300,700 40:500:2000 2>&1 | tail -2 >> gpurun_out/r2m_sweep.txt
done
cat gpurun_out/r2m_sweep.txt
