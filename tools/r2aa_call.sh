#!/bin/bash
# round 2, call AA: scores head start sweep on the new decode kernel, latency microbench (FP64 / shuffle / barrier), timeline with the P2 stamp fixed
mkdir -p gpurun_out
nvcc -O3 -arch=sm_100a -o /tmp/lat tools/microbench/lat.cu && /tmp/lat > gpurun_out/r2aa_lat.txt; cat gpurun_out/r2aa_lat.txt
for hs in 0:2000:500:400:500:0 0:2000:500:400:500:1000 0:2000:500:400:500:1500 0:2000:500:400:500:2000 300:2000:500:400:500:1500 0:2500:500:400:500:1500 0:2000:300:400:300:1500 0:2000:700:600:700:1500; do
  echo "== headstart $hs =="; BARK_B200_HEADSTART=$hs timeout -k 5 200 python tools/decode_bench.py --n-past 300,600,900 40:500:0 2>&1 | tail -3
done
timeout -k 5 300 python tools/decode_timing.py --sweep 480:40:500 300 900 > gpurun_out/r2aa_timing.txt 2>&1; grep -v "layer5 stamp" gpurun_out/r2aa_timing.txt | grep -E "==|\[LN1\]|\[QKV\]" | cut -c1-330
