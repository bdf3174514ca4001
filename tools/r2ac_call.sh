#!/bin/bash
# round 2, call AC: CTA-count sweep of the reworked decode step
mkdir -p gpurun_out
for n in 112 128 136 148; do
  echo "== $n CTAs =="; BARK_B200_DECODE_CTAS=$n timeout -k 5 200 python tools/decode_bench.py --n-past 300,600,900 40:500:0 2>&1 | tail -3
done
echo "== 128 CTAs, fused sampler =="; BARK_B200_FUSE_SAMPLER=1 timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -2
