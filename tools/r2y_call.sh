#!/bin/bash
# round 2, call Y: decode step A/B — base (committed) vs new (LayerNorm butterfly trees + score tasks reduced together) vs new + bulk L2
# prefetch of the next layer's K / V; timeline of the new build; parity subset on the new build
mkdir -p gpurun_out
cp bark.cpp_b200/libbark_b200.so /tmp/new.so
for rep in 1 2; do
  cp bark.cpp_b200/libbark_b200_base.so bark.cpp_b200/libbark_b200.so
  echo "== base =="; timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -2
  cp /tmp/new.so bark.cpp_b200/libbark_b200.so
  echo "== new =="; timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -2
  echo "== new + KV prefetch =="; BARK_B200_KV_PREFETCH=1 timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -2
done
timeout -k 5 300 python tools/decode_timing.py --sweep 480:40:500 300 900 > gpurun_out/r2y_timing.txt 2>&1; grep -v "layer5 stamp" gpurun_out/r2y_timing.txt | cut -c1-200
(BARK_B200_KV_PREFETCH=1 timeout -k 5 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3)
