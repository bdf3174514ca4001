#!/bin/bash
# round 2, call Z: decode step with P2 / P3 out of line, score tasks reduced together, V through cp.async into the staging tail, LayerNorm butterfly
mkdir -p gpurun_out
(timeout -k 5 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3)
cp bark.cpp_b200/libbark_b200.so /tmp/new.so
for rep in 1 2; do
  cp bark.cpp_b200/libbark_b200_base.so bark.cpp_b200/libbark_b200.so
  echo "== base =="; timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -2
  cp /tmp/new.so bark.cpp_b200/libbark_b200.so
  echo "== new =="; timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -2
done
echo "== new + KV prefetch =="; BARK_B200_KV_PREFETCH=1 timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -2
timeout -k 5 300 python tools/decode_timing.py --sweep 480:40:500 300 900 > gpurun_out/r2z_timing.txt 2>&1; grep -v "layer5 stamp" gpurun_out/r2z_timing.txt | grep -E "==|scores|V pref|att arr|LN1 mean" | cut -c1-160
