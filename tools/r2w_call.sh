#!/bin/bash
# round 2, call W: A/B of the decode kernel with P2 / P3 out of line (new) against the committed build (base), then parity subset on new
mkdir -p gpurun_out
for rep in 1 2; do
  echo "== new =="; timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -3
  cp bark.cpp_b200/libbark_b200.so /tmp/new.so; cp bark.cpp_b200/libbark_b200_base.so bark.cpp_b200/libbark_b200.so
  echo "== base =="; timeout -k 5 200 python tools/decode_bench.py --n-past 300,900 40:500:0 2>&1 | tail -3
  cp /tmp/new.so bark.cpp_b200/libbark_b200.so
done
(timeout -k 5 900 python -m pytest tests/test_parity_gpu.py tests/test_true_size_gpu.py -m gpu -q -x 2>&1 | tail -4)
