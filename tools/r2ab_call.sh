#!/bin/bash
# round 2, call AB: parity subset + decode bench + timeline of the current build
mkdir -p gpurun_out
(timeout -k 5 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3)
timeout -k 5 200 python tools/decode_bench.py --n-past 300,600,900 40:500:0 2>&1 | tail -3
timeout -k 5 300 python tools/decode_timing.py --sweep 480:40:500 300 900 > gpurun_out/r2ab_timing.txt 2>&1; grep -v "layer5 stamp" gpurun_out/r2ab_timing.txt | grep -E "==" | cut -c1-200
