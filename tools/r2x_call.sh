#!/bin/bash
# round 2, call X: finer stamps inside the row phases and LayerNorm of the decode step (diagnostic build, TM instantiation only)
mkdir -p gpurun_out
timeout -k 5 300 python tools/decode_timing.py --sweep 480:40:500 300 900 2>&1 | grep -v "layer5 stamp" > gpurun_out/r2x_timing.txt; cat gpurun_out/r2x_timing.txt | head -90
