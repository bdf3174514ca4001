#!/bin/bash
# round 2, call AE: one `ncu --set full` capture of the reworked decode step (launch 300 of a short bench run)
mkdir -p gpurun_out
timeout -k 5 140 ncu --set full --clock-control none --import-source on -k "regex:gpt_decode_step" -s 300 -c 1 -f -o gpurun_out/r2ae_decode python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fast > gpurun_out/r2ae_decode.log 2>&1; tail -1 gpurun_out/r2ae_decode.log | cut -c1-160
