#!/bin/bash
# round 2, call R: ncu source-level capture of the parity tiled GEMM (fine pass fc mat-mul) and of the tiled attention kernels
mkdir -p gpurun_out
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k "regex:lane_gemm_tiled|attn_pv_tiled|attn_scores_tiled" -s 420 -c 7 -f -o gpurun_out/r2r_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fast > gpurun_out/r2r_gemm.log 2>&1; tail -1 gpurun_out/r2r_gemm.log | cut -c1-100
