#!/bin/bash
# usage: tools/ncu_one.sh <kernel-regex> <skip> <count> <out-name>   — one `--set full` capture of a kernel inside a short bench run (GPU box)
set -e
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$1" -s "$2" -c "$3" -f -o "gpurun_out/$4" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > "gpurun_out/$4.log" 2>&1 || { tail -5 "gpurun_out/$4.log"; exit 1; }
tail -2 "gpurun_out/$4.log"
