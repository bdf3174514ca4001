#!/bin/bash
# round 2, call K (8 GPUs): row-sharded fine stage at N = 4 and 8 (BASELINE configs[4]); bark-large f16, 8 replicas (BASELINE configs[2])
mkdir -p gpurun_out
run() { # n config steps extra
  timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1)) bench.py --config $2 --gpus $1 --steps $3 --warmup 3 $4 > gpurun_out/r2k_$2_n$1.json 2> gpurun_out/r2k_$2_n$1.err
  cut -c1-200 gpurun_out/r2k_$2_n$1.json; tail -2 gpurun_out/r2k_$2_n$1.err
}
run 8 fine_only 5 ""
run 4 fine_only 5 ""
run 8 large 2 "--no-fast"
python - <<'PY'
import json
for name in ("fine_only_n8", "fine_only_n4", "large_n8"):
    try:
        d = json.load(open(f"gpurun_out/r2k_{name}.json"))
        print(name, "value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "parity", d.get("parity"), "nvlink", d.get("nvlink", {}).get("achieved_GBps_rank0_out"))
    except Exception as e:
        print(name, "failed:", e)
PY
