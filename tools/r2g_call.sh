#!/bin/bash
# round 2, call G (2 GPUs): row-sharded fine stage over NVLink peer memory (BASELINE configs[4]) at N = 1 and N = 2; two-device thread test
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1; head -8 gpurun_out/r2g_topo.txt
timeout -k 5 300 python bench.py --config fine_only --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2g_fine_n1.json 2> gpurun_out/r2g_fine_n1.err; cut -c1-300 gpurun_out/r2g_fine_n1.json; tail -3 gpurun_out/r2g_fine_n1.err
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --config fine_only --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2g_fine_n2.json 2> gpurun_out/r2g_fine_n2.err; cut -c1-300 gpurun_out/r2g_fine_n2.json; tail -5 gpurun_out/r2g_fine_n2.err
(timeout -k 5 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "one_host_thread" 2>&1 | tail -4) > gpurun_out/r2g_pytest_threads.log; tail -3 gpurun_out/r2g_pytest_threads.log
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.load(open(f"gpurun_out/r2g_fine_n{n}.json"))
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "parity", d["parity"], "nvlink", d.get("nvlink"))
    except Exception as e:
        print(n, "failed:", e)
PY
