#!/usr/bin/env python
"""Average device time of the production decode-step kernel (no stamps) under different exchange knobs, GPU box only.

usage: python tools/decode_bench.py [--n-past 300,900] poll_ns:first_ns:att_ns [poll_ns:first_ns:att_ns ...]
One context per setting; 60 single-token steps of the coarse model of the bark-small f16 bench file, timed with the library's
CUDA-event profiler (bark_b200_profile_report).  Writes gpurun_out/decode_bench.json.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("BARK_B200_QUIET", "1")
import bench  # noqa: E402
import __graft_entry__ as graft  # noqa: E402


def main():
    pkg = graft.load_package()
    path = bench.weights_path()
    args = sys.argv[1:]
    pasts = [300, 900]
    if args and args[0] == "--n-past":
        pasts = [int(v) for v in args[1].split(",")]; args = args[2:]
    out = []
    rng = np.random.default_rng(0)
    for item in args or ["40:500:0"]:
        poll, first, att = (int(v) for v in item.split(":"))
        os.environ["BARK_B200_POLL_NS"], os.environ["BARK_B200_POLL_FIRST_NS"], os.environ["BARK_B200_POLL_ATT_NS"] = str(poll), str(first), str(att)
        with pkg.Bark(path) as b:
            for n_past in pasts:
                toks = rng.integers(10000, 12048, n_past).astype(np.int32)
                _, p = b.gpt_eval(1, toks, 0, False)
                for _ in range(5):
                    _, p = b.gpt_eval(1, np.array([10001], np.int32), p, False)
                pkg.profile_enable(True)
                for _ in range(60):
                    _, p = b.gpt_eval(1, np.array([10001], np.int32), p, False)
                rep = pkg.profile_report()
                pkg.profile_enable(False)
                v = rep.get("gpt_decode_step_kernel") or rep["gpt_decode_cluster_kernel"]
                us = v["ms"] * 1e3 / v["launches"]
                out.append(dict(poll_ns=poll, first_ns=first, att_ns=att, n_kv_start=n_past + 6, us_per_token=round(us, 2), launches=v["launches"]))
                print(f"poll {poll:5d} first {first:5d} att {att:5d}  n_kv {n_past + 6:4d}..{p:4d}: {us:7.2f} us per decode step", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "decode_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
