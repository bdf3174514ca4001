#!/usr/bin/env python
"""Per-phase globaltimer stamps of the persistent decode kernel (BARK_B200_DECODE_TIMING=1), GPU box only.

usage: python tools/decode_timing.py [--sweep tid:poll_ns[:first_ns],...] [n_past ...]   (coarse model, bark-small f16 bench file;
       tid = stamping thread (lane 0 of a warp), poll_ns = back-off between polls of the tagged exchange words)
Prints, per n_past: the time between consecutive stamps on CTA 0 (median over layers) and, at layer 5, the spread over CTAs
of each stamp.  The raw [256][32] dump is saved to gpurun_out/decode_timing_<n_kv>.npy.  Stamp ids: decode_kernels.cu tstamp().
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BARK_B200_DECODE_TIMING"] = "1"
os.environ.setdefault("BARK_B200_QUIET", "1")
import bench  # noqa: E402
import __graft_entry__ as graft  # noqa: E402

NAMES = {0: "layer start", 1: "LN1 mean", 2: "LN1 done", 3: "QKV rows ready (mbarrier)", 4: "QKV rows done", 5: "QKV next rows issued", 6: "K prefetched",
         7: "q arrived", 8: "scores done", 9: "V prefetched + v_new", 10: "scores arrived", 11: "max", 12: "exp", 13: "sum", 14: "probabilities", 15: "PV partials",
         16: "P3 done", 17: "att arrived", 18: "c_proj rows ready", 19: "c_proj rows done", 20: "c_proj next issued", 21: "x arrived", 22: "LN2 mean", 23: "LN2 done",
         24: "fc rows ready", 25: "fc rows done", 26: "fc next issued", 27: "fc block sync", 28: "ff arrived", 29: "proj rows ready", 30: "proj rows done", 31: "proj next issued"}


SUMMARY = []


def measure(pkg, path, pasts, tid, poll, first=0):
    os.environ["BARK_B200_DECODE_TIMING_TID"] = str(tid)
    os.environ["BARK_B200_POLL_NS"] = str(poll)
    os.environ["BARK_B200_POLL_FIRST_NS"] = str(first)
    rng = np.random.default_rng(0)
    with pkg.Bark(path) as b:
        L = int(b.hparams(1)[0])
        for n_past in pasts:
            toks = rng.integers(10000, 12048, n_past).astype(np.int32)
            _, p = b.gpt_eval(1, toks, 0, False)
            for _ in range(int(os.environ.get("WARM_STEPS", "48"))):   # warm (the adaptive head starts settle within ~20 tokens), then keep the last step's stamps
                _, p = b.gpt_eval(1, np.array([10001], np.int32), p, False)
            t = np.zeros(256 * 32, np.uint64)
            pkg.lib().bark_b200_decode_timing(b.ctx, t.ctypes.data_as(C.c_void_p), t.size)
            t = t.reshape(256, 32).astype(np.int64)
            ad = np.zeros(148 * 8, np.uint32)
            if pkg.lib().bark_b200_decode_adapt(b.ctx, 1, ad.ctypes.data_as(C.c_void_p), ad.size) > 0:
                ad = ad.reshape(148, 8)
                print("   adaptive head starts (ns), median over CTAs [q, att, x1, ff, x2, scores]:", np.median(ad[:, :6], axis=0).astype(int).tolist(),
                      " soft_max CTAs (0..47):", np.median(ad[:48, :6], axis=0).astype(int).tolist())
            tag = f"tid{tid}_poll{poll}_first{first}"
            np.save(os.path.join(ROOT, "gpurun_out", f"decode_timing_{p}_{tag}.npy"), t)
            lay = t[:L + 1]
            SUMMARY.append(dict(tid=tid, poll_ns=poll, first_ns=first, n_kv=int(p), us_per_layer=float((lay[L, 0] - lay[0, 0]) / 1e3 / L)))
            print(f"== [{tag}] n_kv {p}: {(lay[L, 0] - lay[0, 0]) / 1e3:.1f} us for {L} layers on CTA 0 ({(lay[L, 0] - lay[0, 0]) / 1e3 / L:.2f} us per layer)")
            used = [i for i in range(32) if lay[1, i] != 0]
            for a, c in zip(used[:-1], used[1:]):
                d = (lay[:L, c] - lay[:L, a]) / 1e3
                print(f"   -> {c:2d} {NAMES[c]:<28s} median {np.median(d):6.2f} us   min {d.min():6.2f}   max {d.max():6.2f}")
            d = (lay[1:L + 1, 0] - lay[:L, used[-1]]) / 1e3
            print(f"   ->  0 {'x arrived (next layer)':<28s} median {np.median(d):6.2f} us   min {d.min():6.2f}   max {d.max():6.2f}")
            # finer stamps (tstamp2): rows 32 + layer = the four row phases (8 slots each), rows 48 + layer = LN1 / LN2 (16 slots each)
            for row0, groups, names in ((32, ((0, "QKV"), (8, "c_proj"), (16, "fc"), (24, "proj")), {0: "entry", 1: "sched loaded, loop start", 2: "pair row_dot done", 3: "pair emitted", 4: "single row_dot done", 5: "single emitted"}),
                                        (48, ((0, "LN1"), (16, "LN2")), {0: "entry", 1: "warp sums done", 2: "barrier 1 passed", 3: "tree done", 5: "variance warp sums done", 6: "barrier 2 passed", 7: "scale known", 8: "act written", 9: "final barrier passed"})):
                sub = t[row0:row0 + L]
                for g0, gname in groups:
                    idx = [i for i in sorted(names) if sub[1, g0 + i] != 0]
                    if not idx: continue
                    print(f"   [{gname}] stamps relative to entry (median over layers, us): " + "  ".join(f"{names[i]}={np.median((sub[1:L, g0 + i] - sub[1:L, g0 + idx[0]]) / 1e3):.2f}" for i in idx))
            cta = t[64:64 + 148]
            base = cta[:, 0].min()
            for c in used:
                col = cta[:, c]; col = col[col != 0]
                v = (col - base) / 1e3
                print(f"   layer5 stamp {c:2d} over {len(v):3d} CTAs: min {v.min():7.2f}  median {np.median(v):7.2f}  max {v.max():7.2f} us   {NAMES[c]}")


def main():
    pkg = graft.load_package()
    path = bench.weights_path()
    args = sys.argv[1:]
    sweep = [(int(os.environ.get("BARK_B200_DECODE_TIMING_TID", "0")), int(os.environ.get("BARK_B200_POLL_NS", "40")))]
    if args and args[0] == "--sweep":                        # --sweep tid:poll,tid:poll,...   one context per setting
        sweep = [tuple(int(v) for v in item.split(":")) for item in args[1].split(",")]
        args = args[2:]
    pasts = [int(a) for a in args] or [300, 900]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for item in sweep:
        measure(pkg, path, pasts, *item)
    import json
    json.dump(SUMMARY, open(os.path.join(ROOT, "gpurun_out", "decode_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
