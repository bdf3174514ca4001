#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  launches <csv> <out.md>      per-kernel launch count / total time / share from a
                               `--metrics gpu__time_duration.sum --csv` launch list
  full <ncu-rep> <out.md>      key metrics of every launch in a `--set full` report (read with `ncu -i ... --page raw --csv`)
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]

def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)

def launches(path, out):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum": continue
        a = agg[short(r["Kernel Name"])]; a[0] += 1; a[1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list: {path}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes)\n\n")
        f.write(f"total {tot/1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches\n\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]/1e6:.3f} | {v[1]/v[0]/1e3:.1f} | {100*v[1]/tot:.1f}% |\n")

def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full: {path}\n\n")
        for r in rows:
            f.write(f"## launch {r[idx['ID']]}: `{short(r[idx['Kernel Name']])}` grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx: f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")

def traffic(path, out):
    """dram bytes per launch of the first kernel in a `--set full` report -> small JSON that bench.py reports as roofline.traffic"""
    import json
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, r = rd[0], rd[1], rd[2]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    def val(k): return float(r[idx[k]].replace(",", "")) * scale[units[idx[k]]]
    d = {"kernel": short(r[idx["Kernel Name"]]), "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
         "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]].replace(",", "")), "source": path,
         "how": "ncu --set full --clock-control none, one launch inside `bench.py --steps 1 --warmup 1` (tools/ncu_one.sh)"}
    d["traffic_bytes"] = d["dram_bytes_read"] + d["dram_bytes_write"]
    json.dump(d, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
