#!/usr/bin/env python
"""Turn ncu captures (gpurun_out/*.ncu-rep, launch-list CSVs) into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py rep  <file.ncu-rep> <out.md>      # key metrics + top stall sites of one capture
    python tools/summarize_ncu.py list <launches.csv> <out.md>      # per-kernel totals / shares of a launch list
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def ncu_csv(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarize_rep(path, out_path):
    rows = ncu_csv(path, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# ncu --set full summary of `{path.split('/')[-1]}`", ""]
    name = vals[hdr.index("Kernel Name")]
    lines += [f"kernel: `{name[:160]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {vals[i]} | {units[i]} |")
    src = ncu_csv(path, "source")
    h = src[1]
    data = src[2:]
    isamp, isrc, iex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp]) for r in data) or 1
    agg = {}
    for r in data:
        for i in stall_cols:
            try:
                agg[h[i]] = agg.get(h[i], 0) + int(r[i])
            except ValueError:
                pass
    lines += ["", f"warp-stall samples: {tot}; static SASS instructions: {len(data)}", "",
              "| stall reason | samples | share |", "|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        lines.append(f"| {k} | {v} | {100.0 * v / tot:.1f}% |")
    lines += ["", "top stall sites (SASS):", "", "| samples | share | executed | main reason | instruction |",
              "|---|---|---|---|---|"]
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:12]:
        st = {h[i]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
        main = max(st, key=st.get) if st else ""
        lines.append(f"| {r[isamp]} | {100.0 * int(r[isamp]) / tot:.1f}% | {r[iex]} | {main} | `{r[isrc].strip()[:70]}` |")
    open(out_path, "w").write("\n".join(lines) + "\n")


def summarize_list(path, out_path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        launches[int(r[iid])] = (r[ik], float(r[iv].replace(",", "")))
    agg = {}
    for _, (k, ns) in sorted(launches.items()):
        short = k.split("(")[0]
        short = short.replace("void ", "").replace("nmpc_b200::", "")
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values()) or 1.0
    lines = [f"# ncu launch list `{path.split('/')[-1]}` (gpu__time_duration.sum, --clock-control none)", "",
             "per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes", "",
             "| kernel | launches | total us | share | mean us |", "|---|---|---|---|---|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:90]}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot:.1f}% | {ns / 1e3 / n:.1f} |")
    lines += ["", "launch order (us): " + ", ".join(f"{ns / 1e3:.0f}" for _, (_, ns) in sorted(launches.items()))]
    open(out_path, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    (summarize_rep if mode == "rep" else summarize_list)(src, dst)
