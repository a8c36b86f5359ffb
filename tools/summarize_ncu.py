#!/usr/bin/env python
"""Turn ncu captures (gpurun_out/*.ncu-rep, launch-list CSVs) into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py rep  <file.ncu-rep> <out.md> [traffic.json]   # per kernel: key metrics + stalls
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


def stall_table(path, kernel_regex, top=10):
    """Warp-stall breakdown and top stall sites (SASS) of one kernel of a capture."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel_regex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    his = [i for i, r in enumerate(rows) if "# Samples" in r]
    if not his:
        return []
    h = rows[his[0]]
    end = his[1] - 1 if len(his) > 1 else len(rows)
    data = [r for r in rows[his[0] + 1:end] if len(r) == len(h)]
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
    lines = ["", f"warp-stall samples: {tot}; static SASS instructions: {len(data)}; executed warp instructions: "
             f"{sum(int(r[iex]) for r in data)}", "", "| stall reason | samples | share |", "|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]:
        lines.append(f"| {k} | {v} | {100.0 * v / tot:.1f}% |")
    lines += ["", "top stall sites (SASS):", "", "| samples | share | executed | main reason | instruction |",
              "|---|---|---|---|---|"]
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
        st = {h[i]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
        main = max(st, key=st.get) if st else ""
        lines.append(f"| {r[isamp]} | {100.0 * int(r[isamp]) / tot:.1f}% | {r[iex]} | {main} | `{r[isrc].strip()[:70]}` |")
    return lines


def summarize_rep(path, out_path):
    """One section per kernel result in the capture; also returns {short kernel name: dram bytes per launch}."""
    rows = ncu_csv(path, "raw")
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of `{path.split('/')[-1]}` (--clock-control none)", ""]
    traffic = {}
    seen = set()
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        short = name.split("<")[0].replace("void ", "").split("::")[-1]
        if short in seen:
            continue
        seen.add(short)
        lines += [f"## `{name[:150]}`", "", "| metric | value | unit |", "|---|---|---|"]
        rec = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} | {vals[i]} | {units[i]} |")
                rec[k] = (vals[i], units[i])

        def to_bytes(key):
            if key not in rec:
                return None
            v, u = float(rec[key][0].replace(",", "")), rec[key][1].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

        rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            traffic[short] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr,
                              "duration_us": float(rec["gpu__time_duration.sum"][0].replace(",", "")) *
                              {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(rec["gpu__time_duration.sum"][1].lower(), 1)}
        lines += stall_table(path, short)
        lines.append("")
    open(out_path, "w").write("\n".join(lines) + "\n")
    return traffic


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
    if mode == "rep":
        import json

        t = summarize_rep(src, dst)
        if len(sys.argv) > 4:  # optional: merge the per-kernel DRAM traffic into a json (profiles/dram_traffic.json)
            path = sys.argv[4]
            try:
                old = json.load(open(path))
            except Exception:
                old = {}
            old.update(t)
            json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    else:
        summarize_list(src, dst)
