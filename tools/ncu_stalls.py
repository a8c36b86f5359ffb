#!/usr/bin/env python
"""Print the warp-stall breakdown and the top stall sites (SASS) of one kernel in an ncu --set full capture.

    python tools/ncu_stalls.py <file.ncu-rep> <kernel-name regex> [top]
"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    his = [i for i, r in enumerate(rows) if "# Samples" in r]
    h = rows[his[0]]
    end = his[1] - 1 if len(his) > 1 else len(rows)
    data = [r for r in rows[his[0] + 1:end] if len(r) == len(h)]
    isamp, isrc, iex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp]) for r in data) or 1
    agg = {}
    for r in data:
        for i in stall:
            try:
                agg[h[i]] = agg.get(h[i], 0) + int(r[i])
            except ValueError:
                pass
    print("kernel:", rows[0][1][:120])
    print("total samples", tot, "static SASS", len(data), "executed(warp-level)", sum(int(r[iex]) for r in data))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k:28s} {v:7d} {100.0 * v / tot:5.1f}%")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
        st = {h[i]: int(r[i]) for i in stall if r[i] not in ("", "0")}
        print(f"{int(r[isamp]):6d} {100.0 * int(r[isamp]) / tot:5.1f}% ex={r[iex]:>8s} {max(st, key=st.get) if st else '':24s} {r[isrc].strip()[:80]}")


if __name__ == "__main__":
    main()
