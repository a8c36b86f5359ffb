#!/usr/bin/env python
"""profiles/r2_sweep_n{1,2,4,8}.json (tools/sweep.py) -> the tables of profiles/r2_sweep.md (printed to stdout)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {n: json.load(open(os.path.join(ROOT, "profiles", f"r2_sweep_n{n}.json")))["rows"] for n in (1, 2, 4, 8)}
hdr = ["| total batch | 1 GPU | 2 GPUs | 4 GPUs | 8 GPUs |", "|---:|---:|---:|---:|---:|"]
out = ["Whole-job trajectories/s, device-resident (`value`); in brackets ms per solve:", ""] + hdr
for i in range(len(rows[1])):
    out.append(f"| {rows[1][i]['total_batch']} | " + " | ".join(
        f"{rows[n][i]['traj_per_s'] / 1e6:.2f} M ({rows[n][i]['ms_per_solve']:.2f})" for n in (1, 2, 4, 8)) + " |")
out += ["", "End to end (host arrays in, first-step controls out, H2D / D2H inside the timed region), M trajectories/s:", ""] + hdr
for i in range(len(rows[1])):
    out.append(f"| {rows[1][i]['total_batch']} | " + " | ".join(f"{rows[n][i]['e2e_traj_per_s'] / 1e6:.2f}" for n in (1, 2, 4, 8)) + " |")
out += ["", f"Compulsory HBM traffic of the whole solve per GPU, GB/s (fraction of the measured peak {rows[1][0]['hbm_peak_gbs']:.0f} GB/s):", ""] + hdr
for i in range(len(rows[1])):
    out.append(f"| {rows[1][i]['total_batch']} | " + " | ".join(
        f"{rows[n][i]['whole_solve_gbs_per_gpu']:.0f} ({rows[n][i]['whole_solve_frac']:.3f})" for n in (1, 2, 4, 8)) + " |")
out += ["", "Per kernel at N = 1 (average launch, compulsory GB/s, fraction of peak):", "",
        "| batch | backward | first line-search candidate | other candidates |", "|---:|---|---|---|"]
for r in rows[1]:
    k = r["kernels"]
    out.append(f"| {r['total_batch']} | " + " | ".join(
        f"{k[n]['ms_per_launch'] * 1e3:.0f} us, {k[n]['achieved_gbs']:.0f} GB/s, {k[n]['frac']:.3f}"
        for n in ("backward", "forward_first", "forward_rest")) + " |")
print("\n".join(out))
