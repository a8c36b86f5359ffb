#!/usr/bin/env python
"""Config 5 of BASELINE.json: cart-pole DDP batch sweep on one GPU (run bench.py per batch size)."""
import json
import subprocess
import sys

batches = [int(b) for b in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256, 1024, 4096, 16384, 65536, 131072]
mode = sys.argv[2] if len(sys.argv) > 2 else "fixed"
rows = []
for B in batches:
    steps = 10 if B <= 16384 else 5
    out = subprocess.run([sys.executable, "bench.py", "--batch", str(B), "--steps", str(steps), "--warmup", "3",
                          "--no-cpu-baseline", "--mode", mode, "--seed", str(B)], capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print("FAILED", B, out.stderr[-500:])
        continue
    d = json.loads(line[-1])
    k = d["roofline"]["kernels"]
    row = {"batch": B, "traj_per_s": d["value"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"],
           "whole_solve_frac": d["roofline"]["whole_solve"]["frac"],
           **{f"{n}_ms": k[n]["ms_per_launch"] for n in k}, **{f"{n}_frac": k[n]["frac"] for n in k},
           "fwd_passes": d["work"]["forward_passes_mean"], "iters": d["work"]["iterations_mean"]}
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(f"gpurun_out/sweep_{mode}.json", "w"), indent=1)
