#!/usr/bin/env python
"""The headline bench (cart-pole DDP, B = 4096, M-fixed) with each kernel variant pinned through its tuning knob's
environment preset: the numbers behind DESIGN.md's "measured and not adopted" table.

    python tools/variants_bench.py [out.json]          (one B200; ~10 s per variant)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [
    ("default (4 lanes per instance + smem exchange; split-role line search)", {}),
    ("backward: 4 lanes per instance, warp-shuffle exchange", {"NMPC_B200_BACKWARD_LANES": "2"}),
    ("backward: 4 lanes per instance, two tiles per CTA", {"NMPC_B200_BACKWARD_LANES_TILES_PER_CTA": "2"}),
    ("backward: thread per instance, fused K1+K2 (round 1's kernel)", {"NMPC_B200_BACKWARD_LANES": "0"}),
    ("backward: three-kernel pipeline (K1, K2 with TMA ring)", {"NMPC_B200_BACKWARD_LANES": "0", "NMPC_B200_BACKWARD_FUSED": "0"}),
    ("line search: phased without split roles (round 1's kernels)", {"NMPC_B200_FORWARD_SPLIT": "0"}),
    ("line search: 16 lanes speculate in one kernel", {"NMPC_B200_FORWARD_LANES": "16"}),
    ("whole solve in one persistent kernel per tile", {"NMPC_B200_SOLVE_TILE": "1"}),
]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "variants.json")
    rows = []
    for name, env in VARIANTS:
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline"],
                           capture_output=True, text=True, env=e, cwd=ROOT)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if not line:
            rows.append({"variant": name, "env": env, "error": r.stderr[-400:]})
            print("FAILED", name, flush=True)
            continue
        d = json.loads(line[-1])
        k = d["roofline"]["kernels"]
        row = {"variant": name, "env": env, "traj_per_s": d["value"], "ms_per_solve": d["ms_per_step"],
               "kernel_us": {n: round(v["ms_per_launch"] * 1e3, 2) for n, v in k.items()}}
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump({"command": "python tools/variants_bench.py", "workload": "bench.py default (cart-pole DDP, B=4096, N=100, 10 iterations, M-fixed)",
               "rows": rows}, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
