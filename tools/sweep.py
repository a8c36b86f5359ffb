#!/usr/bin/env python
"""Config 5 of BASELINE.json: cart-pole DDP batch sweep 256 -> 131072 TOTAL instances on N GPUs of one box, with the
achieved HBM GB/s of every kernel next to the roofline.

    python tools/sweep.py --gpus N [--totals 256,1024,...] [--mode fixed|ref] [--out profiles/r2_sweep_nN.json]

Runs bench.py once per total batch (under torch.distributed.run for N > 1, exactly like the driver), strong scaling:
the total is split evenly over the ranks (bench.py --total-batch), so the N = 1, 2, 4, 8 files compare like for like.
Each row: whole-job trajectories/s (device-resident and end to end), ms per solve, and per kernel the average launch
time, its algorithmic GB/s on rank 0 and the fraction of the measured HBM peak that is."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--totals", default="256,1024,4096,16384,65536,131072")
    ap.add_argument("--mode", default="fixed")
    ap.add_argument("--out", default=None)
    ap.add_argument("--port", type=int, default=29533)
    args = ap.parse_args()
    out_path = args.out or os.path.join(ROOT, "gpurun_out", f"sweep_n{args.gpus}_{args.mode}.json")
    rows = []
    for total in (int(t) for t in args.totals.split(",")):
        if total % args.gpus:
            continue
        per_gpu = total // args.gpus
        steps = 10 if per_gpu <= 16384 else 5
        bench = [os.path.join(ROOT, "bench.py"), "--gpus", str(args.gpus), "--total-batch", str(total), "--steps",
                 str(steps), "--warmup", "3", "--no-cpu-baseline", "--mode", args.mode, "--seed", str(total)]
        if args.gpus > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", str(args.port)] + bench
        else:
            cmd = [sys.executable] + bench
        out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if not line:
            print("FAILED", total, out.stderr[-800:], flush=True)
            continue
        d = json.loads(line[-1])
        k = d["roofline"]["kernels"]
        row = {"total_batch": total, "n_gpus": args.gpus, "batch_per_gpu": per_gpu, "traj_per_s": d["value"],
               "ms_per_solve": d["ms_per_step"], "e2e_traj_per_s": d["e2e"]["value"], "scaling": d["scaling"],
               "whole_solve_gbs_per_gpu": d["roofline"]["whole_solve"]["achieved_gbs"],
               "whole_solve_frac": d["roofline"]["whole_solve"]["frac"], "hbm_peak_gbs": d["roofline"]["peak"],
               "kernels": {n: {"ms_per_launch": k[n]["ms_per_launch"], "achieved_gbs": k[n]["achieved_gbs"],
                               "frac": k[n]["frac"]} for n in k},
               "forward_passes_mean": d["work"]["forward_passes_mean"], "iterations_mean": d["work"]["iterations_mean"],
               "gather_u0": d["work"].get("gather_u0"), "clocks": d.get("clocks")}
        rows.append(row)
        print(json.dumps({a: row[a] for a in ("total_batch", "n_gpus", "traj_per_s", "ms_per_solve", "e2e_traj_per_s",
                                                "whole_solve_gbs_per_gpu")}), flush=True)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump({"command": " ".join(sys.argv), "mode": args.mode, "rows": rows}, open(out_path, "w"), indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
