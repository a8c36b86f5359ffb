#!/usr/bin/env python
"""Time and check kernel variants of the cart-pole DDP solve on one GPU (round-2 experiments).

    python tools/exp_variants.py [--batch 4096] [--out gpurun_out/exp_variants.json] name=ENV1:V1,ENV2:V2 ...

Every variant is an environment setting read by the engine when a solver is created.  For each: parity of a
256-instance M-ref solve and a 256-instance M-fixed solve against the CPU oracle, then device-resident timing of
the B-instance M-fixed solve (CUDA-event stage timers of the engine, median of 7 after 3 warm-ups).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nmpc_b200  # noqa: E402
import oracle_lib as O  # noqa: E402


def rel_u(a, b):
    return float((np.max(np.abs(a - b), axis=(1, 2)) / (1.0 + np.max(np.abs(b), axis=(1, 2)))).max())


def parity(mode, ref_cache):
    B, N = 256, 100
    kw = dict(max_iter=10)
    if mode == "fixed":
        kw.update(k_rel_norm_thre=0.0, cost_update_thre=0.0)
    p = O.default_params("cartpole")
    x0, u0 = O.cartpole_x0(B, 0), np.zeros((B, N, 1))
    if mode not in ref_cache:
        ref_cache[mode] = O.ddp_solve_batch("cartpole", p, O.ddp_config(horizon_steps=N, **kw), x0, u0)
    ref = ref_cache[mode]
    s = nmpc_b200.DDPSolver("cartpole", params=p, batch_capacity=B)
    for k, v in kw.items():
        setattr(s.config(), k, v)
    s.solve_batch(0.0, x0, u0)
    out = {"rel_du": rel_u(s.controlData().u_list, ref["u"]),
           "rel_dcost": float(np.max(np.abs(s.cost() - ref["cost"]) / np.abs(ref["cost"]))),
           "iters_equal": bool(np.array_equal(s.iterations(), ref["iters"])),
           "n_fwd_equal": bool(np.array_equal(s.n_forward(), ref["n_fwd"])),
           "n_bwd_equal": bool(np.array_equal(s.n_backward(), ref["n_bwd"])),
           "status_equal": bool(np.array_equal(s.status(), ref["status"]))}
    s.close()
    return out


def timing(B, N=100, mode="fixed"):
    import torch

    p = O.default_params("cartpole")
    x0 = torch.from_numpy(O.cartpole_x0(B, 0)).cuda()
    u0 = torch.zeros((B, N, 1), dtype=torch.float64, device="cuda")
    s = nmpc_b200.DDPSolver("cartpole", params=p, batch_capacity=B)
    c = s.config()
    c.horizon_steps, c.max_iter = N, 10
    if mode == "fixed":
        c.k_rel_norm_thre, c.cost_update_thre = 0.0, 0.0
    st = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    s.enable_timing(True)
    rows = []
    with torch.cuda.stream(st):
        for r in range(10):
            flush.fill_(1)
            s.solve_batch(0.0, x0, u0, stream=st, read_status=False)
            d = s.computationDuration()
            if r >= 3:
                rows.append([d[k] for k in ("solve", "setup", "backward", "forward")])
    med = np.median(np.array(rows), axis=0)
    out = {"batch": B, "horizon": N, "mode": mode, "iters_mean": float(s.iterations().mean()), "solve_ms": float(med[0]), "setup_ms": float(med[1]), "backward_ms": float(med[2]),
           "forward_ms": float(med[3]), "traj_per_s": B / (med[0] * 1e-3),
           "n_fwd_mean": float(s.n_forward().mean()), "n_bwd_mean": float(s.n_backward().mean())}
    s.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[4096])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "exp_variants.json"))
    ap.add_argument("--horizon", type=int, nargs="+", default=[100])
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--mode", default="fixed", choices=["fixed", "ref"])
    ap.add_argument("variants", nargs="*")
    args = ap.parse_args()
    variants = args.variants or ["default="]
    results, ref_cache = {}, {}
    base_env = dict(os.environ)
    for v in variants:
        name, _, envs = v.partition("=")
        os.environ.clear()
        os.environ.update(base_env)
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition(":")
            os.environ[k] = val
        res = {"env": envs}
        try:
            if not args.no_parity:
                res["parity_ref"] = parity("ref", ref_cache)
                res["parity_fixed"] = parity("fixed", ref_cache)
            res["timing"] = [timing(B, N, args.mode) for B in args.batch for N in args.horizon]
        except Exception as e:  # keep going: one broken variant must not hide the others
            res["error"] = repr(e)[:300]
        results[name] = res
        print(name, json.dumps(res), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
