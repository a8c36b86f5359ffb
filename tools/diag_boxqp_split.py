#!/usr/bin/env python
"""Where and why a control-limited DDP solve on the GPU leaves the reference's iterates (run on the GPU box).

For one golden case of tests/golden/reference_ddp_planar.npz: solve with max_iter = 1, 2, ... on the GPU and with the
CPU oracle (bit-identical to the reference headers on these cases), find the first iteration whose result differs by more
than 1e-6, and print -- for the first horizon step the backward sweep of that iteration gets a different gain at -- the
previous iterate's input next to the limits (the box of that step's QP is [lo - u, hi - u]), and both k / K.

    python tools/diag_boxqp_split.py [case] [instance]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nmpc_b200 as gpu  # noqa: E402
import oracle_lib as O  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "planar_box_cross_fixed"
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_ddp_planar.npz"))
c = {k.split("/", 1)[1]: G[k] for k in G.files if k.startswith(name + "/")}
INT = ("max_iter", "with_input_constraint", "reg_type")
cfg = {k[4:]: (int(v) if k[4:] in INT else float(v)) for k, v in c.items() if k.startswith("cfg_")}
N = cfg["horizon_steps"] = int(c["N"])
lo, hi = c["limits"]
B = len(c["x0"])
ui = np.repeat(c["u_init"][None], B, axis=0)
np.set_printoptions(linewidth=200, precision=17)


def both(mi):
    kw = dict(cfg, max_iter=mi)
    ref = O.ddp_solve_batch("planar_quadrotor", c["params"], O.ddp_config(**kw), c["x0"], ui, u_lo=lo, u_hi=hi)
    s = gpu.DDPSolver("planar_quadrotor", params=c["params"], batch_capacity=B)
    for k, v in kw.items():
        setattr(s.config(), k, bool(v) if k == "with_input_constraint" else v)
    s.setInputLimitsFunc((lo, hi))
    s.solve_batch(0.0, c["x0"], ui)
    out = {"u": s.controlData().u_list.copy(), "k": s.k_list().copy(), "K": s.K_list().copy(), "trace": s.trace().copy()}
    # the oracle stores K column-major (NU x NX per step), K_list() is [B, N, NU, NX]
    ref["K"] = ref["K"].reshape(B, N, out["K"].shape[3], out["K"].shape[2]).transpose(0, 1, 3, 2)
    s.close()
    return ref, out


def rel(a, b):
    ax = tuple(range(1, a.ndim))
    return np.max(np.abs(a - b), axis=ax) / (1.0 + np.max(np.abs(b), axis=ax))


prev = None
done = set()
for mi in range(1, cfg["max_iter"] + 1):
    ref, out = both(mi)
    ru, rk = rel(out["u"], ref["u"]), rel(out["k"], ref["k"])
    rK = rel(out["K"].reshape(B, N, -1), ref["K"].reshape(B, N, -1))
    print(f"max_iter {mi}: rel_u {ru}  rel_k {rk}  rel_K {rK}")
    for b in range(B):
        if b in done or (len(sys.argv) > 2 and b != int(sys.argv[2])):
            continue
        if rK[b] > 1e-6 or rk[b] > 1e-6:
            done.add(b)
            dK = np.abs(out["K"][b].reshape(N, -1) - ref["K"][b].reshape(N, -1)).max(axis=1)
            dk = np.abs(out["k"][b] - ref["k"][b]).max(axis=1)
            # the sweep runs backward: the LAST such step is the first one hit
            i = int(np.max(np.nonzero((dK > 1e-6) | (dk > 1e-6))[0]))
            print(f"  instance {b}: first gain difference at step {i} of iteration {mi} (dK {dK[i]:.3e}, dk {dk[i]:.3e})")
            if prev is not None:
                up_ref, up_gpu = prev[0]["u"][b, i], prev[1]["u"][b, i]
                print("   previous iterate u  (oracle)", up_ref, [float.hex(float(v)) for v in up_ref])
                print("   previous iterate u  (gpu)   ", up_gpu, [float.hex(float(v)) for v in up_gpu])
                print("   box lower  lo - u   (oracle)", lo - up_ref, " (gpu)", lo - up_gpu)
                print("   box upper  hi - u   (oracle)", hi - up_ref, " (gpu)", hi - up_gpu)
            print("   k oracle", ref["k"][b, i], " gpu", out["k"][b, i])
            print("   K oracle", ref["K"][b, i].reshape(-1), "\n   K gpu   ", out["K"][b, i].reshape(-1))
            print("   lambda trace oracle", ref["trace"][b, : mi + 1, 2], " gpu", out["trace"][b, : mi + 1, 2])
    prev = (ref, out)
