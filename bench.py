#!/usr/bin/env python
"""Headline benchmark of the batched DDP / FMPC hot path on N B200s (BASELINE.json metric and configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config ddp|quadrotor|fmpc] [--batch B | --total-batch B] [--mode fixed|ref]

  --config ddp        BASELINE.json configs[1]: cart-pole DDP, n_x=4 n_u=1, horizon 100, 10 iterations, fp64 (default;
                      the configuration the metric is quoted on), 4096 instances per GPU
  --config quadrotor  configs[3]: quadrotor iLQR, n_x=12 n_u=4, horizon 50, 10 iterations, fp32, 8192 instances
  --config fmpc       configs[2]: FMPC cart-pole with box constraints (PDIP + Riccati), horizon 100, 10 iterations,
                      1024 instances
  --total-batch B     configs[4] (strong scaling): B instances split over the --gpus ranks instead of --batch per GPU

A "step" is one batched solve() of the workload's instances.  One process per GPU (torchrun sets RANK / LOCAL_RANK /
WORLD_SIZE); the batch shards with no data-path collective; the only exchange is the gather of the first-step
controls on rank 0 after each solve -- by default one-sided: every rank's gather kernel stores its rows into rank 0's
buffer over NVLink and raises a flag word (--gather peer; --gather nccl is the all-gather, --no-gather none).  Rank 0
prints ONE JSON line; DESIGN.md section 5 defines every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# Same setting torchrun applies to every rank of an N > 1 run: without it the BLAS / OpenMP worker pools that numpy and
# torch start at import compete with the thread that enqueues the solve.  The CPU arms pass their thread count to the
# oracle explicitly (host_threads()).
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MAX_ITER = 10
WORKLOADS = {
    "ddp": dict(kind="ddp", model="cartpole", nx=4, nu=1, ng=0, N=100, batch=4096, dtype="f64", scalar=8,
                metric="DDP trajectories/sec (batch, T=100, 10 iters)", unit="trajectories/s", seed=0,
                text="cart-pole DDP (n_x=4, n_u=1)"),
    "quadrotor": dict(kind="ddp", model="quadrotor", nx=12, nu=4, ng=0, N=50, batch=8192, dtype="f32", scalar=4,
                      metric="iLQR trajectories/sec (quadrotor, batch, T=50, 10 iters, fp32)", unit="trajectories/s",
                      seed=4, text="quadrotor iLQR (n_x=12, n_u=4)"),
    "fmpc": dict(kind="fmpc", model="cartpole", oracle_model="fmpc_cartpole", nx=4, nu=1, ng=4, N=100, batch=1024, dtype="f64", scalar=8,
                 metric="FMPC solves/sec (cart-pole with box constraints, batch, T=100, 10 iters)", unit="solves/s",
                 seed=3, text="FMPC cart-pole (n_x=4, n_u=1, n_g=4; PDIP + Riccati)"),
}


# ------------------------------------------------------------------------------------------------ byte models
def ddp_elements(nx, nu, N):
    """Scalar elements per instance.  D0..D3: SURVEY.md 8(d) / BASELINE.md 4, the three-stage model (K1 writes the
    derivative blocks, K2 reads them back).  The engine's K1+K2 never materialises the blocks, so its COMPULSORY traffic
    per sweep is `fused_bwd`: read (x, u), write (k, K) and the per-instance scalars.  `fan_read` / `fan_write`: one
    listed instance of the second line-search phase reads its operands once for all candidates; only a winner is
    written."""
    blk = 2 * nx * nx + 2 * nx * nu + nx + nu + nu * nu
    traj = (N + 1) * nx + N * nu
    gains = N * nu + N * nu * nx
    return {
        "blk": blk,
        "D0": (nx + N * nu) + (N + 1) * (nx + 1),
        "D1": (N + 1) * nx + N * nu + N * blk + nx + nx * nx,
        "D2": N * blk + nx + nx * nx + 2 * N * nu + N * nu * nx + 4,
        "D3": 2 * (N + 1) * nx + 3 * N * nu + N * nu * nx + (N + 1),
        "fused_bwd": traj + gains + 8,
        "fan_read": traj + gains,
        "fan_write": traj + (N + 1),
    }


# FMPC, SURVEY.md App. D (n_x=4, n_u=1, n_g=4, N=100): elements per instance and iteration of F1 .. F4
FMPC_ELEMENTS = {"coeff": 9632, "backward": 11244, "forward": 9828, "update": 5124}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU
    baseline: the thread count is passed to the oracle explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """Synthetic inputs and the CPU arm of one BASELINE.json config.  Nothing here touches the GPU."""

    def __init__(self, name, mode):
        self.name, self.mode = name, mode
        self.w = WORKLOADS[name]
        import oracle_lib as O  # synthetic-input generator + the CPU legs only

        self.O = O
        self.oracle_model = self.w.get("oracle_model", self.w["model"])
        self.params = O.default_params(self.oracle_model)

    # ---- configuration shared by the GPU solver and the oracle
    def config_kw(self):
        kw = dict(max_iter=MAX_ITER, horizon_steps=self.w["N"])
        if self.w["kind"] == "ddp" and self.mode == "fixed":
            # M-fixed: thresholds 0 => every instance runs exactly 10 iterations
            kw.update(k_rel_norm_thre=0.0, cost_update_thre=0.0)
        return kw

    def inputs(self, B, seed):
        w, N = self.w, self.w["N"]
        if self.name == "quadrotor":
            rng = np.random.default_rng(seed)
            x0 = np.concatenate([rng.uniform(-1, 1, (B, 3)), rng.uniform(-0.5, 0.5, (B, 3)), rng.uniform(-1, 1, (B, 3)),
                                 rng.uniform(-1, 1, (B, 3))], axis=1)
            u = np.zeros((B, N, 4))
            u[:, :, 0] = 9.80665  # hover thrust, m = 1
            return x0, u
        return self.O.cartpole_x0(B, seed), np.zeros((B, N, w["nu"]))

    def fmpc_variable(self, B):
        """Variable.reset(0, 0, 0, 1, 1)."""
        w, N = self.w, self.w["N"]
        return {"x": np.zeros((B, N + 1, w["nx"])), "u": np.zeros((B, N, w["nu"])), "lambda": np.zeros((B, N + 1, w["nx"])),
                "s": np.ones((B, N, w["ng"])), "nu": np.ones((B, N, w["ng"]))}

    # ---- CPU arm: the oracle port of the reference's algorithm on `threads` host threads
    def cpu_solve(self, x0, u_init, threads):
        O = self.O
        if self.w["kind"] == "fmpc":
            cfg = O.fmpc_config(**self.config_kw())
            O.fmpc_solve_batch(self.oracle_model, self.params, cfg, x0, self.fmpc_variable(len(x0)), nthreads=threads,
                               native=True)
        else:
            cfg = O.ddp_config(**self.config_kw())
            O.ddp_solve_batch(self.oracle_model, self.params, cfg, x0, u_init, native=True, outputs=False, nthreads=threads)

    def cpu_baseline(self, B, seed, min_seconds):
        threads = host_threads()
        x0, u = self.inputs(B, seed)
        self.cpu_solve(x0[:64], u[:64], threads)
        done, t0 = 0, time.perf_counter()
        while True:
            self.cpu_solve(x0, u, threads)
            done += B
            el = time.perf_counter() - t0
            if el >= min_seconds:
                break
        n1 = min(128, B)  # single-thread latency of one solve
        t1 = time.perf_counter()
        self.cpu_solve(x0[:n1], u[:n1], 1)
        ms1 = 1e3 * (time.perf_counter() - t1) / n1
        return {"value": done / el, "unit": self.w["unit"], "cores": int(threads), "kind": "port",
                "per_core": done / el / max(int(threads), 1), "single_thread_ms_per_solve": ms1,
                "sample": f"{done} solves ({done // B} x the B={B} workload, mode {self.mode}) in {el:.1f} s, oracle/ built "
                          f"-O3 -march=native -fopenmp, one solver object per thread"}


def reference_headers_sample(O, params, cfg, x0, threads, N, nu, n=256):
    """Informational: the reference's OWN DDPSolver.hpp (oracle/_ref/libnmpc_ref_fast.so, compiled unmodified against
    the heap-backed Eigen stand-in of oracle/ref/eigen_shim) on a small sample.  The stand-in allocates every
    temporary, so this under-states the reference with real Eigen; the headline CPU number is the faster port."""
    import ctypes as C

    path = os.path.join(ROOT, "oracle", "_ref", "libnmpc_ref_fast.so")
    if not os.path.exists(path):
        return None
    try:
        lib = C.CDLL(path)
        n = min(n, len(x0))
        xs = np.ascontiguousarray(x0[:n])
        us = np.zeros((n, N, nu))
        u0, cost, it = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        t0 = time.perf_counter()
        rc = lib.ref_ddp_solve_cartpole_batch(vp(params), C.byref(cfg), n, C.c_double(0.0), vp(xs), vp(us), int(threads),
                                              vp(u0), vp(cost), vp(it))
        el = time.perf_counter() - t0
        return {"value": n / el, "unit": "trajectories/s", "cores": int(threads), "rc": int(rc),
                "iterations_mean": float(it.mean()),
                "sample": f"{n} instances, reference headers + Eigen stand-in (heap-backed), -O3 -march=x86-64-v3 -fopenmp"}
    except Exception as e:  # informational leg only
        return {"error": str(e)[:200]}


def workload_config(args, wl, world, B):
    w = wl.w
    total = B * world
    return {
        "workload": f"{w['text']} batch={B}/GPU x {world} GPU (total {total}), horizon={w['N']}, {MAX_ITER} iterations, "
                    f"{w['dtype']}, mode M-{args.mode}" if w["kind"] == "ddp" else
                    f"{w['text']} batch={B}/GPU x {world} GPU (total {total}), horizon={w['N']}, {MAX_ITER} iterations "
                    f"(fixed), {w['dtype']}, Variable.reset(0,0,0,1,1)",
        "config": args.config, "batch_per_gpu": B, "total_batch": total, "horizon": w["N"], "max_iter": MAX_ITER,
        "mode": args.mode, "seed": args.seed,
        "parallelism": f"batch-sharded x{world}, no data-path collective",
        "gather_u0": bool(world > 1 and not args.no_gather),
        "l2": "L2 flushed between timed steps (256 MiB write)",
    }


def run_reference(args, wl, rank, world, B):
    """--impl reference: the reference's CPU algorithm (oracle port; the Eigen reference cannot be built here,
    DESIGN.md) on this box's host cores.  Rank 0 only."""
    if rank != 0:
        return
    w = wl.w
    sample = min(B, args.ref_sample)  # bounded sample per step so that steps x sample stays within minutes
    x0, u_init = wl.inputs(B, args.seed)
    x0, u_init = x0[:sample], u_init[:sample]
    threads = host_threads()
    for _ in range(max(args.warmup, 1)):
        wl.cpu_solve(x0, u_init, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wl.cpu_solve(x0, u_init, threads)
    el = time.perf_counter() - t0
    value = sample * args.steps / el
    shim = None
    if args.config == "ddp":
        shim = reference_headers_sample(wl.O, wl.params, wl.O.ddp_config(**wl.config_kw()), x0, threads, w["N"], w["nu"])
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.total_batch else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, wl, world, B),
        "cpu_baseline": {"value": value, "unit": w["unit"], "cores": int(threads), "kind": "port",
                         "sample": f"{sample} of the {B} instances per step, {args.steps} steps, all host threads"},
        "e2e": {"value": value, "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_headers_shim": shim,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arms
def ddp_roofline(wl, B, value_per_gpu, solver, iter_ms, launches, peak, peak_src):
    """Per-kernel durations (CUDA events of the engine, median over the timing solves) against two byte models."""
    w = wl.w
    el = ddp_elements(w["nx"], w["nu"], w["N"])
    sz = w["scalar"]
    iters, n_fwd, n_bwd = solver.iterations(), solver.n_forward(), solver.n_backward()
    tr = solver.trace()  # [B, max_iter + 1, 9]
    alpha0 = float(solver.config().alpha_list[0])
    rows = np.arange(1, tr.shape[1])[None, :] <= iters[:, None]
    # the line-search rows that went to the second phase: alpha_list[0] was not accepted (n_alpha > 1)
    listed = rows & (tr[:, 1:, 4] != alpha0) & (tr[:, 1:, 4] != 0.0)
    winners = listed & (tr[:, 1:, 1] != tr[:, :-1, 1])
    n_iter_launch = max(int(iters.max()), 1)
    # an iteration that ends in the small-gradient test runs no line search (M-ref): n_fwd may be below iters
    first_passes = np.minimum(iters, n_fwd)
    fused = launches["derivative"] == 0
    per = {k: float(np.mean(iter_ms[1:n_iter_launch + 1, c])) for k, c in
           (("derivative", 0), ("backward", 1), ("forward_first", 2), ("forward_rest", 3))}
    if fused:
        per["derivative"] = 0.0
    # beyond forward_phased_max_batch the line search is ONE kernel (forward_kernel): the event between the two phases
    # then only brackets a gap, so the whole line search (time and bytes) is reported under forward_first
    single_forward = not (solver.get_tuning("forward_lanes") in (-1, 3) and B <= solver.get_tuning("forward_phased_max_batch"))
    if single_forward:
        per["forward_first"] += per["forward_rest"]
        per["forward_rest"] = 0.0
    total_ms = sum(per.values()) or 1.0
    # algorithmic bytes per launch (all B instances of this GPU), averaged over the launches of one solve
    compulsory = {
        "derivative": sz * el["D1"] * float(iters.sum()) / n_iter_launch,
        "backward": sz * (el["fused_bwd"] if fused else el["D2"]) * float(n_bwd.sum()) / n_iter_launch,
        "forward_first": sz * el["D3"] * float(first_passes.sum()) / n_iter_launch,
        "forward_rest": sz * (el["fan_read"] * float(listed.sum()) + el["fan_write"] * float(winners.sum())) / n_iter_launch,
    }
    if single_forward:
        compulsory["forward_first"] += compulsory["forward_rest"]
        compulsory["forward_rest"] = 0.0
    survey = {
        "derivative": compulsory["derivative"],
        "backward": sz * ((el["D1"] + el["D2"]) if fused else el["D2"]) * float(n_bwd.sum()) / n_iter_launch,
        "forward_first": compulsory["forward_first"],
        "forward_rest": sz * el["D3"] * float((n_fwd - first_passes).sum()) / n_iter_launch,
    }
    if single_forward:
        survey["forward_first"] += survey["forward_rest"]
        survey["forward_rest"] = 0.0
    names = {
        "derivative": "ddp::linearize_kernel",
        "backward": "ddp::backward_lanes_kernel / backward_fused_kernel (K1 + K2 in one kernel: producer warps linearise, "
                    "consumer warps sweep)" if fused else "ddp::backward_kernel",
        "forward_first": "ddp::forward_kernel (the whole line search in one kernel)" if single_forward else
                         "ddp::forward_first_split_kernel / forward_first_kernel (alpha_list[0] of every instance)",
        "forward_rest": "ddp::forward_fanout_split_kernel / forward_fanout_kernel (other candidates of the listed instances)",
    }
    kernels = {}
    for k, ms in per.items():
        if ms <= 0.0 or (k == "derivative" and fused):  # fused: no K1 launch, the slot only holds the event gap
            continue
        ach = compulsory[k] / (ms * 1e-3) / 1e9
        kernels[k] = {"kernel": names[k], "ms_per_launch": ms, "launches_per_step": n_iter_launch,
                      "alg_bytes_per_launch": compulsory[k], "achieved_gbs": ach, "frac": ach / peak,
                      "survey_3stage_bytes_per_launch": survey[k],
                      "frac_survey_3stage_model": survey[k] / (ms * 1e-3) / 1e9 / peak,
                      "share_of_step": ms / total_ms}
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_launch"])
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath) and wl.name == "ddp" and B == 4096 and wl.mode == "fixed":
        try:
            tj = json.load(open(tpath))
            key = {"backward": "backward", "forward_first": "forward_first", "forward_rest": "forward_rest"}[dominant]
            traffic = float(tj["kernels"][key]["dram_bytes_per_launch"])
            traffic_src = f"ncu --set full of this command, build {tj.get('commit', '?')} ({tj.get('file', '?')})"
        except Exception:
            traffic = None
    traj_comp = sz * (el["D0"] + (el["fused_bwd"] if fused else el["D1"] + el["D2"]) * float(n_bwd.sum()) / B
                      + el["D3"] * float(first_passes.sum()) / B
                      + (el["fan_read"] * float(listed.sum()) + el["fan_write"] * float(winners.sum())) / B)
    traj_survey = sz * (el["D0"] + (el["D1"] + el["D2"]) * float(n_bwd.sum()) / B + el["D3"] * float(n_fwd.sum()) / B)
    return {
        "bound": "hbm", "kernel": names[dominant], "achieved": kernels[dominant]["achieved_gbs"], "peak": peak,
        "unit": "GB/s", "frac": kernels[dominant]["frac"], "traffic": traffic, "traffic_source": traffic_src,
        "model": "achieved = COMPULSORY bytes of the kernel as built (K1+K2 fused: read x, u; write k, K; scalars -- the "
                 "derivative blocks never leave the SM) / its average launch duration (CUDA events inside the engine); "
                 "frac_survey_3stage_model charges the SURVEY 8(d) three-stage bytes (D1 + D2 per fused sweep) to the same "
                 "durations and exceeds the compulsory fraction by the traffic the fusion removed",
        "peak_source": peak_src, "kernels": kernels,
        "whole_solve": {"alg_bytes_per_trajectory": traj_comp, "achieved_gbs": traj_comp * value_per_gpu / 1e9,
                        "frac": traj_comp * value_per_gpu / 1e9 / peak,
                        "survey_3stage_bytes_per_trajectory": traj_survey,
                        "frac_survey_3stage_model": traj_survey * value_per_gpu / 1e9 / peak},
    }


def fmpc_roofline(B, value_per_gpu, dur, peak, peak_src):
    stages = ("coeff", "backward", "forward", "update")
    n = {k: max(int(dur["launches"][k]), 1) for k in stages}
    kernels = {}
    tot = sum(dur[k] for k in stages) or 1.0
    names = {"coeff": "fmpc::fmpc_coeff_kernel (F1)", "backward": "fmpc::fmpc_backward_kernel (F2: Riccati sweep)",
             "forward": "fmpc::fmpc_forward_kernel (F3)", "update": "fmpc::fmpc_update_kernel (F4)"}
    for k in stages:
        ms = dur[k] / n[k]
        if ms <= 0:
            continue
        by = 8.0 * FMPC_ELEMENTS[k] * B
        ach = by / (ms * 1e-3) / 1e9
        kernels[k] = {"kernel": names[k], "ms_per_launch": ms, "launches_per_step": n[k], "alg_bytes_per_launch": by,
                      "achieved_gbs": ach, "frac": ach / peak, "share_of_step": dur[k] / tot}
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"])
    per_solve = 8.0 * sum(FMPC_ELEMENTS.values()) * MAX_ITER
    return {"bound": "hbm", "kernel": names[dominant], "achieved": kernels[dominant]["achieved_gbs"], "peak": peak,
            "unit": "GB/s", "frac": kernels[dominant]["frac"], "traffic": None,
            "model": "algorithmic bytes = SURVEY App. D element counts of F1 .. F4 per instance and iteration x 8 B",
            "peak_source": peak_src, "kernels": kernels,
            "whole_solve": {"alg_bytes_per_solve": per_solve, "achieved_gbs": per_solve * value_per_gpu / 1e9,
                            "frac": per_solve * value_per_gpu / 1e9 / peak}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="ddp", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU (default: the config's batch)")
    ap.add_argument("--total-batch", type=int, default=None, help="instances in total, split over the ranks (strong scaling)")
    ap.add_argument("--mode", default="fixed", choices=["fixed", "ref"])
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--no-gather", action="store_true", help="skip the gather of first-step controls (N>1)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: how the first-step controls of all shards reach rank 0 -- 'peer': every rank's gather kernel "
                         "stores its rows into rank 0's buffer over NVLink and raises a flag (no collective); 'nccl': "
                         "all_gather_into_tensor")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ref-sample", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    wl = Workload(args.config, args.mode)
    w = wl.w
    if args.seed is None:
        args.seed = w["seed"]
    if args.total_batch:
        if args.total_batch % world:
            raise SystemExit(f"--total-batch {args.total_batch} is not a multiple of {world} ranks")
        B = args.total_batch // world
    else:
        B = args.batch or w["batch"]

    if args.impl == "reference":
        run_reference(args, wl, rank, world, B)
        return

    import torch
    import torch.distributed as dist

    import nmpc_b200
    from nmpc_b200.ddp import F_COST, F_U

    if nmpc_b200.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: nmpc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, NX, NU, NG = w["N"], w["nx"], w["nu"], w["ng"]
    x0_np, u0_np = wl.inputs(B, args.seed + rank)  # each rank owns a different shard
    is_fmpc = w["kind"] == "fmpc"
    if is_fmpc:
        solver = nmpc_b200.FmpcSolver(w["model"], batch_capacity=B, device=local_rank)
    else:
        solver = nmpc_b200.DDPSolver(w["model"], batch_capacity=B, device=local_rank)
    cfg = solver.config()
    for k, v in wl.config_kw().items():
        setattr(cfg, k, v)

    stream = torch.cuda.Stream(device=dev)
    x0_d = torch.from_numpy(x0_np).to(dev)
    u_init_d = torch.from_numpy(u0_np).to(dev)
    u0_out = torch.empty((B, NU), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * B, NU), dtype=torch.float64, device=dev) if world > 1 else None
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    do_gather = world > 1 and not args.no_gather
    use_peer = do_gather and args.gather == "peer" and not is_fmpc
    peer, peer_step = None, [0]
    if use_peer:
        from nmpc_b200.ddp import F_U0
        from nmpc_b200.sharding import PeerBuffer

        peer = PeerBuffer(world * B, NU * 8, local_rank, owner=0)
    var_np = wl.fmpc_variable(B) if is_fmpc else None

    def make_var(to_torch):
        v = solver.make_variable(B)
        for name, key in (("x_list", "x"), ("u_list", "u"), ("lambda_list", "lambda"), ("s_list", "s"), ("nu_list", "nu")):
            setattr(v, name, to_torch(var_np[key]))
        return v

    var_d = make_var(lambda a: torch.from_numpy(a).to(dev)) if is_fmpc else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        if is_fmpc:
            solver.solve_batch(0.0, x0_d, var_d, stream=stream)
            if do_gather:
                u0_out.copy_(torch.from_numpy(solver.u0()))
        else:
            solver.solve_batch(0.0, x0_d, u_init_d, stream=stream, read_status=False)
            if use_peer:
                # this shard's u_list[0] rows go straight into rank 0's buffer; rank 0's stream waits for all flags
                peer_step[0] += 1
                solver.get_to_device_ptr(F_U0, peer.row_ptr(rank * B), B * NU * 8, stream=stream)
                peer.signal(peer_step[0], stream.cuda_stream)
                if rank == 0:
                    peer.wait(peer_step[0], stream.cuda_stream)
                return
            solver.u0(out=u0_out, stream=stream)
        if do_gather:
            dist.all_gather_into_tensor(gathered, u0_out)

    def timed_loop(step_fn, steps):
        """K steps, each bracketed by CUDA events on the launch stream; L2 flushed between steps."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t_wall = time.perf_counter()
        with torch.cuda.stream(stream):
            for s, e in evs:
                flush_buf.fill_(1)
                s.record(stream)
                step_fn()
                e.record(stream)
        barrier()
        wall = time.perf_counter() - t_wall
        ms = [s.elapsed_time(e) for s, e in evs]
        return float(np.sum(ms)), wall

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            device_step()
    barrier()

    # ---- device-resident throughput (`value`) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, _ = timed_loop(device_step, args.steps)
    if use_peer and rank == 0:
        peer.check(stream.cuda_stream)  # raises if a rank's flag never arrived

    # ---- per-kernel durations for the roofline (CUDA events inside the engine, same stream) ----
    solver.enable_timing(True)
    iter_ms_all, dur = [], None
    with torch.cuda.stream(stream):
        for _ in range(max(3, min(args.steps, 10))):
            flush_buf.fill_(1)
            if is_fmpc:
                solver.solve_batch(0.0, x0_d, var_d, stream=stream)
                dur = solver.computationDuration()
            else:
                solver.solve_batch(0.0, x0_d, u_init_d, stream=stream, read_status=False)
                dur = solver.computationDuration()
                iter_ms_all.append(solver.iterationDurations())
    solver.enable_timing(False)

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    x0_pin = torch.from_numpy(x0_np).pin_memory()
    u_pin = torch.from_numpy(u0_np).pin_memory()
    u0_host = torch.empty((B, NU), dtype=torch.float64).pin_memory()
    ufull_host = torch.empty((B, N, NU), dtype=torch.float64).pin_memory()
    cost_host = torch.empty((B,), dtype=torch.float64).pin_memory()
    x0_pin_np, u_pin_np, u0_host_np = x0_pin.numpy(), u_pin.numpy(), u0_host.numpy()
    var_pin = make_var(lambda a: torch.from_numpy(a).pin_memory().numpy()) if is_fmpc else None

    def e2e_step():
        if is_fmpc:
            solver.solve_batch(0.0, x0_pin_np, var_pin, stream=stream)  # H2D of x0 and the Variable inside
            u0_host_np[...] = solver.u0()
        else:
            solver.solve_batch(0.0, x0_pin_np, u_pin_np, stream=stream, read_status=False)  # H2D inside
            solver.u0(out=u0_host_np, stream=stream)  # D2H of the step's result (first-step controls)

    def e2e_full_step():
        """north_star's outputs: the whole optimal control sequence and the cost of every instance."""
        solver.solve_batch(0.0, x0_pin_np, u_pin_np, stream=stream, read_status=False)
        solver.get_into(F_U, ufull_host.numpy(), stream=stream)
        solver.get_into(F_COST, cost_host.numpy(), stream=stream)

    for _ in range(2):
        e2e_step()
    e2e_steps = max(3, args.steps // 2)
    # per step: CUDA events from just before the H2D copies to just after the D2H copy of the result (the host holds
    # the result when that event has completed); the L2 flush between steps is outside the events.  The wall clock
    # over the whole loop (flushes included) is reported next to it.
    e2e_ms, e2e_wall = timed_loop(e2e_step, e2e_steps)
    e2e_full_ms = None
    if not is_fmpc:
        e2e_full_step()
        e2e_full_ms, _ = timed_loop(e2e_full_step, e2e_steps)
    # the sampler has covered every timed region of this run: `value`, the per-kernel timing solves and `e2e`
    clocks = sampler.stop() if rank == 0 else None
    h2d = x0_np.nbytes + (sum(v.nbytes for v in var_np.values()) if is_fmpc else u0_np.nbytes)
    d2h = u0_host_np.nbytes

    # ---- max over ranks ----
    t = torch.tensor([total_ms, e2e_wall, e2e_ms, e2e_full_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_wall, e2e_ms, e2e_full_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * B * args.steps / (total_ms * 1e-3)
        e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
        peak, peak_src = measured_peaks()
        if is_fmpc:
            roofline = fmpc_roofline(B, value / world, dur, peak, peak_src)
            launches = int(sum(dur["launches"].values())) + 8
            work = {"iterations": MAX_ITER}
        else:
            iter_ms = np.median(np.stack([m for m in iter_ms_all if len(m) == len(iter_ms_all[-1])]), axis=0)
            roofline = ddp_roofline(wl, B, value / world, solver, iter_ms, dur["launches"], peak, peak_src)
            fused = dur["launches"]["derivative"] == 0
            # per step: 2 layout + 1 rollout + iterations x ([derivative,] backward, 2 line-search phases) + 1 extract
            launches = 2 + 1 + (3 if fused else 4) * int(solver.iterations().max()) + 1
            work = {"iterations_mean": float(solver.iterations().mean()),
                    "forward_passes_mean": float(solver.n_forward().mean()),
                    "backward_passes_mean": float(solver.n_backward().mean())}
        work["nccl_collectives_per_step"] = 1 if (do_gather and not use_peer) else 0
        work["gather_u0"] = ("peer stores + flags" if use_peer else "nccl all_gather") if do_gather else "none"

        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = wl.cpu_baseline(B, args.seed, args.cpu_seconds)

        e2e = {"value": e2e_value, "unit": w["unit"], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
               "timing": "CUDA events per step around the public-API call with pinned HOST buffers: H2D of the inputs, "
                         "the solve, D2H of the first-step controls (what an MPC tick needs); max over ranks",
               "wall_clock_value_incl_l2_flush": world * B * e2e_steps / e2e_wall}
        if e2e_full_ms:
            e2e["full_outputs"] = {
                "value": world * B * e2e_steps / (e2e_full_ms * 1e-3), "ms_per_step": e2e_full_ms / e2e_steps,
                "d2h_bytes_per_step": int(B * N * NU * 8 + B * 8),
                "what": "same call, but the D2H read is the whole optimal control sequence u_list and the cost of every "
                        "instance"}
        line = {
            "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.total_batch else "weak", "vs_baseline": None, "dtype": w["dtype"],
            "data": "synthetic", "config": workload_config(args, wl, world, B), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(args.steps * launches), "roofline": roofline, "cpu_baseline": cpu, "work": work,
        }
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.barrier()
        if peer is not None:
            peer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
