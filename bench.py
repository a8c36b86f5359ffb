#!/usr/bin/env python
"""Headline benchmark: batched cart-pole DDP trajectories/sec (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode fixed|ref]

A "step" is one batched DDPSolver::solve() of B independent cart-pole instances (n_x=4, n_u=1,
horizon 100, 10 iterations, fp64, synthetic random initial states) per GPU.  One process per GPU
(torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); the batch shards with no data-path collective; the only
collective is the optional NCCL all-gather of first-step controls after each solve.
Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# Same setting torchrun applies to every rank of an N > 1 run: without it the BLAS / OpenMP worker pools that numpy and
# torch start at import compete with the thread that enqueues the solve (measured on the B200 box: e2e 1.93 ms per
# step with the pools, 1.83 ms without; the device-timed `value` is unaffected).  The CPU arms are not throttled by
# it: they pass their thread count to the oracle explicitly (host_threads()).
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NX, NU, N_STEPS, MAX_ITER = 4, 1, 100, 10
METRIC = "DDP trajectories/sec (batch, T=100, 10 iters)"
UNIT = "trajectories/s"


def algorithmic_elements(nx, nu, N):
    """SURVEY.md 8(d) / BASELINE.md 4: compulsory reads+writes per instance, in scalar elements."""
    blk = 2 * nx * nx + 2 * nx * nu + nx + nu + nu * nu
    D0 = (nx + N * nu) + (N + 1) * (nx + 1)
    D1 = (N + 1) * nx + N * nu + N * blk + nx + nx * nx
    D2 = N * blk + nx + nx * nx + 2 * N * nu + N * nu * nx + 4
    D3 = 2 * (N + 1) * nx + 3 * N * nu + N * nu * nx + (N + 1)
    return {"blk": blk, "D0": D0, "D1": D1, "D2": D2, "D3": D3}


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
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the
    CPU baseline: the thread count is passed to the oracle explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_config(O, mode):
    kw = dict(max_iter=MAX_ITER, horizon_steps=N_STEPS)
    if mode == "fixed":
        # M-fixed: thresholds 0 => every instance runs exactly 10 iterations
        kw.update(k_rel_norm_thre=0.0, cost_update_thre=0.0)
    return kw


def cpu_baseline(O, B, seed, mode, min_seconds, native=True):
    """Oracle (CPU port of the reference algorithm) on all host threads over the same workload."""
    p = O.default_params("cartpole")
    cfg = O.ddp_config(**make_config(O, mode))
    x0 = O.cartpole_x0(B, seed)
    u_init = np.zeros((B, N_STEPS, NU))
    threads = host_threads()
    O.ddp_solve_batch("cartpole", p, cfg, x0[:64], u_init[:64], native=native, outputs=False, nthreads=threads)
    done, t0 = 0, time.perf_counter()
    while True:
        O.ddp_solve_batch("cartpole", p, cfg, x0, u_init, native=native, outputs=False, nthreads=threads)
        done += B
        el = time.perf_counter() - t0
        if el >= min_seconds:
            break
    # single-thread latency of one solve: 256 instances on one thread
    n1 = min(256, B)
    t1 = time.perf_counter()
    O.ddp_solve_batch("cartpole", p, cfg, x0[:n1], u_init[:n1], native=native, outputs=False, nthreads=1)
    ms1 = 1e3 * (time.perf_counter() - t1) / n1
    return {"value": done / el, "unit": UNIT, "cores": int(threads), "kind": "port",
            "per_core": done / el / max(int(threads), 1), "single_thread_ms_per_solve": ms1,
            "sample": f"{done} cart-pole solves ({done // B} x the B={B} workload, mode {mode}) in {el:.1f} s, "
                      f"oracle/ built -O3 -march=native -fopenmp, one solver object per thread"}


def reference_headers_sample(O, params, cfg, x0, threads, n=256):
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
        us = np.zeros((n, N_STEPS, NU))
        u0, cost, it = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        t0 = time.perf_counter()
        rc = lib.ref_ddp_solve_cartpole_batch(vp(params), C.byref(cfg), n, C.c_double(0.0), vp(xs), vp(us), int(threads),
                                              vp(u0), vp(cost), vp(it))
        el = time.perf_counter() - t0
        return {"value": n / el, "unit": UNIT, "cores": int(threads), "rc": int(rc), "iterations_mean": float(it.mean()),
                "sample": f"{n} instances, reference headers + Eigen stand-in (heap-backed), -O3 -march=x86-64-v3 -fopenmp"}
    except Exception as e:  # informational leg only
        return {"error": str(e)[:200]}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Eigen reference cannot be built
    here) on this box's host cores.  Rank 0 only."""
    if rank != 0:
        return
    import oracle_lib as O

    B = args.batch
    p = O.default_params("cartpole")
    cfg = O.ddp_config(**make_config(O, args.mode))
    # bounded sample per step so that steps x sample stays within minutes
    sample = min(B, args.ref_sample)
    x0 = O.cartpole_x0(B, args.seed)[:sample]
    u_init = np.zeros((sample, N_STEPS, NU))
    threads = host_threads()
    for _ in range(max(args.warmup, 1)):
        O.ddp_solve_batch("cartpole", p, cfg, x0, u_init, native=True, outputs=False, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.ddp_solve_batch("cartpole", p, cfg, x0, u_init, native=True, outputs=False, nthreads=threads)
    el = time.perf_counter() - t0
    value = sample * args.steps / el
    shim = reference_headers_sample(O, p, cfg, x0, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(threads), "kind": "port",
                         "sample": f"{sample} of the {B} instances per step, {args.steps} steps, all host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_headers_shim": shim,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": f"cart-pole DDP (n_x=4, n_u=1) batch={args.batch}/GPU x {world} GPU, horizon=100, 10 iterations, "
                    f"fp64, mode M-{args.mode}",
        "batch_per_gpu": args.batch, "horizon": N_STEPS, "max_iter": MAX_ITER, "mode": args.mode, "seed": args.seed,
        "parallelism": f"batch-sharded x{world}, no data-path collective",
        "gather_u0": bool(world > 1 and not args.no_gather),
        "l2": "L2 flushed between timed steps (256 MiB write)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU (BASELINE.json configs[1]: 4096)")
    ap.add_argument("--mode", default="fixed", choices=["fixed", "ref"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-gather", action="store_true", help="skip the NCCL all-gather of first-step controls (N>1)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ref-sample", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import nmpc_b200

    if nmpc_b200.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: nmpc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import oracle_lib as O  # synthetic-input generator + cpu_baseline leg only

    B = args.batch
    x0_np = O.cartpole_x0(B, args.seed + rank)  # each rank owns a different shard
    u0_np = np.zeros((B, N_STEPS, NU))
    solver = nmpc_b200.DDPSolver("cartpole", batch_capacity=B, device=local_rank)
    cfg = solver.config()
    for k, v in make_config(O, args.mode).items():
        setattr(cfg, k, v)

    stream = torch.cuda.Stream(device=dev)
    x0_d = torch.from_numpy(x0_np).to(dev)
    u_init_d = torch.zeros((B, N_STEPS, NU), dtype=torch.float64, device=dev)
    u0_out = torch.empty((B, NU), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * B, NU), dtype=torch.float64, device=dev) if world > 1 else None
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    do_gather = world > 1 and not args.no_gather

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        solver.solve_batch(0.0, x0_d, u_init_d, stream=stream, read_status=False)
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
    iters = solver.iterations()
    n_fwd = solver.n_forward()
    n_bwd = solver.n_backward()

    # ---- per-kernel durations for the roofline (CUDA events inside the engine, same stream) ----
    solver.enable_timing(True)
    stage = {"derivative": [], "backward": [], "forward": [], "setup": [], "solve": []}
    launches = None
    with torch.cuda.stream(stream):
        for _ in range(max(3, min(args.steps, 10))):
            flush_buf.fill_(1)
            solver.solve_batch(0.0, x0_d, u_init_d, stream=stream, read_status=False)
            d = solver.computationDuration()
            for k in stage:
                stage[k].append(d[k])
            launches = d["launches"]
    solver.enable_timing(False)

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    x0_pin = torch.from_numpy(x0_np).pin_memory()
    u_pin = torch.zeros((B, N_STEPS, NU), dtype=torch.float64).pin_memory()
    u0_host = torch.empty((B, NU), dtype=torch.float64).pin_memory()
    x0_pin_np, u_pin_np, u0_host_np = x0_pin.numpy(), u_pin.numpy(), u0_host.numpy()

    def e2e_step():
        solver.solve_batch(0.0, x0_pin_np, u_pin_np, stream=stream, read_status=False)  # H2D inside
        solver.u0(out=u0_host_np, stream=stream)  # D2H of the step's result (first-step controls)

    for _ in range(2):
        e2e_step()
    e2e_steps = max(3, args.steps // 2)
    # per step: CUDA events from just before the H2D copies to just after the D2H copy of the result (the host holds
    # the result when that event has completed); the L2 flush between steps is outside the events.  The wall clock
    # over the whole loop (flushes included) is reported next to it.
    e2e_ms, e2e_wall = timed_loop(e2e_step, e2e_steps)
    # the sampler has covered every timed region of this run: `value`, the per-kernel timing solves and `e2e`
    clocks = sampler.stop() if rank == 0 else None
    h2d = x0_np.nbytes + u0_np.nbytes
    d2h = u0_host_np.nbytes

    # ---- max over ranks ----
    t = torch.tensor([total_ms, e2e_wall, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_wall, e2e_ms = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * B * args.steps / (total_ms * 1e-3)
        e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
        e2e_wall_value = world * B * e2e_steps / e2e_wall

        el = algorithmic_elements(NX, NU, N_STEPS)
        peak, peak_src = measured_peaks()
        med = {k: float(np.median(v)) for k, v in stage.items()}
        # K1 fused into K2 (producer warp + consumer warp, ddp_backward_fused.cuh): no derivative launches; every
        # backward sweep linearises its trajectory first, so the fused kernel is charged D1 + D2 per sweep
        fused = launches["derivative"] == 0
        stages = ("backward", "forward") if fused else ("derivative", "backward", "forward")
        per_launch_ms = {k: med[k] / max(launches[k], 1) for k in stages}
        active_iters = float(iters.sum())
        lin_elems = el["D1"] * (float(n_bwd.sum()) if fused else active_iters)
        alg_bytes = {
            # per launch, averaged over the launches of one solve (all B instances of this GPU)
            "derivative": 8.0 * lin_elems / max(launches["derivative"], 1),
            "backward": 8.0 * (el["D2"] * float(n_bwd.sum()) + (lin_elems if fused else 0.0)) / max(launches["backward"], 1),
            "forward": 8.0 * el["D3"] * float(n_fwd.sum()) / max(launches["forward"], 1),
        }
        kernels = {}
        for k in per_launch_ms:
            ach = alg_bytes[k] / (per_launch_ms[k] * 1e-3) / 1e9 if per_launch_ms[k] > 0 else 0.0
            kernels[k] = {"ms_per_launch": per_launch_ms[k], "launches_per_step": launches[k],
                          "alg_bytes_per_launch": alg_bytes[k], "achieved_gbs": ach, "frac": ach / peak,
                          "share_of_step": med[k] / med["solve"] if med["solve"] > 0 else None}
        dominant = max(kernels, key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"])
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this
        # same command (profiles/r1_v3_stage_kernels.md; tools/profile_kernels.sh), summed over the stage's kernels
        traffic = None
        stage_kernels = {"derivative": ["linearize_kernel"],
                         "backward": ["backward_fused_kernel"] if fused else ["backward_kernel"],
                         "forward": ["forward_first_kernel", "forward_fanout_kernel"]}
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath) and B == 4096 and args.mode == "fixed":
            try:
                tj = json.load(open(tpath))
                traffic = float(sum(tj[k]["dram_bytes_per_launch"] for k in stage_kernels[dominant]))
            except Exception:
                traffic = None
        kernel_names = {"derivative": "ddp::linearize_kernel",
                        "backward": "ddp::backward_fused_kernel (K1 + K2: producer warp linearises, consumer warp sweeps)"
                        if fused else "ddp::backward_kernel",
                        "forward": "ddp::forward_first_kernel + ddp::forward_fanout_kernel (one line search = 2 launches)"}
        roofline = {"bound": "hbm", "kernel": kernel_names[dominant], "achieved": kernels[dominant]["achieved_gbs"],
                    "peak": peak, "unit": "GB/s", "frac": kernels[dominant]["frac"], "traffic": traffic,
                    "model": "algorithmic bytes = SURVEY 8(d) three-stage byte model (K1+K2 fused => D1+D2 per launch); a "
                             "fraction above 1 means the fused kernel does not move the traffic the model charges for",
                    "peak_source": peak_src, "kernels": kernels,
                    "whole_solve": {
                        "alg_bytes_per_trajectory": 8.0 * (el["D0"] + (lin_elems + el["D2"] * float(
                            n_bwd.sum()) + el["D3"] * float(n_fwd.sum())) / B),
                    }}
        ws = roofline["whole_solve"]
        ws["achieved_gbs"] = ws["alg_bytes_per_trajectory"] * (value / world) / 1e9
        ws["frac"] = ws["achieved_gbs"] / peak

        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(O, B, args.seed, args.mode, args.cpu_seconds)

        gathers = 1 if do_gather else 0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "timing": "CUDA events per step around the public-API call with pinned HOST buffers: H2D of x0 and "
                              "initial_u_list, the solve, D2H of the first-step controls; max over ranks",
                    "wall_clock_value_incl_l2_flush": e2e_wall_value},
            # per step: 2 layout + 1 rollout + 10 x ([derivative,] backward, 2 line-search phases) + 1 first-control extract
            "gpu_launches": int(args.steps * (2 + 1 + (3 if fused else 4) * MAX_ITER + 1)),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "work": {"iterations_mean": float(iters.mean()), "forward_passes_mean": float(n_fwd.mean()),
                     "backward_passes_mean": float(n_bwd.mean()), "nccl_collectives_per_step": gathers},
        }
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
