"""Host-side multi-GPU logic on CPU: contiguous batch shards and the first-control gather over a
world_size-2 gloo group (the N>1 path of bench.py uses the same helpers over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from nmpc_b200.sharding import gather_first_controls, shard_range, shard_sizes


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 4096, 4097, 131072):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            for (b0, e0), (b1, e1) in zip(ranges[:-1], ranges[1:]):
                assert e0 == b1 and e0 >= b0
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank solves its own shard (with the CPU oracle standing in for the GPU engine here) ...
        b, e = shard_range(total, world, rank)
        x0 = O.cartpole_x0(total, 5)[b:e]
        cfg = O.ddp_config(max_iter=3, horizon_steps=20)
        r = O.ddp_solve_batch("cartpole", O.default_params("cartpole"), cfg, x0, np.zeros((e - b, 20, 1)), nthreads=1)
        u0_local = torch.from_numpy(r["u"][:, 0, :].copy())
        # ... and only the first-step controls cross ranks
        u0_all = gather_first_controls(u0_local, total)
        ret[rank] = u0_all.numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 9])
def test_gather_first_controls_gloo_world2(total):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, total, ret), nprocs=world, join=True)
        got = [ret[r] for r in range(world)]
    cfg = O.ddp_config(max_iter=3, horizon_steps=20)
    full = O.ddp_solve_batch("cartpole", O.default_params("cartpole"), cfg, O.cartpole_x0(total, 5),
                             np.zeros((total, 20, 1)), nthreads=1)
    for g in got:
        assert g.shape == (total, 1)
        np.testing.assert_array_equal(g, full["u"][:, 0, :])


# ---------------------------------------------------------------------------------------------------------------
# The sharded C ABI (include/nmpc_b200/c_api.h, "several GPUs, one box") on real devices.  With one GPU in the box
# the two shards / two processes share it: every code path (worker threads, chunked staging, direct stores into
# another handle's / another process's buffer, flags) is the same, only the wire is not NVLink.
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHARDED_BIN = os.path.join(ROOT, "tests", "cpp", "test_sharded")


@pytest.fixture(scope="module")
def sharded_bin(nmpc):
    import subprocess

    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "test_sharded.cpp"), "-o", SHARDED_BIN,
           "-L" + os.path.join(ROOT, "nmpc_b200"), "-lnmpc_b200", "-Wl,-rpath," + os.path.join(ROOT, "nmpc_b200"),
           "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.run(cmd, check=True)
    return SHARDED_BIN


def _kv(out):
    return {line.partition(" ")[0]: line.partition(" ")[2] for line in out.splitlines()}


def test_sharded_cpp_caller_fails_loudly_without_gpu(sharded_bin, nmpc):
    import subprocess

    if nmpc.device_count() > 0:
        pytest.skip("a GPU is present: covered by test_sharded_cpp_caller_world2")
    r = subprocess.run([sharded_bin], capture_output=True, text=True)
    d = _kv(r.stdout)
    assert r.returncode == 0 and d["no_device_error"].split()[0] != "0" and "no CPU fallback" in d["no_device_error"]


@pytest.mark.gpu
def test_sharded_cpp_caller_world2(sharded_bin, gpu):
    """tests/cpp/test_sharded.cpp: 37 cart-pole solves over two shards == the same solves on one handle."""
    import subprocess

    r = subprocess.run([sharded_bin], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    d = _kv(r.stdout)
    assert d["done"] == "1" and d["num_shards"] == "2"
    assert d["range0"].split()[:2] == ["0", "19"] and d["range1"].split()[:2] == ["19", "37"]
    assert float(d["u_max_rel_diff"]) <= 1e-9
    for key in ("u_identical", "cost_identical", "iters_identical", "u0_peer_identical", "small_identical",
                "facade_identical"):
        assert d[key] == "1", (key, r.stdout)
    assert int(d["max_iters"]) >= 3
    assert d["too_large"] != "0" and d["short_dst"] != "0"
    assert "initial_u_list length should be 50 but 49." in d["bad_horizon"]


@pytest.mark.gpu
def test_sharded_solver_python_against_oracle(gpu):
    """ShardedDDPSolver over every visible device (twice device 0 on a one-GPU box), ragged batch, input limits."""
    import torch
    from nmpc_b200.sharding import ShardedDDPSolver

    n_dev = gpu.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    B, N = 8 * len(devices) + 3, 40
    x0 = O.cartpole_x0(B, 17)
    lo, hi = np.array([-12.0]), np.array([12.0])
    s = ShardedDDPSolver("cartpole", total_capacity=B, devices=devices)
    c = s.config()
    c.horizon_steps, c.max_iter, c.with_input_constraint = N, 8, True
    s.setInputLimitsFunc((lo, hi))
    ok = s.solve_batch(0.0, x0, np.zeros((B, N, 1)))
    ref = O.ddp_solve_batch("cartpole", O.default_params("cartpole"),
                            O.ddp_config(horizon_steps=N, max_iter=8, with_input_constraint=1), x0, np.zeros((B, N, 1)),
                            u_lo=lo, u_hi=hi)
    assert np.array_equal(ok, ref["status"] == 1) and np.array_equal(s.iterations(), ref["iters"])
    u = s.u_list()
    assert (np.max(np.abs(u - ref["u"]), axis=(1, 2)) / (1 + np.max(np.abs(ref["u"]), axis=(1, 2)))).max() <= 1e-9
    np.testing.assert_allclose(s.cost(), ref["cost"], rtol=1e-11)
    assert sum(s.shard_range(B, i)[1] - s.shard_range(B, i)[0] for i in range(s.num_shards())) == B
    # every shard stores its first-step controls straight into one tensor on device 0
    u0 = torch.zeros((B, 1), dtype=torch.float64, device="cuda:0")
    s.u0(out=u0, dst_device=0)
    assert np.array_equal(u0.cpu().numpy(), u[:, 0, :])
    s.close()


def _peer_worker(rank, world, port, total, ret):
    import torch

    import nmpc_b200 as gpu
    from nmpc_b200.sharding import PeerBuffer

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = rank % gpu.device_count()
        torch.cuda.set_device(dev)
        b, e = shard_range(total, world, rank)
        N = 30
        x0 = O.cartpole_x0(total, 23)[b:e]
        s = gpu.DDPSolver("cartpole", batch_capacity=e - b, device=dev)
        s.config().horizon_steps, s.config().max_iter = N, 5
        buf = PeerBuffer(total, 8, dev, owner=0)  # u0 rows of 1 double
        for step in (1, 2):  # two rounds: the flags count up
            s.solve_batch(0.0, x0 * step, np.zeros((e - b, N, 1)), read_status=False)
            s.get_to_device_ptr(11, buf.row_ptr(b), (e - b) * 8)  # NMPC_B200_DDP_U0 -> rank 0's memory
            buf.signal(step)
            if rank == 0:
                buf.wait(step)
                buf.check()
                ret[f"u0_step{step}"] = buf.read(total)
            ret[f"local{rank}_step{step}"] = s.u0()[:, 0].copy()
            dist.barrier()
        if rank == 0:
            # nobody raises the flags to 9: the wait gives up instead of hanging the device
            buf.wait(9, timeout_ms=200)
            try:
                buf.check()
                ret["timeout_reported"] = False
            except gpu.NmpcB200Error as err:
                ret["timeout_reported"] = "timed out" in str(err)
        dist.barrier()
        buf.close()
        s.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_buffer_gather_between_two_processes(gpu):
    """One process per shard: each rank's gather kernel stores its u0 rows into rank 0's device buffer (CUDA IPC
    mapping) and raises its flag; rank 0's stream waits on the flags -- no collective."""
    world, total = 2, 11
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_peer_worker, args=(r, world, port, total, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        for step in (1, 2):
            want = np.concatenate([ret[f"local{r}_step{step}"] for r in range(world)])
            assert np.array_equal(ret[f"u0_step{step}"], want)
        assert ret["timeout_reported"] is True


@pytest.mark.gpu
def test_sharded_fmpc_matches_one_handle(gpu):
    """nmpc_b200_fmpc_create_sharded: a ragged FMPC batch over two shards == the same batch on one handle, bit for bit."""
    from nmpc_b200.sharding import ShardedFmpcSolver

    n_dev = gpu.device_count()
    devices = [0, 1] if n_dev > 1 else [0, 0]
    B, N = 21, 30
    x0 = O.cartpole_x0(B, 8)
    one = gpu.FmpcSolver("cartpole", batch_capacity=B)
    one.config().horizon_steps, one.config().max_iter = N, 4
    v = one.make_variable(B)
    v.reset(0.0, 0.0, 0.0, 1.0, 1.0)
    status_one = np.asarray(one.solve_batch(0.0, x0, v))
    out = one.variable()
    sh = ShardedFmpcSolver("cartpole", total_capacity=B, devices=devices)
    sh.config().horizon_steps, sh.config().max_iter = N, 4
    w = one.make_variable(B)
    w.reset(0.0, 0.0, 0.0, 1.0, 1.0)
    status = sh.solve_batch(0.0, x0, w)
    assert sh.num_shards() == 2 and np.array_equal(status, status_one.astype(np.int32))
    assert np.array_equal(sh.get(1, (B, N, 1)), out.u_list)
    assert np.array_equal(sh.get(0, (B, N + 1, 4)), out.x_list)
    assert np.array_equal(sh.get(4, (B, N, 4)), out.nu_list)
    assert np.array_equal(sh.get(10, (B, 1)), out.u_list[:, 0, :])
    with pytest.raises(gpu.NmpcB200Error):
        sh.solve_batch(0.0, O.cartpole_x0(B + 1, 1), one.make_variable(B + 1))
    sh.close()
    one.close()
