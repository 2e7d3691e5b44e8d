#!/usr/bin/env python
"""Benchmark of the ONE hot path: env-steps/s of the batched air-hockey frame loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], and configs[3] sharded over N GPUs): per GPU 2^20 environments, FP32
throughput build, one bench "step" = ONE launch of the fused step kernel = 8 frames per env with per-frame
discrete actions read from an int8 [8, n] tensor, the in-kernel computerAI opponent and Rewards reward,
writing reward/done/score/game_over [8, n] and the next observation [n, 8]; statistics are warp/block
reduced in-kernel and (N > 1) all-reduced over NCCL once per step.  1 env-step = 1 frame of Game.play
(3 physics sub-updates + reward + score / game-over handling).

`value`  : device-resident throughput (actions already in HBM).
`e2e`    : same call, but every step's actions come from pinned HOST memory (H2D inside the timed region) and
           the step's results (reward, done, next observation) are copied back to pinned host memory.
`cpu_baseline` / `--impl reference`: the C restatement of the reference (oracle/, kind "port") on the host
           cores with OpenMP -- the Python reference itself cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAMES_PER_STEP = 8          # K fused frames per launch (BASELINE config 3)
ENVS_PER_GPU = 1 << 20
METRIC = "env-steps/sec"
UNIT = "env-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_port_throughput(seconds: float = 10.0, envs: int = 1 << 18, threads: int = 0):
    """The C oracle (a port of the reference's step) on the host cores, same workload shape: 8 frames per
    env per call, per-frame int8 actions, Philox resets.  Returns (env-steps/s, threads, sample description)."""
    import numpy as np
    from oracle import oracle
    threads = threads or len(os.sched_getaffinity(0))
    rng = np.random.default_rng(0)
    st, sc = oracle.initial_state(envs)
    cur = np.zeros(envs, np.int32)
    acts = rng.integers(0, 4, (FRAMES_PER_STEP, envs)).astype(np.int8)
    kw = dict(actions=acts, reset_cursor=cur, seed=1234, nthreads=threads, want=("reward", "done", "score", "game_over"))
    oracle.run_frames(st, sc, FRAMES_PER_STEP, **kw)                 # warm-up (thread pool, page faults)
    calls, t0 = 0, time.perf_counter()
    while True:
        oracle.run_frames(st, sc, FRAMES_PER_STEP, **kw)
        calls += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return envs * FRAMES_PER_STEP * calls / dt, threads, f"{envs} envs x {FRAMES_PER_STEP} frames x {calls} calls ({dt:.1f} s)"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle
    threads = len(os.sched_getaffinity(0))
    envs = 1 << 18
    rng = np.random.default_rng(0)
    st, sc = oracle.initial_state(envs)
    cur = np.zeros(envs, np.int32)
    acts = rng.integers(0, 4, (FRAMES_PER_STEP, envs)).astype(np.int8)
    kw = dict(actions=acts, reset_cursor=cur, seed=1234, nthreads=threads, want=("reward", "done", "score", "game_over"))
    for _ in range(args.warmup):
        oracle.run_frames(st, sc, FRAMES_PER_STEP, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.run_frames(st, sc, FRAMES_PER_STEP, **kw)
    dt = time.perf_counter() - t0
    v = envs * FRAMES_PER_STEP * args.steps / dt
    sample = f"{envs} envs x {FRAMES_PER_STEP} frames per step (bounded sample of the {ENVS_PER_GPU}-env-per-GPU workload)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(gpus: int) -> str:
    return (f"{ENVS_PER_GPU} envs/GPU x {gpus} GPU, FP32, K={FRAMES_PER_STEP} fused frames/launch, per-frame int8 actions, "
            "computerAI opponent + Rewards reward, reward/done/score/game_over[K,n] + obs_next[n,8] written, "
            "in-kernel stats" + (" + NCCL all-reduce of stats per step" if gpus > 1 else ""))


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=260)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from air_hockey_training_environment_b200 import BatchedAirHockey

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, K = args.envs_per_gpu, FRAMES_PER_STEP
    W = max(3, args.warmup)

    env = BatchedAirHockey(n, dtype=torch.float32, device=dev, seed=1234, env_id0=rank * n)
    out = env.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next"))
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_act = 4                                                    # rotate several action tensors
    acts = [torch.randint(0, 4, (K, n), dtype=torch.int8, device=dev, generator=gen) for _ in range(n_act)]
    stats = torch.zeros(16, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def one_step(i: int):
        env.step(acts[i % n_act], K=K, out=out)
        if world > 1:                                            # per-rollout episode/score statistics
            env.stats_tensor(clear=True, out=stats)
            dist.all_reduce(stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # warm-up: >= 2,000 frames so the games decorrelate (mean game length ~1,900 frames)
    for i in range(W):
        one_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        one_step(i)
        ev[i + 1].record(stream)
    barrier()
    t_wall1 = time.time()
    total_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # kernel-only duration (N = 1: step == one kernel launch; events sit on the launching stream)
    kernel_ms = None
    if world == 1:
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i, (a, b) in enumerate(kev):
            a.record(stream)
            env.step(acts[i % n_act], K=K, out=out)
            b.record(stream)
        torch.cuda.synchronize(dev)
        kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev)

    value = world * n * K * args.steps / (total_ms * 1e-3)

    # ---- e2e: host actions in, host results out, copies inside the timed region (double-buffered streams)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(env, dev, n, K, max(3, min(W, 10)), max(10, min(args.steps, 50)), world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src, sm_max = peaks()
    # algorithmic bytes per launch: state 68 B read + 68 B written per env; per frame 1 B action,
    # 4 B reward, 1 B done, 1 B score, 1 B game_over; 32 B obs_next per env per launch
    state_b = 3 * 16 + 4 + 16
    bytes_per_launch = n * (2 * state_b + 32 + K * (1 + 4 + 1 + 1 + 1))
    info = env.launch_info()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(world), "envs_per_gpu": n, "frames_per_step": K,
                   "sub_updates_per_sec": 3 * value,
                   "l2": "no flush: each step touches %.0f MB of distinct HBM lines (> 126 MB L2)" % (bytes_per_launch / 1e6 - 0),
                   "launch": info},
        "gpu_launches": args.steps * (1 if world == 1 else 2),
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if kernel_ms is not None:
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_step_kernel.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                            "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                            "kernel": "ah::step_kernel<float,1>", "kernel_ms": kernel_ms,
                            "algorithmic_bytes_per_launch": bytes_per_launch,
                            "bytes_per_env_step": bytes_per_launch / (n * K)}
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_port_throughput(args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(env, dev, n, K, warm, steps, world):
    """Public-API call with HOST buffers: per step, H2D of the int8 [K,n] actions from pinned memory, the fused
    step, and D2H of reward [K,n], done [K,n] and obs_next [n,8] into pinned memory.  Two slots, three streams:
    step i's D2H overlaps step i+1's kernel and step i+2's H2D."""
    import torch
    import torch.distributed as dist
    slots = 2
    h_act = [torch.randint(0, 4, (K, n), dtype=torch.int8).pin_memory() for _ in range(slots)]
    d_act = [torch.empty(K, n, dtype=torch.int8, device=dev) for _ in range(slots)]
    outs = [env.make_outputs(K, ("reward", "done", "obs_next")) for _ in range(slots)]
    h_out = [(torch.empty(K, n, dtype=torch.float32).pin_memory(), torch.empty(K, n, dtype=torch.uint8).pin_memory(),
              torch.empty(n, 8, dtype=torch.float32).pin_memory()) for _ in range(slots)]
    s_in, s_run, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    e_in = [torch.cuda.Event() for _ in range(slots)]
    e_run = [torch.cuda.Event() for _ in range(slots)]
    e_out = [torch.cuda.Event() for _ in range(slots)]

    def step(i):
        s = i % slots
        with torch.cuda.stream(s_in):
            s_in.wait_event(e_run[s])                     # slot's previous kernel has consumed its actions
            d_act[s].copy_(h_act[s], non_blocking=True)
            e_in[s].record(s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(e_in[s])
            s_run.wait_event(e_out[s])                    # slot's previous results have left the device
            env.step(d_act[s], K=K, out=outs[s])
            e_run[s].record(s_run)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e_run[s])
            h_out[s][0].copy_(outs[s].reward, non_blocking=True)
            h_out[s][1].copy_(outs[s].done, non_blocking=True)
            h_out[s][2].copy_(outs[s].obs_next, non_blocking=True)
            e_out[s].record(s_out)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(warm):
        step(i)
    sync()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    sync()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    h2d = K * n
    d2h = K * n * 4 + K * n + n * 8 * 4
    return {"value": world * n * K * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "note": "pinned host buffers, 2 slots / 3 streams; timed with the host clock around a "
                                    "barrier + synchronize on both sides"}


if __name__ == "__main__":
    main()
