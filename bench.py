#!/usr/bin/env python
"""Benchmark of the ONE hot path: env-steps/s of the batched air-hockey frame loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--parts P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]; configs[3] = the same per-GPU work sharded over N GPUs): per GPU 2^20
environments, FP32 throughput build, one bench "step" = 8 fused frames for every env in the PACKED rollout mode
(csrc/ah_step_packed.cuh): per-frame discrete actions read as quad-major uint8 [2, n, 4], the in-kernel computerAI
opponent and Rewards reward, one event byte (reward, done, score, game_over) per env-frame and the next observation
[n, 8] written, statistics reduced in-kernel.  A step is issued as `--parts` (default 2) disjoint env slices on as many
streams (sub-range launches of the same kernel; consecutive steps overlap tail-to-head); the statistics are copied out --
and at N > 1 all-reduced (by the library's fused exchange over NVLink peer memory, `--collective p2p`, or NCCL,
`--collective nccl`) -- once per rollout of 16 steps and at the end of the timed region, and the reduced payload is verified.  1 env-step = 1 frame of Game.play (3 physics sub-updates + reward + score / game-over handling).

`value`        device-resident throughput (actions already in HBM), CUDA events, max over ranks.
`e2e`          the public HOST-buffer API (host.HostStepper): every step's actions come from pinned HOST memory (H2D inside
               the timed region) and the step's results (event bytes, next observation) go back to pinned host memory;
               `e2e.roofline` = the same copies without the kernels (the host link's measured ceiling).
`roofline`     the fused K=8 kernel is bound by FP32 instruction issue, not by HBM: `achieved` counts
               the ALGORITHMIC lane-ops (254 per env-step) per second against 148 SMs x 128 lanes x SM
               clock, per GPU; `single_launch_*` = one whole-n launch per step (N = 1); the HBM view of the same step is
               beside it.
`extras`       BASELINE configs 4 (the packed step with the fused ring append; HBM roofline) and 5 (PPO rollouts) at this
               world size; at N = 1 also the HBM-bound K=1 mode, round 1's unpacked kernel, the FP64 validation build.
`cpu_baseline` the reference's OWN Python implementation (oracle/_ref: byte-compiled from /root/reference by
               oracle/fetch_ref.py) on this box's host cores: BASELINE config 1, one process and a pool over all cores
               (kind "reference"); the C port (kind "port") beside it.  `--impl reference` times the same reference.
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
LANE_OPS_PER_ENV_STEP = 254  # algorithmic FP32-pipe lane-ops per frame (88 FLOP + 166 cmp/sel/cvt)
STATE_BYTES = 3 * 16 + 4 + 16   # puck, robot, opponent float4 + prev_dist + meta int4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def profile_fact(name, default=None):
    """Numbers that only ncu can provide (committed under profiles/ from the last capture of this kernel)."""
    for f in ("step_kernel_facts.json", "ring_kernel_facts.json"):
        p = os.path.join(ROOT, "profiles", f)
        if os.path.exists(p):
            facts = json.load(open(p))
            if name in facts:
                return facts[name]
    return default


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0: float, t1: float):
        sm, mx, pw, reasons = [], [], [], set()
        for t, r in self.rows:
            if not (t0 <= t <= t1):
                continue
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_port_run(steps: int, warmup: int, envs: int = 1 << 18, threads: int = 0, seconds: float = 0.0):
    """The C oracle (a port of the reference's step) on the host cores, same workload shape: 8 frames per env per
    call, per-frame int8 actions, Philox resets.  Either `steps` calls or as many as fit in `seconds`."""
    import numpy as np
    from oracle import oracle
    threads = threads or len(os.sched_getaffinity(0))
    rng = np.random.default_rng(0)
    st, sc = oracle.initial_state(envs)
    cur = np.zeros(envs, np.int32)
    acts = rng.integers(0, 4, (FRAMES_PER_STEP, envs)).astype(np.int8)
    kw = dict(actions=acts, reset_cursor=cur, seed=1234, nthreads=threads, want=("reward", "done", "score", "game_over"))
    for _ in range(max(1, warmup)):
        oracle.run_frames(st, sc, FRAMES_PER_STEP, **kw)
    calls, t0 = 0, time.perf_counter()
    while True:
        oracle.run_frames(st, sc, FRAMES_PER_STEP, **kw)
        calls += 1
        dt = time.perf_counter() - t0
        if (seconds > 0 and dt >= seconds) or (seconds <= 0 and calls >= steps):
            break
    return envs * FRAMES_PER_STEP * calls / dt, threads, dt, calls, envs


def cpu_baseline(port_seconds: float):
    """The reference's CPU path on THIS box's host cores, timed beside the GPU numbers (rank 0, N = 1).  `value` is the
    reference's own Python implementation (kind "reference": oracle/_ref, BASELINE config 1 = 10,000 frames per
    process) with a multiprocessing pool over all cores, pure-simulation variant (the most favourable to it); the
    one-process figure, the socket-stub variant and the C port (kind "port") are reported beside it."""
    v, cores, dt, calls, envs = cpu_port_run(0, 1, seconds=port_seconds)
    v1, _, dt1, calls1, envs1 = cpu_port_run(0, 1, envs=1 << 15, threads=1, seconds=min(2.0, port_seconds))
    port = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{envs} envs x {FRAMES_PER_STEP} frames x {calls} calls ({dt:.1f} s), OpenMP over {cores} threads "
                      "(oracle/ah_oracle.c: C restatement, bit-exact vs the Python reference)",
            "one_thread_value": v1,
            "one_thread_sample": f"{envs1} envs x {FRAMES_PER_STEP} frames x {calls1} calls ({dt1:.1f} s), 1 thread"}
    ref = python_reference_timing(10000)
    if "pure_sim" not in ref:
        return dict(port, python_reference=ref)
    ps, ss = ref["pure_sim"], ref["socket_stub"]
    return {"value": ps["pool_sum_of_rates"], "unit": UNIT, "cores": ref["host_cores"], "kind": "reference",
            "language": "python", "reference_build": ref["reference"],
            "sample": f"BASELINE config 1: single AirHockey env, random robot vs computerAI, {ref['frames_per_process']} frames "
                      f"per process; multiprocessing.Pool({ref['host_cores']}) = all host cores, one env per worker, sum of "
                      "per-worker rates; Redis posts + CSV rows no-op'd (pure simulation, most favourable to the reference)",
            "one_process_value": ps["one_process"], "pool_wall_clock_value": ps["pool_wall_clock_rate"],
            "socket_stub": {"one_process_value": ss["one_process"], "pool_sum_of_rates": ss["pool_sum_of_rates"],
                            "pool_wall_clock_value": ss["pool_wall_clock_rate"],
                            "what": "Redis stubbed at the client class: the JSON payload of every sub-update is still built "
                                    "and encoded, the per-done CSV row still written (closest to shipped behaviour)"},
            "sub_updates_per_s_one_process": ps["sub_updates_per_s_one_process"],
            "goals_per_1000_frames": ps["goals_per_1000_frames"], "python": ref["python"], "numpy": ref["numpy"],
            "port": port}


def python_reference_timing(frames: int = 10000, timeout: float = 240.0):
    """the reference's own Python loop (oracle/_ref, byte-compiled from /root/reference by
    oracle/fetch_ref.py; the live sources where present) on THIS host: one process, then a pool over all cores, in the
    socket-stub and pure-simulation variants.  Separate process tree: no fork after CUDA initialisation."""
    try:
        res = subprocess.run([sys.executable, "-m", "oracle.time_reference", "--frames", str(frames)], cwd=ROOT,
                             capture_output=True, text=True, timeout=timeout)
        if res.returncode != 0:
            return {"unavailable": (res.stderr or "").strip().splitlines()[-1:]}
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:                                   # reference absent: the C port is reported alone
        return {"unavailable": repr(e)}


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores.  The Python
    reference (oracle/_ref) with a multiprocessing pool over every core, one env per worker, each bench "step" =
    `--ref-frames` frames of the Game.play loop on every worker (a bounded sample of the workload: the reference has no
    batched form, so its throughput does not depend on how many envs the GPU arm carries).  Variant "pure_sim" (Redis
    posts and CSV rows no-op'd) is the line's value: the most favourable to the reference.  The C port's number (kind
    "port", ~10^3 x faster per core than the reference) is reported beside it, never as the value."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import refshim
    W = max(1, min(args.warmup, 5))
    port_v, port_threads, _, _, port_envs = cpu_port_run(0, 2, seconds=2.0)
    port = {"value": port_v, "unit": UNIT, "cores": port_threads, "kind": "port",
            "sample": f"{port_envs} envs x {FRAMES_PER_STEP} frames per call for 2 s, OpenMP (oracle/ah_oracle.c)"}
    if refshim.reference_available():
        from oracle.time_reference import ReferencePool
        frames = args.ref_frames
        rates = {}
        for variant in ("socket_stub", "pure_sim"):
            pool = ReferencePool(variant=variant)
            for _ in range(W):
                pool.step(frames)
            steps = args.steps if variant == "pure_sim" else max(2, min(args.steps, 5))
            done, t0 = 0, time.perf_counter()
            for _ in range(steps):
                done += pool.step(frames)
            dt = time.perf_counter() - t0
            rates[variant] = (done / dt, dt, steps, pool.procs)
            pool.close()
        v, dt, steps, cores = rates["pure_sim"]
        kind = "reference"
        sample = (f"{cores} worker processes (all host cores) x 1 AirHockey env each x {frames} frames per step, {steps} steps; "
                  f"the reference's own Python classes ({refshim.reference_kind()} from {os.path.relpath(refshim.REFERENCE_ROOT, ROOT) if refshim.REFERENCE_ROOT.startswith(ROOT) else refshim.REFERENCE_ROOT}), "
                  "Game.play loop of BASELINE config 1, side channels no-op'd (most favourable variant)")
        extra = {"language": "python", "variant": "pure_sim", "socket_stub_value": rates["socket_stub"][0], "port": port}
    else:                                                     # should not happen: oracle/_ref ships with the snapshot
        v, cores, dt, steps, envs = cpu_port_run(args.steps, W)
        kind, sample, extra = "port", port["sample"], {"port": port, "note": "oracle/_ref absent: C port timed instead"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": W, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "sample": sample},
        "cpu_baseline": dict({"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, **extra),
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_name(gpus: int) -> str:
    return (f"{ENVS_PER_GPU} envs/GPU x {gpus} GPU, FP32, K={FRAMES_PER_STEP} fused frames/launch, per-frame uint8 actions "
            "(quad-major), computerAI opponent + Rewards reward; written per env-frame: one event byte = reward, done, score, "
            "game_over; per env: obs_next[n,8]; in-kernel stats, copied out"
            + (" and all-reduced over NCCL" if gpus > 1 else "") + f" once per rollout ({ROLLOUT_STEPS} steps)")


# ----------------------------------------------------------------------------------------------- GPU arm
ROLLOUT_STEPS = 16           # bench steps per "rollout": the episode/score statistics are all-reduced once per rollout


class StatsReducer:
    """The path's only exchange: once per rollout (every ROLLOUT_STEPS steps, and once more at the
    end of a timed region) the int64[AH_NSTATS] statistics vector the step kernel accumulated on the device is copied out
    and cleared (`ah_stats_bank`, one tiny launch) and, at N > 1, summed over the ranks with one asynchronous NCCL
    all-reduce (160 B over NVLink).  Nothing in the step path waits for it: with an InterleavedStepper the copy runs on a
    private stream behind exactly the steps of the finished rollout while later steps accumulate into the other
    statistics bank (`close_rollout`); on a single stream NCCL's stream waits for the step kernel and the next step's
    kernel does not wait for NCCL.  The reduced vectors are accumulated so that the payload can be checked:
    frames == world x envs x K x steps."""

    def __init__(self, env, world, dev, every=ROLLOUT_STEPS, stepper=None, peer=None):
        import torch
        from air_hockey_training_environment_b200 import STAT_NAMES
        self.env, self.world, self.every, self.stepper, self.dev = env, world, every, stepper, dev
        self.peer = peer if (peer is not None and peer.available and world > 1) else None   # fused P2P all-reduce
        self.bufs = [torch.zeros(len(STAT_NAMES), dtype=torch.int64, device=dev) for _ in range(2)]
        self.acc = torch.zeros(len(STAT_NAMES), dtype=torch.int64, device=dev)
        self.pending, self.j, self.count, self.reductions = None, 0, 0, 0

    def _collect(self):
        import torch
        if self.pending is not None:
            work, buf, side = self.pending
            if work is not None:
                work.wait()                    # stream-level wait, no host sync
            if side is not None:
                torch.cuda.current_stream(self.dev).wait_stream(side)
            self.acc += buf
            self.pending = None

    def reduce(self):
        import torch
        import torch.distributed as dist
        self._collect()
        buf = self.bufs[self.j]
        self.j ^= 1
        side = None
        if self.stepper is not None:
            self.stepper.stats_stream.wait_stream(torch.cuda.current_stream(self.dev))   # buf's previous reader is done
            side = self.stepper.close_rollout(buf, peer=self.peer)
            work = None
            if self.peer is None and self.world > 1:
                with torch.cuda.stream(side):
                    work = dist.all_reduce(buf, async_op=True)
        elif self.peer is not None:
            self.peer.allreduce(buf, clear=True)
            work = None
        else:
            self.env.stats_tensor(clear=True, out=buf)
            work = dist.all_reduce(buf, async_op=True) if self.world > 1 else None
        self.pending = (work, buf, side)
        self.reductions += 1

    def after_step(self):
        self.count += 1
        if self.count % self.every == 0:
            self.reduce()

    def flush(self):
        self._collect()

    def restart(self):
        """Start a measured window: drain what earlier steps accumulated and zero the totals."""
        self.reduce()
        self._collect()
        self.acc.zero_()
        self.count, self.reductions = 0, 0

    def finish(self):
        if self.count % self.every != 0:
            self.reduce()
        self._collect()

    def check(self, expected_frames: int):
        v = self.acc.tolist()
        ok = v[0] == expected_frames
        if not ok:
            raise SystemExit(f"all-reduced statistics are wrong: frames {v[0]} != {expected_frames}")
        return {"frames": v[0], "expected_frames": expected_frames, "ok": ok, "reductions": self.reductions,
                "robot_goals": v[1], "opponent_goals": v[2], "robot_wins": v[3], "opponent_wins": v[4], "dones": v[5],
                "goals_per_1000_frames": 1000.0 * (v[1] + v[2]) / max(1, v[0])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=260)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--cpu-seconds", type=float, default=4.0)
    ap.add_argument("--ref-frames", type=int, default=1000, help="frames per worker per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="statistics all-reduce at N > 1: fused over NVLink peer memory (default) or copy + NCCL")
    ap.add_argument("--parts", type=int, default=2, help="env slices (streams) per step; 1 = a single launch per step")
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
    # host placement for the host-buffer (e2e) leg: cores and pinned staging memory next to this rank's GPU
    from air_hockey_training_environment_b200.dist import bind_host_to_gpu
    affinity0 = os.sched_getaffinity(0)
    host_binding = bind_host_to_gpu(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, K = args.envs_per_gpu, FRAMES_PER_STEP
    # warm-up: never fewer than 260 steps (2,080 frames): all games start from the same constructor state and stay
    # correlated (warp-uniform branches => optimistic timings) until about one mean game length (~1,900 frames) has passed
    W = max(260, args.warmup)

    env = BatchedAirHockey(n, dtype=torch.float32, device=dev, seed=1234, env_id0=rank * n)
    # the packed rollout mode: quad-major uint8 actions in, one event byte (reward, done, score, game_over) per
    # env-frame + the next observation out (csrc/ah_step_packed.cuh)
    out = env.make_outputs(K, ("events", "obs_next"))
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_act = 4                                                    # rotate several action tensors
    acts = [torch.randint(0, 4, ((K + 3) // 4, n, 4), dtype=torch.uint8, device=dev, generator=gen) for _ in range(n_act)]
    stream = torch.cuda.current_stream(dev)
    # Every step is issued as `--parts` disjoint env slices on as many streams (InterleavedStepper): the slices are
    # independent, so the GPU does not drain between steps -- the tail of one slice's launch overlaps the head of the
    # next one's.  --parts 1 = one launch per step on the current stream.
    from air_hockey_training_environment_b200 import InterleavedStepper
    stepper = InterleavedStepper(env, parts=args.parts) if args.parts > 1 else None
    # the statistics exchange: fused over NVLink peer memory (ah_stats_allreduce) when the ranks can map each other's
    # receive buffers (CUDA IPC), else a copy kernel + an NCCL all-reduce; --collective nccl forces the latter
    from air_hockey_training_environment_b200.dist import PeerStats
    peer = PeerStats(env) if (world > 1 and args.collective == "p2p") else None
    red = StatsReducer(env, world, dev, stepper=stepper, peer=peer)

    def one_step(i: int):
        if stepper is not None:
            stepper.step(acts[i % n_act], K, out)
        else:
            env.step(acts[i % n_act], K=K, out=out)
        red.after_step()

    def barrier():
        if stepper is not None:
            stepper.join()
        red.flush()
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
        time.sleep(0.3)
    barrier()
    for i in range(50):                       # the GPU idled while the sampler started: bring the clocks back up (untimed)
        one_step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    red.restart()
    align_start(dev, world)
    ev0.record(stream)
    t_issue0 = time.perf_counter()
    for i in range(args.steps):
        one_step(i)
    host_issue_us = 1e6 * (time.perf_counter() - t_issue0) / args.steps
    red.finish()                              # the last (partial) rollout's all-reduce is inside the timed region too
    if stepper is not None:
        stepper.join()
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    total_ms = ev0.elapsed_time(ev1)
    per_rank_ms = [total_ms]
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        every = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x.item()) for x in every]
        total_ms = max(per_rank_ms)                              # the MAX over ranks is the job's time
    value = world * n * K * args.steps / (total_ms * 1e-3)
    stats_check = red.check(world * n * K * args.steps)          # the collective's payload, verified on every rank

    clocks = None
    if rank == 0:
        clocks = sampler.window(t_wall0, t_wall1)
        clocks["window"] = "timed region"
    if total_ms < 600.0:
        # the timed region was shorter than a few nvidia-smi sampling periods: keep the same launches going for
        # ~1.2 s (untimed; the same iteration count on every rank) so that the recorded clocks are clocks UNDER THIS
        # LOAD, and say so
        iters = int(1200.0 / (total_ms / args.steps)) + 1
        t0 = time.time()
        for i in range(iters):
            one_step(i)
        barrier()
        if rank == 0:
            clocks = sampler.window(t0 + 0.1, time.time())
            clocks["window"] = "1.2 s of the same launches right after the (shorter) timed region"
    if rank == 0:
        sampler.stop()

    # ---- duration of the dominant kernel.  The timed region is a back-to-back queue of step-kernel launches and nothing
    # else (plus one tiny stats copy per rollout), so CUDA-event time / steps is the kernel time one step costs: with
    # --parts 1 that IS the average launch duration; with slices the launches of consecutive steps overlap tail-to-head
    # and it is the time per step's worth of launches.  Beside it: a single whole-n launch timed back-to-back on one
    # stream (`single_launch_ms`) and with an event pair around every launch (`kernel_ms_isolated`).
    kernel_ms_isolated = single_launch_ms = None
    kernel_ms = total_ms / args.steps                # at N > 1: the slowest rank's; the roofline below is per GPU
    if world == 1:
        barrier()
        single_launch_ms = _time_loop(dev, lambda i: env.step(acts[i % n_act], K=K, out=out), 20, max(args.steps, 100))
        reps = min(args.steps, 200)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i, (a, b) in enumerate(kev):
            a.record(stream)
            env.step(acts[i % n_act], K=K, out=out)
            b.record(stream)
        torch.cuda.synchronize(dev)
        kernel_ms_isolated = sum(a.elapsed_time(b) for a, b in kev) / len(kev)

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(env, dev, n, K, 5, max(10, min(args.steps, 40)), world)

    extras = {}
    if not args.no_extras:
        # BASELINE configs 4 and 5 at THIS world size (collective inside the timed region, payload checked)
        extras["ring_append_k8_cfg4"] = run_ring_append(env, dev, n, K, world)
        extras["ppo_rollout_cfg5"] = run_ppo_rollout(dev, world, rank)
    if world == 1 and not args.no_extras:
        extras["unpacked_outputs_k8"] = run_unpacked(env, dev, n, K)
        extras["rollout_k1"] = run_rollout_k1(env, dev, n)
        extras["rollout_k1_4m"] = run_rollout_k1(None, dev, 1 << 22)
        extras["philox_sim_k8"] = run_philox_sim(env, dev, n, K)
        extras["fp64_validation_cfg2"] = run_fp64(dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src, sm_max = peaks()
    bytes_per_launch = n * (2 * STATE_BYTES + 32 + K * (1 + 1))     # state in + out, obs_next, 1 action + 1 event byte per frame
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(world), "envs_per_gpu": n, "frames_per_step": K,
                   "sub_updates_per_sec": 3 * value,
                   "l2": "no flush: each step touches %.0f MB of distinct HBM lines (> 126 MB L2)" % (bytes_per_launch / 1e6),
                   # packed rollout kernel (csrc/ah_step_packed.cuh): one env per thread, 128-thread blocks, 9 resident per SM
                   "launch": {"kernel": "step_packed_kernel<float>", "threads": 128, "blocks_per_step": -(-n // 128),
                              "resident_blocks_per_sm": 9, "sms": env.launch_info()["sms"]},
                   "launches_per_step": max(1, args.parts), "host_issue_us_per_step": host_issue_us,
                   "per_rank_ms_per_step": [round(x / args.steps, 5) for x in per_rank_ms],
                   "interleaving": (f"each step = {args.parts} disjoint env slices on {args.parts} streams (sub-range launches of the same "
                                    "kernel); consecutive steps overlap tail-to-head" if args.parts > 1 else "one launch per step")},
        # per rank: one step kernel per step + one ah_stats copy kernel per rollout (every 16 steps and at the end)
        "gpu_launches": args.steps * max(1, args.parts) + stats_check["reductions"],
        "stats_allreduce": dict(stats_check, every_steps=ROLLOUT_STEPS,
                                collective=("none (1 GPU)" if world == 1 else
                                            ("fused P2P over NVLink peer memory (ah_stats_allreduce)" if red.peer is not None
                                             else "ah_stats_bank + NCCL all-reduce" + (f" [p2p unavailable: {peer.error}]" if peer is not None else ""))),
                                what="int64[20] episode/score statistics, summed over ranks (NCCL) once per rollout, "
                                     "inside the timed region; frames verified against world x envs x K x steps"),
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if kernel_ms is not None:
        sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
        lane_peak = 148 * 128 * sm_mhz * 1e6 / 1e9               # Glane-op/s at the SM clock seen under this load
        lane_ach = LANE_OPS_PER_ENV_STEP * n * K / (kernel_ms * 1e-3) / 1e9
        hbm_ach = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        line["roofline"] = {
            "bound": "fp32", "achieved": lane_ach, "peak": lane_peak, "unit": "Glane-op/s", "frac": lane_ach / lane_peak,
            "traffic": profile_fact("dram_bytes_per_launch"), "per": "GPU",
            "kernel": "ah::step_packed_kernel<float>", "kernel_ms": kernel_ms,
            "kernel_ms_isolated": kernel_ms_isolated, "single_launch_ms": single_launch_ms,
            "single_launch_frac": (LANE_OPS_PER_ENV_STEP * n * K / (single_launch_ms * 1e-3) / 1e9 / lane_peak) if single_launch_ms else None,
            "note": "K=8 fused frames: bound by FP32 instruction issue, not HBM (north_star: the slower of FP32 peak and "
                    "state bytes over HBM bandwidth). achieved = 254 algorithmic lane-ops/env-step x env-steps/s; "
                    "peak = 148 SMs x 128 FP32 lanes x SM clock under load. kernel_ms = timed region / steps (each step = "
                    "`launches_per_step` sub-range launches whose tails and heads overlap across steps); single_launch_* = one "
                    "whole-n launch per step on one stream. The HBM view of the same step follows.",
            "algorithmic_lane_ops_per_env_step": LANE_OPS_PER_ENV_STEP,
            "executed_warp_insts_per_warp_frame": profile_fact("warp_insts_per_warp_frame"),
            "alu_pipe_active_pct_ncu": profile_fact("alu_pipe_active_pct"),
            "issue_active_pct_ncu": profile_fact("issue_active_pct"),
            "hbm_achieved": hbm_ach, "hbm_peak": hbm_peak, "hbm_unit": "GB/s", "hbm_frac": hbm_ach / hbm_peak,
            "hbm_peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
            "bytes_per_env_step": bytes_per_launch / (n * K),
        }
    line["host_binding"] = host_binding
    if not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, affinity0)                       # the CPU baseline gets every host core
        line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
    if extras:
        line["extras"] = extras
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_align_buf = {}


def align_start(dev, world):
    """Queued right before a start event.  (1) The GPU spins (0.1 ms; ~1 ms at N > 1) so that the host's first launches
    are already queued when the region opens: the region measures K steps of device work, not the host's latency to
    issue the first one.  (2) At N > 1, AFTER the spin (whose length depends on each GPU's momentary clock), a
    one-element NCCL all-reduce: it completes on every GPU at the same moment, so the ranks' timed regions open together
    ON THE DEVICE -- the host barrier before it only aligns the hosts to a few hundred microseconds, which the collective
    that closes a ~2 ms timed region would otherwise expose as rank skew in the max-over-ranks time."""
    import torch
    import torch.distributed as dist
    torch.cuda._sleep(2_000_000 if world > 1 else 200_000)
    if world > 1:
        if dev not in _align_buf:
            _align_buf[dev] = torch.zeros(1, device=dev)
        dist.all_reduce(_align_buf[dev])


def _time_loop(dev, fn, warm, reps, world=1, red=None, join=None):
    """Average ms per call of fn(i): CUDA events on the current stream around `reps` calls; at world > 1 bracketed by
    barriers and reduced with MAX over the ranks.  `red` (a StatsReducer) is stepped after every call and its final
    all-reduce is inside the timed region.  `join` (an InterleavedStepper's): makes the current stream wait for the work fn
    queued on other streams, before the closing event."""
    import torch
    import torch.distributed as dist

    def sync():
        if join is not None:
            join()
        if red is not None:
            red.flush()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(warm):
        fn(i)
        if red is not None:
            red.after_step()
    sync()
    if red is not None:
        red.restart()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.current_stream(dev)
    align_start(dev, world)
    a.record(s)
    for i in range(reps):
        fn(i)
        if red is not None:
            red.after_step()
    if red is not None:
        red.finish()
    if join is not None:
        join()
    b.record(s)
    sync()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / reps


def run_rollout_k1(env, dev, n):
    """K = 1: one policy decision per launch, every output written -- the HBM-bound mode a PPO collector uses.
    `env` None => a fresh env of n games (used for the n = 2^22 point whose state does not fit in L2)."""
    import torch
    if env is None:
        from air_hockey_training_environment_b200 import BatchedAirHockey
        env = BatchedAirHockey(n, dtype=torch.float32, device=dev, seed=99)
        env.step(None, K=2000, out=env.make_outputs(2000, ()))        # decorrelate the games
    out = env.make_outputs(1, ("reward", "done", "score", "game_over", "obs_next"))
    acts = [torch.randint(0, 4, (n,), dtype=torch.int8, device=dev) for _ in range(4)]
    ms = _time_loop(dev, lambda i: env.step(acts[i % 4], K=1, out=out), 20, 200)
    hbm_peak, src, _ = peaks()
    b = n * (2 * STATE_BYTES + 32 + (1 + 4 + 1 + 1 + 1))
    ach = b / (ms * 1e-3) / 1e9
    return {"value": n / (ms * 1e-3), "unit": UNIT, "kernel_ms": ms, "envs": n,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": 127331072.0 if n == (1 << 20) else None,
                         "algorithmic_bytes_per_launch": b, "bytes_per_env_step": b / n,
                         "note": ("state (71 MB) < L2 (126 MB): ncu shows 127 MB of DRAM traffic for the 184.5 MB algorithmic "
                                  "bytes (profiles/r1g_other_variants.md); see rollout_k1_4m for the L2-proof point")
                         if n == (1 << 20) else "state 285 MB > 126 MB L2: algorithmic bytes are DRAM bytes"}}


def run_ring_append(env, dev, n, K, world=1):
    """BASELINE config 4: per GPU 2^20 envs, the headline step PLUS the fused replay-ring append of the Observation tuple
    (state, action, reward, done, new_state: 70 B per env-frame as reals, 51 B in the compact layout; replaces MemoryBuffer.append,
    lib/buffer.py:24-32), sharded over `world` GPUs with the per-rollout NCCL all-reduce of the statistics inside the
    timed region and its payload checked.  HBM-bound (write-heavy): the roofline is per GPU."""
    import torch
    from air_hockey_training_environment_b200 import InterleavedStepper, ReplayRing
    from air_hockey_training_environment_b200.env import pack_actions
    out = env.make_outputs(K, ("events", "obs_next"))
    acts = [pack_actions(torch.randint(0, 4, (K, n), dtype=torch.int8, device=dev)) for _ in range(4)]
    hbm_peak, src, _ = peaks()
    res = {}
    # the compact tuple layout (AH_RING_COMPACT: int16 locations + f32 velocities + int8 reward = 51 B, whole-sector stores) is
    # the one config 4 is quoted on; the rows layout (the reference's tuple as 70 B of reals) is timed beside it
    for key, compact, tuple_bytes in (("", True, 51), ("rows_layout_", False, 70)):
        ring = ReplayRing(2 * K, n, env.dtype, dev, compact=compact)
        reps = 96
        if compact:          # the step as 2 env slices on 2 streams, each with its own copy of the ring's counter
            st = InterleavedStepper(env, parts=2)
            red = StatsReducer(env, world, dev, stepper=st)
            ms = _time_loop(dev, lambda i: st.step(acts[i % 4], K, out, ring=ring), 16, reps, world, red, join=st.join)
            chk = red.check(world * n * K * reps)
            red1 = StatsReducer(env, world, dev)
            ms_one = _time_loop(dev, lambda i: env.step(acts[i % 4], K=K, out=out, ring=ring), 16, reps, world, red1)
            red1.check(world * n * K * reps)
        else:
            red = StatsReducer(env, world, dev)
            ms = _time_loop(dev, lambda i: env.step(acts[i % 4], K=K, out=out, ring=ring), 16, reps, world, red)
            chk = red.check(world * n * K * reps)
        assert ring.counter.tolist() == [ring.appended] * 4
        b = n * (2 * STATE_BYTES + 32 + K * (1 + 1) + K * tuple_bytes)      # state in + out, obs_next, action + event byte, the tuple
        ach = b / (ms * 1e-3) / 1e9
        del ring
        torch.cuda.empty_cache()
        if compact:
            res = {"value": world * n * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "envs_total": world * n,
                   "ring_capacity_frames": 2 * K, "ring_layout": "compact (51 B per tuple)", "launches_per_step": 2,
                   "single_launch_ms_per_step": ms_one, "stats_allreduce": chk,
                   "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                "per": "GPU", "traffic": profile_fact("ringc_dram_bytes_per_launch"),
                                "kernel": "ah::step_packed_kernel<float, IO_RING_COMPACT>",
                                "algorithmic_bytes_per_launch": b, "bytes_per_env_step": b / (n * K)}}
        else:
            res["rows_layout"] = {"value": world * n * K / (ms * 1e-3), "ms_per_step": ms, "stats_frames_ok": chk["ok"],
                                  "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                               "traffic": profile_fact("ring_dram_bytes_per_launch"),
                                               "kernel": "ah::step_packed_kernel<float, IO_RING>",
                                               "algorithmic_bytes_per_launch": b, "bytes_per_env_step": b / (n * K)}}
    return res


def run_unpacked(env, dev, n, K):
    """The same K = 8 launch through the general API's tensors: time-major int8 actions in, four separate per-frame outputs
    (reward f32, done u8, score i8, game_over u8, [K, n] each) out -- what ``BatchedAirHockey.step(actions)`` does by default.
    Runs on the packed-core kernel (IO_UNPACKED front-end); round 1's kernel (``packed_core=False``:
    ah::step_kernel<float, 1, LEAN>, round 1's headline) is timed beside it."""
    import torch
    from air_hockey_training_environment_b200 import BatchedAirHockey
    acts = [torch.randint(0, 4, (K, n), dtype=torch.int8, device=dev) for _ in range(4)]
    peak = 148 * 128 * peaks()[2] * 1e6
    res = {"unit": UNIT}
    for key, e in (("", env), ("round1_kernel_", None)):
        if e is None:
            e = BatchedAirHockey(n, dtype=torch.float32, device=dev, seed=1234, packed_core=False)
            e.step(None, K=2080, out=e.make_outputs(2080, ()))
        out = e.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next"))
        ms = _time_loop(dev, lambda i: e.step(acts[i % 4], K=K, out=out), 10, 100)
        res.update({key + "value": n * K / (ms * 1e-3), key + "kernel_ms": ms,
                    key + "lane_op_roofline_frac": LANE_OPS_PER_ENV_STEP * n * K / (ms * 1e-3) / peak})
    return res


def run_philox_sim(env, dev, n, K):
    """Pure simulation: actions drawn in-kernel (Philox), only the final observation written (packed-core kernel,
    IO_PHILOX front-end)."""
    out = env.make_outputs(K, ("obs_next",))
    ms = _time_loop(dev, lambda i: env.step(None, K=K, out=out), 10, 100)
    return {"value": n * K / (ms * 1e-3), "unit": UNIT, "kernel_ms": ms}


def run_ppo_rollout(dev, world=1, rank=0):
    """BASELINE config 5: per GPU 65,536 envs x 128 frames under a torch policy (the reference-shaped actor MLP), the
    rollout tensors written zero-copy on the device; one all-reduce of the statistics per rollout, inside the timed
    region.  Variants: (a) ANY torch callable evaluated by torch + the library's fused categorical sampler + env.step per
    frame, captured in one CUDA graph; (a') the same torch module handed to the collector as it is: recognised as a small
    dense network, exported (checked against the module) and run in ONE launch per rollout; (b) the fused actor+sampling kernel + env.step per frame in one CUDA graph;
    (c) the whole closed loop in ONE launch (in-kernel actor); (d) the same single launch for ANY small torch MLP exported
    with policy.DevicePolicy (the reference's PPO actor shape; the dueling Q-net with epsilon-greedy exploration).  At
    world > 1 only (a), (a') and (c) are timed."""
    import torch
    from air_hockey_training_environment_b200 import BatchedAirHockey
    from air_hockey_training_environment_b200.rollout import RolloutCollector
    n, T = 65536, 128
    res = {"envs_per_gpu": n, "frames": T, "unit": UNIT, "n_gpus": world}
    from air_hockey_training_environment_b200 import policy as ahp
    variants = [("torch_policy_graph", dict(fused_actor=False, auto_export=False)),
                # a plain torch module handed to the collector: exported after a numerical check, then ONE launch per rollout
                ("torch_module_auto_exported", dict(fused_actor=False)),
                ("fused_actor_graph", dict(single_launch=False)),
                ("single_launch", dict(single_launch=True)),
                # ANY small torch MLP, exported (policy.DevicePolicy): network + exploration strategy + physics in ONE launch
                ("exported_ppo_actor_single_launch", lambda: dict(device_policy=ahp.DevicePolicy(ahp.ppo_actor_discrete().to(dev), dev))),
                ("exported_dueling_qnet_eps_greedy_single_launch",
                 lambda: dict(device_policy=ahp.DevicePolicy(ahp.DuelingQNet(4).to(dev), dev, select="eps_greedy", epsilon=0.1,
                                                             initial_epsilon=0.1, final_epsilon=0.1)))]
    if world > 1:
        variants = [variants[0], variants[1], variants[3]]
    for name, kw in variants:
        env = BatchedAirHockey(n, dtype=torch.float32, device=dev, seed=7, env_id0=rank * n)
        env.step(None, K=2000, out=env.make_outputs(2000, ()))        # decorrelate the games
        if callable(kw):
            kw = kw()
        col = RolloutCollector(env, T, seed=7, **kw)
        red = StatsReducer(env, world, dev, every=1)
        reps = 10
        ms = _time_loop(dev, lambda i: col.collect(), 3, reps, world, red)
        chk = red.check(world * n * T * reps)
        res[name] = {"value": world * n * T / (ms * 1e-3), "ms_per_rollout": ms, "stats_frames_ok": chk["ok"]}
        del col, env
    res["value"] = res["single_launch"]["value"]
    res["note"] = ("obs[T+1,n,8], actions, probs[T,n,4], rewards, dones, scores, game_over written zero-copy into the "
                   "rollout tensors in all variants; one stats all-reduce per rollout in the timed region")
    return res


def run_fp64(dev):
    """BASELINE config 2 shape (4,096 envs, FP64 validation build), timed: 1,000 frames in launches of 100."""
    import torch
    from air_hockey_training_environment_b200 import BatchedAirHockey
    n = 4096
    env = BatchedAirHockey(n, dtype=torch.float64, device=dev, seed=3)
    env.step(None, K=2000, out=env.make_outputs(2000, ()))            # decorrelate the games
    out = env.make_outputs(100, ("reward", "done", "score", "game_over", "obs_next"))
    acts = torch.randint(0, 4, (100, n), dtype=torch.int8, device=dev)
    ms = _time_loop(dev, lambda i: env.step(acts, K=100, out=out), 2, 10)
    return {"value": n * 100 / (ms * 1e-3), "unit": UNIT, "ms_per_100_frames": ms, "envs": n,
            "parity": "bit-exact vs the reference (tests/test_gpu_parity.py)"}


def run_e2e(env, dev, n, K, warm, steps, world):
    """The public HOST-buffer API (air_hockey_training_environment_b200.host.HostStepper): per step, H2D of the uint8
    quad-major actions from pinned host memory, the fused packed step, and D2H of the step's results (one event byte per
    env-frame = reward, done, score, game_over; the next observation as 4 x int16 + 4 x f32) into pinned host memory --
    all inside the timed region; 4 slots / 3 streams so the copies overlap the kernels.  Every result is waited for on
    the host before its slot is reused.  Its roofline is the host link: the same copies WITHOUT the kernels, on the same
    streams and buffers, concurrently on every rank (`roofline.peak`)."""
    import torch
    import torch.distributed as dist
    from air_hockey_training_environment_b200.host import HostStepper
    slots = 4
    st = HostStepper(env, K=K, slots=slots)
    h_act = [torch.randint(0, 4, ((K + 3) // 4, n, 4), dtype=torch.uint8).pin_memory() for _ in range(slots)]

    def run(count):
        inflight = []
        for i in range(count):
            inflight.append(st.submit(h_act[i % slots]))
            if len(inflight) == slots:
                inflight.pop(0).wait()                   # host sees step i - (slots - 1) before its slot is reused
        for r in inflight:
            r.wait()

    def copies_only(count):
        for i in range(count):
            s = i % slots
            with torch.cuda.stream(st.s_in):
                st.d_act[s].copy_(h_act[s], non_blocking=True)
            with torch.cuda.stream(st.s_out):
                st.h_buf[s].copy_(st.d_buf[s], non_blocking=True)     # events + compact observation: one buffer per slot

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, count):
        sync()
        t0 = time.perf_counter()
        fn(count)
        sync()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    run(warm)
    dt = timed(run, steps)
    copies_only(warm)
    dt_copy = timed(copies_only, steps)
    last = st.h[(steps - 1) % slots]
    checksum = int(last.reward.sum(dtype=torch.int64)) + int(last.done.sum())     # the host really holds decodable results
    per_step = st.h2d_bytes_per_step + st.d2h_bytes_per_step
    return {"value": world * n * K * steps / dt, "unit": UNIT, "h2d_bytes_per_step": st.h2d_bytes_per_step,
            "d2h_bytes_per_step": st.d2h_bytes_per_step, "steps": steps, "host_checksum_last_step": checksum,
            "roofline": {"bound": "pcie", "unit": "GB/s (H2D + D2H, whole job)", "achieved": world * per_step * steps / dt / 1e9,
                         "peak": world * per_step * steps / dt_copy / 1e9, "frac": dt_copy / dt,
                         "peak_source": "measured in this run: the same pinned copies on the same streams without the kernels, "
                                        "all ranks concurrently"},
            "note": "HostStepper: pinned host actions in (1 B/frame/env); event byte (1 B/frame/env) and the next observation "
                    "(4 x int16 + 4 x f32 per env) out, lossless; 4 slots / 3 streams; host clock around barrier + synchronize"}


if __name__ == "__main__":
    main()
