"""GPU tests of the PACKED rollout mode (csrc/ah_step_packed.cuh): quad-major uint8 actions in, one event byte per
env-frame out.  It must be the SAME simulation as the unpacked step kernel -- bitwise equal state, the same
rewards / dones / scores / game-over flags, the same statistics -- in both arithmetic builds, and the FP64 build must
stay bit-exact against the live-reference goldens and the C oracle through this path too."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def make_env(n, dtype=torch.float32, **kw):
    from air_hockey_training_environment_b200 import BatchedAirHockey
    return BatchedAirHockey(n, dtype=dtype, device="cuda", **kw)


def snapshot(env):
    return [t.clone() for t in (env.puck, env.robot, env.opponent, env.prev_dist, env.meta)]


def same(a, b):
    return all(torch.equal(x.view(torch.uint8), y.view(torch.uint8)) for x, y in zip(a, b))


def test_pack_unpack_helpers_roundtrip():
    from air_hockey_training_environment_b200.env import pack_actions
    a = torch.randint(0, 4, (10, 7), dtype=torch.int8, device="cuda")
    p = pack_actions(a)
    assert p.shape == (3, 7, 4) and p.dtype == torch.uint8
    back = p.permute(0, 2, 1).reshape(12, 7)[:10]
    assert torch.equal(back, a.to(torch.uint8))
    assert int(p.permute(0, 2, 1).reshape(12, 7)[10:].sum()) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("K", [8, 4, 1, 3, 13])
def test_packed_equals_unpacked_bitwise(dtype, K):
    """Same seed, same actions: the packed launch and the unpacked (LEAN) launch leave bit-identical state, the
    events decode to the unpacked outputs, obs_next and the statistics are equal.  Includes partial quads and
    out-of-range action bytes (no nudge, airhockey.py:74-84)."""
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    n, launches = 8192, 40
    a, b = make_env(n, dtype, seed=5), make_env(n, dtype, seed=5)
    for env in (a, b):                                     # get past the correlated start, with goals and game overs
        env.step(None, K=1500, out=env.make_outputs(1500, ()))
        env.stats(clear=True)
    gen = torch.Generator(device="cuda").manual_seed(K)
    for it in range(launches):
        acts = torch.randint(-1, 6, (K, n), dtype=torch.int8, device="cuda", generator=gen)
        out_a = a.step(acts, K=K, out=a.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next")))
        out_b = b.step(pack_actions(acts), K=K, out=b.make_outputs(K, ("events", "obs_next")))
        ev = unpack_events(out_b.events, K)
        assert torch.equal(ev["reward"].to(dtype), out_a.reward)
        assert torch.equal(ev["done"].to(torch.uint8), out_a.done)
        assert torch.equal(ev["score"], out_a.score)
        assert torch.equal(ev["game_over"], out_a.game_over)
        assert torch.equal(out_a.obs_next.view(torch.uint8), out_b.obs_next.view(torch.uint8))
    assert same(snapshot(a), snapshot(b))
    sa, sb = a.stats(), b.stats()
    assert sa == sb, (sa, sb)
    assert sa["frames"] == n * K * launches and sa["robot_goals"] + sa["opponent_goals"] > 0


@pytest.mark.parametrize("name,K", [("canonical_16x1000", 8), ("long_256x1000", 20), ("canonical_16x1000", 5)])
def test_packed_fp64_bit_exact_vs_golden(name, K):
    """The live reference's committed trajectories (tests/golden; `long` includes finished games) replayed through the
    packed path in launches of K frames: state bit-exact wherever the golden recorded it, events equal at every frame,
    obs_next bit-exact."""
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    T, n = g["actions"].shape
    env = make_env(n, torch.float64)
    table = torch.as_tensor(g["reset_table"], dtype=torch.float64).cuda().contiguous()
    acts = torch.as_tensor(g["actions"]).cuda()
    rec = int(g["record_every"]) if "record_every" in g.files else 1
    for t0 in range(0, T, K):
        out = env.step(pack_actions(acts[t0:t0 + K]), K=K, out=env.make_outputs(K, ("events", "obs_next")), reset_table=table)
        ev = unpack_events(out.events, K)
        assert np.array_equal(ev["reward"].cpu().numpy(), g["reward"][t0:t0 + K])
        assert np.array_equal(ev["score"].cpu().numpy(), g["score"][t0:t0 + K])
        assert np.array_equal(ev["done"].cpu().numpy().astype(np.uint8), g["done"][t0:t0 + K])
        assert np.array_equal(ev["game_over"].cpu().numpy(), g["game_over"][t0:t0 + K])
        if "obs_next" in g.files:
            assert np.array_equal(out.obs_next.cpu().numpy().view(np.uint64), g["obs_next"][t0 + K - 1].view(np.uint64))
        if (t0 + K) % rec == 0:
            st = torch.cat([env.puck, env.robot, env.opponent, env.prev_dist[:, None]], 1).cpu().numpy()
            assert np.array_equal(st.view(np.uint64), g["states"][(t0 + K) // rec - 1].view(np.uint64)), t0
    assert np.array_equal(env.reset_count.cpu().numpy(), g["reset_cursor"])
    s = env.stats()
    assert s["frames"] == T * n and s["games"] == (g["game_over"] != 0).sum()
    assert s["done_ambiguous"] == 0          # `done` (libm pow in the reference) is provably exact for this run


def test_packed_fp64_bit_exact_vs_oracle_philox_resets():
    """BASELINE config 2 shape through the packed path: 4,096 envs x 1,000 frames, in-kernel Philox resets, against
    the C oracle: final state, scores, reset counters bit-exact; every frame's events equal."""
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    n, T, K = 4096, 1000, 8
    rng = np.random.default_rng(11)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    st, sc = oracle.initial_state(n)
    cur = np.zeros(n, np.int32)
    ref = oracle.run_frames(st, sc, T, actions=actions, seed=99, reset_cursor=cur, nthreads=8,
                            want=("reward", "score", "done", "game_over"))
    env = make_env(n, torch.float64, seed=99)
    acts = torch.as_tensor(actions).cuda()
    for t0 in range(0, T, K):
        out = env.step(pack_actions(acts[t0:t0 + K]), K=K, out=env.make_outputs(K, ("events",)))
        ev = unpack_events(out.events, K)
        assert np.array_equal(ev["reward"].cpu().numpy().astype(np.int32), ref["reward"][t0:t0 + K])
        assert np.array_equal(ev["score"].cpu().numpy(), ref["score"][t0:t0 + K])
        assert np.array_equal(ev["done"].cpu().numpy().astype(np.uint8), ref["done"][t0:t0 + K])
        assert np.array_equal(ev["game_over"].cpu().numpy(), ref["game_over"][t0:t0 + K])
    got = torch.cat([env.puck, env.robot, env.opponent, env.prev_dist[:, None]], 1).cpu().numpy()
    assert np.array_equal(got.view(np.uint64), st.view(np.uint64))
    assert np.array_equal(env.scores.cpu().numpy().astype(np.int32), sc)
    assert np.array_equal(env.reset_count.cpu().numpy(), cur)
    s = env.stats()
    o = dict(zip(oracle.STAT_NAMES, ref["stats"]))
    for k in ("frames", "robot_goals", "opponent_goals", "robot_wins", "opponent_wins", "dones", "reward_sum", "reward_sq_sum"):
        assert s[k] == o[k], k
    assert s["done_ambiguous"] == 0


def test_packed_table_resets_beyond_speed_limit_and_injected_state():
    """Reset tables with |v| > 10 (incl. the dy < -10 -> +10 quirk, puck.py:102-103) and injected out-of-limit /
    about-to-finish states: the packed path's speed-limit bookkeeping (threshold carried as +inf) equals the unpacked
    kernel's flag, bitwise, in both builds."""
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    n, K = 4096, 8
    gen = torch.Generator(device="cuda").manual_seed(3)
    table = (torch.rand(n, 24, 2, dtype=torch.float64, device="cuda", generator=gen) * 40 - 20).contiguous()
    for dtype in (torch.float32, torch.float64):
        a, b = make_env(n, dtype, seed=1), make_env(n, dtype, seed=1)
        vel = (torch.rand(n, 2, device="cuda", generator=gen) * 50 - 25).to(dtype)                      # out-of-limit velocities
        pos = (torch.rand(n, device="cuda", generator=gen) * 860 + 20).to(dtype)
        for env in (a, b):
            env.puck[:, 2:4] = vel
            env.puck[:, 0] = pos
            env.meta[:, 0] = torch.where(torch.arange(n, device="cuda") % 3 == 0, 10, 9 | (9 << 16)).to(torch.int32)   # 10-goal flag carried in
        assert same(snapshot(a), snapshot(b))
        for it in range(12):
            acts = torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda", generator=gen)
            oa = a.step(acts, K=K, out=a.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next")), reset_table=table)
            ob = b.step(pack_actions(acts), K=K, out=b.make_outputs(K, ("events", "obs_next")), reset_table=table)
            ev = unpack_events(ob.events, K)
            assert torch.equal(ev["game_over"], oa.game_over) and torch.equal(ev["score"], oa.score)
            assert torch.equal(ev["reward"].to(dtype), oa.reward)
        assert same(snapshot(a), snapshot(b))
        assert a.stats() == b.stats()


def test_packed_mode_argument_errors():
    from air_hockey_training_environment_b200 import AhError
    from air_hockey_training_environment_b200.env import pack_actions
    env = make_env(256)
    acts = torch.randint(0, 4, (8, 256), dtype=torch.int8, device="cuda")
    with pytest.raises((AhError, ValueError)):             # events need packed actions
        env.step(acts, K=8, out=env.make_outputs(8, ("events",)))
    with pytest.raises((AhError, ValueError)):             # and exclude the unpacked per-frame outputs
        env.step(pack_actions(acts), K=8, out=env.make_outputs(8, ("events", "reward")))
    fr = make_env(256, friction=True)
    with pytest.raises(AhError):
        fr.step(pack_actions(acts), K=8, out=fr.make_outputs(8, ("events",)))


def test_packed_full_size_properties():
    """BASELINE config 3 (2^20 envs, FP32, K = 8) through the packed path: tallies == sums of the decoded events,
    K-fusion invariance (one launch of 8 == two launches of 4, bitwise), and the reference's goal-rate band."""
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    n, K = 1 << 20, 8
    a, b = make_env(n, seed=1234), make_env(n, seed=1234)
    for env in (a, b):
        env.step(None, K=2000, out=env.make_outputs(2000, ()))
        env.stats(clear=True)
    tot = {"reward": 0, "done": 0, "rg": 0, "og": 0, "go": 0}
    gen = torch.Generator(device="cuda").manual_seed(9)
    for it in range(6):
        acts = torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda", generator=gen)
        oa = a.step(pack_actions(acts), K=K, out=a.make_outputs(K, ("events", "obs_next")))
        for h in range(2):
            b.step(pack_actions(acts[4 * h:4 * h + 4]), K=4, out=b.make_outputs(4, ("events", "obs_next")))
        ev = unpack_events(oa.events, K)
        tot["reward"] += int(ev["reward"].sum(dtype=torch.int64)); tot["done"] += int(ev["done"].sum())
        tot["rg"] += int((ev["score"] > 0).sum()); tot["og"] += int((ev["score"] < 0).sum())
        tot["go"] += int((ev["game_over"] > 0).sum())
    assert same(snapshot(a), snapshot(b))
    s = a.stats()
    assert s["frames"] == 6 * K * n and s == b.stats()
    assert (s["reward_sum"], s["dones"], s["robot_goals"], s["opponent_goals"], s["games"]) == \
        (tot["reward"], tot["done"], tot["rg"], tot["og"], tot["go"])
    rate = 1e3 * (s["robot_goals"] + s["opponent_goals"]) / s["frames"]
    assert 7.0 < rate < 9.0, rate                            # reference: 2.79 + 5.14 per 1,000 frames


@pytest.mark.parametrize("parts", [2, 3])
def test_interleaved_slices_equal_one_launch(parts):
    """A step issued as disjoint env slices on several streams (env_range launches) is the same step: bitwise equal
    state, events, obs_next and statistics; the frame counter advances once per step."""
    from air_hockey_training_environment_b200 import InterleavedStepper
    n, K = 50000, 8                                         # ragged: the last slice is shorter
    a, b = make_env(n, seed=9), make_env(n, seed=9)
    for env in (a, b):
        env.step(None, K=600, out=env.make_outputs(600, ()))
        env.stats(clear=True)
    f0 = a.counters()["frame"]
    st = InterleavedStepper(b, parts=parts)
    assert sum(c for _, c in st.ranges) == n and len(st.ranges) == parts
    gen = torch.Generator(device="cuda").manual_seed(1)
    outs_a, outs_b = a.make_outputs(K, ("events", "obs_next")), b.make_outputs(K, ("events", "obs_next"))
    for it in range(10):
        acts = torch.randint(0, 4, (2, n, 4), dtype=torch.uint8, device="cuda", generator=gen)
        a.step(acts, K=K, out=outs_a)
        st.step(acts, K, outs_b)
        st.join()
        assert torch.equal(outs_a.events, outs_b.events) and torch.equal(outs_a.obs_next, outs_b.obs_next)
    assert same(snapshot(a), snapshot(b)) and a.stats() == b.stats()
    assert a.counters()["frame"] == b.counters()["frame"] == f0 + 10 * K
    from air_hockey_training_environment_b200 import AhError
    with pytest.raises(AhError):                            # sub-ranges exist in the packed mode only
        b.step(None, K=1, env_range=(0, 100))


def test_close_rollout_banks_partition_the_statistics():
    """close_rollout(): the statistics of the steps issued so far land in one bank, later steps in the other -- without a
    join in between.  Per-rollout vectors must add up to the single-stream totals and each carry exactly its own frames."""
    from air_hockey_training_environment_b200 import InterleavedStepper, STAT_NAMES
    n, K = 1 << 16, 8
    a, b = make_env(n, seed=4), make_env(n, seed=4)
    for env in (a, b):
        env.step(None, K=400, out=env.make_outputs(400, ()))
        env.stats(clear=True)
    st = InterleavedStepper(b, parts=2)
    gen = torch.Generator(device="cuda").manual_seed(2)
    oa, ob = a.make_outputs(K, ("events",)), b.make_outputs(K, ("events",))
    bufs = [torch.zeros(len(STAT_NAMES), dtype=torch.int64, device="cuda") for _ in range(3)]
    per_rollout = [5, 3, 4]
    for r, steps in enumerate(per_rollout):
        for _ in range(steps):
            acts = torch.randint(0, 4, (2, n, 4), dtype=torch.uint8, device="cuda", generator=gen)
            a.step(acts, K=K, out=oa)
            st.step(acts, K, ob)
        st.close_rollout(bufs[r])                            # no join: the next rollout's steps follow immediately
    st.join()
    torch.cuda.synchronize()
    for r, steps in enumerate(per_rollout):
        assert int(bufs[r][0]) == steps * K * n              # each rollout's vector holds exactly its own frames
    total = dict(zip(STAT_NAMES, (bufs[0] + bufs[1] + bufs[2]).tolist()))
    assert total == a.stats()
    assert b.stats(clear=False)["frames"] == 0 and int(b.stats_tensor(bank=1)[0]) == 0
    assert same(snapshot(a), snapshot(b))


def test_interleaved_steps_inside_a_cuda_graph():
    """A rollout of interleaved steps (2 slices on 2 streams per step, fork / join through events) captured in ONE CUDA
    graph and replayed: same state, events and statistics as the eager single-stream path."""
    from air_hockey_training_environment_b200 import InterleavedStepper
    n, K, steps = 1 << 15, 8, 4
    a, b = make_env(n, seed=6), make_env(n, seed=6)
    for env in (a, b):
        env.step(None, K=500, out=env.make_outputs(500, ()))
        env.stats(clear=True)
    gen = torch.Generator(device="cuda").manual_seed(4)
    acts = [torch.randint(0, 4, (2, n, 4), dtype=torch.uint8, device="cuda", generator=gen) for _ in range(steps)]
    outs_a = [a.make_outputs(K, ("events", "obs_next")) for _ in range(steps)]
    outs_b = [b.make_outputs(K, ("events", "obs_next")) for _ in range(steps)]
    st = InterleavedStepper(b, parts=2)
    saved = b.state_dict()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                            # warm-up outside the capture
        st.step(acts[0], K, outs_b[0]); st.join()
    torch.cuda.current_stream().wait_stream(side)
    b.load_state_dict(saved)
    b.stats(clear=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(steps):
            st.step(acts[k], K, outs_b[k])
        st.join()
    for rep in range(3):
        for k in range(steps):
            a.step(acts[k], K=K, out=outs_a[k])
        g.replay()
        torch.cuda.synchronize()
        for k in range(steps):
            assert torch.equal(outs_a[k].events, outs_b[k].events) and torch.equal(outs_a[k].obs_next, outs_b[k].obs_next)
    assert same(snapshot(a), snapshot(b)) and a.stats() == b.stats()
    assert a.counters() == b.counters()


# ---------------------------------------------------------------- the packed core behind the general API's tensors
@pytest.mark.parametrize("dtype,K", [(torch.float32, 8), (torch.float32, 5), (torch.float64, 7), (torch.float32, 3)])
def test_general_api_runs_on_the_packed_core_with_identical_results(dtype, K):
    """``env.step(int8 [K, N] actions)`` with the default four per-frame outputs (K > 2) runs on the packed-core kernel
    (IO_UNPACKED front-end: same arithmetic, the event byte decoded into the caller's tensors); ``packed_core=False`` keeps
    round 1's step kernel.  Bitwise the same outputs, state, statistics and counters, incl. out-of-range action bytes."""
    n = 3001
    a, b = make_env(n, dtype, seed=31), make_env(n, dtype, seed=31, packed_core=False)
    for env in (a, b):
        env.step(None, K=600, out=env.make_outputs(600, ()))
    g = torch.Generator(device="cuda").manual_seed(9)
    for launch in range(5):
        acts = torch.randint(-2, 7, (K, n), dtype=torch.int8, device="cuda", generator=g)
        oa, ob = a.step(acts, K=K), b.step(acts, K=K)
        for f in ("reward", "done", "score", "game_over", "obs_next"):
            assert torch.equal(getattr(oa, f), getattr(ob, f)), (launch, f)
    for x, y in zip((a.puck, a.robot, a.opponent, a.prev_dist, a.meta), (b.puck, b.robot, b.opponent, b.prev_dist, b.meta)):
        assert torch.equal(x, y)
    assert a.stats() == b.stats() and a.counters() == b.counters()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_philox_simulation_runs_on_the_packed_core_with_identical_results(dtype):
    """``env.step(None, K)`` without per-frame outputs (pure simulation, in-kernel Philox actions keyed on the device
    frame counter) on the packed core == round 1's kernel, across launches (the frame counter carries over)."""
    n = 2500
    a, b = make_env(n, dtype, seed=44), make_env(n, dtype, seed=44, packed_core=False)
    for K in (70, 1, 9, 130):                          # crosses the 64-frame Philox action blocks at odd offsets
        oa = a.step(None, K=K, out=a.make_outputs(K, ("obs_next",)))
        ob = b.step(None, K=K, out=b.make_outputs(K, ("obs_next",)))
        assert torch.equal(oa.obs_next, ob.obs_next)
        assert a.counters() == b.counters()
    for x, y in zip((a.puck, a.robot, a.opponent, a.prev_dist, a.meta), (b.puck, b.robot, b.opponent, b.prev_dist, b.meta)):
        assert torch.equal(x, y)
    assert a.stats() == b.stats()
