"""GPU tests of the drop-in API surface: the reference's own unit-test pins re-expressed against the CUDA
path, error behaviour, K-fusion / sharding / replay invariance, the fused ring buffer, CUDA-graph capture,
and size-independent properties at BASELINE's full sizes (1M envs FP32, K=8)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_env(n, dtype=torch.float32, **kw):
    from air_hockey_training_environment_b200 import BatchedAirHockey
    return BatchedAirHockey(n, dtype=dtype, device="cuda", **kw)


def snapshot(env):
    return [t.clone() for t in (env.puck, env.robot, env.opponent, env.prev_dist, env.meta)]


def same(a, b):
    return all(torch.equal(x.view(torch.uint8), y.view(torch.uint8)) for x, y in zip(a, b))


# ---------------------------------------------------------------- the reference's unit pins, on the GPU
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_initial_state_and_observation(dtype):
    env = make_env(8, dtype)
    assert env.get_state("robot")[0].tolist() == [350, 240, 450, 240, 0, 0, -3, 0]
    assert env.get_state("opponent")[0].tolist() == [550, 240, 450, 240, 0, 0, -3, 0]
    assert env.scores.sum().item() == 0 and torch.isinf(env.prev_dist).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_environment_pin(dtype):
    # reference tests/test_environment.py:9-16 (with "computer" for the stale "opponent"): puck at (438, 240)
    env = make_env(4, dtype)
    d = env.device
    env.update_state(torch.tensor([[80.0, 100.0]] * 4, dtype=dtype, device=d), "robot")
    env.update_state(agent_name="computer")
    env.update_state(torch.tensor([[200.0, 234.0]] * 4, dtype=dtype, device=d), "robot")
    env.update_state(agent_name="computer")
    assert env.get_state()[0, 2:4].tolist() == [438, 240]


def test_invalid_agent_raises():
    from air_hockey_training_environment_b200 import InvalidAgentError
    env = make_env(4)
    with pytest.raises(InvalidAgentError):       # airhockey.py:108-110 ; tests/test_environment.py uses "opponent"
        env.update_state(torch.zeros(4, 2, device=env.device), "opponent")
    with pytest.raises(InvalidAgentError):
        env.update_state()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_mallet_limits(dtype):
    # reference tests/test_mallet.py:31-63: robot x in [20,430], opponent x in [470,880], y in [20,460]
    env = make_env(2, dtype)
    env.robot[:, 0:2] = torch.tensor([[-1.0, -1.0], [1000.0, 1000.0]], dtype=dtype)
    env.update_state(agent_name="computer")      # tail of update_state runs Mallet.update on both mallets
    assert env.robot[:, 0:2].tolist() == [[20, 20], [430, 460]]
    env.opponent[:, 2:4] = 0                     # computerAI left a velocity behind; the tail advances by it
    env.update_state(torch.tensor([[-1.0, -1.0], [1000.0, 1000.0]], dtype=dtype, device=env.device), "human")
    assert env.opponent[:, 0:2].tolist() == [[470, 20], [880, 460]]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_reward_predicates(dtype):
    # reference tests/test_puck.py:18-63 and tests/test_goal.py:14-40 through Rewards.__call__
    env = make_env(8, dtype)
    P = torch.tensor([[340, 230], [300, 190],        # overlap robot mallet at (350,240): yes / no
                      [10, 300], [100, 300],         # left wall: yes / no
                      [890, 300], [800, 300],        # right wall: yes / no
                      [10, 240], [900, 240]],        # in left goal / in right goal
                     dtype=dtype)
    env.puck[:, 0:2] = P.to(env.device)
    reward, score, done = env.rewards()
    assert done.tolist() == [1, 0, 0, 0, 0, 0, 0, 0]
    assert score.tolist() == [0, 0, -1, 0, 1, 0, -1, 1]          # y = 300 is inside the goal mouth (|240 - y| < 95)
    # reward = (3 | -1) + wall(-2 | +2 | 0) + (+1: closer than prev (inf))
    assert reward.tolist() == [4, 0, -2, 0, 2, 0, -2, 0]         # x = 900 is 18 px from right_wall = 882: no wall term
    assert not torch.isinf(env.prev_dist).any()
    # goal test fails away from the mouth: (10, 30) and (900, 30)   (tests/test_goal.py:25,39)
    env.puck[0:2, 0:2] = torch.tensor([[10, 30], [900, 30]], dtype=dtype).to(env.device)
    assert env.rewards()[1][0:2].tolist() == [0, 0]


def test_update_score_and_reset():
    env = make_env(6, torch.float64, seed=5)
    env.robot[:, 2:4] = 3.0
    rob = env.robot.clone()
    env.update_score(torch.tensor([1, -1, 0, 1, 0, -1], dtype=torch.int8, device=env.device))
    assert env.scores.tolist() == [[1, 0], [0, 1], [0, 0], [1, 0], [0, 0], [0, 1]]
    hit = torch.tensor([1, 1, 0, 1, 0, 1], dtype=torch.bool, device=env.device)
    assert (env.puck[hit, 0:2] == torch.tensor([450.0, 240.0], device=env.device, dtype=torch.float64)).all()
    assert (env.puck[hit, 2:4].abs() <= 10).all() and (env.puck[~hit, 2] == -3).all()
    assert (env.robot[hit, 2:4] == 0).all() and torch.equal(env.robot[:, 0:2], rob[:, 0:2])   # positions kept
    assert (env.robot[~hit, 2:4] == 3).all()
    assert env.reset_count.tolist() == [1, 1, 0, 1, 0, 1]
    env.reset(total=True)
    assert env.scores.sum().item() == 0 and env.reset_count.tolist() == [2, 2, 1, 2, 1, 2]
    # Philox draws equal the oracle's restatement
    from oracle import oracle
    want = oracle.rng_reset_pair(5, 0, 1)
    assert env.puck[0, 2:4].tolist() == want.tolist()


def test_fp32_reset_draws_equal_the_oracle_restatement():
    """The FP32 throughput build draws reset velocities from pcg3d keyed (splitmix64(seed), GLOBAL env id, reset index)
    (csrc/ah_rng.cuh); the oracle restates it (tests/test_reference_pins.py pins and tests that restatement on the CPU).
    Through the verbs (update_score, reset) and through the packed kernel's goal branch, also with env ids >= 2^32."""
    from oracle import oracle
    for seed, id0 in ((5, 0), (2 ** 63 + 11, 2 ** 40 + 5)):
        n = 64
        env = make_env(n, torch.float32, seed=seed, env_id0=id0)
        env.update_score(torch.ones(n, dtype=torch.int8, device=env.device))
        env.reset(total=False)
        env.reset(total=True)
        want = np.stack([oracle.rng_reset_pair_f32(seed, id0 + i, 2) for i in range(n)]).astype(np.float32)
        assert np.array_equal(env.puck[:, 2:4].cpu().numpy().view(np.uint32), want.view(np.uint32))
        assert env.reset_count.tolist() == [3] * n
    # in the frame loop: put the puck in the robot's goal mouth so that every env scores in its first frame
    from air_hockey_training_environment_b200.env import pack_actions
    n, seed, id0 = 4096, 77, 123456
    env = make_env(n, torch.float32, seed=seed, env_id0=id0)
    env.puck[:, 0] = 20.0; env.puck[:, 2:4] = 0.0
    acts = torch.full((1, n), 4, dtype=torch.int8, device=env.device)                  # no nudge
    for packed in (True, False):
        e = make_env(n, torch.float32, seed=seed, env_id0=id0, packed_core=packed)
        e.puck.copy_(env.puck)
        if packed:
            e.step(pack_actions(acts), K=1, out=e.make_outputs(1, ("events", "obs_next")))
        else:
            e.step(acts, K=1, out=e.make_outputs(1, ("reward", "score")))
        assert e.reset_count.tolist() == [1] * n and e.scores[:, 1].tolist() == [1] * n
        want = np.stack([oracle.rng_reset_pair_f32(seed, id0 + i, 0) for i in range(n)]).astype(np.float32)
        # the computer's sub-update follows the reset in the same frame: undo nothing, compare what it cannot change
        # (computerAI only adds its +2 nudge when the puck is within 40 px of the opponent mallet: not at the centre spot)
        got = e.puck[:, 2:4].cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), packed


def test_bad_shapes_are_rejected():
    env = make_env(16)
    with pytest.raises(ValueError):
        env.step(torch.zeros(3, 16, dtype=torch.int8, device=env.device), K=8)
    with pytest.raises(ValueError):
        env.step(torch.zeros(8, 16, dtype=torch.int8), K=8)          # CPU tensor
    with pytest.raises(ValueError):
        env.get_state(out=torch.zeros(16, 8, dtype=torch.float64, device=env.device))


# ---------------------------------------------------------------- invariances (bit-exact, both builds)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_k_fusion_invariance(dtype):
    """K fused frames in one launch == K launches of one frame."""
    n, K = 4096, 8
    a = make_env(n, dtype, seed=11)
    b = make_env(n, dtype, seed=11)
    for e in (a, b):
        e.step(None, K=200, out=e.make_outputs(200, ()))          # decorrelate
    acts = torch.randint(0, 4, (K, n), dtype=torch.int8, device=a.device)
    oa = a.step(acts, K=K, out=a.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next"), True))
    rs = []
    for k in range(K):
        ob = b.step(acts[k], K=1, out=b.make_outputs(1, ("reward", "done", "score", "game_over", "obs_next"), True))
        rs.append(ob)
    assert same(snapshot(a), snapshot(b))
    for f in ("reward", "done", "score", "game_over", "obs_next"):
        assert torch.equal(getattr(oa, f), torch.cat([getattr(o, f) for o in rs]))
    assert a.stats() == b.stats()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_sharding_invariance(dtype):
    """Results are keyed on the GLOBAL env id: one env of N == two shards of N/2 with env_id0 offsets."""
    n = 2048
    full = make_env(n, dtype, seed=21)
    lo = make_env(n // 2, dtype, seed=21, env_id0=0)
    hi = make_env(n // 2, dtype, seed=21, env_id0=n // 2)
    for e in (full, lo, hi):
        for _ in range(6):
            e.step(None, K=64, out=e.make_outputs(64, ()))
    for name in ("puck", "robot", "opponent", "prev_dist", "meta"):
        assert torch.equal(getattr(full, name), torch.cat([getattr(lo, name), getattr(hi, name)]))
    sf, sl, sh = full.stats(), lo.stats(), hi.stats()
    assert all(sf[k] == sl[k] + sh[k] for k in sf)
    assert sf["robot_goals"] + sf["opponent_goals"] > 100


def test_checkpoint_resume_is_exact():
    n = 1024
    a = make_env(n, seed=3)
    a.step(None, K=100, out=a.make_outputs(100, ()))
    sd = a.state_dict()
    a.step(None, K=100, out=a.make_outputs(100, ()))
    b = make_env(n, seed=3)
    b.load_state_dict(sd)
    b.step(None, K=100, out=b.make_outputs(100, ()))
    assert same(snapshot(a), snapshot(b))


def test_repeat_action_equals_tiled_actions():
    n, K = 512, 4
    a, b = make_env(n, seed=9), make_env(n, seed=9)
    act = torch.randint(0, 4, (n,), dtype=torch.int8, device=a.device)
    a.step(act, K=K, repeat=True)
    b.step(act[None].repeat(K, 1).contiguous(), K=K)
    assert same(snapshot(a), snapshot(b))


def test_friction_flag_changes_trajectories_only_when_enabled():
    a, b, c = make_env(64, seed=1), make_env(64, seed=1), make_env(64, seed=1, friction=True)
    for e in (a, b, c):
        e.step(None, K=50)
    assert same(snapshot(a), snapshot(b)) and not same(snapshot(a), snapshot(c))


# (the fused ring buffer, lib/buffer.py, has its own file: tests/test_gpu_ring.py)


# ---------------------------------------------------------------- CUDA graph capture
def test_step_is_graph_capturable_and_counters_advance_on_device():
    n, K = 8192, 8
    a, b = make_env(n, seed=13), make_env(n, seed=13)
    out = a.make_outputs(K, ("reward",))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        a.step(None, K=K, out=out)                       # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        a.step(None, K=K, out=out)
    for _ in range(5):
        g.replay()
    for _ in range(6):
        b.step(None, K=K)
    torch.cuda.synchronize()
    assert same(snapshot(a), snapshot(b))                # Philox frame counter advanced inside the graph
    assert a.counters()["frame"] == 6 * K


# ---------------------------------------------------------------- properties at BASELINE's full size
def test_full_size_properties_1m_envs_fp32_k8():
    n, K = 1 << 20, 8
    env = make_env(n, seed=1234)
    twin = make_env(n, seed=1234)
    out = env.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next"))
    tot = {"robot_goals": 0, "opponent_goals": 0, "dones": 0, "reward_sum": 0, "games": 0}
    for it in range(40):
        env.step(None, K=K, out=out)
        twin.step(None, K=K, out=twin.make_outputs(K, ()))
        tot["robot_goals"] += int((out.score > 0).sum()); tot["opponent_goals"] += int((out.score < 0).sum())
        tot["dones"] += int(out.done.sum()); tot["reward_sum"] += int(out.reward.sum().item())
        tot["games"] += int((out.game_over != 0).sum())
    st = env.stats()
    assert st["frames"] == 40 * K * n and st["nonfinite"] == 0
    for k, v in tot.items():                             # checksum of checksums: device tallies == output sums
        assert st[k] == v, (k, st[k], v)
    assert same(snapshot(env), snapshot(twin))           # determinism, independent of which outputs are written
    # table / limit invariants after any frame
    assert env.robot[:, 0].min() >= 20 and env.robot[:, 0].max() <= 430
    assert env.opponent[:, 0].min() >= 470 and env.opponent[:, 0].max() <= 880
    for m in (env.robot, env.opponent):
        assert m[:, 1].min() >= 20 and m[:, 1].max() <= 460
    assert env.puck[:, 2:4].abs().max() <= 10            # speed clamp precedes every advance
    assert env.puck[:, 0].min() >= 15 - 10 and env.puck[:, 0].max() <= 885 + 10
    assert env.scores.min() >= 0 and env.scores.max() <= 9
    assert set(torch.unique(out.reward).tolist()) <= {-4.0, -2.0, 0.0, 2.0, 4.0, 6.0}
    o = out.obs_next
    assert torch.equal(o[:, 0:4], o[:, 0:4].trunc())     # positions are int()-truncated in the observation
    rate = (st["robot_goals"] + st["opponent_goals"]) / st["frames"]
    assert 0.004 < rate < 0.012                          # reference: 7.9 goals per 1,000 frames


# ---------------------------------------------------------------- host-buffer front end
@pytest.mark.parametrize("K", [8, 6])
def test_host_stepper_round_trip_is_lossless(K):
    from air_hockey_training_environment_b200.host import HostStepper
    n = 8192
    env, twin = make_env(n, seed=31), make_env(n, seed=31)
    for e in (env, twin):
        e.step(None, K=300, out=e.make_outputs(300, ()))              # decorrelate, get some dones / goals (caller's stream:
    st = HostStepper(env, K=K, slots=3)                               #  submit() orders itself after it)
    acts = [torch.randint(0, 4, (K, n), dtype=torch.int8) for _ in range(5)]
    results = []
    for a in acts:
        r = st.submit(HostStepper.pack_actions_host(a)).wait()
        results.append((r.reward.clone(), r.done.clone(), r.score.clone(), r.game_over.clone(), r.obs.clone()))
    st.join()                                                         # env state / stats are read on the caller's stream
    dones = 0
    for a, (rw, dn, sc, go, ob) in zip(acts, results):
        out = twin.step(a.cuda(), K=K)
        assert torch.equal(rw.float(), out.reward.cpu()) and torch.equal(dn, out.done.cpu().bool())
        assert torch.equal(sc, out.score.cpu()) and torch.equal(go, out.game_over.cpu())
        assert torch.equal(ob, out.obs_next.cpu())
        dones += int(dn.sum())
    assert same(snapshot(env), snapshot(twin)) and env.stats() == twin.stats()
    q = (K + 3) // 4
    assert st.d2h_bytes_per_step == n * (4 * q + 8 + 16) and st.h2d_bytes_per_step == n * 4 * q
    assert dones > 0
    with pytest.raises(ValueError):
        st.submit(acts[0])                                            # time-major int8 is not the wire format


# ---------------------------------------------------------------- the HBM-bound small-K variant (persistent grid + prefetch)
@pytest.mark.parametrize("K", [1, 2])
def test_prefetch_variant_is_bitwise_identical(K):
    """For K <= 2 and n >= 65,536 the FP32 LEAN launch uses a persistent grid that prefetches the next env's rows;
    it must produce exactly what the one-env-per-thread kernel produces (`prefetch=False` selects the latter)."""
    n = (1 << 17) + 1000                                   # not a multiple of the grid: exercises the loop tail
    a, b = make_env(n, seed=17), make_env(n, seed=17, prefetch=False)
    fields = ("reward", "done", "score", "game_over", "obs_next")
    gen = torch.Generator(device=a.device).manual_seed(3)
    for it in range(40):
        acts = torch.randint(0, 4, (K, n), dtype=torch.int8, device=a.device, generator=gen)
        acts = acts[0].contiguous() if K == 1 else acts
        oa = a.step(acts, K=K, out=a.make_outputs(K, fields))
        ob = b.step(acts, K=K, out=b.make_outputs(K, fields))
        for f in fields:
            assert torch.equal(getattr(oa, f), getattr(ob, f)), (it, f)
    assert same(snapshot(a), snapshot(b))
    assert a.stats() == b.stats() and a.stats()["frames"] == 40 * K * n


# ---------------------------------------------------------------- edge cases and error paths
def test_invalid_k_and_pointer_errors():
    from air_hockey_training_environment_b200 import AhError
    env = make_env(64)
    for K in (0, 4097):
        with pytest.raises(AhError):
            env.step(None, K=K, out=env.make_outputs(max(K, 1), ()))
    with pytest.raises(ValueError):
        env.step(torch.zeros(64, dtype=torch.int8, device=env.device), K=1, policy_weights=torch.zeros(114, device=env.device))
    with pytest.raises(ValueError):
        env.step(None, K=1, out=env.make_outputs(1, ("probs",)))      # probs only exist in policy mode
    with pytest.raises(ValueError):
        make_env(16, dtype=torch.float16)


def test_reset_table_overflow_and_nonfinite_are_counted():
    env = make_env(256, torch.float64, seed=2)
    table = torch.zeros(256, 1, 2, dtype=torch.float64, device=env.device)       # room for ONE reset per env
    env.step(None, K=2000, out=env.make_outputs(2000, ()), reset_table=table)
    st = env.stats()
    assert st["reset_table_overflow"] > 0 and st["nonfinite"] == 0
    env2 = make_env(64, torch.float64, seed=3)      # (the FP32 build's FMNMX clamps turn a NaN into a table bound)
    env2.puck[5, 0] = float("nan")
    env2.step(None, K=4)
    assert env2.stats()["nonfinite"] == 1


def test_ragged_sizes_and_single_env():
    """n = 1, 31, 33, 257: partial warps and partial blocks; each env must equal the same env inside a big batch."""
    big = make_env(257, torch.float64, seed=8)
    big.step(None, K=300, out=big.make_outputs(300, ()))
    for n in (1, 31, 33, 257):
        e = make_env(n, torch.float64, seed=8)
        e.step(None, K=300, out=e.make_outputs(300, ()))
        assert torch.equal(e.puck, big.puck[:n]) and torch.equal(e.meta, big.meta[:n])


def test_continuous_repeat_and_ring_with_continuous_actions():
    from air_hockey_training_environment_b200 import ReplayRing
    n, K = 128, 3
    a, b = make_env(n, seed=6), make_env(n, seed=6)
    act = (torch.rand(n, 2, device=a.device) * 6 - 3).contiguous()              # game.py:122 range
    ring = ReplayRing(4, n, a.dtype, a.device, continuous=True)
    a.step(act, K=K, repeat=True, ring=ring, out=a.make_outputs(K, ("reward",)))
    b.step(act[None].repeat(K, 1, 1).contiguous(), K=K, out=b.make_outputs(K, ("reward",)))
    assert same(snapshot(a), snapshot(b))
    assert len(ring) == 3 and torch.equal(ring.retreive()["action"][0], act)
    with pytest.raises(ValueError):
        a.step(torch.zeros(n, dtype=torch.int8, device=a.device), K=1, ring=ring)   # discrete actions into a continuous ring


# ---------------------------------------------------------------- out-of-bounds writes: guard bands (compute-sanitizer is closed on this pool)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_guard_bands_survive_every_kernel(dtype):
    """Every state and output buffer is carved out of a larger allocation whose margins hold a sentinel; after running
    all kernels / variants on a ragged n the margins must be untouched and the kernels must have written the interior."""
    n, K, G = 1000, 5, 64                                  # G guard elements on both sides
    env = make_env(n, dtype, seed=4)
    dev = env.device
    guards = []

    def carve(shape, dt, fill=None):
        numel = int(np.prod(shape))
        el = torch.empty((), dtype=dt).element_size()
        pad = (G * 16) // el                              # keep 16-byte alignment of the interior
        raw = torch.full((numel + 2 * pad,), 77, dtype=dt, device=dev)
        inner = raw[pad:pad + numel].view(*shape)
        if fill is not None:
            inner.copy_(fill)
        guards.append((raw, pad, numel))
        return inner

    env.puck, env.robot, env.opponent = carve((n, 4), dtype, env.puck), carve((n, 4), dtype, env.robot), carve((n, 4), dtype, env.opponent)
    env.prev_dist, env.meta = carve((n,), dtype, env.prev_dist), carve((n, 4), torch.int32, env.meta)
    env._bind()
    from air_hockey_training_environment_b200 import ReplayRing, StepOutputs
    out = StepOutputs(obs_state=carve((K, n, 8), dtype), obs_new_state=carve((K, n, 8), dtype), reward=carve((K, n), dtype),
                      done=carve((K, n), torch.uint8), score=carve((K, n), torch.int8), game_over=carve((K, n), torch.uint8),
                      obs_next=carve((K, n, 8), dtype), actions=carve((K, n), torch.int8))
    ring = ReplayRing(3, n, dtype, dev)
    ring.state, ring.new_state = carve((3, n, 8), dtype), carve((3, n, 8), dtype)
    ring.action, ring.reward, ring.done = carve((3, n), torch.int8), carve((3, n), dtype), carve((3, n), torch.uint8)
    from air_hockey_training_environment_b200 import _capi
    ring.counter = carve((4,), torch.int64, ring.counter)            # AH_RING_COUNTERS copies
    ring._c = _capi.AhRing(ring.state.data_ptr(), ring.new_state.data_ptr(), ring.action.data_ptr(), ring.reward.data_ptr(),
                           ring.done.data_ptr(), 3, ring.counter.data_ptr())
    acts = torch.randint(0, 4, (K, n), dtype=torch.int8, device=dev)
    for _ in range(40):
        env.step(acts, K=K, out=out, ring=ring)                                   # GENERIC
    lean = StepOutputs(reward=out.reward, done=out.done, score=out.score, game_over=out.game_over, obs_next=carve((n, 8), dtype))
    env.step(acts, K=K, out=lean)                                                 # LEAN
    env.step(None, K=K, out=StepOutputs(obs_next=lean.obs_next))                  # SIM + Philox
    env.update_state(acts[0].contiguous(), "robot"); env.update_state(agent_name="computer")
    obs = carve((n, 8), dtype); env.get_state("robot", out=obs)
    r, s, d = env.rewards(); env.update_score(s); env.reset(total=False, mask=(s != 0))
    pw = torch.randn(114, device=dev) * 0.1
    pol = StepOutputs(reward=out.reward, done=out.done, obs_next=out.obs_next, actions=out.actions, probs=carve((K, n, 4), dtype))
    env.step(None, K=K, out=pol, policy_weights=pw)                               # in-kernel policy
    from air_hockey_training_environment_b200.env import pack_actions
    packed = StepOutputs(events=carve(((K + 3) // 4, n, 4), torch.uint8), obs_next=lean.obs_next)
    env.step(carve(((K + 3) // 4, n, 4), torch.uint8, pack_actions(acts)), K=K, out=packed)   # PACKED (partial last quad)
    smp = {"state": carve((500, 8), dtype), "new_state": carve((500, 8), dtype), "action": carve((500,), torch.int8),
           "reward": carve((500,), dtype), "done": carve((500,), torch.uint8), "index": carve((500,), torch.int64)}
    ring.sample(500, env, out=smp)                                                # ring sampler
    torch.cuda.synchronize()
    for raw, pad, numel in guards:
        assert bool((raw[:pad] == 77).all()) and bool((raw[pad + numel:] == 77).all()), "guard band overwritten"
    assert not bool((out.reward == 77).any()) and not bool((pol.probs == 77).any())    # interiors were written
    assert env.stats()["frames"] == n * K * 44 and env.stats()["nonfinite"] == 0


def test_8m_envs_on_one_gpu():
    """BASELINE config 4's TOTAL size (2^23 envs) on a single GPU: 64-bit row offsets, stats, invariants."""
    n, K = 1 << 23, 8
    env = make_env(n, seed=5)
    out = env.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next"))
    for _ in range(6):
        env.step(None, K=K, out=out)
    st = env.stats()
    assert st["frames"] == 6 * K * n and st["nonfinite"] == 0
    assert int((out.score != 0).sum()) > 0 and int(out.reward[-1, -1000:].abs().sum()) > 0      # the far end was written
    small = make_env(4096, seed=5)                         # the first 4096 games of the big batch, alone
    for _ in range(6):
        small.step(None, K=K)
    assert torch.equal(env.puck[:4096], small.puck) and torch.equal(env.meta[:4096], small.meta)
    tail = make_env(4096, seed=5, env_id0=n - 4096)        # and the last 4096
    for _ in range(6):
        tail.step(None, K=K)
    assert torch.equal(env.puck[-4096:], tail.puck) and torch.equal(env.meta[-4096:], tail.meta)
