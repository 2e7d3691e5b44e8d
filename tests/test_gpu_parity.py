"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against
  (1) the committed trajectories of the LIVE Python reference (tests/golden/*.npz), and
  (2) the C oracle (oracle/ah_oracle.c, itself pinned bit-exact to the reference) on the same seeded inputs.

FP64 validation build: BIT-EXACT (uint64 patterns) -- stronger than the 1e-9 the north star asks for.
FP32 throughput build: teacher-forced per-frame tolerance 1e-3 px (positions) / 1e-4 (velocities).
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu

F64 = torch.float64


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def make_env(n, dtype=F64, **kw):
    from air_hockey_training_environment_b200 import BatchedAirHockey
    return BatchedAirHockey(n, dtype=dtype, device="cuda", **kw)


def put_state(env, state13, scores=None, reset_cursor=None):
    dev, dt = env.device, env.dtype
    s = torch.as_tensor(state13, dtype=torch.float64)
    env.puck.copy_(s[:, 0:4].to(dt))
    env.robot.copy_(s[:, 4:8].to(dt))
    env.opponent.copy_(s[:, 8:12].to(dt))
    env.prev_dist.copy_(s[:, 12].to(dt))
    meta = torch.zeros(env.n, 4, dtype=torch.int32)
    if scores is not None:
        sc = torch.as_tensor(scores, dtype=torch.int32)
        meta[:, 0] = sc[:, 0] | (sc[:, 1] << 16)
    if reset_cursor is not None:
        meta[:, 1] = torch.as_tensor(reset_cursor, dtype=torch.int32)
    env.meta.copy_(meta.to(dev))


def get_state13(env):
    return torch.cat([env.puck, env.robot, env.opponent, env.prev_dist[:, None]], dim=1).double().cpu().numpy()


def get_scores(env):
    return env.scores.cpu().numpy().astype(np.int32)


ALL = ("obs_state", "obs_new_state", "reward", "done", "score", "game_over", "obs_next")


def run_gpu(env, T, actions, table, chunk):
    """T frames in launches of `chunk` fused frames; returns stacked per-frame outputs."""
    dev = env.device
    acts = None if actions is None else torch.as_tensor(actions).to(dev)
    tab = None if table is None else torch.as_tensor(table, dtype=torch.float64).to(dev).contiguous()
    outs = {k: [] for k in ALL}
    states, scores = [], []
    for t0 in range(0, T, chunk):
        out = env.make_outputs(chunk, ALL, obs_next_per_frame=True)
        a = None if acts is None else acts[t0:t0 + chunk].contiguous()
        env.step(a, K=chunk, out=out, reset_table=tab)
        for k in ALL:
            outs[k].append(getattr(out, k).cpu().numpy())
        states.append(get_state13(env))
        scores.append(get_scores(env))
    res = {k: np.concatenate(v, axis=0) for k, v in outs.items()}
    res["states"] = np.stack(states)
    res["scores"] = np.stack(scores)
    return res


def check_events(g, res):
    assert np.array_equal(np.asarray(g["reward"], np.float64), res["reward"].astype(np.float64))
    assert np.array_equal(g["score"], res["score"])
    assert np.array_equal(g["done"], res["done"])
    assert np.array_equal(g["game_over"], res["game_over"])


# ------------------------------------------------------------------------------ FP64 vs the live reference
def test_build_has_no_fma_contraction():
    assert make_env(32).compiled_without_fma_contraction()


@pytest.mark.parametrize("chunk", [1, 1000])
def test_fp64_golden_canonical(golden_dir, chunk):
    g = np.load(os.path.join(golden_dir, "canonical_16x1000.npz"))
    T, n = g["actions"].shape
    env = make_env(n)
    res = run_gpu(env, T, g["actions"], g["reset_table"], chunk)
    check_events(g, res)
    for k in ("obs_state", "obs_new_state", "obs_next"):
        assert np.array_equal(bits(g[k]), bits(res[k])), k
    if chunk == 1:
        assert np.array_equal(bits(g["states"]), bits(res["states"]))
        assert np.array_equal(g["scores"], res["scores"])
    else:
        assert np.array_equal(bits(g["states"][-1]), bits(res["states"][-1]))
        assert np.array_equal(g["scores"][-1], res["scores"][-1])
    assert np.array_equal(g["reset_cursor"], env.reset_count.cpu().numpy())


def test_fp64_golden_long(golden_dir):
    g = np.load(os.path.join(golden_dir, "long_256x1000.npz"))
    T, n = g["actions"].shape
    every = int(g["record_every"])
    env = make_env(n)
    res = run_gpu(env, T, g["actions"], g["reset_table"], every)
    check_events(g, res)
    assert np.array_equal(bits(g["states"]), bits(res["states"]))
    assert np.array_equal(g["scores"], res["scores"])
    assert (g["game_over"] != 0).sum() > 0
    st = env.stats()
    assert st["frames"] == T * n
    assert st["robot_goals"] == (g["score"] > 0).sum() and st["opponent_goals"] == (g["score"] < 0).sum()
    assert st["robot_wins"] == (g["game_over"] == 1).sum() and st["opponent_wins"] == (g["game_over"] == 2).sum()
    assert st["dones"] == g["done"].sum() and st["reward_sum"] == g["reward"].astype(np.int64).sum()
    assert st["reward_sq_sum"] == (g["reward"].astype(np.int64) ** 2).sum()
    assert st["games"] == (g["game_over"] != 0).sum() and st["nonfinite"] == 0 and st["reset_table_overflow"] == 0


def test_fp64_golden_friction(golden_dir):
    """`ah_config.friction` against the live reference with its own (dead) Puck.friction_on_puck hooked in before
    limit_puck_speed (tests/golden/friction_24x150.npz, generated by oracle/gen_golden.py: gen_friction): bit-exact."""
    g = np.load(os.path.join(golden_dir, "friction_24x150.npz"))
    T, n = g["actions"].shape
    for chunk in (1, 6):
        env = make_env(n, friction=True)
        put_state(env, g["init_state"])
        res = run_gpu(env, T, g["actions"], g["reset_table"], chunk)
        check_events(g, res)
        assert np.array_equal(bits(g["states"][chunk - 1::chunk]), bits(res["states"]))
        assert np.array_equal(g["scores"][chunk - 1::chunk], res["scores"])


def test_fp64_golden_scrambled(golden_dir):
    g = np.load(os.path.join(golden_dir, "scrambled_96x200.npz"))
    T, n = g["actions"].shape
    env = make_env(n)
    put_state(env, g["init_state"], g["init_scores"])
    res = run_gpu(env, T, g["actions"], g["reset_table"], 1)
    check_events(g, res)
    assert np.array_equal(bits(g["states"]), bits(res["states"]))
    assert np.array_equal(g["scores"], res["scores"])
    for k in ("obs_state", "obs_new_state", "obs_next"):
        assert np.array_equal(bits(g[k]), bits(res[k])), k


def test_fp64_golden_continuous(golden_dir):
    g = np.load(os.path.join(golden_dir, "continuous_16x300.npz"))
    T, n = g["actions"].shape[:2]
    env = make_env(n)
    res = run_gpu(env, T, g["actions"], g["reset_table"], 50)
    check_events(g, res)
    for k in ("obs_state", "obs_new_state", "obs_next"):
        assert np.array_equal(bits(g[k]), bits(res[k])), k
    assert np.array_equal(bits(g["states"][49::50]), bits(res["states"]))


def test_fp64_golden_substeps(golden_dir):
    """Raw AirHockey.update_state calls mixing robot (discrete / continuous), human and computer."""
    g = np.load(os.path.join(golden_dir, "substeps_24x40.npz"))
    n_seq, L = g["agents"].shape
    env = make_env(n_seq)
    put_state(env, g["init_state"])
    dev = env.device
    names = ("robot", "human", "computer")
    # all sequences advance in lockstep, but each has its own agent at step t: apply per-agent with masks by
    # running every agent kind on a scratch copy and merging -- simpler: one env per sequence, one call each.
    for s in range(n_seq):
        e1 = make_env(1)
        put_state(e1, g["init_state"][s:s + 1])
        for t in range(L):
            ag = int(g["agents"][s, t])
            if ag == 0 and g["kinds"][s, t] == 0:
                e1.update_state(torch.tensor([int(g["ia"][s, t])], dtype=torch.int8, device=dev), "robot")
            elif ag == 0:
                e1.update_state(torch.tensor(g["axy"][s, t:t + 1], dtype=F64, device=dev), "robot")
            elif ag == 1:
                e1.update_state(torch.tensor(g["axy"][s, t:t + 1], dtype=F64, device=dev), "human")
            else:
                e1.update_state(agent_name="computer")
            got = get_state13(e1)[0, :12]
            assert np.array_equal(bits(got), bits(g["states"][s, t])), (s, t, names[ag])


# ------------------------------------------------------------------------------ FP64 vs the C oracle, config 2
def test_fp64_config2_4096x1000_vs_oracle():
    """BASELINE config 2: 4,096 envs, FP64 validation build, 1,000 frames, trajectory + score diff."""
    n, T = 4096, 1000
    rng = np.random.default_rng(2024)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 64, 2))
    st, sc = oracle.initial_state(n)
    ref = oracle.run_frames(st, sc, T, actions=actions, reset_table=table, record_every=100, nthreads=8,
                            want=("reward", "score", "done", "game_over", "obs_next"))
    assert not ref["reset_overflow"]
    env = make_env(n)
    res = run_gpu(env, T, actions, table, 100)
    assert np.array_equal(ref["reward"], res["reward"].astype(np.int32))
    for k in ("score", "done", "game_over"):
        assert np.array_equal(ref[k], res[k]), k
    assert np.array_equal(bits(ref["obs_next"]), bits(res["obs_next"]))
    assert np.array_equal(bits(ref["states_rec"]), bits(res["states"]))
    assert np.array_equal(ref["scores_rec"], res["scores"])
    d = np.abs(np.nan_to_num(ref["states_rec"], posinf=0) - np.nan_to_num(res["states"], posinf=0))
    assert d.max() <= 1e-9          # the north star's tolerance (we are bit-exact)
    assert (ref["score"] != 0).sum() > 20000 and (ref["game_over"] != 0).sum() > 50


def test_fp64_philox_resets_and_actions_vs_oracle():
    """No tables at all: resets and actions drawn in-kernel (Philox) must equal the oracle's restatement."""
    n, T, seed, id0 = 2048, 640, 991, 12345
    st, sc = oracle.initial_state(n)
    ref = oracle.run_frames(st, sc, T, seed=seed, env_id0=id0, frame0=0, record_every=64, nthreads=8,
                            want=("reward", "score", "done", "game_over"))
    env = make_env(n, seed=seed, env_id0=id0)
    res = run_gpu(env, T, None, None, 64)
    assert np.array_equal(ref["reward"], res["reward"].astype(np.int32))
    for k in ("score", "done", "game_over"):
        assert np.array_equal(ref[k], res[k]), k
    assert np.array_equal(bits(ref["states_rec"]), bits(res["states"]))
    assert env.counters()["frame"] == T


# ------------------------------------------------------------------------------ FP32 throughput build
def teacher_forced_errors(n=4096, T=100, seed=5):
    """Per env-frame error of the FP32 build against the FP64 oracle when both start every frame from the SAME
    binary32-representable state (teacher forcing).  Returns flat arrays over the env-frames whose events agree."""
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 16, 2)).astype(np.float32).astype(np.float64)
    env = make_env(n, dtype=torch.float32)
    dev = env.device
    tab = torch.as_tensor(table).to(dev)
    st, sc = oracle.initial_state(n)
    cur = np.zeros(n, np.int32)
    warm = rng.integers(0, 4, (400, n)).astype(np.int8)                     # decorrelate with the oracle alone
    oracle.run_frames(st, sc, 400, actions=warm, reset_table=None, seed=3, reset_cursor=cur, want=())
    pos_cols, vel_cols = [0, 1, 4, 5, 8, 9], [2, 3, 6, 7, 10, 11]
    res = {k: [] for k in ("pos", "vel", "dist", "contacts", "dmin")}
    mismatch = 0
    for t in range(T):
        st = st.astype(np.float32).astype(np.float64)      # the state both sides start the frame from
        cur[:] = 0
        put_state(env, st, sc, cur)
        out = env.make_outputs(1, ("reward", "done", "score", "game_over"))
        env.step(torch.as_tensor(actions[t]).to(dev), K=1, out=out, reset_table=tab)
        ref = oracle.run_frames(st, sc, 1, actions=actions[t:t + 1], reset_table=table, reset_cursor=cur,
                                want=("reward", "score", "done", "game_over", "contacts", "contact_dmin"))
        got = get_state13(env)
        same = ((ref["score"][0] == out.score[0].cpu().numpy()) & (ref["done"][0] == out.done[0].cpu().numpy())
                & (ref["game_over"][0] == out.game_over[0].cpu().numpy())
                & (ref["reward"][0] == out.reward[0].cpu().numpy().astype(np.int32)))
        mismatch += int((~same).sum())
        assert np.array_equal(get_scores(env)[same], sc[same])              # scores follow the oracle
        ok = same & np.isfinite(st[:, 12])
        res["pos"].append(np.abs(got[ok][:, pos_cols] - st[ok][:, pos_cols]).max(axis=1))
        res["vel"].append(np.abs(got[ok][:, vel_cols] - st[ok][:, vel_cols]).max(axis=1))
        res["dist"].append(np.abs(got[ok][:, 12] - st[ok][:, 12]))
        res["contacts"].append(ref["contacts"][0][ok])
        res["dmin"].append(ref["contact_dmin"][0][ok])
    return {k: np.concatenate(v) for k, v in res.items()}, mismatch, n * T


def test_fp32_teacher_forced_tolerance():
    """FP32 throughput build, single-frame error (3 physics sub-updates) against the FP64 oracle.

    Stated tolerance, per frame:
      * frames WITHOUT a mallet-puck contact (97 % of frames):  |d position| <= 1e-3 px, |d velocity| <= 1e-4,
        |d prev_dist| <= 2e-3 -- a few binary32 ulps at table scale (ulp(900) = 6e-5);
      * frames WITH a contact: the collision response divides by the centre distance d and the reference's
        separation term ((max(35 - d_x, 35 - d_y) - 0.01) / 12 * 0.2 * n * 10, up to 12 px) is applied along the
        normalised direction, so binary32 input rounding (<= 1e-4 px) is amplified by ~12/d: the bound is
        scaled by (1 + 35/d_min) and 99 % of contact frames must still meet 10x the contact-free bound;
      * events (reward, score, done, game_over) identical, except when a comparison sits within binary32
        rounding of its threshold: at most 1 frame in 10,000.
    """
    r, mismatch, frames = teacher_forced_errors()
    free = r["contacts"] == 0
    hit = ~free
    assert free.sum() > 0.9 * frames and hit.sum() > 0.005 * frames
    assert r["pos"][free].max() <= 1e-3, r["pos"][free].max()
    assert r["vel"][free].max() <= 1e-4, r["vel"][free].max()
    assert r["dist"][free].max() <= 2e-3, r["dist"][free].max()
    scale = 1.0 + 35.0 / np.maximum(r["dmin"][hit], 1e-3)
    assert (r["pos"][hit] <= 1e-3 * scale).all(), (r["pos"][hit] / scale).max()
    assert (r["vel"][hit] <= 1e-3 * scale).all(), (r["vel"][hit] / scale).max()
    assert np.quantile(r["pos"][hit], 0.99) <= 1e-2 and np.quantile(r["vel"][hit], 0.99) <= 1e-2
    assert mismatch <= frames // 10000, mismatch


def test_fp32_tolerance_on_baseline_config3_launches():
    """The stated FP32 tolerance ON BASELINE's own configuration: 2^20 envs, FP32, K = 8 fused frames per launch (the
    packed rollout mode bench.py times), games decorrelated.  A contiguous block of 4,096 envs is followed: before a
    launch its binary32 state is handed (exactly) to the FP64 oracle, which plays the same 8 frames with the same
    actions and the same reset stream (the FP32 build's pcg3d draws, restated by the oracle); after the launch the two are compared.  Within one launch nothing is
    re-synchronised, so the per-frame bounds of test_fp32_teacher_forced_tolerance are scaled by the 8 frames:
      * env-launches without a contact: 99 % within |d position| <= 8e-3 px and |d velocity| <= 8e-4 (8 x the per-frame
        bound); the worst case is bounded by 5e-2 px / 2e-2 (a difference of one binary32 ulp in a position decides
        on which side of a clamp / reflection threshold a later sub-update falls, which moves the result by a
        velocity-sized amount times that ulp's lever; measured worst 2.6e-2 px / 7.9e-3);
      * with a contact (errors scaled by 1 / (1 + 35 / d_min) as in the per-frame test): 99 % within 1e-2 px / 2e-3 and
        99.9 % within 5e-2 px / 2e-2 (measured 2.0e-3 / 2.4e-4 and 1.5e-2 / 2.1e-3); the worst case over ~8,000 contact
        env-launches depends on the sample (a contact sequence amplifies a one-ulp difference through up to 24
        sub-updates: measured 2.6e-2 px with one reset stream, 2.6e-1 px with another) and is bounded loosely (1 px / 0.2);
      * all 8 frames' events (reward, score, done, game_over) identical except for <= 2 env-launches in 10,000
        (measured: 0 of 49,152)."""
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    n, K, lo, m, seed = 1 << 20, 8, 300_000, 4096, 4321
    env = make_env(n, dtype=torch.float32, seed=seed)
    env.step(None, K=2000, out=env.make_outputs(2000, ()))              # decorrelate (mean game length ~1,900 frames)
    gen = torch.Generator(device=env.device).manual_seed(7)
    out = env.make_outputs(K, ("events", "obs_next"))
    pos_cols, vel_cols = [0, 1, 4, 5, 8, 9], [2, 3, 6, 7, 10, 11]
    worst = {"pos_free": 0.0, "vel_free": 0.0, "pos_hit_scaled": 0.0, "vel_hit_scaled": 0.0}
    free_pos, free_vel, hit_pos, hit_vel = [], [], [], []
    mismatch = launches = hits = 0
    for it in range(12):
        acts = torch.randint(0, 4, (K, n), dtype=torch.int8, device=env.device, generator=gen)
        st = get_state13(env)[lo:lo + m]                                 # binary32 values, exact in binary64
        sc = get_scores(env)[lo:lo + m].copy()
        cur = env.reset_count[lo:lo + m].cpu().numpy().astype(np.int32)
        env.step(pack_actions(acts), K=K, out=out)
        ref = oracle.run_frames(st, sc, K, actions=acts[:, lo:lo + m].cpu().numpy(), seed=seed, env_id0=lo, reset_cursor=cur,
                                reset_rng="f32",
                                want=("reward", "score", "done", "game_over", "contacts", "contact_dmin"))
        ev = unpack_events(out.events, K)
        same = np.ones(m, bool)
        for k, key in (("reward", "reward"), ("score", "score"), ("done", "done"), ("game_over", "game_over")):
            same &= (ev[k][:, lo:lo + m].cpu().numpy().astype(np.int32) == ref[key].astype(np.int32)).all(axis=0)
        mismatch += int((~same).sum())
        launches += m
        got = get_state13(env)[lo:lo + m]
        ok = same & np.isfinite(st[:, 12])
        assert np.array_equal(get_scores(env)[lo:lo + m][same], sc[same])
        dpos = np.abs(got[:, pos_cols] - st[:, pos_cols]).max(axis=1)
        dvel = np.abs(got[:, vel_cols] - st[:, vel_cols]).max(axis=1)
        contact = ref["contacts"].sum(axis=0) > 0
        free, hit = ok & ~contact, ok & contact
        hits += int(hit.sum())
        worst["pos_free"] = max(worst["pos_free"], dpos[free].max())
        worst["vel_free"] = max(worst["vel_free"], dvel[free].max())
        free_pos.append(dpos[free]); free_vel.append(dvel[free])
        if hit.any():
            scale = 1.0 + 35.0 / np.maximum(ref["contact_dmin"].min(axis=0)[hit], 1e-3)
            worst["pos_hit_scaled"] = max(worst["pos_hit_scaled"], (dpos[hit] / scale).max())
            worst["vel_hit_scaled"] = max(worst["vel_hit_scaled"], (dvel[hit] / scale).max())
            hit_pos.append(dpos[hit] / scale); hit_vel.append(dvel[hit] / scale)
    q_pos, q_vel = np.quantile(np.concatenate(free_pos), 0.99), np.quantile(np.concatenate(free_vel), 0.99)
    print("fp32 K=8 launch errors on config 3:", worst, "99% free:", q_pos, q_vel, "event mismatches", mismatch, "of", launches,
          "contact env-launches", hits)
    assert q_pos <= 8e-3 and q_vel <= 8e-4, (q_pos, q_vel)
    assert worst["pos_free"] <= 5e-2 and worst["vel_free"] <= 2e-2, worst
    hp, hv = np.concatenate(hit_pos), np.concatenate(hit_vel)
    print("contact env-launches, scaled errors: 99 % / 99.9 % / max  pos", np.quantile(hp, 0.99), np.quantile(hp, 0.999), hp.max(),
          " vel", np.quantile(hv, 0.99), np.quantile(hv, 0.999), hv.max())
    assert np.quantile(hp, 0.99) <= 1e-2 and np.quantile(hv, 0.99) <= 2e-3, (np.quantile(hp, 0.99), np.quantile(hv, 0.99))
    assert np.quantile(hp, 0.999) <= 5e-2 and np.quantile(hv, 0.999) <= 2e-2, (np.quantile(hp, 0.999), np.quantile(hv, 0.999))
    assert worst["pos_hit_scaled"] <= 1.0 and worst["vel_hit_scaled"] <= 0.2, worst
    assert hits > 0.05 * launches and mismatch <= max(2, 2 * launches // 10000), (hits, mismatch)


def test_fp32_score_distribution_matches_oracle():
    """Free-running FP32 vs the FP64 oracle: per-frame rates of goals / dones / wins and the reward histogram
    must agree within sampling error (both are random-action rollouts of ~13M frames)."""
    n, T = 16384, 800
    env = make_env(n, dtype=torch.float32, seed=77)
    hist = torch.zeros(16, dtype=torch.int64, device=env.device)
    for _ in range(T // 8):
        out = env.make_outputs(8, ("reward",))
        env.step(None, K=8, out=out)
        hist += torch.bincount((out.reward.flatten().long() + 8), minlength=16)
    g = env.stats()
    st, sc = oracle.initial_state(n)
    ref = oracle.run_frames(st, sc, T, seed=78, nthreads=8, want=("reward",))
    o = dict(zip(oracle.STAT_NAMES, ref["stats"]))
    frames = n * T
    assert g["frames"] == frames == o["frames"]
    for k in ("robot_goals", "opponent_goals", "dones"):
        pg, po = g[k] / frames, o[k] / frames
        sigma = np.sqrt(po * (1 - po) / frames) * np.sqrt(2)
        # goal events are positively correlated within an env (cluster factor ~3): allow 8 sigma
        assert abs(pg - po) < 8 * sigma + 0.02 * po, (k, pg, po, sigma)
    assert abs(g["reward_sum"] / frames - o["reward_sum"] / frames) < 0.02
    ref_hist = np.bincount(ref["reward"].flatten() + 8, minlength=16) / frames
    gpu_hist = hist.cpu().numpy() / frames
    assert np.abs(ref_hist - gpu_hist).max() < 5e-3, (ref_hist, gpu_hist)
    wins_g = (g["robot_wins"] + g["opponent_wins"]) / frames
    wins_o = (o["robot_wins"] + o["opponent_wins"]) / frames
    assert abs(wins_g - wins_o) < 0.15 * wins_o + 1e-5


# ------------------------------------------------------------------------------ configuration switches vs the oracle
@pytest.mark.parametrize("dot_fma,friction", [(False, False), (True, True), (False, True)])
def test_fp64_switches_match_oracle(dot_fma, friction):
    """`dot_fma = 0` (a numpy build whose ddot tail is not FMA-contracted) and `friction = 1` (the reference's dead
    Puck.friction_on_puck, opt-in) are restated in the oracle too: the FP64 build must stay bit-exact under both."""
    n, T = 512, 400
    rng = np.random.default_rng(11)
    actions = rng.integers(0, 4, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 32, 2))
    st, sc = oracle.initial_state(n)
    ref = oracle.run_frames(st, sc, T, actions=actions, reset_table=table, dot_fma=int(dot_fma), friction=friction,
                            record_every=50, want=("reward", "score", "done", "game_over"))
    env = make_env(n, dot_fma=dot_fma, friction=friction)
    res = run_gpu(env, T, actions, table, 50)
    assert np.array_equal(bits(ref["states_rec"]), bits(res["states"]))
    assert np.array_equal(ref["reward"], res["reward"].astype(np.int32))
    assert np.array_equal(ref["score"], res["score"]) and np.array_equal(ref["game_over"], res["game_over"])
    base, _ = oracle.initial_state(n)
    sc0 = np.zeros((n, 2), np.int32)
    oracle.run_frames(base, sc0, T, actions=actions, reset_table=table, want=())
    assert not np.array_equal(bits(base), bits(st)), "the switch must actually change the trajectories"


def test_fp64_contact_statistics_match_oracle():
    n, T = 2048, 300
    st, sc = oracle.initial_state(n)
    ref = oracle.run_frames(st, sc, T, seed=5, want=())
    env = make_env(n, seed=5)
    env.step(None, K=T, out=env.make_outputs(T, ()))
    g, o = env.stats(), dict(zip(oracle.STAT_NAMES, ref["stats"]))
    for k in ("frames", "robot_goals", "opponent_goals", "dones", "reward_sum", "reward_sq_sum", "contacts_robot", "contacts_opponent"):
        assert g[k] == o[k], k
    assert g["contacts_robot"] > 0 and g["contacts_opponent"] > 0


def test_fp32_long_rollout_matches_reference_behaviour_statistics():
    """Free-running FP32 build, uniform-random robot vs computerAI, 65,536 envs x 6,000 frames (3.9e8 frames), against
    the statistics of the LIVE Python reference measured in the survey (BASELINE.md §5, 8 seeds x 200,000 frames;
    per-seed spread in brackets there).  The first 2,000 frames are discarded on our side: all envs start from the
    same constructor state, the reference's numbers are long-run averages."""
    n = 65536
    env = make_env(n, dtype=torch.float32, seed=2718)
    env.step(None, K=2000, out=env.make_outputs(2000, ()))
    env.stats(clear=True)
    hist = torch.zeros(16, dtype=torch.int64, device=env.device)
    for _ in range(4000 // 40):
        out = env.make_outputs(40, ("reward",))
        env.step(None, K=40, out=out)
        hist += torch.bincount((out.reward.flatten().long() + 8), minlength=16)
    s = env.stats()
    f = s["frames"]
    assert f == n * 4000
    per_k = lambda k: 1e3 * s[k] / f
    assert 2.6 < per_k("robot_goals") < 3.0, per_k("robot_goals")               # reference 2.79 (2.69-2.90)
    assert 4.9 < per_k("opponent_goals") < 5.5, per_k("opponent_goals")         # reference 5.14 (4.99-5.41)
    assert 2.6 < per_k("dones") < 3.2, per_k("dones")                           # reference 2.89
    assert 10.5 < per_k("contacts_robot") < 14.0, per_k("contacts_robot")       # reference 12.2
    assert 14.0 < per_k("contacts_opponent") < 17.5, per_k("contacts_opponent") # reference 15.7
    games = s["games"]
    assert games > 50000
    assert 0.05 < s["robot_wins"] / games < 0.11, s["robot_wins"] / games       # reference 67 / (67 + 769) = 8.0 %
    assert 1800 < s["game_len_sum"] / games < 2010, s["game_len_sum"] / games   # reference mean game length 1,905 frames
    h = (hist.cpu().numpy() / f)
    ref = {-4: 0.0099, -2: 0.5048, 0: 0.4788, 2: 0.0038, 4: 0.0027}             # reference reward histogram
    for r, p in ref.items():
        assert abs(h[r + 8] - p) < 0.01 + 0.1 * p, (r, h[r + 8], p)


def test_fp64_max_k_single_launch_vs_oracle():
    """K = 4096 (the ABI's maximum) in ONE launch, Philox actions and resets: bit-exact vs the oracle."""
    n, K, seed = 384, 4096, 77
    st, sc = oracle.initial_state(n)
    ref = oracle.run_frames(st, sc, K, seed=seed, nthreads=4, want=("reward", "score", "game_over"))
    env = make_env(n, seed=seed)
    out = env.make_outputs(K, ("reward", "score", "game_over", "actions"))
    env.step(None, K=K, out=out)
    assert np.array_equal(bits(st), bits(get_state13(env))) and np.array_equal(sc, get_scores(env))
    assert np.array_equal(ref["reward"], out.reward.cpu().numpy().astype(np.int32))
    assert np.array_equal(ref["score"], out.score.cpu().numpy()) and np.array_equal(ref["game_over"], out.game_over.cpu().numpy())
    want_actions = np.array([[oracle.rng_action(seed, e, f) for e in range(4)] for f in range(0, K, 97)], np.int8)
    assert np.array_equal(want_actions, out.actions.cpu().numpy()[::97, :4])      # the Philox action stream itself
    assert (ref["game_over"] != 0).sum() > 100


def test_fp64_soak_65536_envs_2000_frames_vs_oracle():
    """1.3e8 env-steps, Philox actions + resets, FP64: final state, scores, RNG counters and every tally bit-exact."""
    n, T, seed = 65536, 2000, 4242
    st, sc = oracle.initial_state(n)
    ref = oracle.run_frames(st, sc, T, seed=seed, nthreads=16, want=())
    env = make_env(n, seed=seed)
    for _ in range(T // 250):
        env.step(None, K=250, out=env.make_outputs(250, ()))
    assert np.array_equal(bits(st), bits(get_state13(env)))
    assert np.array_equal(sc, get_scores(env))
    assert np.array_equal(ref["reset_cursor"], env.reset_count.cpu().numpy())
    g, o = env.stats(), dict(zip(oracle.STAT_NAMES, ref["stats"]))
    for k in ("frames", "robot_goals", "opponent_goals", "robot_wins", "opponent_wins", "dones", "reward_sum",
              "reward_sq_sum", "contacts_robot", "contacts_opponent"):
        assert g[k] == o[k], k
    assert g["games"] == g["robot_wins"] + g["opponent_wins"] > 10000


def test_fp64_scrambled_states_at_scale_vs_oracle():
    """65,536 random non-integer initial states (mallets overlapping the puck, coincident centres, speeds beyond the
    limits, scores one goal from game over, out-of-range actions): every rare branch many times; bit-exact for 300 frames."""
    from oracle.gen_golden import scrambled_states
    n, T = 65536, 300
    rng = np.random.default_rng(909)
    init = scrambled_states(rng, n)
    init_scores = rng.integers(0, 10, (n, 2)).astype(np.int32)
    init_scores[: n // 4] = rng.integers(8, 10, (n // 4, 2))
    actions = rng.integers(-2, 7, (T, n)).astype(np.int8)
    table = rng.uniform(-12, 12, (n, 24, 2))                  # beyond the speed limits on purpose (speed_dirty path)
    st, sc = init.copy(), init_scores.copy()
    ref = oracle.run_frames(st, sc, T, actions=actions, reset_table=table, record_every=100, nthreads=16,
                            want=("reward", "score", "done", "game_over"))
    assert not ref["reset_overflow"]
    env = make_env(n)
    put_state(env, init, init_scores)
    res = run_gpu(env, T, actions, table, 100)
    assert np.array_equal(bits(ref["states_rec"]), bits(res["states"]))
    assert np.array_equal(ref["scores_rec"], res["scores"])
    assert np.array_equal(ref["reward"], res["reward"].astype(np.int32))
    for k in ("score", "done", "game_over"):
        assert np.array_equal(ref[k], res[k]), k
    o = dict(zip(oracle.STAT_NAMES, ref["stats"]))
    g = env.stats()
    assert g["contacts_robot"] == o["contacts_robot"] > 20000 and g["contacts_opponent"] == o["contacts_opponent"] > 20000
    assert (ref["game_over"] != 0).sum() > 5000


# ------------------------------------------------------------------------------ the LIVE reference, on this box
def test_cuda_fp64_vs_the_live_reference_on_this_box():
    """The reference itself (oracle/_ref: byte-compiled from /root/reference by oracle/fetch_ref.py; the live sources in
    the development container) is run HERE, next to the GPU, on fresh seeded inputs -- not a committed fixture -- and the
    FP64 CUDA path (unpacked and packed), the C oracle and the reference must agree bit for bit."""
    from oracle import ref_loop, refshim
    if not refshim.reference_available():
        pytest.skip("oracle/_ref has not been built (run oracle/fetch_ref.py where /root/reference exists)")
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    n, T = 12, 320
    seed = int.from_bytes(os.urandom(4), "little")                                # fresh inputs every run
    print("seed of this run:", seed)
    rng = np.random.default_rng(seed)
    actions = rng.integers(-1, 5, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 48, 2))
    ref = ref_loop.run_reference(actions, table)
    st, sc = oracle.initial_state(n)
    orc = oracle.run_frames(st, sc, T, actions=actions, reset_table=table, record_every=1, want=("reward", "score", "done", "game_over"))
    assert np.array_equal(bits(ref["states"]), bits(orc["states_rec"])) and np.array_equal(ref["reward"], orc["reward"])
    env = make_env(n)
    res = run_gpu(env, T, actions, table, 8)
    assert np.array_equal(bits(ref["states"][7::8]), bits(res["states"]))
    assert np.array_equal(ref["reward"].astype(np.float64), res["reward"].astype(np.float64))
    assert np.array_equal(ref["score"], res["score"]) and np.array_equal(ref["done"], res["done"])
    assert np.array_equal((ref["robot_win"] + 2 * ref["opponent_win"]).astype(np.uint8), res["game_over"])
    for k in ("obs_state", "obs_new_state", "obs_next"):
        assert np.array_equal(bits(ref[k]), bits(res[k]))
    pk = make_env(n)
    tab = torch.as_tensor(table, dtype=torch.float64).cuda().contiguous()
    acts = torch.as_tensor(actions).cuda()
    for t0 in range(0, T, 8):
        out = pk.step(pack_actions(acts[t0:t0 + 8]), K=8, out=pk.make_outputs(8, ("events", "obs_next")), reset_table=tab)
        ev = unpack_events(out.events, 8)
        assert np.array_equal(ev["reward"].cpu().numpy().astype(np.int32), ref["reward"][t0:t0 + 8])
        assert np.array_equal(ev["done"].cpu().numpy().astype(np.uint8), ref["done"][t0:t0 + 8])
        assert np.array_equal(get_state13(pk).view(np.uint64), ref["states"][t0 + 7].view(np.uint64))
    assert pk.stats()["done_ambiguous"] == 0


def test_cuda_fp64_vs_the_live_reference_from_scrambled_states():
    """Same as above from random NON-integer states (contacts, wall clamps, both speed-clamp signs, goals, scores one goal
    from game over, out-of-range actions) drawn fresh every run: reference == oracle == CUDA FP64 (unpacked, K = 8 packed)."""
    from oracle import ref_loop, refshim
    from oracle.gen_golden import scrambled_states
    if not refshim.reference_available():
        pytest.skip("oracle/_ref has not been built (run oracle/fetch_ref.py where /root/reference exists)")
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    n, T = 24, 160
    seed = int.from_bytes(os.urandom(4), "little")
    print("seed of this run:", seed)
    rng = np.random.default_rng(seed)
    actions = rng.integers(-1, 6, (T, n)).astype(np.int8)
    table = rng.uniform(-10, 10, (n, 40, 2))
    init = scrambled_states(rng, n)
    init_scores = rng.integers(0, 10, (n, 2)).astype(np.int32)
    init_scores[: n // 3] = rng.integers(8, 10, (n // 3, 2))
    ref = ref_loop.run_reference(actions, table, init_state=init, init_scores=init_scores)
    assert ref["reset_cursor"].max() < 40
    env = make_env(n)
    put_state(env, init, init_scores)
    res = run_gpu(env, T, actions, table, 1)
    assert np.array_equal(bits(ref["states"]), bits(res["states"])) and np.array_equal(ref["scores"], res["scores"])
    assert np.array_equal(ref["reward"].astype(np.float64), res["reward"].astype(np.float64))
    assert np.array_equal(ref["score"], res["score"]) and np.array_equal(ref["done"], res["done"])
    assert np.array_equal((ref["robot_win"] + 2 * ref["opponent_win"]).astype(np.uint8), res["game_over"])
    for k in ("obs_state", "obs_new_state", "obs_next"):
        assert np.array_equal(bits(ref[k]), bits(res[k])), k
    assert env.stats()["done_ambiguous"] == 0
    pk = make_env(n)
    put_state(pk, init, init_scores)
    tab = torch.as_tensor(table, dtype=torch.float64).cuda().contiguous()
    acts = torch.as_tensor(actions).cuda()
    for t0 in range(0, T, 8):
        out = pk.step(pack_actions(acts[t0:t0 + 8]), K=8, out=pk.make_outputs(8, ("events", "obs_next")), reset_table=tab)
        ev = unpack_events(out.events, 8)
        assert np.array_equal(ev["reward"].cpu().numpy().astype(np.int32), ref["reward"][t0:t0 + 8])
        assert np.array_equal(ev["score"].cpu().numpy().astype(np.int32), ref["score"][t0:t0 + 8].astype(np.int32))
        assert np.array_equal(get_state13(pk).view(np.uint64), ref["states"][t0 + 7].view(np.uint64))
