"""GPU tests of ``ReplayRing``, the device-resident replacement of ``MemoryBuffer`` (lib/buffer.py:11-66): the fused
append from inside the step kernel, the ring-owned device counter (purge, CUDA-graph replay, several rings on one
env), the one-kernel sampler ``ah_ring_sample``, and the reference's own tests/test_buffer.py pins re-expressed."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


def make_env(n, dtype=torch.float32, **kw):
    from air_hockey_training_environment_b200 import BatchedAirHockey
    return BatchedAirHockey(n, dtype=dtype, device="cuda", **kw)


FIELDS = ("obs_state", "obs_new_state", "reward", "done")


def step_with_ring(env, ring, K, gen=None):
    acts = torch.randint(0, 4, (K, env.n), dtype=torch.int8, device=env.device, generator=gen)
    out = env.step(acts, K=K, out=env.make_outputs(K, FIELDS), ring=ring)
    return [(out.obs_state[k].clone(), acts[k].clone(), out.reward[k].clone(), out.done[k].clone(), out.obs_new_state[k].clone())
            for k in range(K)]


def assert_newest_first(ring, hist):
    got = ring.retreive()                                # newest first, like MemoryBuffer.retreive()
    assert got["state"].shape[0] == min(len(hist), ring.capacity)
    for j in range(got["state"].shape[0]):
        s, a, r, d, s2 = hist[-1 - j]
        assert torch.equal(got["state"][j], s) and torch.equal(got["new_state"][j], s2)
        assert torch.equal(got["action"][j], a) and torch.equal(got["reward"][j], r) and torch.equal(got["done"][j], d)


# ---------------------------------------------------------------- fused append
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_ring_append_matches_step_outputs_and_wraps(dtype):
    from air_hockey_training_environment_b200 import ReplayRing
    n, cap = 256, 6
    env = make_env(n, dtype, seed=4)
    ring = ReplayRing(cap, n, env.dtype, env.device)
    assert len(ring) == 0
    hist = []
    for _ in range(4):                                   # 4 launches x 2 frames = 8 appends into 6 slots
        hist += step_with_ring(env, ring, 2)
    assert len(ring) == cap and ring.appended == 8
    assert_newest_first(ring, hist)


def test_ring_purge_then_append_reads_back_what_was_written():
    """Round-1 bug: purge() reset only a host mirror while the kernel kept writing at the old device offset."""
    from air_hockey_training_environment_b200 import ReplayRing
    n, cap = 128, 5
    env = make_env(n, seed=8)
    ring = ReplayRing(cap, n, env.dtype, env.device)
    for _ in range(3):
        step_with_ring(env, ring, 3)                     # 9 appends: the write slot is mid-ring
    ring.purge()
    assert len(ring) == 0 and ring.retreive()["state"].shape[0] == 0
    hist = step_with_ring(env, ring, 2)
    assert len(ring) == 2
    assert_newest_first(ring, hist)
    hist += step_with_ring(env, ring, 4)                 # wraps again after the purge
    assert len(ring) == cap
    assert_newest_first(ring, hist)


def test_second_ring_on_an_env_that_already_appended_and_checkpoint():
    from air_hockey_training_environment_b200 import ReplayRing
    n = 64
    env = make_env(n, seed=2)
    first = ReplayRing(4, n, env.dtype, env.device)
    step_with_ring(env, first, 3)
    second = ReplayRing(7, n, env.dtype, env.device)      # a fresh ring starts at its own slot 0
    hist = step_with_ring(env, second, 5)
    assert first.appended == 3 and second.appended == 5
    assert_newest_first(second, hist)
    sd = second.state_dict()
    hist2 = step_with_ring(env, second, 4)
    assert_newest_first(second, hist + hist2)
    second.load_state_dict(sd)                            # counter and contents travel together
    assert second.appended == 5
    assert_newest_first(second, hist)


def test_ring_counter_advances_under_cuda_graph_replay():
    from air_hockey_training_environment_b200 import ReplayRing
    n, K, cap = 4096, 2, 5
    env = make_env(n, seed=21)
    ring = ReplayRing(cap, n, env.dtype, env.device)
    acts = torch.randint(0, 4, (K, n), dtype=torch.int8, device=env.device)
    out = env.make_outputs(K, FIELDS)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        env.step(acts, K=K, out=out, ring=ring)           # warm-up outside the capture
    torch.cuda.current_stream().wait_stream(side)
    env.init_state()
    ring.purge()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        env.step(acts, K=K, out=out, ring=ring)
    hist = []
    for rep in range(4):                                  # 8 appends through replays only
        g.replay()
        hist += [(out.obs_state[k].clone(), acts[k].clone(), out.reward[k].clone(), out.done[k].clone(), out.obs_new_state[k].clone())
                 for k in range(K)]
    assert ring.appended == 8 and len(ring) == cap
    assert_newest_first(ring, hist)
    ring.purge()
    g.replay()                                            # purge is seen by the captured launch
    assert ring.appended == K
    assert_newest_first(ring, hist[:0] + [(out.obs_state[k].clone(), acts[k].clone(), out.reward[k].clone(), out.done[k].clone(),
                                           out.obs_new_state[k].clone()) for k in range(K)])


# ---------------------------------------------------------------- the sampler (lib/buffer.py:47-50)
def _ring_round(v, key):
    M = 0xFFFFFFFF
    v = ((v + key) * 0x9E3779B1) & M
    v ^= v >> 15
    v = (v * 0x85EBCA77) & M
    v ^= v >> 13
    v = (v * 0xC2B2AE3D) & M
    v ^= v >> 16
    return v


def host_sample_indices(seed, draw, population, batch):
    """Restatement of ah_ring_sample's index map (include/ah_b200.h): 6-round Feistel over the next even power of two,
    cycle-walked; round keys = Philox4x32-10(counter = (draw, 'RING', 0/1), key = seed)."""
    bits = 2
    while (1 << bits) < population:
        bits += 2
    half = bits // 2
    mask = (1 << half) - 1
    ks = []
    for blk in (0, 1):
        ks += [int(x) for x in oracle.philox4x32_10([draw & 0xFFFFFFFF, draw >> 32, 0x52494E47, blk], [seed & 0xFFFFFFFF, seed >> 32])]
    out = []
    for s in range(batch):
        f = s % population
        while True:
            L, R = (f >> half) & mask, f & mask
            for r in range(6):
                L, R = R, L ^ (_ring_round(R, ks[r]) & mask)
            f = (L << half) | R
            if f < population:
                break
        out.append(f)
    return np.array(out, np.int64)


def test_ring_sample_is_without_replacement_and_matches_host_restatement():
    from air_hockey_training_environment_b200 import ReplayRing
    n, cap = 300, 7
    env = make_env(n, seed=77)
    ring = ReplayRing(cap, n, env.dtype, env.device)
    hist = []
    for _ in range(3):
        hist += step_with_ring(env, ring, 3)             # 9 appends into 7 slots
    pop = cap * n
    full = ring.sample(pop, env, draw=5, with_index=True)
    idx = full["index"].cpu().numpy()
    assert np.array_equal(np.sort(idx), np.arange(pop))   # a whole-population sample is a permutation
    assert np.array_equal(idx[:500], host_sample_indices(77, 5, pop, 500))
    small = ring.sample(64, env, draw=5, with_index=True)
    assert torch.equal(small["index"], full["index"][:64])    # the sample is a prefix of the permutation
    other = ring.sample(64, env, draw=6, with_index=True)
    assert not torch.equal(other["index"], small["index"])
    # every sampled tuple is the tuple at (age, env) in newest-first order
    got = ring.retreive()
    age, e = full["index"] // n, full["index"] % n
    for k in ("state", "new_state", "action", "reward", "done"):
        assert torch.equal(full[k], got[k][age, e]), k
    with pytest.raises(ValueError):
        ring.sample(pop + 1, env)                         # random.sample raises ValueError too
    # partially filled ring: only held tuples are sampled
    ring.purge()
    step_with_ring(env, ring, 2)
    part = ring.sample(2 * n, env, with_index=True)
    assert np.array_equal(np.sort(part["index"].cpu().numpy()), np.arange(2 * n))


def test_ring_sample_is_uniform():
    from air_hockey_training_environment_b200 import ReplayRing
    n, cap, B = 1000, 10, 2000
    env = make_env(n, seed=3)
    ring = ReplayRing(cap, n, env.dtype, env.device)
    step_with_ring(env, ring, cap)
    counts = torch.zeros(cap * n, dtype=torch.int64, device=env.device)
    draws = 200
    for d in range(draws):
        counts += torch.bincount(ring.sample(B, env, draw=d, check=False, with_index=True)["index"], minlength=cap * n)
    # each of the 10,000 tuples is expected B * draws / pop = 40 times; binomial sd ~ 5.7
    c = counts.double()
    assert abs(c.mean().item() - 40.0) < 1e-9 and 4.5 < c.std().item() < 7.0, (c.mean().item(), c.std().item())
    assert c.max().item() < 80 and c.min().item() > 10
    by_age = c.view(cap, n).sum(1)
    assert (by_age - by_age.mean()).abs().max().item() < 0.03 * by_age.mean().item()


# ---------------------------------------------------------------- reference tests/test_buffer.py, re-expressed
def test_memory_buffer_default_fill():
    # tests/test_buffer.py:11-16: MemoryBuffer(3, default) holds `capacity` default entries
    from air_hockey_training_environment_b200 import ReplayRing
    ring = ReplayRing(3, 4, torch.float32, "cuda", default={"state": 0, "reward": 0})
    assert len(ring) == 3
    got = ring.retreive()
    assert got["state"].shape == (3, 4, 8) and float(got["state"].abs().sum()) == 0 and float(got["reward"].abs().sum()) == 0


def test_memory_buffer_purge_and_sample_size():
    # tests/test_buffer.py:18-43: a full buffer samples `k` distinct held items; purge() -> len 0
    from air_hockey_training_environment_b200 import ReplayRing
    env = make_env(1, seed=0)
    cap = 10
    ring = ReplayRing(cap, 1, env.dtype, env.device)
    rng = np.random.default_rng(0)
    data = rng.integers(0, 11, (cap, 2))
    for item in data:
        s = torch.zeros(1, 8, device="cuda")
        s[0, :2] = torch.as_tensor(item, dtype=torch.float32)
        ring.append(s, torch.zeros(1, dtype=torch.int8, device="cuda"), torch.zeros(1, device="cuda"),
                    torch.zeros(1, dtype=torch.uint8, device="cuda"), s)
    assert len(ring) == cap
    batch = ring.sample(4, env, with_index=True)
    assert batch["state"].shape == (4, 8) and len(set(batch["index"].tolist())) == 4
    held = {tuple(int(v) for v in row) for row in data}
    assert all(tuple(int(v) for v in row[:2]) in held for row in batch["state"].cpu().numpy())
    ring.purge()
    assert len(ring) == 0


def test_memory_buffer_retrieve_average():
    # tests/test_buffer.py:45-54: column average of [(3,4),(5,6),(4,6),(2,1)] is ((3, 4),)  (int(sum / len))
    from air_hockey_training_environment_b200 import ReplayRing
    ring = ReplayRing(4, 1, torch.float32, "cuda")
    for item in [(3, 4), (5, 6), (4, 6), (2, 1)]:
        s = torch.zeros(1, 8, device="cuda")
        s[0, :2] = torch.tensor(item, dtype=torch.float32)
        ring.append(s, torch.zeros(1, dtype=torch.int8, device="cuda"), torch.zeros(1, device="cuda"),
                    torch.zeros(1, dtype=torch.uint8, device="cuda"), s)
    avg = ring.retreive(average=True)
    assert avg["state"].shape == (1, 1, 8) and avg["state"][0, 0, :2].tolist() == [3, 4]
    newest = ring.retreive()["state"][:, 0, :2].tolist()
    assert newest == [[2, 1], [4, 6], [5, 6], [3, 4]]        # appendleft: newest first (lib/buffer.py:27)


# ---------------------------------------------------------------- the packed rollout mode with the fused append
@pytest.mark.parametrize("dtype,K,cap", [(torch.float32, 8, 16), (torch.float64, 6, 5), (torch.float32, 3, 7)])
def test_packed_mode_ring_append_equals_the_general_kernel(dtype, K, cap):
    """``step(packed actions, out.events, ring=...)`` (ah::step_packed_kernel<R, RING>) appends exactly what the general
    kernel appends -- state, action, reward, done, new_state of every frame, slot by slot, across launches that wrap the
    ring -- and leaves the same env state, event bytes and ring counter."""
    from air_hockey_training_environment_b200 import ReplayRing
    from air_hockey_training_environment_b200.env import pack_actions, unpack_events
    n = 1000                                             # not a multiple of the block size
    a, b = make_env(n, dtype, seed=12), make_env(n, dtype, seed=12)
    for env in (a, b):
        env.step(None, K=700, out=env.make_outputs(700, ()))         # contacts, goals and resets are happening
    ra, rb = ReplayRing(cap, n, a.dtype, a.device), ReplayRing(cap, n, b.dtype, b.device)
    g = torch.Generator(device="cuda").manual_seed(5)
    for launch in range(4):
        acts = torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda", generator=g)
        oa = a.step(pack_actions(acts), K=K, out=a.make_outputs(K, ("events", "obs_next")), ring=ra)
        ob = b.step(acts, K=K, out=b.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next")), ring=rb)
        ev = unpack_events(oa.events, K)
        assert torch.equal(ev["reward"].to(b.dtype), ob.reward) and torch.equal(ev["done"], ob.done.bool())
        assert torch.equal(oa.obs_next, ob.obs_next)
        assert ra.appended == rb.appended == (launch + 1) * K
        for x, y in ((ra.state, rb.state), (ra.new_state, rb.new_state), (ra.action, rb.action), (ra.reward, rb.reward),
                     (ra.done, rb.done)):
            assert torch.equal(x, y)
    for x, y in zip((a.puck, a.robot, a.opponent, a.prev_dist, a.meta), (b.puck, b.robot, b.opponent, b.prev_dist, b.meta)):
        assert torch.equal(x, y)
    assert a.stats() == b.stats() and a.counters() == b.counters()
    got, want = ra.retreive(), rb.retreive()
    assert all(torch.equal(got[k], want[k]) for k in want)


def test_packed_mode_ring_append_under_graph_replay_and_argument_errors():
    from air_hockey_training_environment_b200 import AhError, ReplayRing
    from air_hockey_training_environment_b200.env import pack_actions
    n, K, cap = 512, 4, 6
    env, ref = make_env(n, seed=3), make_env(n, seed=3)
    ring, rref = ReplayRing(cap, n, env.dtype, env.device), ReplayRing(cap, n, ref.dtype, ref.device)
    acts = pack_actions(torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda"))
    out = env.make_outputs(K, ("events", "obs_next"))
    plan = env.plan_step(acts, K=K, out=out, ring=ring)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        env.launch(plan)                                  # warm-up launch outside the capture
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        env.launch(plan)
    for _ in range(3):
        g.replay()
    for _ in range(4):
        ref.step(acts, K=K, out=ref.make_outputs(K, ("events", "obs_next")), ring=rref)
    assert ring.appended == rref.appended == 16           # the device counter advanced inside the replays
    assert torch.equal(ring.state, rref.state) and torch.equal(ring.reward, rref.reward)
    got, want = ring.retreive(), rref.retreive()
    assert all(torch.equal(got[k], want[k]) for k in want)
    with pytest.raises(AhError):                          # the append needs the whole env range in one launch
        env.step(acts, K=K, out=out, ring=ring, env_range=(0, 256))


# ---------------------------------------------------------------- the compact ring layout (AH_RING_COMPACT)
@pytest.mark.parametrize("K,cap", [(8, 16), (3, 7), (5, 4)])
def test_compact_ring_holds_the_same_tuples_as_the_rows_layout(K, cap):
    """``ReplayRing(compact=True)`` (int16 locations + float32 velocities + int8 reward: 51 instead of 70 bytes per tuple,
    written by ah::step_packed_kernel<float, IO_RING_COMPACT>) against the rows layout on the same games: retreive() and
    sample() return identical tensors, launch after launch, across wrap-arounds; env state, events and statistics equal."""
    from air_hockey_training_environment_b200 import ReplayRing
    from air_hockey_training_environment_b200.env import pack_actions
    n = 1000
    a, b = make_env(n, torch.float32, seed=21), make_env(n, torch.float32, seed=21)
    for env in (a, b):
        env.step(None, K=700, out=env.make_outputs(700, ()))
    ra = ReplayRing(cap, n, a.dtype, a.device, compact=True)
    rb = ReplayRing(cap, n, b.dtype, b.device)
    assert ra.state.shape == (cap, n, 4) and ra.state_pos.dtype == torch.int16 and ra.reward.dtype == torch.int8
    g = torch.Generator(device="cuda").manual_seed(6)
    for launch in range(5):
        acts = pack_actions(torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda", generator=g))
        oa = a.step(acts, K=K, out=a.make_outputs(K, ("events", "obs_next")), ring=ra)
        ob = b.step(acts, K=K, out=b.make_outputs(K, ("events", "obs_next")), ring=rb)
        assert torch.equal(oa.events, ob.events) and torch.equal(oa.obs_next, ob.obs_next)
        assert ra.appended == rb.appended == (launch + 1) * K
        got, want = ra.retreive(), rb.retreive()
        for k in want:
            assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), (launch, k)
        B = min(4096, len(rb) * n)
        sa, sb = ra.sample(B, a, draw=launch, with_index=True), rb.sample(B, b, draw=launch, with_index=True)
        for k in sb:
            assert sa[k].dtype == sb[k].dtype and torch.equal(sa[k], sb[k]), (launch, k)
    for x, y in zip((a.puck, a.robot, a.opponent, a.prev_dist, a.meta), (b.puck, b.robot, b.opponent, b.prev_dist, b.meta)):
        assert torch.equal(x, y)
    assert a.stats() == b.stats()


def test_compact_ring_host_verbs_checkpoint_and_argument_errors():
    """fill / append / purge / state_dict on the compact layout behave like the rows layout; the layout is refused where
    no kernel writes it (the general kernel, the FP64 build, continuous actions)."""
    from air_hockey_training_environment_b200 import AhError, ReplayRing
    from air_hockey_training_environment_b200.env import pack_actions
    n, cap = 64, 5
    env = make_env(n, torch.float32, seed=2)
    rc, rr = ReplayRing(cap, n, env.dtype, env.device, compact=True), ReplayRing(cap, n, env.dtype, env.device)
    default = {"state": 3, "action": 1, "reward": -2, "done": 0, "new_state": 7}
    for r in (rc, rr):
        r.fill(default)
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(7):                                   # wraps
        st = torch.floor(torch.rand(n, 8, device="cuda", generator=g) * 400)
        st[:, 4:] = torch.rand(n, 4, device="cuda", generator=g) * 20 - 10
        ns = st + 1
        act = torch.randint(0, 4, (n,), dtype=torch.int8, device="cuda", generator=g)
        rew = (torch.randint(0, 6, (n,), device="cuda", generator=g) * 2 - 4).to(torch.float32)
        dn = torch.randint(0, 2, (n,), dtype=torch.uint8, device="cuda", generator=g)
        for r in (rc, rr):
            r.append(st, act, rew, dn, ns)
        got, want = rc.retreive(), rr.retreive()
        assert all(torch.equal(got[k], want[k]) for k in want)
    avg_c, avg_r = rc.retreive(average=True), rr.retreive(average=True)
    assert all(torch.equal(avg_c[k], avg_r[k]) for k in avg_r)
    sd = rc.state_dict()
    rc.purge()
    assert len(rc) == 0
    rc.load_state_dict(sd)
    got, want = rc.retreive(), rr.retreive()
    assert all(torch.equal(got[k], want[k]) for k in want)
    acts = torch.randint(0, 4, (4, n), dtype=torch.int8, device="cuda")
    with pytest.raises(ValueError):                      # the general kernel does not write the compact layout
        env.step(acts, K=4, out=env.make_outputs(4, ("reward", "done")), ring=rc)
    with pytest.raises(ValueError):
        ReplayRing(cap, n, torch.float64, env.device, compact=True)
    with pytest.raises(ValueError):
        ReplayRing(cap, n, torch.float32, env.device, continuous=True, compact=True)
    env.step(pack_actions(acts), K=4, out=env.make_outputs(4, ("events", "obs_next")), ring=rc)      # the packed mode does
    assert rc.appended == int(sd["counter"][0]) + 4


@pytest.mark.parametrize("compact,parts", [(True, 2), (False, 3), (True, 4)])
def test_ring_append_from_interleaved_slices_equals_whole_launches(compact, parts):
    """A step issued as P concurrent env slices (InterleavedStepper) appends to the ring too: every slice reads and
    advances its own copy of the ring's counter (ah_step_io.ring_counter_first / _count), so no launch waits for or races
    with another.  Same ring contents, counters, env state and statistics as whole-n launches; whole-n and sliced steps
    can alternate."""
    from air_hockey_training_environment_b200 import InterleavedStepper, ReplayRing
    from air_hockey_training_environment_b200.env import pack_actions
    n, K, cap = 5000, 8, 20                               # ragged slices
    a, b = make_env(n, torch.float32, seed=31), make_env(n, torch.float32, seed=31)
    for env in (a, b):
        env.step(None, K=600, out=env.make_outputs(600, ()))
    ra = ReplayRing(cap, n, a.dtype, a.device, compact=compact)
    rb = ReplayRing(cap, n, b.dtype, b.device, compact=compact)
    st = InterleavedStepper(a, parts=parts, align=128)
    oa, ob = a.make_outputs(K, ("events", "obs_next")), b.make_outputs(K, ("events", "obs_next"))
    g = torch.Generator(device="cuda").manual_seed(8)
    for launch in range(6):
        acts = pack_actions(torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda", generator=g))
        if launch == 3:                                   # a whole-n launch in between keeps every counter copy in step
            a.step(acts, K=K, out=oa, ring=ra)
        else:
            st.step(acts, K, oa, ring=ra)
            st.join()
        b.step(acts, K=K, out=ob, ring=rb)
        torch.cuda.synchronize()
        assert ra.counter.tolist() == rb.counter.tolist() == [(launch + 1) * K] * 4
        assert torch.equal(oa.events, ob.events) and torch.equal(oa.obs_next, ob.obs_next)
        got, want = ra.retreive(), rb.retreive()
        assert all(torch.equal(got[k], want[k]) for k in want), launch
    for x, y in zip((a.puck, a.robot, a.opponent, a.prev_dist, a.meta), (b.puck, b.robot, b.opponent, b.prev_dist, b.meta)):
        assert torch.equal(x, y)
    assert a.stats() == b.stats() and a.counters() == b.counters()
