"""CPU tests of the replay ring's host verbs (air_hockey_training_environment_b200.env.ReplayRing: fill / append / retreive /
purge / state_dict -- plain torch, no kernel): the compact tuple layout (AH_RING_COMPACT: int16 locations + float32 velocities
+ int8 reward) must behave exactly like the rows layout, and both like ``MemoryBuffer`` (lib/buffer.py:11-66: appendleft +
retreive = newest first, purge -> len 0, column average truncated).  The kernels' side of both layouts is tested on the GPU
(tests/test_gpu_ring.py)."""
import pytest
import torch


def rings(cap, n):
    from air_hockey_training_environment_b200.env import ReplayRing
    return (ReplayRing(cap, n, torch.float32, "cpu", compact=True), ReplayRing(cap, n, torch.float32, "cpu"))


def random_tuples(n, g):
    st = torch.floor(torch.rand(n, 8, generator=g) * 800)
    st[:, 4:] = torch.rand(n, 4, generator=g) * 20 - 10
    ns = st.clone(); ns[:, :4] += 3; ns[:, 4:] *= 0.5
    act = torch.randint(0, 4, (n,), dtype=torch.int8, generator=g)
    rew = (torch.randint(0, 6, (n,), generator=g) * 2 - 4).to(torch.float32)
    dn = torch.randint(0, 2, (n,), dtype=torch.uint8, generator=g)
    return st, act, rew, dn, ns


def test_compact_layout_equals_rows_layout_through_the_host_verbs():
    n, cap = 37, 5
    rc, rr = rings(cap, n)
    assert rc.state.shape == (cap, n, 4) and rc.state_pos.dtype == torch.int16 and rc.reward.dtype == torch.int8
    assert rr.state.shape == (cap, n, 8) and rc.counter.numel() == rr.counter.numel() == 4
    g = torch.Generator().manual_seed(3)
    newest = None
    for k in range(8):                                   # wraps the ring
        tup = random_tuples(n, g)
        for r in (rc, rr):
            r.append(*tup)
        newest = tup
        assert len(rc) == len(rr) == min(k + 1, cap) and rc.appended == k + 1
        got, want = rc.retreive(), rr.retreive()
        for key in want:
            assert got[key].dtype == want[key].dtype and torch.equal(got[key], want[key]), (k, key)
        assert torch.equal(got["state"][0], newest[0]) and torch.equal(got["new_state"][0], newest[4])      # newest first
        assert torch.equal(got["reward"][0], newest[2]) and torch.equal(got["action"][0], newest[1])
    a, b = rc.retreive(average=True), rr.retreive(average=True)
    assert all(torch.equal(a[k], b[k]) for k in b)
    sd = rc.state_dict()
    rc.purge()
    assert len(rc) == 0 and rc.counter.tolist() == [0, 0, 0, 0]
    rc.load_state_dict(sd)
    got, want = rc.retreive(), rr.retreive()
    assert all(torch.equal(got[k], want[k]) for k in want)


def test_default_fill_and_purge_like_memory_buffer():
    # reference tests/test_buffer.py:11-54: a buffer created with a default is full of it; purge -> empty (or default again)
    n, cap = 4, 6
    default = {"state": 3, "action": 1, "reward": -2, "done": 0, "new_state": 7}
    for r in rings(cap, n):
        r.fill(default)
        assert len(r) == cap
        got = r.retreive()
        assert (got["state"] == 3).all() and (got["new_state"] == 7).all() and (got["reward"] == -2).all()
        assert got["state"].shape == (cap, n, 8) and got["reward"].dtype == torch.float32
        avg = r.retreive(average=True)
        assert avg["state"].shape == (1, n, 8) and (avg["state"] == 3).all() and (avg["action"] == 1).all()
        r.purge()
        assert len(r) == 0
        r.purge(default)
        assert len(r) == cap and (r.retreive()["new_state"] == 7).all()


def test_compact_layout_argument_errors():
    from air_hockey_training_environment_b200.env import ReplayRing
    with pytest.raises(ValueError):
        ReplayRing(4, 8, torch.float64, "cpu", compact=True)
    with pytest.raises(ValueError):
        ReplayRing(4, 8, torch.float32, "cpu", continuous=True, compact=True)
