"""GPU tests of the device policies (csrc/ah_policy.cuh, policy.py): the generic dense network evaluated in the kernels
against the torch module, every exploration strategy of the reference (lib/exploration/strategies.py,
lib/exploration/random.py) against a numpy restatement fed the SAME Philox uniforms, and the closed loop in one launch
(AH_ACT_POLICY_MLP) against the frame-by-frame path."""
import math

import numpy as np
import pytest
import torch
from torch import nn

from oracle import oracle

pytestmark = pytest.mark.gpu


def make_env(n, dtype=torch.float32, **kw):
    from air_hockey_training_environment_b200 import BatchedAirHockey
    return BatchedAirHockey(n, dtype=dtype, device="cuda", **kw)


def uniforms(seed, env_ids, frame, stream):
    """The kernels' per-frame uniform (csrc/ah_rng.cuh: sample_block / sample_block2 + sample_uniform)."""
    out = np.zeros(len(env_ids), np.float32)
    for k, e in enumerate(env_ids):
        e = int(e)
        o = oracle.philox4x32_10([e & 0xFFFFFFFF, e >> 32, (frame >> 2) & 0xFFFFFFFF, stream ^ ((frame >> 34) << 8)],
                                 [seed & 0xFFFFFFFF, seed >> 32])
        out[k] = np.float32(int(o[frame & 3]) >> 8) * np.float32(1.0 / 16777216.0)
    return out


def randomize(module, std=0.6):
    torch.manual_seed(11)
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, std=std)
            nn.init.normal_(m.bias, std=0.3)
        if isinstance(m, nn.BatchNorm1d):
            m.running_mean.uniform_(-0.5, 0.5); m.running_var.uniform_(0.5, 2.0)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.uniform_(-0.3, 0.3)
    return module


def scaled_obs(n, device):
    """Observation-shaped inputs (positions in px, velocities in px/frame), scaled so the random nets do not saturate."""
    g = torch.Generator(device=device).manual_seed(5)
    o = torch.rand(n, 8, device=device, generator=g)
    o[:, :4] = torch.floor(o[:, :4] * 400) * 0.01
    o[:, 4:] = o[:, 4:] * 20 - 10
    return o.contiguous()


# ---------------------------------------------------------------- the network
@pytest.mark.parametrize("which", ["ppo", "a2c", "dueling", "continuous", "wide"])
def test_network_in_kernel_equals_torch_module(which):
    from air_hockey_training_environment_b200 import policy
    n = 5000
    env = make_env(n, seed=3)
    net = {"ppo": policy.ppo_actor_discrete, "a2c": policy.a2c_actor, "dueling": policy.DuelingQNet,
           "continuous": policy.ppo_actor_continuous,
           "wide": lambda: nn.Sequential(nn.Linear(8, 32), nn.Tanh(), nn.Linear(32, 32), nn.ReLU(), nn.Linear(32, 17), nn.ReLU(),
                                         nn.Linear(17, 7)).eval()}[which]()
    net = randomize(net).cuda().eval()
    head = {"ppo": None, "a2c": None, "dueling": None, "continuous": "continuous", "wide": "values"}[which]
    dp = policy.DevicePolicy(net, env.device, head=head, select="greedy")
    obs = scaled_obs(n, env.device)
    vals = torch.empty(n, dp.n_actions, device=env.device)
    acts = env.policy_act(dp, obs=obs, values_out=vals)
    with torch.no_grad():
        want = net(obs)
    assert (vals - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())
    if which == "continuous":
        assert torch.equal(acts, vals)                               # greedy on the continuous head: the mean itself
    else:
        assert torch.equal(acts.long(), vals.argmax(1))              # np.argmax: first maximum


@pytest.mark.parametrize("which", ["ppo", "a2c", "continuous", "odd"])
def test_register_path_of_narrow_networks_equals_the_general_path_bitwise(which, monkeypatch):
    """Networks whose widths are all <= 8 are evaluated from registers (policy_forward_narrow: exact unrolled bodies for
    the reference's layer shapes, a guarded body for any other, here 8-5-7-3); AH_POLICY_NO_NARROW forces the general
    shared-memory path.  Same multiply-add order per output: the head values must be the same BITS."""
    from air_hockey_training_environment_b200 import policy
    n = 3000
    env = make_env(n, seed=3)
    net = {"ppo": policy.ppo_actor_discrete, "a2c": policy.a2c_actor, "continuous": policy.ppo_actor_continuous,
           "odd": lambda: nn.Sequential(nn.Linear(8, 5), nn.Tanh(), nn.Linear(5, 7), nn.ReLU(), nn.Linear(7, 3)).eval()}[which]()
    net = randomize(net).cuda().eval()
    head = {"ppo": None, "a2c": None, "continuous": "continuous", "odd": "values"}[which]
    dp = policy.DevicePolicy(net, env.device, head=head, select="greedy")
    obs = scaled_obs(n, env.device)
    got = []
    for force_general in (False, True):
        if force_general:
            monkeypatch.setenv("AH_POLICY_NO_NARROW", "1")
        vals = torch.empty(n, dp.n_actions, device=env.device)
        env.policy_act(dp, obs=obs, values_out=vals)
        got.append(vals.clone())
    monkeypatch.delenv("AH_POLICY_NO_NARROW")
    assert torch.equal(got[0].view(torch.int32), got[1].view(torch.int32))
    with torch.no_grad():
        want = net(obs)
    assert (got[0] - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())


# ---------------------------------------------------------------- discrete strategies vs a numpy restatement
def restate_discrete(select, v, u0, u1, t, kw):
    """lib/exploration/strategies.py on ONE env: v = head values, (u0, u1) the frame's uniforms, t = step() call number."""
    A = len(v)

    def choice(p, u):
        c = np.float32(0)
        for q in range(A):
            c = np.float32(c + p[q])
            if u < c:
                return q
        return A - 1

    def boltz():
        e = np.exp(np.clip(v / np.float32(kw["tau"]), kw["clip"][0], kw["clip"][1]).astype(np.float32)).astype(np.float32)
        return e / e.sum(dtype=np.float32)

    def epsilon():
        delta = np.float32((kw["initial_epsilon"] - kw["final_epsilon"]) / kw["explore"])
        m_max = math.ceil((kw["epsilon"] - kw["final_epsilon"]) / float(delta))
        return np.float32(kw["epsilon"]) - np.float32(min(t // kw["observe"], m_max)) * delta

    if select == "sample":
        return choice(v, u0)
    if select == "greedy":
        return int(np.argmax(v))
    if select == "boltzmann":
        return choice(boltz(), u0)
    if u0 < epsilon():
        return min(int(u1 * np.float32(A)), A - 1) if select == "eps_greedy" else choice(boltz(), u1)
    return int(np.argmax(v))


@pytest.mark.parametrize("select", ["sample", "greedy", "eps_greedy", "boltzmann", "max_boltzmann"])
def test_exploration_strategies_match_numpy_restatement(select):
    from air_hockey_training_environment_b200 import policy
    n, seed, id0 = 1500, 77, 1_000_003
    env = make_env(n, seed=seed, env_id0=id0)
    net = randomize(policy.a2c_actor() if select == "sample" else policy.DuelingQNet(4), 0.3).cuda().eval()
    kw = dict(epsilon=0.9, initial_epsilon=1.0, final_epsilon=0.05, observe=5, explore=40, tau=0.7, clip=(-3.0, 3.0))
    dp = policy.DevicePolicy(net, env.device, select=select, **kw)
    obs = scaled_obs(n, env.device)
    picked = set()
    for frame in (0, 3, 4, 37, 250, 100_000):                       # epsilon: before / during / after the schedule
        env.set_counters(frame=frame)
        vals = torch.empty(n, 4, device=env.device)
        acts = env.policy_act(dp, obs=obs, values_out=vals).cpu().numpy()
        v = vals.cpu().numpy()
        ids = id0 + np.arange(n)
        u0, u1 = uniforms(seed, ids, frame, 2), uniforms(seed, ids, frame, 3)
        want = np.array([restate_discrete(select, v[i], u0[i], u1[i], frame + 1, kw) for i in range(n)])
        # a cumulative sum / epsilon within one float ulp of the uniform may round the other way in the kernel's order
        assert (acts != want).sum() <= 2, (frame, (acts != want).sum())
        picked |= set(acts.tolist())
    assert picked == {0, 1, 2, 3}
    if select in ("eps_greedy", "max_boltzmann"):                   # after the schedule: epsilon = final -> almost always greedy
        greedy = (acts == v.argmax(1)).mean()
        assert greedy > 0.9


def test_selection_only_policy_takes_head_values_from_any_network():
    """ah_policy_act with d_values_in: the caller's own torch network (here 51-atom C51-style heads reduced to Q-values)
    produces the head values; only the exploration strategy runs on the device."""
    from air_hockey_training_environment_b200 import policy
    n, seed = 4096, 5
    env = make_env(n, seed=seed)
    dp = policy.DevicePolicy(None, env.device, head="values", select="eps_greedy", n_actions=4, epsilon=0.3,
                             initial_epsilon=0.3, final_epsilon=0.3, observe=10, explore=10)
    q = torch.randn(n, 4, device=env.device)
    env.set_counters(frame=12)
    acts = env.policy_act(dp, values=q).cpu().numpy()
    u0, u1 = uniforms(seed, np.arange(n), 12, 2), uniforms(seed, np.arange(n), 12, 3)
    want = np.where(u0 < np.float32(0.3), np.minimum((u1 * np.float32(4)).astype(np.int64), 3), q.argmax(1).cpu().numpy())
    assert np.array_equal(acts, want)
    assert 0.2 < (u0 < 0.3).mean() < 0.4


# ---------------------------------------------------------------- continuous head: Gaussian / OU noise
@pytest.mark.parametrize("select", ["gaussian", "ou"])
def test_continuous_noise_processes(select):
    from air_hockey_training_environment_b200 import policy
    n, seed = 20000, 9
    env = make_env(n, seed=seed)
    net = randomize(policy.ppo_actor_continuous(), 0.2).cuda().eval()
    kw = dict(mu=0.1, sigma=0.8, sigma_min=0.2, n_steps_annealing=100, theta=0.15, dt=1e-2)
    dp = policy.DevicePolicy(net, env.device, head="continuous", select=select, n_envs=n, **kw)
    obs = scaled_obs(n, env.device)
    with torch.no_grad():
        mean = net(obs)
    x_prev = dp.noise_state.clone() if select == "ou" else None
    for frame in (0, 50, 400):
        env.set_counters(frame=frame)
        vals = torch.empty(n, 2, device=env.device)
        act = env.policy_act(dp, obs=obs, values_out=vals)
        assert (vals - mean).abs().max().item() < 1e-4
        u0, u1 = uniforms(seed, np.arange(2000), frame, 2), uniforms(seed, np.arange(2000), frame, 3)
        rad, ang = np.sqrt(-2.0 * np.log(1.0 - u0.astype(np.float64))), 2 * np.pi * u1.astype(np.float64)
        z = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
        sigma = max(kw["sigma_min"], kw["sigma"] - (kw["sigma"] - kw["sigma_min"]) * frame / kw["n_steps_annealing"])
        noise = (act - vals)[:2000].double().cpu().numpy()
        if select == "gaussian":
            want = kw["mu"] + sigma * z                                                  # random.py:38-41
        else:
            xp = x_prev[:2000].double().cpu().numpy()                                    # random.py:56-63
            want = xp + kw["theta"] * (kw["mu"] - xp) * kw["dt"] + sigma * math.sqrt(kw["dt"]) * z
            x_prev = dp.noise_state.clone()
            assert np.abs(x_prev[:2000].double().cpu().numpy() - want).max() < 2e-3
        assert np.abs(noise - want).max() < 2e-3 * max(1.0, np.abs(want).max())           # __logf / __cosf are approximate
    if select == "gaussian":                                                                # moments of the last frame's noise
        nz = (act - vals).double()
        assert abs(nz.mean().item() - kw["mu"]) < 0.01 and abs(nz.std().item() - 0.2) < 0.01


# ---------------------------------------------------------------- the closed loop in ONE launch
@pytest.mark.parametrize("which,select", [("a2c", "sample"), ("dueling", "eps_greedy"), ("continuous", "gaussian"), ("ppo", "max_boltzmann")])
def test_single_launch_policy_rollout_equals_frame_by_frame(which, select):
    """T frames with the policy evaluated in the step kernel (AH_ACT_POLICY_MLP, one launch) == T x (ah_policy_act on the
    current observation + a K = 1 step): same actions, same head values, bit-identical state and statistics."""
    from air_hockey_training_environment_b200 import policy
    n, T = 3000, 24
    a, b = make_env(n, seed=21), make_env(n, seed=21)
    for env in (a, b):
        env.step(None, K=300, out=env.make_outputs(300, ()))
        env.set_counters(frame=0)
        env.stats(clear=True)
    net = {"a2c": policy.a2c_actor, "dueling": policy.DuelingQNet, "continuous": policy.ppo_actor_continuous,
           "ppo": policy.ppo_actor_discrete}[which]()
    net = randomize(net, 0.05 if which != "continuous" else 0.02).cuda().eval()
    head = "continuous" if which == "continuous" else ("values" if which == "ppo" else None)   # (max-Boltzmann on logits-as-values)
    if which == "ppo":
        net = nn.Sequential(*list(net)[:-1]).eval()                   # drop the Softmax: Q-style head
    dp = policy.DevicePolicy(net, a.device, head=head, select=select, epsilon=0.5, initial_epsilon=0.5, final_epsilon=0.1,
                             observe=3, explore=10, tau=0.5)
    nv = dp.n_actions
    fields = ("reward", "done", "score", "game_over", "obs_next", "probs") + (("actions_real",) if dp.continuous else ("actions",))
    out = a.make_outputs(T, fields, obs_next_per_frame=True, n_values=nv)
    a.step(None, K=T, out=out, policy=dp)
    for t in range(T):
        vals = torch.empty(n, nv, device=b.device)
        act = b.policy_act(dp, values_out=vals)
        o1 = b.step(act, K=1, out=b.make_outputs(1, ("reward", "done", "score", "game_over", "obs_next")))
        assert torch.equal(vals, out.probs[t]), t
        assert torch.equal(act, out.actions_real[t] if dp.continuous else out.actions[t]), t
        assert torch.equal(o1.reward[0], out.reward[t]) and torch.equal(o1.obs_next, out.obs_next[t])
    for x, y in zip((a.puck, a.robot, a.opponent, a.prev_dist, a.meta), (b.puck, b.robot, b.opponent, b.prev_dist, b.meta)):
        assert torch.equal(x, y)
    assert a.stats() == b.stats() and a.counters() == b.counters()


def test_policy_argument_errors():
    from air_hockey_training_environment_b200 import AhError, policy
    env = make_env(256)
    dp = policy.DevicePolicy(policy.a2c_actor().cuda(), env.device)
    with pytest.raises(ValueError):
        env.step(torch.zeros(256, dtype=torch.int8, device="cuda"), K=1, policy=dp)          # actions AND policy
    with pytest.raises(ValueError):
        env.step(None, K=2, out=env.make_outputs(2, ("actions_real",)), policy=dp)            # discrete policy
    with pytest.raises(AhError):
        bad = policy.DevicePolicy(policy.a2c_actor().cuda(), env.device, select="gaussian")   # noise on a discrete head
        env.policy_act(bad)


# ---------------------------------------------------------------- a plain torch module handed to the rollout collector
def test_rollout_collector_exports_dense_torch_policies_and_follows_the_optimiser():
    """``RolloutCollector(policy=<torch module>)``: a small dense network (here a wrapper whose forward is a softmax over an
    nn.Sequential, like the reference-shaped actors) is exported after a numerical check against the module and the whole
    rollout runs in one launch; its probabilities equal the module's; parameters written by an optimiser reach the
    kernel at the next collect(); a callable the library cannot see into stays on the torch path."""
    from air_hockey_training_environment_b200 import policy
    from air_hockey_training_environment_b200.learners import A2CActor
    from air_hockey_training_environment_b200.rollout import RolloutCollector
    n, T = 2048, 12
    torch.manual_seed(3)
    actor = randomize(A2CActor(), 0.3).cuda().eval()
    a, b, c = make_env(n, seed=5), make_env(n, seed=5), make_env(n, seed=5)
    col = RolloutCollector(a, T, policy=actor)
    assert col.auto_exported and col.single_launch and col.device_policy.head == "softmax"
    ref = RolloutCollector(b, T, device_policy=policy.DevicePolicy(actor.net, b.device, head="softmax"))
    r, rr = col.collect(), ref.collect()
    for k in ("obs", "actions", "rewards", "dones", "probs"):
        assert torch.equal(r[k], rr[k]), k
    with torch.no_grad():
        assert torch.allclose(r["probs"][T - 1], actor(r["obs"][T - 1]), rtol=1e-4, atol=2e-5)
    # an optimiser step (in-place parameter write) is picked up without an explicit refresh
    opt = torch.optim.SGD(actor.parameters(), lr=0.02)
    sum(p.sum() for p in actor.parameters()).backward()                    # every parameter moves by -0.02
    opt.step()
    r = col.collect()
    with torch.no_grad():
        want = actor(r["obs"][3])
    assert torch.allclose(r["probs"][3], want, rtol=1e-4, atol=2e-5)
    assert not col.device_policy.refresh_if_changed()                      # nothing written since
    # not exportable: an opaque callable, a module that does more than its chain, a network wider than the kernel's limit
    opaque = RolloutCollector(c, T, policy=lambda obs: actor(obs))
    assert not opaque.auto_exported and not opaque.single_launch

    class Tempered(A2CActor):
        def forward(self, obs):
            return torch.softmax(self.net(obs) / 3.0, dim=-1)
    assert policy.export_if_dense(randomize(Tempered(), 0.3).cuda().eval(), c) is None
    assert policy.export_if_dense(randomize(A2CActor(), 0.3).cuda().eval(), c) is not None
    wide = nn.Sequential(nn.Linear(8, 64), nn.ReLU(), nn.Linear(64, 4), nn.Softmax(dim=-1)).cuda().eval()
    assert policy.export_if_dense(wide, c) is None
    assert policy.export_if_dense(policy.DuelingQNet(4).cuda().eval(), c).head == "dueling"
