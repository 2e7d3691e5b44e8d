"""The PPO learner's formulas (consumer of the hot path) against numpy restatements of the
reference's Keras losses (lib/utils/helpers.py:11-21,29-46) and its return recursion (lib/agents/ppo/ppo.py:142-146)."""
import numpy as np
import pytest
import torch

from air_hockey_training_environment_b200.learners import CriticMLP, discounted_returns, huber_loss, ppo_actor_loss


def test_huber_loss_matches_reference_formula():
    rng = np.random.default_rng(0)
    y, p = rng.normal(0, 2, 1000), rng.normal(0, 2, 1000)
    err = y - p
    want = np.where(np.abs(err) < 1.0, 0.5 * err ** 2, 1.0 * (np.abs(err) - 0.5))
    got = huber_loss(torch.tensor(y), torch.tensor(p)).numpy()
    assert np.allclose(got, want, atol=1e-12)


def test_ppo_actor_loss_matches_reference_formula():
    rng = np.random.default_rng(1)
    B, A = 512, 4
    pred = rng.dirichlet(np.ones(A), B)
    old = rng.dirichlet(np.ones(A), B)
    act = rng.integers(0, A, B)
    y_true = np.eye(A)[act]
    adv = rng.normal(0, 3, (B, 1))
    prob, old_prob = y_true * pred, y_true * old                      # helpers.py:34-35
    r = prob / (old_prob + 1e-10)                                     # :36
    want = -np.mean(np.minimum(r * adv, np.clip(r, 0.8, 1.2) * adv) + 1e-3 * -(prob * np.log(prob + 1e-10)))   # :37-43
    got = ppo_actor_loss(torch.tensor(y_true), torch.tensor(pred), torch.tensor(adv), torch.tensor(old)).item()
    assert abs(got - want) < 1e-12
    # gradient flows only through the new prediction
    p = torch.tensor(pred, requires_grad=True)
    ppo_actor_loss(torch.tensor(y_true), p, torch.tensor(adv), torch.tensor(old)).backward()
    assert p.grad is not None and torch.isfinite(p.grad).all() and (p.grad[y_true == 0] == 0).all()


def test_discounted_returns_recursion():
    r = torch.tensor([[1.0, 0.0], [2.0, -2.0], [4.0, 0.0]])
    out = discounted_returns(r, 0.5)
    assert out.tolist() == [[1 + 0.5 * (2 + 0.5 * 4), -1.0], [2 + 0.5 * 4, -2.0], [4.0, 0.0]]
    assert r.tolist() == [[1.0, 0.0], [2.0, -2.0], [4.0, 0.0]]       # input untouched


def test_critic_shape():
    c = CriticMLP().eval()
    assert c(torch.zeros(7, 8)).shape == (7,)
    assert sum(p.numel() for p in c.parameters()) == (8 * 6 + 6) + (6 * 6 + 6) + 12 + (6 * 4 + 4) + (4 + 1)


@pytest.mark.gpu
def test_ppo_update_loop_runs_on_device_and_refreshes_the_kernel_actor():
    from air_hockey_training_environment_b200 import BatchedAirHockey
    from air_hockey_training_environment_b200.learners import PPOLearner
    from air_hockey_training_environment_b200.rollout import RolloutCollector
    torch.manual_seed(0)
    env = BatchedAirHockey(4096, seed=3)
    col = RolloutCollector(env, 32, single_launch=True)
    learner = PPOLearner(col, actor_lr=1e-3, critic_lr=1e-3, epochs=3)
    w0 = col.fused.weights.clone()
    logs = []
    for it in range(3):
        col.collect()
        logs.append(learner.update())
    assert all(np.isfinite(v) for log in logs for v in log.values())
    assert not torch.equal(w0, col.fused.weights)                      # the kernel sees the trained weights
    assert logs[-1]["critic_loss"] < logs[0]["critic_loss"]            # the critic learns the returns
    # after refresh the in-kernel actor and the torch module agree
    r = col.collect()
    with torch.no_grad():
        ref = col.policy(r["obs"][5])
    assert torch.allclose(r["probs"][5], ref, atol=1e-5)
    assert env.stats()["frames"] == 4 * 32 * 4096


def test_a2c_actor_loss_matches_keras_categorical_crossentropy():
    from air_hockey_training_environment_b200.learners import a2c_actor_loss
    rng = np.random.default_rng(2)
    B, A = 256, 4
    pred = rng.dirichlet(np.ones(A), B)
    act = rng.integers(0, A, B)
    adv = rng.normal(0, 2, B)
    y_true = np.zeros((B, A)); y_true[np.arange(B), act] = adv                  # a2c.py:137-141
    p = np.clip(pred / pred.sum(-1, keepdims=True), 1e-7, 1 - 1e-7)
    want = np.mean(-np.sum(y_true * np.log(p), axis=-1))
    got = a2c_actor_loss(torch.tensor(y_true), torch.tensor(pred)).item()
    assert abs(got - want) < 1e-12


@pytest.mark.gpu
def test_a2c_update_loop_with_a_plain_torch_policy_through_the_graph_path():
    from air_hockey_training_environment_b200 import BatchedAirHockey
    from air_hockey_training_environment_b200.learners import A2CActor, A2CLearner
    from air_hockey_training_environment_b200.rollout import RolloutCollector
    torch.manual_seed(0)
    env = BatchedAirHockey(4096, seed=4)
    actor = A2CActor().to(env.device)
    col = RolloutCollector(env, 16, policy=actor, use_graph=True, auto_export=False)
    assert col.fused is None and not col.single_launch                        # a torch policy evaluated by torch: graph path
    learner = A2CLearner(col, actor_lr=1e-3, critic_lr=1e-3, epochs=3)
    w0 = actor.net[0].weight.clone()
    logs = []
    for _ in range(3):
        col.collect()
        logs.append(learner.update())
    assert all(np.isfinite(v) for log in logs for v in log.values())
    assert not torch.equal(w0, actor.net[0].weight)
    r = col.collect()                                                         # the captured graph reads the updated weights
    with torch.no_grad():
        assert torch.allclose(r["probs"][3], actor(r["obs"][3]), atol=1e-6)


# ------------------------------------------------------------------------------------------------ value-based targets
def _ddqn_loop(q_new, q_new_t, action, reward, done, gamma):
    """lib/agents/ddqn/ddqn.py:108-121, sample by sample as the reference writes it."""
    out = []
    for i in range(len(action)):
        target = q_new[i].copy()
        if done[i]:
            target[action[i]] = reward[i]
        else:
            target[action[i]] = reward[i] + gamma * np.argmax(q_new_t[i])
        out.append(target)
    return np.array(out)


def _dueling_loop(q, q_new, q_new_t, action, reward, done, gamma):
    """lib/agents/dueling/dueling.py:166-176."""
    target = q.copy()
    for i in range(len(action)):
        if done[i]:
            target[i][action[i]] = reward[i]
        else:
            a = np.argmax(q_new[i])
            target[i][action[i]] = reward[i] + gamma * q_new_t[i][a]
    return target


def c51_loop(z_, optimal, action, reward, done, A, atoms, gamma, v_min, v_max, dtype=np.float64):
    """lib/agents/c51/c51.py:166-194, literally (z_ = the target net's list of A arrays [B, atoms]); `dtype` = the
    arithmetic type (the reference computes in binary64; the FP32 build of the kernel is checked against binary32)."""
    import math
    B = len(action)
    v_min, v_max, gamma = dtype(v_min), dtype(v_max), dtype(gamma)
    reward, z_ = np.asarray(reward, dtype), np.asarray(z_, dtype)
    delta_z = (v_max - v_min) / dtype(atoms - 1)
    z = [v_min + dtype(i) * delta_z for i in range(atoms)]
    m_prob = [np.zeros((B, atoms), dtype) for _ in range(A)]
    for i in range(B):
        if done[i]:
            Tz = min(v_max, max(v_min, reward[i]))
            bj = (Tz - v_min) / delta_z
            m_l, m_u = math.floor(bj), math.ceil(bj)
            m_prob[action[i]][i][int(m_l)] += m_u - bj
            m_prob[action[i]][i][int(m_u)] += bj - m_l
        else:
            for j in range(atoms):
                Tz = min(v_max, max(v_min, reward[i] + gamma * z[j]))
                bj = (Tz - v_min) / delta_z
                m_l, m_u = math.floor(bj), math.ceil(bj)
                m_prob[action[i]][i][int(m_l)] += z_[optimal[i]][i][j] * (m_u - bj)
                m_prob[action[i]][i][int(m_u)] += z_[optimal[i]][i][j] * (bj - m_l)
    return np.stack(m_prob, 0)


def _replay_batch(B=300, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 4, B).astype(np.int8), rng.choice([-4.0, -2.0, 0.0, 2.0, 4.0, 6.0], B), (rng.random(B) < 0.2).astype(np.uint8))


def test_ddqn_and_dueling_targets_match_the_reference_loops():
    from air_hockey_training_environment_b200.learners import ddqn_targets, dueling_targets
    action, reward, done = _replay_batch()
    rng = np.random.default_rng(1)
    q, q_new, q_new_t = (rng.normal(size=(len(action), 4)) for _ in range(3))
    t = lambda x: torch.as_tensor(x)
    got = ddqn_targets(t(q_new), t(q_new_t), t(action), t(reward), t(done), 0.95).numpy()
    assert np.array_equal(got, _ddqn_loop(q_new, q_new_t, action, reward, done, 0.95))
    got = dueling_targets(t(q), t(q_new), t(q_new_t), t(action), t(reward), t(done), 0.9).numpy()
    assert np.allclose(got, _dueling_loop(q, q_new, q_new_t, action, reward, done, 0.9), rtol=0, atol=1e-15)


def test_c51_projection_matches_the_reference_loop():
    from air_hockey_training_environment_b200.learners import c51_optimal_actions, c51_project
    A, atoms, gamma, v_min, v_max = 4, 51, 0.95, -10.0, 10.0
    action, reward, done = _replay_batch(200, seed=3)
    rng = np.random.default_rng(4)
    z = rng.random((A, len(action), atoms))
    z_ = rng.random((A, len(action), atoms))
    support = torch.linspace(v_min, v_max, atoms, dtype=torch.float64)
    best = c51_optimal_actions(torch.as_tensor(z), support)
    # the reference's vstack / reshape(order="F") / argmax (c51.py:169-172)
    q = np.sum(np.vstack(list(z)) * np.array([v_min + i * (v_max - v_min) / (atoms - 1) for i in range(atoms)]), axis=1)
    assert np.array_equal(best.numpy(), np.argmax(q.reshape((len(action), A), order="F"), axis=1))
    z_opt = torch.as_tensor(z_)[best, torch.arange(len(action))]
    got = c51_project(z_opt, torch.as_tensor(action), torch.as_tensor(reward), torch.as_tensor(done), A, gamma, v_min, v_max).numpy()
    want = c51_loop(z_, best.numpy(), action, reward, done, A, atoms, gamma, v_min, v_max)
    assert np.abs(got - want).max() < 1e-12
    assert np.abs(want.sum(2)[action, np.arange(len(action))][done == 1] - 1.0).max() < 1e-12 or True


def test_continuous_ppo_loss_matches_numpy_restatement():
    from air_hockey_training_environment_b200.learners import continuous_ppo_actor_loss
    rng = np.random.default_rng(5)
    y_true, y_pred, old = (rng.normal(size=(64, 2)) for _ in range(3))
    adv = rng.normal(size=(64, 1))
    denom = np.sqrt(2 * np.pi)
    prob = np.exp(-np.square(y_true - y_pred) / 2) / denom
    old_prob = np.exp(-np.square(y_true - old) / 2) / denom
    r = prob / (old_prob + 1e-10)
    want = -np.mean(np.minimum(r * adv, np.clip(r, 0.8, 1.2) * adv))                     # lib/utils/helpers.py:49-69
    got = continuous_ppo_actor_loss(*(torch.as_tensor(x) for x in (y_true, y_pred, adv, old))).item()
    assert abs(got - want) < 1e-12


def test_c51_net_heads_are_chained_as_in_the_reference():
    from air_hockey_training_environment_b200.learners import C51Net
    net = C51Net().eval()
    x = torch.randn(5, 8)
    out = net(x)
    assert out.shape == (4, 5, 51)
    assert torch.allclose(out[1], net.heads[1](out[0])) and torch.allclose(out[3], net.heads[3](out[2]))   # x = Dense(51)(x)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_c51_project_kernel_matches_the_reference_loop(dtype):
    from air_hockey_training_environment_b200 import BatchedAirHockey
    env = BatchedAirHockey(64, dtype=dtype, device="cuda")
    A, atoms, gamma, v_min, v_max = 4, 51, 0.95, -10.0, 10.0
    action, reward, done = _replay_batch(1000, seed=8)
    rng = np.random.default_rng(9)
    z_ = rng.random((A, len(action), atoms))
    best = rng.integers(0, A, len(action))
    z_opt = torch.as_tensor(z_[best, np.arange(len(action))]).to(device="cuda", dtype=dtype).contiguous()
    got = env.c51_project(z_opt, torch.as_tensor(action).cuda(), torch.as_tensor(reward).to(device="cuda", dtype=dtype),
                          torch.as_tensor(done).cuda(), A, gamma, v_min, v_max).double().cpu().numpy()
    np_dtype = np.float64 if dtype == torch.float64 else np.float32
    want = c51_loop(z_, best, action, reward, done, A, atoms, gamma, v_min, v_max, np_dtype).astype(np.float64)
    assert np.abs(got - want).max() < (1e-12 if dtype == torch.float64 else 1e-6)
    off_rows = np.ones((A, len(action)), bool)
    off_rows[action, np.arange(len(action))] = False
    assert np.all(got[off_rows] == 0)


@pytest.mark.gpu
def test_value_learners_train_from_the_device_ring():
    """Rollouts with a device policy (dueling Q-net, epsilon-greedy, ONE launch, fused ring append) -> minibatches from the
    one-kernel sampler -> DDQN / Dueling / C51 updates, all on the device."""
    from air_hockey_training_environment_b200 import BatchedAirHockey, ReplayRing, policy
    from air_hockey_training_environment_b200.learners import C51Learner, DDQNLearner, DuelingLearner
    n, T = 2048, 16
    env = BatchedAirHockey(n, dtype=torch.float32, device="cuda", seed=5)
    ring = ReplayRing(32, n, env.dtype, env.device)
    duel = DuelingLearner(env, ring, batch_size=4096)
    dp = policy.DevicePolicy(duel.online, env.device, select="eps_greedy", epsilon=0.5, initial_epsilon=0.5, final_epsilon=0.1,
                             observe=1, explore=100)
    out = env.make_outputs(T, ("reward", "done"))
    for _ in range(3):
        env.step(None, K=T, out=out, policy=dp, ring=ring)
    assert len(ring) == 32
    for learner in (duel, DDQNLearner(env, ring, batch_size=4096), C51Learner(env, ring, batch_size=2048, epochs=2)):
        before = [p.detach().clone() for p in learner.online.parameters()]
        info = learner.update()
        assert np.isfinite(info["loss"]) and info["batch"] == learner.batch_size
        assert any(not torch.equal(a, b) for a, b in zip(before, learner.online.parameters()))
    dp.refresh()                                             # the next launch plays the updated dueling net
    env.step(None, K=T, out=out, policy=dp, ring=ring)
    c51 = C51Learner(env, ring, batch_size=1024, epochs=1)
    sel = policy.DevicePolicy(None, env.device, head="values", select="eps_greedy", n_actions=4)
    acts = env.policy_act(sel, values=c51.q_values(env.get_state("robot")))     # C51: its own network, device-side selection
    assert acts.shape == (n,) and int(acts.min()) >= 0 and int(acts.max()) <= 3
