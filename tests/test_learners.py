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
    col = RolloutCollector(env, 16, policy=actor, use_graph=True)
    assert col.fused is None and not col.single_launch                        # an arbitrary torch policy: graph path
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
