"""A PPO learner in torch that consumes the zero-copy rollout tensors of ``rollout.RolloutCollector`` (SURVEY.md §8f,
row N2).  This is a CONSUMER of the hot path, not part of it: tiny MLPs trained with autograd/Adam on the device.

Restated from the reference (paths relative to the reference root):
    actor  8 -> 4 relu -> BatchNorm -> 6 relu -> 4 relu -> 4 softmax        lib/agents/ppo/model.py:56-74  (rollout.ActorMLP)
    critic 8 -> 6 tanh -> 6 tanh -> BatchNorm -> 4 tanh -> 1 linear         lib/agents/ppo/model.py:113-137
    clipped-surrogate actor loss, eps = 0.2, entropy 1e-3                    lib/utils/helpers.py:29-46
    Huber critic loss, delta = 1                                             lib/utils/helpers.py:11-21
    Monte-Carlo discounted returns, advantage = return - V(s)                lib/agents/ppo/ppo.py:142-160
    Adam, lr 1e-5, gamma 0.8, 10 epochs                                      lib/agents/ppo/model.py:30-44
The reference's Keras/TF 1.12 learner cannot run in this image, so this module is checked against numpy
restatements of the two loss formulas (tests/test_learners.py), not against the reference itself.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from .rollout import ActorMLP, RolloutCollector


class CriticMLP(nn.Module):
    """lib/agents/ppo/model.py:113-137 (``kernel_initializer="normal"`` = N(0, 0.05))."""

    def __init__(self, state_size: int = 8):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(state_size, 6), nn.Tanh(), nn.Linear(6, 6), nn.Tanh(), nn.BatchNorm1d(6),
                                 nn.Linear(6, 4), nn.Tanh(), nn.Linear(4, 1))
        for m in self.net:
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.05)
                nn.init.zeros_(m.bias)

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.net(obs).squeeze(-1)


def huber_loss(y_true: torch.Tensor, y_pred: torch.Tensor, clip_delta: float = 1.0) -> torch.Tensor:
    """Element-wise Huber loss, lib/utils/helpers.py:11-21."""
    error = y_true - y_pred
    quad = 0.5 * error * error
    lin = clip_delta * (error.abs() - 0.5 * clip_delta)
    return torch.where(error.abs() < clip_delta, quad, lin)


def ppo_actor_loss(y_true: torch.Tensor, y_pred: torch.Tensor, advantage: torch.Tensor, old_prediction: torch.Tensor,
                   loss_clipping: float = 0.2, entropy_loss: float = 1e-3) -> torch.Tensor:
    """The reference's discrete PPO loss exactly as written (lib/utils/helpers.py:29-46): ``y_true`` is the one-hot
    action matrix [B, A], ``y_pred`` / ``old_prediction`` the new / old action probabilities [B, A], ``advantage`` [B, 1].
    The product with the one-hot mask is element-wise and the mean runs over ALL B*A entries (the A-1 masked entries
    of a row have r = 0, so they contribute min(0, (1 - eps) * advantage)) -- reproduced, not corrected."""
    prob = y_true * y_pred
    old_prob = y_true * old_prediction
    r = prob / (old_prob + 1e-10)
    surrogate = torch.minimum(r * advantage, r.clamp(1 - loss_clipping, 1 + loss_clipping) * advantage)
    return -(surrogate + entropy_loss * -(prob * torch.log(prob + 1e-10))).mean()


def discounted_returns(rewards: torch.Tensor, gamma: float) -> torch.Tensor:
    """``rewards[j] += rewards[j + 1] * gamma`` for j = T-2 .. 0 over a time-major [T, N] tensor
    (lib/agents/ppo/ppo.py:142-146).  NOTE: the reference feeds this its buffer NEWEST-FIRST (``MemoryBuffer.retreive``,
    lib/buffer.py:27,45), i.e. it discounts backwards in time; here index 0 is the OLDEST frame, so these are the
    ordinary reward-to-go returns."""
    ret = rewards.clone()
    for t in range(ret.shape[0] - 2, -1, -1):
        ret[t] += gamma * ret[t + 1]
    return ret


class PPOLearner:
    """Rollout (device tensors, zero-copy) -> advantages -> ``epochs`` full-batch Adam steps on actor and critic ->
    refreshed weights for the in-kernel actor."""

    def __init__(self, collector: RolloutCollector, gamma: float = 0.8, actor_lr: float = 1e-5, critic_lr: float = 1e-5,
                 epochs: int = 10, minibatch: Optional[int] = None):
        if not isinstance(collector.policy, ActorMLP):
            raise ValueError("PPOLearner trains the reference-shaped ActorMLP")
        self.col, self.gamma, self.epochs, self.minibatch = collector, gamma, epochs, minibatch
        self.actor = collector.policy
        dev, dt = collector.env.device, collector.env.dtype
        self.critic = CriticMLP().to(device=dev, dtype=dt)
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=actor_lr)
        self.opt_critic = torch.optim.Adam(self.critic.parameters(), lr=critic_lr)

    def update(self) -> Dict[str, float]:
        """One PPO update from the collector's current rollout tensors."""
        c = self.col
        T, n = c.T, c.env.n
        obs = c.obs[:T].reshape(T * n, 8)
        returns = discounted_returns(c.rewards, self.gamma).reshape(T * n)
        old_pred = c.probs.reshape(T * n, c.A).detach()
        onehot = torch.nn.functional.one_hot(c.actions.reshape(T * n).long(), c.A).to(obs.dtype)
        self.critic.eval()
        with torch.no_grad():
            advantage = (returns - self.critic(obs)).unsqueeze(1)
        self.actor.train(); self.critic.train()
        B = T * n
        mb = self.minibatch or B
        la = lc = 0.0
        for _ in range(self.epochs):
            perm = torch.randperm(B, device=obs.device) if mb < B else None
            for s in range(0, B, mb):
                idx = slice(s, s + mb) if perm is None else perm[s:s + mb]
                loss_a = ppo_actor_loss(onehot[idx], self.actor(obs[idx]), advantage[idx], old_pred[idx])
                self.opt_actor.zero_grad(set_to_none=True)
                loss_a.backward()
                self.opt_actor.step()
                loss_c = huber_loss(returns[idx], self.critic(obs[idx])).mean()
                self.opt_critic.zero_grad(set_to_none=True)
                loss_c.backward()
                self.opt_critic.step()
                la, lc = float(loss_a.detach()), float(loss_c.detach())
        self.actor.eval(); self.critic.eval()
        if c.fused is not None:
            c.fused.refresh()                  # the in-kernel actor reads the new weights from the next launch on
        return {"actor_loss": la, "critic_loss": lc, "mean_return": float(returns.mean()), "mean_reward": float(c.rewards.mean())}


# ------------------------------------------------------------------------------------------------ A2C
class A2CActor(nn.Module):
    """lib/agents/a2c/model.py:50-60: Dense(8, relu) - BatchNorm - Dense(4, softmax)."""

    def __init__(self, state_size: int = 8, action_size: int = 4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(state_size, state_size), nn.ReLU(), nn.BatchNorm1d(state_size),
                                 nn.Linear(state_size, action_size))
        for m in self.net:
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.05)
                nn.init.zeros_(m.bias)
        self.eval()

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return torch.softmax(self.net(obs), dim=-1)


class A2CCritic(nn.Module):
    """lib/agents/a2c/model.py:64-75: Dense(8, relu) - BatchNorm - Dense(1, linear)."""

    def __init__(self, state_size: int = 8):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(state_size, state_size), nn.ReLU(), nn.BatchNorm1d(state_size), nn.Linear(state_size, 1))

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.net(obs).squeeze(-1)


def a2c_actor_loss(advantage_targets: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """Keras ``categorical_crossentropy(y_true, y_pred)`` with the reference's targets: a one-hot row whose "1" is replaced
    by the advantage (lib/agents/a2c/a2c.py:137-143).  Keras renormalises y_pred, clips it to [1e-7, 1 - 1e-7] and
    returns -sum_j y_true_j * log(y_pred_j) per sample; the fit minimises its mean."""
    p = y_pred / y_pred.sum(dim=-1, keepdim=True)
    p = p.clamp(1e-7, 1 - 1e-7)
    return -(advantage_targets * torch.log(p)).sum(dim=-1).mean()


class A2CLearner:
    """Monte-Carlo returns (gamma 0.95, a2c/model.py:33), advantage-weighted cross-entropy for the actor, Huber for the
    critic, Adam(eps=1e-3) -- lib/agents/a2c/a2c.py:107-148.  The policy is an ordinary torch module, so the collector
    runs it through the CUDA-graph path (any torch policy works zero-copy; only the PPO-shaped actor is in-kernel)."""

    def __init__(self, collector: RolloutCollector, gamma: float = 0.95, actor_lr: float = 1e-5, critic_lr: float = 1e-5,
                 epochs: int = 10):
        self.col, self.gamma, self.epochs = collector, gamma, epochs
        self.actor = collector.policy
        dev, dt = collector.env.device, collector.env.dtype
        self.critic = A2CCritic().to(device=dev, dtype=dt)
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=actor_lr, eps=1e-3)
        self.opt_critic = torch.optim.Adam(self.critic.parameters(), lr=critic_lr, eps=1e-3)

    def update(self) -> Dict[str, float]:
        c = self.col
        T, n = c.T, c.env.n
        obs = c.obs[:T].reshape(T * n, 8)
        returns = discounted_returns(c.rewards, self.gamma).reshape(T * n)
        self.critic.eval()
        with torch.no_grad():
            adv = returns - self.critic(obs)
        targets = torch.zeros(T * n, c.A, dtype=obs.dtype, device=obs.device)
        targets.scatter_(1, c.actions.reshape(T * n, 1).long(), adv.unsqueeze(1))
        self.actor.train(); self.critic.train()
        la = lc = 0.0
        for _ in range(self.epochs):
            loss_a = a2c_actor_loss(targets, self.actor(obs))
            self.opt_actor.zero_grad(set_to_none=True); loss_a.backward(); self.opt_actor.step()
            loss_c = huber_loss(returns, self.critic(obs)).mean()
            self.opt_critic.zero_grad(set_to_none=True); loss_c.backward(); self.opt_critic.step()
            la, lc = float(loss_a.detach()), float(loss_c.detach())
        self.actor.eval(); self.critic.eval()
        return {"actor_loss": la, "critic_loss": lc, "mean_reward": float(c.rewards.mean())}
