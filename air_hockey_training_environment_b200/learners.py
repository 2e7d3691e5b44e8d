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
restatements of the loss formulas and literal loops of the reference's target / projection code (tests/test_learners.py),
not against the reference itself.

Deliberate deviations of the policy-gradient learners (the value-based ones list theirs at each function):
  * returns: the reference discounts over ``MemoryBuffer.retreive()``, which is NEWEST-FIRST (lib/buffer.py:27,45), i.e.
    backwards in time; here rollouts are time-major with index 0 the OLDEST frame and ``discounted_returns`` is the
    ordinary reward-to-go (the one function used by both learners; the collector has no copy of it);
  * inputs: the reference trains on ``Observation.state``, the observation AFTER the first application of the action
    (lib/agents/agent.py:61); PPO / A2C here train on ``obs[t]``, the policy's own input at the start of frame t
    (lib/agents/agent.py:46) -- the state the stored action probabilities were computed from;
  * BatchNorm uses Keras' defaults (eps 1e-3, momentum 0.99 = torch 0.01) so that exported / folded statistics match.
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
        self.net = nn.Sequential(nn.Linear(state_size, 6), nn.Tanh(), nn.Linear(6, 6), nn.Tanh(), nn.BatchNorm1d(6, eps=1e-3, momentum=0.01),
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
        self.net = nn.Sequential(nn.Linear(state_size, state_size), nn.ReLU(), nn.BatchNorm1d(state_size, eps=1e-3, momentum=0.01),
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
        self.net = nn.Sequential(nn.Linear(state_size, state_size), nn.ReLU(), nn.BatchNorm1d(state_size, eps=1e-3, momentum=0.01), nn.Linear(state_size, 1))

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


# ------------------------------------------------------------------------------------------------ value-based learners
# DDQN / Dueling DDQN / C51 on minibatches drawn by the device sampler (``ReplayRing.sample`` -> ``ah_ring_sample``).  The
# reference's update rules are restated AS WRITTEN (their quirks are listed at each function and reproduced, not
# corrected); the one structural change is that a minibatch is fitted in one Adam step instead of the reference's
# sample-by-sample ``model.fit`` calls (lib/agents/ddqn/ddqn.py:103-125).
def continuous_ppo_actor_loss(y_true: torch.Tensor, y_pred: torch.Tensor, advantage: torch.Tensor,
                              old_prediction: torch.Tensor, loss_clipping: float = 0.2, noise: float = 1.0) -> torch.Tensor:
    """The reference's continuous PPO loss (lib/utils/helpers.py:49-69): Gaussian likelihood ratio with a FIXED sigma = 1
    around the predicted mean, clipped surrogate, mean over all B x 2 entries.  y_true = the action taken [B, 2]."""
    var = noise * noise
    denom = (2 * torch.pi * var) ** 0.5
    prob = torch.exp(-(y_true - y_pred) ** 2 / (2 * var)) / denom
    old_prob = torch.exp(-(y_true - old_prediction) ** 2 / (2 * var)) / denom
    r = prob / (old_prob + 1e-10)
    return -torch.minimum(r * advantage, r.clamp(1 - loss_clipping, 1 + loss_clipping) * advantage).mean()


def ddqn_targets(q_new_online: torch.Tensor, q_new_target: torch.Tensor, action: torch.Tensor, reward: torch.Tensor,
                 done: torch.Tensor, gamma: float) -> torch.Tensor:
    """lib/agents/ddqn/ddqn.py:108-121, as written: the regression target starts from the ONLINE net's prediction for the
    NEW state (:108, not the current one), and its ``action`` entry becomes ``reward`` when done, else
    ``reward + gamma * np.argmax(t)`` -- the INDEX of the target net's best action, not its value (:121)."""
    target = q_new_online.clone()
    boot = reward + gamma * q_new_target.argmax(dim=1).to(reward.dtype)
    target.scatter_(1, action.long().unsqueeze(1), torch.where(done.bool(), reward, boot).unsqueeze(1))
    return target


def dueling_targets(q_online: torch.Tensor, q_new_online: torch.Tensor, q_new_target: torch.Tensor, action: torch.Tensor,
                    reward: torch.Tensor, done: torch.Tensor, gamma: float) -> torch.Tensor:
    """lib/agents/dueling/dueling.py:148-176 (double DQN): the online net picks the next action, the target net values it."""
    target = q_online.clone()
    a = q_new_online.argmax(dim=1, keepdim=True)
    boot = reward + gamma * q_new_target.gather(1, a).squeeze(1)
    target.scatter_(1, action.long().unsqueeze(1), torch.where(done.bool(), reward, boot).unsqueeze(1))
    return target


def c51_optimal_actions(z_online: torch.Tensor, support: torch.Tensor) -> torch.Tensor:
    """lib/agents/c51/c51.py:169-172: q[i, a] = sum_j z[a][i][j] * support[j]; argmax over a.  z_online: [A, B, atoms]."""
    return (z_online * support).sum(dim=2).argmax(dim=0)


def c51_project(z_target_opt: torch.Tensor, action: torch.Tensor, reward: torch.Tensor, done: torch.Tensor, n_actions: int,
                gamma: float, v_min: float, v_max: float) -> torch.Tensor:
    """The categorical projection of lib/agents/c51/c51.py:175-190, vectorised in torch (the one-kernel device version
    is ``BatchedAirHockey.c51_project`` / ``ah_c51_project``).  z_target_opt [B, atoms] = the target net's distribution of
    the optimal next action; returns m_prob [A, B, atoms], non-zero only in row ``action[i]``.
    As written in the reference: ``m_l, m_u = floor(bj), ceil(bj)`` -- when bj is an integer both weights (m_u - bj) and
    (bj - m_l) are 0 and that atom's mass is dropped."""
    B, atoms = z_target_opt.shape
    dz = (v_max - v_min) / (atoms - 1)
    support = v_min + dz * torch.arange(atoms, dtype=z_target_opt.dtype, device=z_target_opt.device)
    dn = done.bool()
    tz = (reward.unsqueeze(1) + gamma * support.unsqueeze(0)).clamp(v_min, v_max)            # [B, atoms]
    tz = torch.where(dn.unsqueeze(1), reward.clamp(v_min, v_max).unsqueeze(1).expand(B, atoms), tz)
    mass = torch.where(dn.unsqueeze(1), torch.zeros_like(z_target_opt), z_target_opt)
    mass[:, 0] = torch.where(dn, torch.ones_like(reward), mass[:, 0])                       # terminal: a single unit point
    bj = (tz - v_min) / dz
    ml, mu = bj.floor(), bj.ceil()
    row = torch.zeros(B, atoms, dtype=z_target_opt.dtype, device=z_target_opt.device)
    row.scatter_add_(1, ml.long(), mass * (mu - bj))
    row.scatter_add_(1, mu.long(), mass * (bj - ml))
    m = torch.zeros(n_actions, B, atoms, dtype=z_target_opt.dtype, device=z_target_opt.device)
    m[action.long(), torch.arange(B, device=m.device)] = row
    return m


class QNet(nn.Module):
    """The reference's DDQN network: Dense(8, relu) - BatchNorm - Dense(4, softmax) (lib/agents/ddqn/model.py:39-53; yes,
    a softmax on Q-values -- as written)."""

    def __init__(self, state_size: int = 8, action_size: int = 4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(state_size, state_size), nn.ReLU(), nn.BatchNorm1d(state_size, eps=1e-3, momentum=0.01),
                                 nn.Linear(state_size, action_size), nn.Softmax(dim=-1))

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.net(obs)


class C51Net(nn.Module):
    """lib/agents/c51/model.py:44-63 as written: Dense(32, relu) - BatchNorm, then the A "heads" are CHAINED
    (``x = Dense(atoms)(x)`` reuses x), linear, un-normalised.  forward -> [A, B, atoms]."""

    def __init__(self, action_size: int = 4, atoms: int = 51, state_size: int = 8):
        super().__init__()
        self.trunk = nn.Sequential(nn.Linear(state_size, 32), nn.ReLU(), nn.BatchNorm1d(32, eps=1e-3, momentum=0.01))
        self.heads = nn.ModuleList([nn.Linear(32 if k == 0 else atoms, atoms) for k in range(action_size)])

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        x, outs = self.trunk(obs), []
        for h in self.heads:
            x = h(x)
            outs.append(x)
        return torch.stack(outs, 0)


class ReplayLearner:
    """Shared machinery: minibatch from the device ring, online / target networks, Adam, target sync."""

    def __init__(self, env, ring, online: nn.Module, target: nn.Module, gamma: float, lr: float, batch_size: int):
        self.env, self.ring, self.online, self.target = env, ring, online, target
        self.gamma, self.batch_size = gamma, int(batch_size)
        self.opt = torch.optim.Adam(online.parameters(), lr=lr)
        self.update_target_model()

    def update_target_model(self) -> None:
        self.target.load_state_dict(self.online.state_dict())
        self.target.eval()

    def minibatch(self) -> Dict[str, torch.Tensor]:
        held = len(self.ring) * self.ring.n
        return self.ring.sample(min(self.batch_size, held), self.env, check=False)      # lib/agents/ddqn/ddqn.py:101-102

    def fit(self, loss: torch.Tensor) -> float:
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return float(loss.detach())


class DDQNLearner(ReplayLearner):
    """lib/agents/ddqn/ddqn.py:95-125 (gamma 0.95, lr 1e-3; Huber loss, ddqn/model.py:27-53)."""

    def __init__(self, env, ring, gamma: float = 0.95, lr: float = 1e-3, batch_size: int = 10000):
        dev, dt = env.device, env.dtype
        super().__init__(env, ring, QNet().to(device=dev, dtype=dt), QNet().to(device=dev, dtype=dt), gamma, lr, batch_size)

    def update(self) -> Dict[str, float]:
        b = self.minibatch()
        self.online.eval()
        with torch.no_grad():
            target = ddqn_targets(self.online(b["new_state"]), self.target(b["new_state"]), b["action"], b["reward"], b["done"],
                                  self.gamma)
        if bool(b["done"].any()):                      # "if observation.done: self.update_target_model()"  (:111-113)
            self.update_target_model()
        self.online.train()
        loss = self.fit(huber_loss(target, self.online(b["state"])).mean())
        self.online.eval()
        return {"loss": loss, "batch": int(b["reward"].numel())}


class DuelingLearner(ReplayLearner):
    """lib/agents/dueling/dueling.py:125-178 (gamma 0.9, lr 1e-5; the target net is synced on every update, :131)."""

    def __init__(self, env, ring, gamma: float = 0.9, lr: float = 1e-5, batch_size: int = 10000):
        from .policy import DuelingQNet
        dev, dt = env.device, env.dtype
        super().__init__(env, ring, DuelingQNet().to(device=dev, dtype=dt), DuelingQNet().to(device=dev, dtype=dt), gamma, lr, batch_size)

    def update(self) -> Dict[str, float]:
        self.update_target_model()
        b = self.minibatch()
        with torch.no_grad():
            target = dueling_targets(self.online(b["state"]), self.online(b["new_state"]), self.target(b["new_state"]),
                                     b["action"], b["reward"], b["done"], self.gamma)
        loss = self.fit(huber_loss(target, self.online(b["state"])).mean())
        return {"loss": loss, "batch": int(b["reward"].numel())}


class C51Learner(ReplayLearner):
    """lib/agents/c51/c51.py:139-195 (51 atoms on [-10, 10], gamma 0.95, 10 epochs per update); the projection runs as one
    kernel on the device (``ah_c51_project``)."""

    def __init__(self, env, ring, gamma: float = 0.95, lr: float = 1e-4, batch_size: int = 10000, atoms: int = 51,
                 v_min: float = -10.0, v_max: float = 10.0, epochs: int = 10, device_projection: bool = True):
        dev, dt = env.device, env.dtype
        super().__init__(env, ring, C51Net(atoms=atoms).to(device=dev, dtype=dt), C51Net(atoms=atoms).to(device=dev, dtype=dt),
                         gamma, lr, batch_size)
        self.atoms, self.v_min, self.v_max, self.epochs, self.device_projection = atoms, v_min, v_max, epochs, device_projection
        self.support = torch.linspace(v_min, v_max, atoms, dtype=dt, device=dev)

    def q_values(self, obs: torch.Tensor) -> torch.Tensor:
        """lib/agents/c51/c51.py:124-131: Q[i, a] = sum_j z[a][i][j] * support[j] -- the head values for ``policy_act``."""
        return (self.online(obs) * self.support).sum(dim=2).t().contiguous()

    def update(self) -> Dict[str, float]:
        self.update_target_model()
        b = self.minibatch()
        self.online.eval()
        with torch.no_grad():
            best = c51_optimal_actions(self.online(b["new_state"]), self.support)
            z_t = self.target(b["new_state"])                                              # [A, B, atoms]
            z_opt = z_t[best, torch.arange(best.numel(), device=best.device)].contiguous()
            if self.device_projection:
                m_prob = self.env.c51_project(z_opt, b["action"], b["reward"], b["done"], 4, self.gamma, self.v_min, self.v_max)
            else:
                m_prob = c51_project(z_opt, b["action"], b["reward"], b["done"], 4, self.gamma, self.v_min, self.v_max)
        self.online.train()
        loss = 0.0
        for _ in range(self.epochs):
            loss = self.fit(huber_loss(m_prob, self.online(b["state"])).mean())
        self.online.eval()
        return {"loss": loss, "batch": int(b["reward"].numel())}
