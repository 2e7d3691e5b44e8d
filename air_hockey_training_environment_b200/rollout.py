"""Zero-copy rollout collection: the step kernel writes observations, rewards and dones straight into the
preallocated ``[T(+1), N, ...]`` rollout tensors that a PPO / A2C / DQN learner consumes on the device -- no host
round-trip anywhere in the loop (BASELINE config 5: 65,536 envs x 128 steps feeding a torch policy).

What it replaces in the reference (paths relative to the reference root):
    Agent.get_action -> env.get_state -> serialize_state -> policy          lib/agents/agent.py:42-48
    Agent.move / Agent.step / env.update_score / computer sub-update        game.py:133-177 (fused: env.step)
    self.memory.append(observation)                                          lib/agents/ppo/ppo.py:144, lib/buffer.py:24-32
The learners themselves (Keras PPO/A2C/DDQN, lib/agents/**) are consumers of these tensors and are out of scope
here; ``ActorMLP`` only mirrors the reference's discrete PPO actor so the rollout runs a policy of the right shape.

The whole T-step loop (policy forward, action sampling, env.step) can be captured in ONE CUDA graph: every C-ABI
call enqueues on the caller's stream without allocating or synchronising, and the frame / ring counters live on
the device.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
from torch import nn

from . import _capi
from .env import BatchedAirHockey, StepOutputs


class ActorMLP(nn.Module):
    """The reference's discrete PPO actor: Dense(4, relu) - BatchNorm - Dense(6, relu) - Dense(4, relu) -
    Dense(4, softmax)  (lib/agents/ppo/model.py:63-74; ``kernel_initializer="normal"`` is N(0, 0.05) in Keras; BatchNorm
    with Keras' defaults eps = 1e-3, momentum = 0.99, i.e. torch momentum 0.01).  Inference only (BatchNorm in eval mode)."""

    def __init__(self, state_size: int = 8, action_size: int = 4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(state_size, 4), nn.ReLU(), nn.BatchNorm1d(4, eps=1e-3, momentum=0.01), nn.Linear(4, 6), nn.ReLU(),
                                 nn.Linear(6, 4), nn.ReLU(), nn.Linear(4, action_size))
        for m in self.net:
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.05)
                nn.init.zeros_(m.bias)
        self.eval()

    def forward(self, obs: torch.Tensor) -> torch.Tensor:      # obs [N, 8] -> action probabilities [N, 4]
        return torch.softmax(self.net(obs), dim=-1)


class FusedActor:
    """Inference-only twin of ``ActorMLP`` running in ONE kernel together with the action sampling
    (``ah_actor_sample``): obs -> probabilities + sampled action, no intermediate tensors.  The torch module stays
    the owner of the parameters (a learner trains it with autograd); ``refresh()`` re-reads them, folding the
    BatchNorm's affine map into the second Dense layer."""

    def __init__(self, module: ActorMLP, env: BatchedAirHockey):
        self.module, self.env = module, env
        self.weights = torch.zeros(_capi.AH_ACTOR_NWEIGHTS, dtype=torch.float32, device=env.device)
        self._lib = _capi.lib()
        self.refresh()

    @torch.no_grad()
    def refresh(self) -> None:
        l1, _, bn, l2, _, l3, _, l4 = self.module.net
        if l4.out_features != 4:
            raise ValueError("the fused actor implements the reference's 4-action discrete actor")
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        t = bn.bias - bn.running_mean * s
        w2 = l2.weight * s[None, :]
        b2 = l2.bias + l2.weight @ t
        flat = torch.cat([l1.weight.flatten(), l1.bias, w2.flatten(), b2, l3.weight.flatten(), l3.bias,
                          l4.weight.flatten(), l4.bias]).to(torch.float32)
        assert flat.numel() == _capi.AH_ACTOR_NWEIGHTS
        self.weights.copy_(flat)

    def sample(self, obs: torch.Tensor, probs: Optional[torch.Tensor], actions: torch.Tensor, greedy: bool = False) -> None:
        """obs real [N,8] -> probs real [N,4] (optional), actions int8 [N]; sampled with the env's Philox stream."""
        env = self.env
        env._check_tensor(obs, (env.n, 8), env.dtype, "obs")
        env._check_tensor(actions, (env.n,), torch.int8, "actions")
        if probs is not None:
            env._check_tensor(probs, (env.n, 4), env.dtype, "probs")
        _capi.check(self._lib.ah_actor_sample(env._h, self.weights.data_ptr(), obs.data_ptr(),
                                              None if probs is None else probs.data_ptr(), actions.data_ptr(), int(greedy),
                                              env._stream()), "ah_actor_sample", env._h)


class RolloutCollector:
    """Collects ``T`` frames from ``env`` under ``policy`` into preallocated device tensors.

    obs      real [T+1, N, 8]   obs[t] is the policy input of frame t (agent.py:46); obs[T] bootstraps the value
    actions  int8 [T, N]
    rewards  real [T, N], dones u8 [T, N], scores i8 [T, N], game_over u8 [T, N]
    probs    real [T, N, A]     the policy's action distribution at frame t (PPO's ``old_prediction``)
    """

    def __init__(self, env: BatchedAirHockey, T: int, policy: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 action_size: int = 4, use_graph: bool = True, seed: int = 0, fused_actor: bool = True,
                 single_launch: Optional[bool] = None, device_policy=None, device_select: bool = True,
                 auto_export: bool = True):
        """policy:        any callable obs [N,8] -> action probabilities [N,A] (a torch module); default: ``ActorMLP``
        device_policy: a ``policy.DevicePolicy`` exported from a torch module: the WHOLE T-frame closed loop (network,
                       exploration strategy, physics, rewards, resets) runs in ONE launch (AH_ACT_POLICY_MLP)
        auto_export:   a torch ``policy`` that is a small dense network (``policy.export_if_dense``: a Linear / ReLU / Tanh /
                       BatchNorm chain, or a module wrapping one, verified numerically against the module on the envs'
                       observations) is exported and run as a ``device_policy``; its parameters are re-exported at the
                       next ``collect()`` whenever an optimiser has written them.  ``single_launch=False`` opts out
        device_select: for a torch ``policy`` that cannot be exported, sample the action from its probabilities with
                       ``ah_policy_act`` (Philox, on the device) instead of ``torch.multinomial`` + dtype copies"""
        self.env, self.T, self.A = env, int(T), int(action_size)
        n, dt, dev = env.n, env.dtype, env.device
        self.auto_exported = False
        if device_policy is None and policy is None:
            policy = ActorMLP(8, action_size).to(device=dev, dtype=dt)
        if (device_policy is None and auto_export and single_launch is not False and isinstance(policy, nn.Module)
                and not (fused_actor and isinstance(policy, ActorMLP))):
            from .policy import export_if_dense
            device_policy = export_if_dense(policy, env)
            self.auto_exported = device_policy is not None
        self.device_policy = device_policy
        if device_policy is not None:
            if device_policy.continuous:
                raise ValueError("RolloutCollector collects discrete-action rollouts")
            self.A = action_size = device_policy.n_actions
            policy = policy if policy is not None else device_policy.module
        self.policy = policy
        self._selector = None
        if device_select and device_policy is None:
            from .policy import DevicePolicy
            self._selector = DevicePolicy(None, dev, head="values", select="sample", n_actions=action_size)
        # the reference-shaped actor runs fused with the sampling in one kernel; any other callable goes through torch
        self.fused = (FusedActor(self.policy, env)
                      if (fused_actor and device_policy is None and isinstance(self.policy, ActorMLP)) else None)
        self.obs = torch.zeros(T + 1, n, 8, dtype=dt, device=dev)
        self.actions = torch.zeros(T, n, dtype=torch.int8, device=dev)
        self.rewards = torch.zeros(T, n, dtype=dt, device=dev)
        self.dones = torch.zeros(T, n, dtype=torch.uint8, device=dev)
        self.scores = torch.zeros(T, n, dtype=torch.int8, device=dev)
        self.game_over = torch.zeros(T, n, dtype=torch.uint8, device=dev)
        self.probs = torch.zeros(T, n, action_size, dtype=dt, device=dev)
        # views handed to the kernel: it writes frame t's outputs directly into row t of the rollout tensors
        self._outs = [StepOutputs(reward=self.rewards[t:t + 1], done=self.dones[t:t + 1], score=self.scores[t:t + 1],
                                  game_over=self.game_over[t:t + 1], obs_next=self.obs[t + 1]) for t in range(T)]
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        # single_launch: the whole T-frame closed loop (actor forward, sampling, physics, reward, resets) runs in ONE
        # kernel launch with the env state in registers (BatchedAirHockey.step(policy_weights=...)); needs the fused actor
        if single_launch is None:                 # default: the fastest path the policy allows
            single_launch = self.fused is not None
        self.single_launch = bool((single_launch and self.fused is not None) or device_policy is not None)
        self._out_all = StepOutputs(reward=self.rewards, done=self.dones, score=self.scores, game_over=self.game_over,
                                    obs_next=self.obs[1:], actions=self.actions, probs=self.probs)
        self.use_graph = use_graph and not self.single_launch
        self._graph = None
        self.env.get_state("robot", out=self.obs[0])

    def _loop(self, collect_stats: bool = True) -> None:
        with torch.no_grad():
            for t in range(self.T):
                if self.fused is not None:
                    self.fused.sample(self.obs[t], self.probs[t], self.actions[t])
                    self.env.step(self.actions[t], K=1, out=self._outs[t], collect_stats=collect_stats)
                    continue
                p = self.policy(self.obs[t])
                self.probs[t].copy_(p)
                if self._selector is not None:                   # SoftmaxPolicy on the device, straight into the rollout
                    self.env.policy_act(self._selector, values=self.probs[t], actions=self.actions[t])
                else:
                    a = torch.multinomial(p.float(), 1, generator=self.gen).squeeze(1)
                    self.actions[t].copy_(a)                     # int64 -> int8, into the rollout tensor
                self.env.step(self.actions[t], K=1, out=self._outs[t], collect_stats=collect_stats)

    def _capture(self) -> None:
        dev = self.env.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        saved = self.env.state_dict()
        with torch.cuda.stream(side):                            # warm-up outside the capture (lazy init of kernels)
            self._loop(collect_stats=False)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.env.load_state_dict(saved)                          # the warm-up must not advance the games
        self.env.get_state("robot", out=self.obs[0])
        g = torch.cuda.CUDAGraph()
        g.register_generator_state(self.gen)
        # thread_local: other threads (NCCL's watchdog at world_size > 1) may issue CUDA calls during the capture
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._loop()
        self._graph = g
        self.env.load_state_dict(saved)                          # the capture pass itself ran nothing, but be explicit
        self.env.get_state("robot", out=self.obs[0])

    def collect(self) -> Dict[str, torch.Tensor]:
        """Run T frames.  The previous rollout's last observation becomes obs[0]."""
        if self.single_launch:
            if self._ran:
                self.obs[0].copy_(self.obs[self.T])
            if self.device_policy is not None:
                if self.auto_exported:
                    self.device_policy.refresh_if_changed()      # the torch module stays the owner of the parameters
                self.env.step(None, K=self.T, out=self._out_all, policy=self.device_policy)
            else:
                self.env.step(None, K=self.T, out=self._out_all, policy_weights=self.fused.weights)
        elif self.use_graph:
            if self._graph is None:
                self._capture()
            else:
                self.obs[0].copy_(self.obs[self.T])
            self._graph.replay()
        else:
            if self._ran:
                self.obs[0].copy_(self.obs[self.T])
            self._loop()
        self._ran = True
        return {"obs": self.obs, "actions": self.actions, "rewards": self.rewards, "dones": self.dones,
                "scores": self.scores, "game_over": self.game_over, "probs": self.probs}

    _ran = False
