"""Device policies: a small dense network + an exploration strategy, evaluated per env INSIDE the kernels
(``include/ah_b200.h``: ``ah_policy``, ``ah_policy_act``, ``AH_ACT_POLICY_MLP``).

What it replaces in the reference, every frame and on the host (paths relative to the reference root):
    Agent.get_action -> serialize_state -> model.predict          lib/agents/agent.py:42-48, lib/agents/*/.py
    SoftmaxPolicy / EpsilonGreedy / BoltzmannQPolicy / MaxBoltzmannQPolicy     lib/exploration/strategies.py:7-123
    GaussianWhiteNoiseProcess / OrnsteinUhlenbeckProcess                       lib/exploration/random.py:11-67

The torch module stays the owner of the parameters (a learner trains it with autograd); ``DevicePolicy.refresh()``
re-exports them -- BatchNorm folded into the neighbouring dense layer, weights transposed and padded the way the kernel
reads them.  Networks of the reference, as torch modules with the reference's layer shapes:
    ``ppo_actor_discrete``    8-4-BN-6-4-4 softmax      lib/agents/ppo/model.py:56-74
    ``ppo_actor_continuous``  8-6-6-4-2 tanh/linear      lib/agents/ppo/model.py:84-110
    ``a2c_actor``             8-8-BN-4 softmax            lib/agents/a2c/model.py:50-60   (= the DDQN net, ddqn/model.py:39-53)
    ``DuelingQNet``           8-32-(8|8)-(1|4)            lib/agents/dueling/model.py:50-66
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _capi

HEADS = {"softmax": _capi.AH_HEAD_SOFTMAX, "values": _capi.AH_HEAD_VALUES, "dueling": _capi.AH_HEAD_DUELING,
         "continuous": _capi.AH_HEAD_CONTINUOUS}
SELECTS = {"sample": _capi.AH_SELECT_SAMPLE, "greedy": _capi.AH_SELECT_GREEDY, "eps_greedy": _capi.AH_SELECT_EPS_GREEDY,
           "boltzmann": _capi.AH_SELECT_BOLTZMANN, "max_boltzmann": _capi.AH_SELECT_MAX_BOLTZMANN,
           "gaussian": _capi.AH_SELECT_GAUSSIAN, "ou": _capi.AH_SELECT_OU}
_ACTS = {nn.ReLU: _capi.AH_ACTV_RELU, nn.Tanh: _capi.AH_ACTV_TANH, nn.Identity: _capi.AH_ACTV_LINEAR}


def _normal_(m: nn.Linear, std: float = 0.05) -> None:        # Keras kernel_initializer="normal" is N(0, 0.05)
    nn.init.normal_(m.weight, std=std)
    nn.init.zeros_(m.bias)


def _bn(width: int) -> nn.BatchNorm1d:                        # Keras BatchNormalization defaults: eps 1e-3, momentum 0.99
    return nn.BatchNorm1d(width, eps=1e-3, momentum=0.01)


def ppo_actor_discrete(action_size: int = 4) -> nn.Sequential:
    net = nn.Sequential(nn.Linear(8, 4), nn.ReLU(), _bn(4), nn.Linear(4, 6), nn.ReLU(), nn.Linear(6, 4), nn.ReLU(),
                        nn.Linear(4, action_size), nn.Softmax(dim=-1))
    for m in net:
        if isinstance(m, nn.Linear):
            _normal_(m)
    return net.eval()


def ppo_actor_continuous(action_size: int = 2) -> nn.Sequential:
    net = nn.Sequential(nn.Linear(8, 6), nn.Tanh(), nn.Linear(6, 6), nn.Tanh(), nn.Linear(6, 4), nn.Tanh(),
                        nn.Linear(4, action_size))
    for m in list(net)[:-1]:
        if isinstance(m, nn.Linear):
            _normal_(m)
    nn.init.uniform_(net[-1].weight, -0.05, 0.05)             # kernel_initializer="random_uniform"
    nn.init.zeros_(net[-1].bias)
    return net.eval()


def a2c_actor(action_size: int = 4) -> nn.Sequential:
    net = nn.Sequential(nn.Linear(8, 8), nn.ReLU(), _bn(8), nn.Linear(8, action_size), nn.Softmax(dim=-1))
    for m in net:
        if isinstance(m, nn.Linear):
            _normal_(m)
    return net.eval()


class DuelingQNet(nn.Module):
    """Dense(32, relu) -> {V: Dense(8, relu), Dense(1)} + {A: Dense(8, relu), Dense(A)} -> q = v + a - mean(a)
    (lib/agents/dueling/model.py:50-66)."""

    def __init__(self, action_size: int = 4, state_size: int = 8):
        super().__init__()
        self.trunk = nn.Linear(state_size, 32)
        self.v1, self.v2 = nn.Linear(32, state_size), nn.Linear(state_size, 1)
        self.a1, self.a2 = nn.Linear(32, state_size), nn.Linear(state_size, action_size)
        for m in (self.trunk, self.v1, self.a1):
            _normal_(m)
        for m in (self.v2, self.a2):
            nn.init.uniform_(m.weight, -0.05, 0.05)
            nn.init.zeros_(m.bias)
        self.action_size = action_size

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = torch.relu(self.trunk(x))
        v = self.v2(torch.relu(self.v1(h)))
        a = self.a2(torch.relu(self.a1(h)))
        return v + a - a.mean(dim=-1, keepdim=True)

    def dense_chain(self) -> List[Tuple[torch.Tensor, torch.Tensor, int]]:
        """The two towers as ONE block-structured dense chain: 8-32, 32-16 (V | A hidden), 16-(1 + A) (block diagonal)."""
        hv = self.v1.out_features
        w2 = torch.cat([self.v1.weight, self.a1.weight], 0)
        b2 = torch.cat([self.v1.bias, self.a1.bias], 0)
        A = self.action_size
        w3 = torch.zeros(1 + A, w2.shape[0], dtype=w2.dtype, device=w2.device)
        w3[0, :hv] = self.v2.weight[0]
        w3[1:, hv:] = self.a2.weight
        b3 = torch.cat([self.v2.bias, self.a2.bias], 0)
        return [(self.trunk.weight, self.trunk.bias, _capi.AH_ACTV_RELU), (w2, b2, _capi.AH_ACTV_RELU),
                (w3, b3, _capi.AH_ACTV_LINEAR)]


def dense_chain_of(module: nn.Module) -> Tuple[List[Tuple[torch.Tensor, torch.Tensor, int]], Optional[str]]:
    """``nn.Sequential`` of Linear / ReLU / Tanh / BatchNorm1d (eval) / Flatten / Softmax -> [(W[out,in], b[out], act)]
    with every BatchNorm folded into a neighbouring Linear, and the head implied by a trailing Softmax."""
    if isinstance(module, DuelingQNet):
        return module.dense_chain(), "dueling"
    if not isinstance(module, nn.Sequential):
        raise TypeError("export needs an nn.Sequential of Linear/ReLU/Tanh/BatchNorm1d/Softmax layers or a DuelingQNet")
    layers: List[List] = []            # [W, b, act]
    pending = None                     # affine (scale, shift) waiting to be folded into the NEXT linear's input
    head = None
    mods = [m for m in module if not isinstance(m, (nn.Flatten, nn.Dropout))]
    for k, m in enumerate(mods):
        if isinstance(m, nn.Linear):
            W, b = m.weight, m.bias if m.bias is not None else torch.zeros(m.out_features, device=m.weight.device)
            if pending is not None:    # y = W (s x + t) + b
                s, t = pending
                b = b + W @ t
                W = W * s[None, :]
                pending = None
            layers.append([W, b, _capi.AH_ACTV_LINEAR])
        elif type(m) in _ACTS:
            if not layers or pending is not None:
                raise ValueError("an activation must follow a Linear layer")
            if layers[-1][2] != _capi.AH_ACTV_LINEAR:
                raise ValueError("two activations in a row")
            layers[-1][2] = _ACTS[type(m)]
        elif isinstance(m, nn.BatchNorm1d):
            if m.training:
                raise ValueError("BatchNorm must be in eval mode to be folded (call module.eval())")
            s = m.weight / torch.sqrt(m.running_var + m.eps) if m.affine else 1.0 / torch.sqrt(m.running_var + m.eps)
            t = (m.bias if m.affine else 0.0) - m.running_mean * s
            if layers and layers[-1][2] == _capi.AH_ACTV_LINEAR and pending is None:
                layers[-1][0] = layers[-1][0] * s[:, None]          # BN right after a Linear: fold into it
                layers[-1][1] = layers[-1][1] * s + t
            else:
                pending = (s, t) if pending is None else (pending[0] * s, pending[1] * s + t)
        elif isinstance(m, nn.Softmax):
            if k != len(mods) - 1:
                raise ValueError("Softmax is only supported as the last layer")
            head = "softmax"
        else:
            raise TypeError(f"unsupported layer {type(m).__name__}")
    if pending is not None:
        raise ValueError("a trailing BatchNorm cannot be folded")
    return [tuple(l) for l in layers], head


class DevicePolicy:
    """A torch policy exported for the kernels, with its exploration strategy.

    head:   "softmax" | "values" | "dueling" | "continuous"   (default: what the module implies)
    select: "sample" | "greedy" | "eps_greedy" | "boltzmann" | "max_boltzmann" | "gaussian" | "ou"
    The keyword defaults are the reference's (strategies.py:31-37,66-70,92-101; random.py:32,46)."""

    def __init__(self, module: Optional[nn.Module], device, head: Optional[str] = None, select: str = "sample",
                 n_actions: Optional[int] = None, n_envs: Optional[int] = None, *, epsilon: float = 1.0,
                 initial_epsilon: float = 1.0, final_epsilon: float = 0.001, observe: int = 2000, explore: int = 50000,
                 tau: float = 1.0, clip: Sequence[float] = (-500.0, 500.0), mu: float = 0.0, sigma: float = 1.0,
                 sigma_min: Optional[float] = None, n_steps_annealing: int = 1000, theta: float = 0.15, dt: float = 1e-2):
        self.module, self.device = module, torch.device(device)
        implied = None
        if module is not None:
            _, implied = dense_chain_of(module)
        self.head = head or implied or "values"
        self.select = select
        if self.head == "continuous":
            n_actions = 2
        self.c = _capi.AhPolicy()
        self.c.head, self.c.select = HEADS[self.head], SELECTS[select]
        self.c.epsilon, self.c.initial_epsilon, self.c.final_epsilon = epsilon, initial_epsilon, final_epsilon
        self.c.observe, self.c.explore = int(observe), int(explore)
        self.c.tau, self.c.clip_lo, self.c.clip_hi = tau, clip[0], clip[1]
        self.c.mu, self.c.sigma, self.c.n_steps_annealing = mu, sigma, int(n_steps_annealing)
        self.c.sigma_min = -1.0 if sigma_min is None else sigma_min
        self.c.theta, self.c.dt = theta, dt
        self.weights: Optional[torch.Tensor] = None
        self.noise_state: Optional[torch.Tensor] = None
        self.source, self._seen_version = module, None        # export_if_dense: the module whose parameters are watched
        if select == "ou":                                   # OrnsteinUhlenbeckProcess.reset_states (random.py:66-67)
            if n_envs is None:
                raise ValueError("the OU process keeps one state per env: pass n_envs")
            self.noise_state = (mu + sigma * torch.randn(n_envs, 2, device=self.device)).float().contiguous()
            self.c.noise_state = self.noise_state.data_ptr()
        self._n_actions = n_actions
        self.refresh()

    @torch.no_grad()
    def refresh(self) -> None:
        """Re-export the module's current parameters (call after every optimiser step that should reach the envs)."""
        if self.module is None:                              # selection only: head values come from the caller's network
            if self._n_actions is None:
                raise ValueError("n_actions is required without a module")
            self.c.n_actions = int(self._n_actions)
            return
        chain, _ = dense_chain_of(self.module)
        if not 1 <= len(chain) <= _capi.AH_POLICY_MAX_LAYERS:
            raise ValueError(f"1..{_capi.AH_POLICY_MAX_LAYERS} dense layers are supported, got {len(chain)}")
        flat, widths = [], [chain[0][0].shape[1]]
        for l, (W, b, act) in enumerate(chain):
            out_f, in_f = W.shape
            if in_f != widths[-1] or out_f > _capi.AH_POLICY_MAX_WIDTH:
                raise ValueError("layer shapes do not chain, or a layer is wider than %d" % _capi.AH_POLICY_MAX_WIDTH)
            outp = (out_f + 3) // 4 * 4
            Wt = torch.zeros(in_f, outp, dtype=torch.float32, device=W.device)
            Wt[:, :out_f] = W.t().float()
            bp = torch.zeros(outp, dtype=torch.float32, device=W.device)
            bp[:out_f] = b.float()
            flat += [Wt.flatten(), bp]
            widths.append(out_f)
            self.c.activation[l] = act
        if widths[0] != 8:
            raise ValueError("the policy input is the 8-number observation")
        last = widths[-1]
        n_actions = {"dueling": last - 1, "continuous": 2}.get(self.head, last)
        if self._n_actions is not None and self.head != "continuous" and n_actions != self._n_actions:
            raise ValueError("n_actions does not match the network's last layer")
        self.c.n_layers, self.c.n_actions = len(chain), n_actions
        for l, w in enumerate(widths):
            self.c.width[l] = w
        packed = torch.cat(flat).to(self.device)
        if self.weights is None or self.weights.numel() != packed.numel():
            self.weights = packed.contiguous()
        else:
            self.weights.copy_(packed)                       # same storage: captured graphs keep seeing the new weights
        self.c.weights = self.weights.data_ptr()

    def refresh_if_changed(self) -> bool:
        """``refresh()`` only if a parameter / buffer of the source module was written since the last export (torch bumps a
        tensor's version counter on every in-place write, which is how optimisers update)."""
        if self.source is None:
            return False
        v = _param_version(self.source)
        if v == self._seen_version:
            return False
        self.refresh()
        self._seen_version = v
        return True

    @property
    def n_actions(self) -> int:
        return int(self.c.n_actions)

    @property
    def continuous(self) -> bool:
        return self.head == "continuous"

    def ref(self):
        return C.byref(self.c)


def _param_version(module: nn.Module) -> int:
    """Changes whenever an optimiser (or anyone) writes a parameter or buffer of ``module`` in place."""
    return sum(int(t._version) for t in list(module.parameters()) + list(module.buffers()))


def export_if_dense(policy, env, select: str = "sample", rtol: float = 1e-4, atol: float = 2e-5) -> Optional[DevicePolicy]:
    """Try to export a torch policy ``obs [N,8] -> action probabilities / values [N,A]`` for the kernels, and VERIFY the
    export: the device network is evaluated on the envs' current observations (``ah_policy_act``) and must reproduce
    ``policy(obs)``.  Candidates: the module itself when it is a dense chain (``dense_chain_of``), else its only
    ``nn.Sequential`` child (a wrapper whose ``forward`` is the chain, or a softmax over it -- ``rollout.ActorMLP``,
    ``learners.A2CActor``).  Returns None when the policy is not a module, has no exportable chain, is too wide / deep
    for the kernel, or does anything in ``forward`` the export cannot see -- callers then evaluate it with torch."""
    if not isinstance(policy, nn.Module) or policy.training and any(isinstance(m, nn.BatchNorm1d) for m in policy.modules()):
        return None
    candidates = []
    try:
        _, implied = dense_chain_of(policy)
        candidates.append((policy, implied or "values"))
    except (TypeError, ValueError):
        seqs = [m for m in policy.children() if isinstance(m, nn.Sequential)]
        if len(seqs) == 1 and len(list(policy.children())) == 1:
            candidates += [(seqs[0], "softmax"), (seqs[0], "values")]
    if not candidates:
        return None
    # two probes: the envs' current observations, and the same scaled down (networks saturate on raw pixel coordinates,
    # and a saturated softmax hides what happened to the logits)
    real = env.get_state("robot")
    probes = [real, (real * 0.01).contiguous()]
    try:
        with torch.no_grad():
            wants = [policy(o) for o in probes]
    except RuntimeError:                           # e.g. a float32 module on a float64 env: not ours to run
        return None
    if any(not torch.is_tensor(w) or w.dim() != 2 or w.shape[0] != env.n for w in wants):
        return None
    for module, head in candidates:
        try:
            dp = DevicePolicy(module, env.device, head=head, select=select)
        except (TypeError, ValueError):
            continue
        if dp.continuous or dp.n_actions != wants[0].shape[1]:
            continue
        got = torch.empty(env.n, dp.n_actions, dtype=env.dtype, device=env.device)
        same = True
        for o, want in zip(probes, wants):
            env.policy_act(dp, obs=o, values_out=got)
            same = same and torch.allclose(got.to(want.dtype), want, rtol=rtol, atol=atol)
        if same:
            dp.source, dp._seen_version = policy, _param_version(policy)
            return dp
    return None
