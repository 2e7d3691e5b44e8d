"""Single-game shim with the reference's object protocol, backed by a 1-env ``BatchedAirHockey`` on the GPU.

It exists so that ``Agent``-style code written against the reference (lib/agents/agent.py:20,36,46,61,67,70 and
game.py:39,93-106,142,166,169-177) runs unchanged: same class and attribute names (``env.puck.x``,
``env.robot.location()``, ``env.robot_score`` ...), same verbs, same error (``InvalidAgentError``).  Every
attribute access is a device round-trip, so this is a compatibility surface, not a fast path -- throughput comes
from ``BatchedAirHockey.step``.

Mirrors: environment/airhockey.py:24-187, environment/puck.py:120-133, environment/mallet.py:86-99,
environment/table.py:8-15, environment/goal.py:6-18, lib/rewards.py:15-36,118-142, lib/types.py:6.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Optional, Tuple, Union

import torch

from .env import BatchedAirHockey
from .exceptions import InvalidAgentError

State = namedtuple("state", ["agent_location", "puck_location", "agent_velocity", "puck_velocity"])   # lib/types.py:6
Observation = namedtuple("observation", ["state", "action", "reward", "done", "new_state"])          # lib/types.py:7


class Table:                                   # environment/table.py:8-15
    size = (900, 480)
    midpoints = (450, 240)
    left_wall = 13
    right_wall = 882


class Goal:                                    # environment/goal.py:6-11 (the predicate runs in the kernel)
    def __init__(self, x: int, y: int):
        self.x, self.y = x, y


class _Body:
    """A row of one of the SoA state tensors, with the reference's attribute names."""

    def __init__(self, env: "AirHockey", tensor: torch.Tensor, name: str, radius: int, mass: float):
        self._env, self._t, self.name, self.radius, self.mass, self.imass = env, tensor, name, radius, mass, 1.0 / mass

    def _get(self, j: int) -> float:
        return float(self._t[0, j].item())

    def _set(self, j: int, v) -> None:
        self._t[0, j] = float(v)

    x = property(lambda s: s._get(0), lambda s, v: s._set(0, v))
    y = property(lambda s: s._get(1), lambda s, v: s._set(1, v))
    dx = property(lambda s: s._get(2), lambda s, v: s._set(2, v))
    dy = property(lambda s: s._get(3), lambda s, v: s._set(3, v))

    def location(self) -> Tuple[int, int]:     # puck.py:120-123, mallet.py:86-89: int() truncation
        return int(self.x), int(self.y)

    def velocity(self) -> Tuple[float, float]:
        return self.dx, self.dy


class AirHockey:
    """Drop-in for ``environment.AirHockey`` (one game)."""

    def __init__(self, dtype: torch.dtype = torch.float64, device: Union[str, torch.device] = "cuda", seed: int = 0,
                 reset_table: Optional[torch.Tensor] = None):
        self.batched = BatchedAirHockey(1, dtype=dtype, device=device, seed=seed)
        b = self.batched
        self.table = Table()
        self.left_goal = Goal(x=0, y=self.table.midpoints[1])                    # airhockey.py:35
        self.right_goal = Goal(x=self.table.size[0], y=self.table.midpoints[1])  # airhockey.py:36
        self.puck = _Body(self, b.puck, "puck", 15, 0.1)
        self.robot = _Body(self, b.robot, "robot", 20, 0.5)
        self.opponent = _Body(self, b.opponent, "opponent", 20, 0.5)
        self.mallets = [self.robot, self.opponent]                               # airhockey.py:54
        self.step_size = 10                                                      # airhockey.py:63
        self.ticks_to_friction = 60                                              # airhockey.py:60 (unused there too)
        self.reset_table = reset_table          # optional float64 [1, R, 2]: replaces the np.random draws of Puck.reset

    # scores live in the packed meta word
    @property
    def robot_score(self) -> int:
        return int(self.batched.robot_score[0].item())

    @robot_score.setter
    def robot_score(self, v: int) -> None:
        self.batched.scores[0, 0] = int(v)

    @property
    def opponent_score(self) -> int:
        return int(self.batched.opponent_score[0].item())

    @opponent_score.setter
    def opponent_score(self, v: int) -> None:
        self.batched.scores[0, 1] = int(v)

    def update_state(self, action="", agent_name: str = "") -> None:             # airhockey.py:96-134
        b = self.batched
        if agent_name not in ("robot", "human", "computer"):
            raise InvalidAgentError(agent_name)
        if agent_name == "computer":
            return b.update_state(agent_name="computer")
        if isinstance(action, (tuple, list)):                                     # continuous / human location
            a = torch.tensor([[float(action[0]), float(action[1])]], dtype=b.dtype, device=b.device)
        elif isinstance(action, int) and not isinstance(action, bool):
            a = torch.tensor([action if -128 <= action <= 127 else -1], dtype=torch.int8, device=b.device)
        else:
            a = None                                                              # e.g. "": no isinstance branch matches
        b.update_state(a, agent_name)

    def get_state(self, agent_name: str = "robot") -> State:                      # airhockey.py:136-147
        o = self.batched.get_state(agent_name)[0].tolist()
        return State(agent_location=(int(o[0]), int(o[1])), puck_location=(int(o[2]), int(o[3])),
                     agent_velocity=(o[4], o[5]), puck_velocity=(o[6], o[7]))

    def update_score(self, score: int) -> None:                                   # airhockey.py:149-169
        s = torch.tensor([1 if score > 0 else (-1 if score < 0 else 0)], dtype=torch.int8, device=self.batched.device)
        self.batched.update_score(s, reset_table=self.reset_table)

    def reset(self, total: bool = False) -> None:                                 # airhockey.py:171-187
        self.batched.reset(total=total, reset_table=self.reset_table)


class Rewards:
    """Drop-in for ``lib.rewards.Rewards`` for the robot: ``tracker(env.puck, env.robot) -> (reward, score, done)``.
    ``prev_mallet_to_puck_distance`` lives in the environment's state (rewards.py:31,96-100)."""

    def __init__(self, name: str, left_goal: Goal, right_goal: Goal, table: Table):
        if name != "robot":
            raise NotImplementedError("the reference only ever tracks the robot's reward (game.py:50, agent.py:20)")
        self.agent_name, self.left_goal, self.right_goal, self.table = name, left_goal, right_goal, table

    def __call__(self, puck: _Body, mallet: _Body):
        reward, score, done = puck._env.batched.rewards()
        return int(reward[0].item()), int(score[0].item()), bool(done[0].item())
