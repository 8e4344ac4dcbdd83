"""Host-side mirror of gold's `pkg/v1/agent/her` (DQN + Hindsight Experience Replay) over the
same C ABI: SURVEY.md §8(f) rank 1.  The learn step is the deepq one on `state‖goal` inputs
(reference: pkg/v1/agent/her/agent.go:129-189), so it runs on the same kernels; what differs is
host logic: a slice-backed memory (her/memory.go:75-125), target sync every
`UpdateTargetEpisodes` Learn calls (agent.go:180-201: `episodes` is incremented per Learn),
epsilon decayed per Action (agent.go:205), and Hindsight relabelling (agent.go:246-263).
"""
import random

import numpy as np

from . import _capi as G
from . import synth
from .deepq import (AdamSolver, DecaySchedule, MSE, PolicyConfig, Sequential, _dims_from_layers,  # noqa: F401
                    default_decay_schedule, default_policy_config, handle_options)


class Event:
    """her.Event (her/memory.go:15-32): state, goal, action and the outcome fields."""

    __slots__ = ("state", "goal", "action", "observation", "reward", "done", "i")

    def __init__(self, state, goal, action, observation, reward, done):
        self.state = np.asarray(state, np.float32).reshape(-1)
        self.goal = np.asarray(goal, np.float32).reshape(-1)
        self.action = int(action)
        self.observation = np.asarray(observation, np.float32).reshape(-1)
        self.reward = float(np.float32(reward))
        self.done = bool(done)
        self.i = -1

    def copy(self):
        return Event(self.state.copy(), self.goal.copy(), self.action, self.observation.copy(), self.reward, self.done)


class Hyperparameters:
    """her.Hyperparameters and defaults (her/agent.go:43-63)."""

    def __init__(self, gamma=0.9, epsilon=None, update_target_episodes=50):
        self.gamma = gamma
        self.epsilon = epsilon if epsilon is not None else default_decay_schedule()
        self.update_target_episodes = update_target_episodes


class AgentConfig:
    def __init__(self, hyperparameters=None, policy_config=None, successful_reward=0.0, memory_size=int(1e5),
                 device=0, seed=None):
        self.hyperparameters = hyperparameters or Hyperparameters()
        self.policy_config = policy_config or default_policy_config()
        self.successful_reward = successful_reward
        self.memory_size = memory_size
        self.device, self.seed = device, seed


class Agent:
    """her.Agent: Learn / Action / Remember / Hindsight."""

    def __init__(self, config, state_dim, goal_dim, n_actions, sample_action=None):
        c = config or AgentConfig()
        hp, pc = c.hyperparameters, c.policy_config
        self.state_dim, self.goal_dim = int(state_dim), int(goal_dim)
        dims = _dims_from_layers(pc.layer_builder(self.state_dim + self.goal_dim, n_actions))
        # loss / solver mapping and validation shared with deepq.Agent (PseudoHuber and RMSProp included)
        self._h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3],
                           batch_size=pc.batch_size, replay_capacity=c.memory_size, gamma=hp.gamma,
                           device=c.device, **handle_options(pc))
        seed = c.seed if c.seed is not None else random.randrange(1 << 30)
        self._h.set_params(G.NET_ONLINE, synth.glorot_params(dims, seed))
        self._h.set_params(G.NET_TARGET, synth.glorot_params(dims, seed + 1))
        self.policy = Sequential("online", self._h, G.NET_ONLINE)
        self.target_policy = Sequential("target", self._h, G.NET_TARGET)
        self.hyperparameters = hp
        self.epsilon_schedule = hp.epsilon
        self.epsilon = hp.epsilon.initial()
        self.batch_size = pc.batch_size
        self.successful_reward = c.successful_reward
        self.steps = 0
        self.episodes = 0
        self._pending = []
        self._len = 0
        self._rng = random.Random(seed)
        self._sample_action = sample_action or (lambda: self._rng.randrange(n_actions))

    def memory_len(self):
        return self._len

    def remember(self, *events):
        """Agent.Remember -> Memory.Remember (her/memory.go:89-94): append in arrival order."""
        self._pending.extend(events)
        self._len = min(self._len + len(events), int(self._h.cfg.replay_capacity))

    def _flush(self):
        if self._pending:
            ev = self._pending
            s = np.stack([np.concatenate([e.state, e.goal]) for e in ev])          # State.Concat(1, Goal)
            s2 = np.stack([np.concatenate([e.observation, e.goal]) for e in ev])   # Observation.Concat(1, Goal)
            self._h.replay_push(s, np.array([e.action for e in ev], np.int32),
                                np.array([e.reward for e in ev], np.float32), s2,
                                np.array([e.done for e in ev], np.uint8))
            self._pending = []

    def learn(self, indices=None):
        """Agent.Learn (her/agent.go:129-189)."""
        self._flush()
        if self._len < self.batch_size:
            return None
        if indices is None:
            self._h.learn(None, seed=self._rng.getrandbits(63))
        else:
            # her.Memory indexes OLDEST first (events[value], her/memory.go:75-117); the device ring counts
            # from the newest event (deepq's deque.At): convert
            idx = np.asarray(indices, np.int64)
            self._h.learn((self._len - 1 - idx).astype(np.int32))
        self.episodes += 1                                                           # agent.go:180
        if self.episodes % self.hyperparameters.update_target_episodes == 0:         # agent.go:192
            self._h.sync_target()
        return None

    def action(self, state, goal):
        """Agent.Action (her/agent.go:203-219): epsilon decays after every call."""
        self.steps += 1
        try:
            if self._rng.random() < self.epsilon:
                return self._sample_action()
            sg = np.concatenate([np.asarray(state, np.float32).reshape(-1), np.asarray(goal, np.float32).reshape(-1)])
            return int(self._h.greedy_action(sg.reshape(1, -1))[0])
        finally:
            self.epsilon = self.epsilon_schedule.value()

    def hindsight(self, episode_events):
        """Agent.Hindsight (her/agent.go:246-263): replay the episode with the goal replaced by
        the state actually reached; the last event becomes a success.  One Learn per event."""
        alt = [e.copy() for e in episode_events]
        final = alt[-1]
        for k, e in enumerate(alt):
            e.goal = final.observation.copy()
            if k == len(alt) - 1:
                e.reward = float(np.float32(self.successful_reward))
                e.done = True
            self.remember(e)
            self.learn()

    def last_loss(self):
        return float(self._h.last_loss()[0])

    def close(self):
        self._h.close()
