"""Host-side mirror of gold's `pkg/v1/agent/deepq` API over the C ABI (include/goldq.h).

Same names, argument meaning and error behaviour as the reference package
(reference: pkg/v1/agent/deepq/{agent,memory,policy}.go), Python spelling:

    reference (Go)                         here
    -------------------------------------  ------------------------------------------
    deepq.NewAgent(cfg, env)               Agent(cfg, env)
    (*Agent).Learn() error                 Agent.learn()
    (*Agent).Action(state) (int, error)    Agent.action(state) -> int
    (*Agent).Remember(event)               Agent.remember(event)
    deepq.NewEvent / Event / Memory        new_event / Event / Memory
    Memory.Sample(n) ([]*Event, error)     Memory.sample(n) -> [Event]   (raises if Len < n)
    PolicyConfig / DefaultPolicyConfig     PolicyConfig / default_policy_config()
    goro model.Model (Predict, Fit, ...)   Sequential (predict, fit, fit_batch, ...)
    common.DecaySchedule                   DecaySchedule

The compute is entirely in libgoldq.so (CUDA); there is no CPU fallback here: the
constructors raise if the library or a B200 is missing.  The Go package a gold
maintainer would actually link (go/deepq, cgo) has the same structure; it cannot
be compiled in this image (no Go toolchain), this module is its executable twin.
"""
import collections
import math
import random

import numpy as np

from . import _capi as G
from . import synth


# --------------------------------------------------------------------------- common
class DecaySchedule:
    """common.DecaySchedule (reference: pkg/v1/common/schedule.go:77-141)."""

    def __init__(self, decay_rate=0.995, initial_value=1.0, min_value=0.01):
        self.decay_rate = float(np.float32(decay_rate))
        self.initial_value = float(np.float32(initial_value))
        self.min_value = float(np.float32(min_value))
        self.current_value = self.initial_value

    def value(self):
        """Value for the given step. Decays with each call (schedule.go:101-105)."""
        self.current_value *= self.decay_rate
        self.current_value = max(self.min_value, self.current_value)
        return float(np.float32(self.current_value))

    def initial(self):
        return float(np.float32(self.initial_value))


def default_decay_schedule(decay_rate=None):
    """common.DefaultDecaySchedule (schedule.go:130-141): 0.995 -> min 0.01 from 1.0."""
    s = DecaySchedule(0.995, 1.0, 0.01)
    s.decay_rate, s.initial_value, s.min_value, s.current_value = 0.995, 1.0, 0.01, 1.0   # float64 literals there
    if decay_rate is not None:
        s.decay_rate = float(np.float32(decay_rate))
    return s


# --------------------------------------------------------------------------- env (shapes only)
class Outcome:
    """envv1.Outcome (reference: pkg/v1/env/env.go:129-141)."""

    __slots__ = ("observation", "action", "reward", "done")

    def __init__(self, observation, action, reward, done):
        self.observation = np.asarray(observation, np.float32)
        self.action = int(action)
        self.reward = float(np.float32(reward))
        self.done = bool(done)


class ShapesEnv:
    """What NewAgent/MakePolicy read from *envv1.Env: the observation shape and the number
    of discrete actions (policy.go:75-79), plus SampleAction for exploration
    (env.go:161-168).  No gym server involved (BASELINE configs: "no Sphere env")."""

    def __init__(self, obs_dim, n_actions, seed=None):
        """obs_dim: an int (flat observations) or a (C, H, W) tuple (image observations of the Atari wrapper)."""
        self.obs_shape = [int(obs_dim)] if np.isscalar(obs_dim) else [int(v) for v in obs_dim]
        self.obs_dim, self.n_actions = int(np.prod(self.obs_shape)), int(n_actions)
        self._rng = random.Random(seed)

    def observation_space_shape(self):
        return list(self.obs_shape)

    def sample_action(self):
        return self._rng.randrange(self.n_actions)


# --------------------------------------------------------------------------- memory
class Event:
    """deepq.Event (memory.go:15-25): Outcome + the state and action that led to it."""

    __slots__ = ("state", "action", "outcome", "i")

    def __init__(self, state, action, outcome):
        self.state = np.asarray(state, np.float32)
        self.action = int(action)
        self.outcome = outcome
        self.i = -1

    # Go embeds *Outcome: event.Reward, event.Done, event.Observation
    @property
    def reward(self):
        return self.outcome.reward

    @property
    def done(self):
        return self.outcome.done

    @property
    def observation(self):
        return self.outcome.observation


def new_event(state, action, outcome):
    return Event(state, action, outcome)


class Memory:
    """deepq.Memory (memory.go:41-69).  PushFront makes an event logical index 0; the
    device ring (goldq_replay_push) follows the same convention.  A bounded host-side
    deque of Event objects is kept only so that sample() can hand Events back like the
    reference does; learn() never reads it."""

    def __init__(self, handle=None, host_events=4096):
        self._h = handle
        self._events = collections.deque(maxlen=host_events)
        self._len = 0
        self._pending = []
        self._rng = random.Random()

    def len(self):
        """Memory.Len(): events on the device ring plus those remembered since the last flush, never more
        than the ring holds (Remember pops the back past BufferSize, agent.go:226-228)."""
        if not self._h:
            return self._len
        return min(self._h.replay_len() + len(self._pending), int(self._h.cfg.replay_capacity))

    def __len__(self):
        return self.len()

    def push_front(self, event):
        self._events.appendleft(event)
        self._len += 1
        if self._h:
            self._pending.append(event)

    def pop_back(self):
        # eviction past BufferSize is done by the device ring itself (replay_capacity)
        if not self._h:
            self._len -= 1
            if len(self._events) > self._len:
                self._events.pop()

    def flush(self):
        """Send the events remembered since the last flush to the device ring."""
        if self._h and self._pending:
            ev = self._pending
            s = np.stack([e.state.reshape(-1) for e in ev])
            s2 = np.stack([e.observation.reshape(-1) for e in ev])
            a = np.array([e.action for e in ev], np.int32)
            r = np.array([e.reward for e in ev], np.float32)
            d = np.array([e.done for e in ev], np.uint8)
            self._h.replay_push(s, a, r, s2, d)
            self._pending = []

    def at(self, i):
        return self._events[i]

    def sample_indices(self, batchsize):
        """The first `batchsize` entries of a permutation of Len (memory.go:59-61)."""
        n = self.len()
        if n < batchsize:
            raise ValueError(f"queue size {n} is less than batch size {batchsize}")
        return np.array(self._rng.sample(range(n), batchsize), np.int32)

    def sample(self, batchsize):
        """Memory.Sample: errors when Len < batchsize (memory.go:55-57), otherwise returns
        `batchsize` distinct events with event.i set (memory.go:62-64)."""
        n = self.len()
        if n < batchsize:
            raise ValueError(f"queue size {n} is less than batch size {batchsize}")
        # Events older than the host-side mirror exist only on the device: the Events handed back are drawn
        # from the mirrored range (the newest `host_events`).  Learn never goes through here -- it samples on
        # the device over the whole ring.
        m = min(n, len(self._events))
        if m < batchsize:
            raise IndexError(f"only {m} events are mirrored on the host (host_events) for a sample of {batchsize}")
        out = []
        for i in self._rng.sample(range(m), batchsize):
            e = self._events[int(i)]
            e.i = int(i)
            out.append(e)
        return out


# --------------------------------------------------------------------------- policy / model
class AdamSolver:
    """gorgonia.AdamSolver's settable fields (vendor/gorgonia.org/gorgonia/solvers.go:399-438)."""

    def __init__(self, learn_rate=0.001, eps=1e-8, beta1=0.9, beta2=0.999):
        self.eta, self.eps, self.beta1, self.beta2 = learn_rate, eps, beta1, beta2


class RMSPropSolver:
    """gorgonia.RMSPropSolver's settable fields (solvers.go:216-246): the solver of gold's Atari policy
    config (policy.go:41-47).  Runs on the generic CUDA path."""

    def __init__(self, learn_rate=0.001, eps=1e-8, rho=0.999):
        self.eta, self.eps, self.decay = learn_rate, eps, rho


MSE = "mse"  # modelv1.MSE (goro model/loss.go:22)


class PseudoHuberLoss:
    """model.PseudoHuberLoss (goro model/loss.go:90-111).  Reduced by Mean on the device; the reference's
    own Compute ends in Sum(.., 1) and cannot be differentiated (loss.go:94 "!blocked")."""

    def __init__(self, delta=1.0):
        self.delta = float(delta)


PSEUDO_HUBER = PseudoHuberLoss(1.0)  # model.PseudoHuber (loss.go:101-104)


class FC:
    """layer.FC config (goro layer/fc.go:13-40); activation None -> ReLU (fc.go:72-75)."""

    def __init__(self, input, output, activation=None):
        self.input, self.output, self.activation = int(input), int(output), activation


LINEAR = "linear"


class Conv2D:
    """layer.Conv2D (goro layer/conv2d.go:14-50): Input/Output channels, kernel Height x Width, Stride; the defaults the
    layer applies -- pad (1, 1), ReLU, no bias -- are the only ones the device path implements."""

    def __init__(self, input, output, width, height, stride=(1, 1), pad=(1, 1), activation=None):
        self.input, self.output, self.width, self.height = int(input), int(output), int(width), int(height)
        self.stride, self.pad, self.activation = tuple(stride), tuple(pad), activation


class Flatten:
    """layer.Flatten (goro layer/flatten.go): reshape (batch, C*H*W)."""


def default_atari_layer_builder(x_dim, y_dim):
    """deepq.DefaultAtariLayerBuilder (policy.go:62-71)."""
    return [Conv2D(1, 32, 8, 8, stride=(4, 4)), Conv2D(32, 64, 4, 4, stride=(2, 2)), Conv2D(64, 64, 3, 3, stride=(1, 1)),
            Flatten(), FC(6400, 512), FC(512, y_dim, activation=LINEAR)]


def default_fc_layer_builder(x_dim, y_dim):
    """deepq.DefaultFCLayerBuilder (policy.go:53-59)."""
    return [FC(x_dim, 24), FC(24, 24), FC(24, y_dim, activation=LINEAR)]


class PolicyConfig:
    """deepq.PolicyConfig (policy.go:14-29)."""

    def __init__(self, loss=MSE, optimizer=None, layer_builder=default_fc_layer_builder, batch_size=20, track=True):
        self.loss = loss
        self.optimizer = optimizer if optimizer is not None else AdamSolver(learn_rate=0.0005)
        self.layer_builder = layer_builder
        self.batch_size = batch_size
        self.track = track


def default_atari_policy_config():
    """deepq.DefaultAtariPolicyConfig (policy.go:41-47): MSE, RMSPropSolver with gorgonia's defaults
    (eta 0.001, rho 0.999, eps 1e-8; WithBatchSize(20) does nothing for RMSProp), the conv layer builder, batch 20."""
    return PolicyConfig(loss=MSE, optimizer=RMSPropSolver(), layer_builder=default_atari_layer_builder, batch_size=20)


def default_policy_config():
    """deepq.DefaultPolicyConfig (policy.go:32-38) — a fresh solver per call: the reference's
    package-level singleton shares one iter counter between agents."""
    return PolicyConfig()


class Sequential:
    """The slice of goro's model.Model a deepq policy uses (goro model/model.go:17-50),
    bound to one network (online or target) of a handle."""

    def __init__(self, name, handle, net, replica=0):
        self.name, self._h, self._net, self._replica = name, handle, net, replica

    def predict(self, x):
        """Sequential.Predict: one row (model.go:500-512); returns a fresh (1, A) array."""
        x = np.asarray(x, np.float32).reshape(1, -1)
        return self._h.predict(self._net, x, self._replica)

    def predict_batch(self, x):
        return self._h.predict(self._net, x, self._replica)

    def fit(self, x, y):
        """Sequential.Fit: one (1,D)/(1,A) pair (model.go:528-550)."""
        self._need_online()
        self._h.fit_batch(np.asarray(x, np.float32).reshape(1, -1), np.asarray(y, np.float32).reshape(1, -1))

    def fit_batch(self, x, y):
        """Sequential.FitBatch: exactly batch_size rows (model.go:552-573; io.go:138-153)."""
        self._need_online()
        self._h.fit_batch(x, y)

    def resize_batch(self, n):
        """Sequential.ResizeBatch (model.go:485-497): both networks of the agent share the scratch."""
        self._h.resize_batch(n)

    def learnables(self):
        """Flat [W1,b1,W2,b2,W3,b3] host copy of the parameters."""
        P = self._h.P
        return self._h.get_params(self._net)[self._replica * P:(self._replica + 1) * P]

    def clone_learnables_to(self, other):
        """Sequential.CloneLearnablesTo (model.go:606-636): online -> target."""
        if not (self._net == G.NET_ONLINE and other._net == G.NET_TARGET and other._h is self._h):
            raise ValueError("only online -> target of the same agent is supported")
        self._h.sync_target()

    def _need_online(self):
        if self._net != G.NET_ONLINE:
            raise ValueError("the target network is never fitted (agent.go:167 fits Policy only)")


def handle_options(pc):
    """PolicyConfig.Loss / .Optimizer -> goldq_config fields (shared by deepq.Agent and her.Agent).  Anything the
    CUDA path does not implement is an error: there is no CPU fallback."""
    if pc.loss != MSE and not isinstance(pc.loss, PseudoHuberLoss):
        raise ValueError("the CUDA path implements model.MSE and model.PseudoHuber; no CPU fallback exists")
    if isinstance(pc.optimizer, AdamSolver):
        opts = dict(solver=G.SOLVER_ADAM, lr=pc.optimizer.eta, beta1=pc.optimizer.beta1, beta2=pc.optimizer.beta2,
                    eps=pc.optimizer.eps)
    elif isinstance(pc.optimizer, RMSPropSolver):
        opts = dict(solver=G.SOLVER_RMSPROP, lr=pc.optimizer.eta, beta2=pc.optimizer.decay, eps=pc.optimizer.eps)
    else:
        raise ValueError("the CUDA path implements gorgonia.AdamSolver and RMSPropSolver; no CPU fallback exists")
    if isinstance(pc.loss, PseudoHuberLoss):
        opts.update(loss=G.LOSS_PSEUDO_HUBER, huber_delta=pc.loss.delta)
    else:
        opts.update(loss=G.LOSS_MSE)
    return opts


def _atari_from_layers(layers, obs_shape):
    """The conv stack the device implements is exactly DefaultAtariLayerBuilder's: three Conv2D (32x8x8/4, 64x4x4/2,
    64x3x3/1, pad 1, ReLU), Flatten, FC(F, 512, ReLU), FC(512, A, linear).  Returns (C, H, W, A) or None when the
    layers are not convolutional; anything else convolutional is an error (no CPU fallback)."""
    if not layers or not isinstance(layers[0], Conv2D):
        return None
    want = [(32, 8, 8, (4, 4)), (64, 4, 4, (2, 2)), (64, 3, 3, (1, 1))]
    ok = (len(layers) == 6 and all(isinstance(l, Conv2D) for l in layers[:3]) and isinstance(layers[3], Flatten)
          and isinstance(layers[4], FC) and isinstance(layers[5], FC))
    if ok:
        for l, (o, hh, ww, st) in zip(layers[:3], want):
            ok = ok and (l.output, l.height, l.width, l.stride) == (o, hh, ww, st) and l.pad == (1, 1) and l.activation in (None, "relu")
        ok = ok and layers[1].input == 32 and layers[2].input == 64 and layers[4].output == 512 and layers[5].input == 512
        ok = ok and layers[4].activation in (None, "relu") and layers[5].activation == LINEAR
    if not ok or len(obs_shape) != 3 or obs_shape[0] != layers[0].input:
        raise ValueError("the CUDA path implements deepq.DefaultAtariLayerBuilder's stack on (C, H, W) states only; "
                         "no CPU fallback exists for other convolutional graphs")
    C, H, W = (int(v) for v in obs_shape)
    h1, w1 = (H + 2 - 8) // 4 + 1, (W + 2 - 8) // 4 + 1
    h2, w2 = (h1 + 2 - 4) // 2 + 1, (w1 + 2 - 4) // 2 + 1
    if layers[4].input != 64 * h2 * w2:          # conv 3 keeps the size (3x3, stride 1, pad 1)
        raise ValueError(f"FC input {layers[4].input} does not match the {64 * h2 * w2} features of a {H}x{W} state")
    return C, H, W, layers[5].output


def _dims_from_layers(layers):
    if (len(layers) != 3 or layers[0].activation not in (None, "relu") or layers[1].activation not in (None, "relu")
            or layers[2].activation != LINEAR or layers[0].output != layers[1].input
            or layers[1].output != layers[2].input):
        raise ValueError("the CUDA path implements FC(relu) -> FC(relu) -> FC(linear) chains "
                         "(deepq.DefaultFCLayerBuilder); no CPU fallback exists for other graphs")
    return (layers[0].input, layers[0].output, layers[1].output, layers[2].output)


# --------------------------------------------------------------------------- agent
class Hyperparameters:
    """deepq.Hyperparameters (agent.go:44-58)."""

    def __init__(self, gamma=0.95, epsilon=None, update_target_steps=100, buffer_size=int(10e6)):
        self.gamma = gamma
        self.epsilon = epsilon if epsilon is not None else default_decay_schedule()
        self.update_target_steps = update_target_steps
        self.buffer_size = buffer_size


class AgentConfig:
    """deepq.AgentConfig (agent.go:68-78)."""

    def __init__(self, hyperparameters=None, policy_config=None, precision="fp32", device=0, seed=None):
        self.hyperparameters = hyperparameters or Hyperparameters()
        self.policy_config = policy_config or default_policy_config()
        self.precision = precision
        self.device = device
        self.seed = seed


class Agent:
    """deepq.Agent (agent.go:19-42): Learn / Action / Remember."""

    def __init__(self, config=None, env=None):
        c = config or AgentConfig()
        if env is None:
            raise ValueError("environment cannot be nil")            # agent.go:96-98
        hp, pc = c.hyperparameters, c.policy_config
        opts = handle_options(pc)
        obs_shape = tuple(env.observation_space_shape())
        layers = pc.layer_builder(obs_shape[0], env.n_actions)
        seed = c.seed if c.seed is not None else random.randrange(1 << 30)
        atari = _atari_from_layers(layers, obs_shape)
        if atari is not None:
            C, H, W, A = atari
            self._h = G.Handle(arch=G.ARCH_ATARI_CONV, in_channels=C, in_height=H, in_width=W, obs_dim=C * H * W, n_actions=A,
                               batch_size=pc.batch_size, replay_capacity=hp.buffer_size, gamma=hp.gamma, device=c.device,
                               **opts)
            self._h.set_params(G.NET_ONLINE, synth.atari_params(C, H, W, A, seed))
            self._h.set_params(G.NET_TARGET, synth.atari_params(C, H, W, A, seed + 1))
        else:
            dims = _dims_from_layers(layers)
            self._h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3],
                               batch_size=pc.batch_size, replay_capacity=hp.buffer_size, gamma=hp.gamma,
                               device=c.device, precision=G.PREC_BF16 if c.precision == "bf16" else G.PREC_FP32,
                               **opts)
            # two independently initialised networks, GlorotN(1) (agent.go:101-109; goro fc.go:76-81);
            # the target is first synced when steps % UpdateTargetSteps == 0
            self._h.set_params(G.NET_ONLINE, synth.glorot_params(dims, seed))
            self._h.set_params(G.NET_TARGET, synth.glorot_params(dims, seed + 1))
        self.hyperparameters = hp
        self.policy = Sequential("online", self._h, G.NET_ONLINE)
        self.target_policy = Sequential("target", self._h, G.NET_TARGET)
        self.epsilon_schedule = hp.epsilon
        self.env = env
        self.epsilon = hp.epsilon.initial()
        self.update_target_steps = hp.update_target_steps
        self.batch_size = pc.batch_size
        self.memory = Memory(self._h)
        self.steps = 0
        self._rng = random.Random(seed)
        self._learn_calls = 0
        self._stepped_over = False     # action_batch advanced `steps` past a multiple of update_target_steps

    def action_batch(self, states):
        """Agent.Action (agent.go:193-221) for a batch of states from vectorised environments: one
        steps++ / epsilon track per state, exploration decided per row on the device (uniform random
        action where the reference asks env.SampleAction), greedy = first-argmax of the online Q."""
        states = np.ascontiguousarray(states, np.float32).reshape(-1, self._h.D)
        before = self.steps
        self.steps += len(states)
        if before // self.update_target_steps != self.steps // self.update_target_steps:
            self._stepped_over = True   # the reference's steps++ would have landed on the multiple (see _update_target)
        return self._h.epsilon_greedy_action(states, float(self.epsilon), self._rng.getrandbits(63))

    def learn(self, indices=None):
        """Agent.Learn (agent.go:126-178).  `indices` injects Memory.Sample's draw (tests);
        by default the indices are drawn on the device."""
        self.memory.flush()
        if self.memory.len() < self.batch_size:
            return None                                            # agent.go:127-129
        self._learn_calls += 1
        if indices is None:
            self._h.learn(None, seed=self._rng.getrandbits(63))
        else:
            self._h.learn(indices)
        self.epsilon = self.epsilon_schedule.value()               # agent.go:171
        self._update_target()                                      # agent.go:173
        return None

    def _update_target(self):
        """agent.go:181-190: sync when steps sits on a multiple of UpdateTargetSteps at Learn time.  action_batch
        advances steps by a whole batch of states, so a multiple it stepped OVER counts too -- with 64 environments
        and a period of 100 the exact multiple is hit only every 1600 steps.  With one Action per Learn this is
        exactly the reference's `steps % N == 0`."""
        if self.steps % self.update_target_steps == 0 or self._stepped_over:   # agent.go:182
            self.policy.clone_learnables_to(self.target_policy)
        self._stepped_over = False

    def action(self, state):
        """Agent.Action (agent.go:193-205): epsilon-greedy; steps is counted here."""
        self.steps += 1
        if self._rng.random() < self.epsilon:
            return self.env.sample_action()
        return self._greedy(state)

    def _greedy(self, state):
        """Agent.action (agent.go:207-221): first-argmax of the online prediction."""
        return int(self._h.greedy_action(np.asarray(state, np.float32).reshape(1, -1))[0])

    def remember(self, event):
        """Agent.Remember (agent.go:224-229): PushFront, evict past BufferSize."""
        self.memory.push_front(event)

    def last_loss(self):
        """What gold's tracker reads from the `<name>_train_batch_loss` node."""
        return float(self._h.last_loss()[0])

    def close(self):
        self._h.close()


def new_agent(config, env):
    return Agent(config, env)
