// Package deepq is a drop-in for github.com/aunum/gold/pkg/v1/agent/deepq whose learn step runs on a
// B200 through libgoldq (cgo).  The exported surface -- types, fields, functions, argument meaning and
// error behaviour -- is the reference's (pkg/v1/agent/deepq/agent.go, memory.go, policy.go;
// tests/test_go_surface.py holds it to that); everything behind it is new.
//
// NOT COMPILED HERE (no Go toolchain in the build image); its executable twin is
// gold_b200/deepq.py, which tests/test_host_api.py drives through the same C ABI.
package deepq

import (
	"fmt"
	"math/rand"

	agentv1 "github.com/aunum/gold/pkg/v1/agent"
	"github.com/aunum/gold/pkg/v1/common"
	"github.com/aunum/gold/pkg/v1/common/num"
	envv1 "github.com/aunum/gold/pkg/v1/env"
	"github.com/aunum/goro/pkg/v1/model"
	"github.com/aunum/log"
	"gorgonia.org/tensor"

	"github.com/aunum/gold-b200/go/goldq"
)

// Agent is a dqn agent (reference: agent.go:19-42).
type Agent struct {
	// Base for the agent.
	*agentv1.Base

	// Hyperparameters for the dqn agent.
	*Hyperparameters

	// Policy for the agent.
	Policy model.Model

	// Target policy for double Q learning.
	TargetPolicy model.Model

	// Epsilon is the rate at which the agent explores vs exploits.
	Epsilon common.Schedule

	env               *envv1.Env
	epsilon           float32
	updateTargetSteps int
	batchSize         int
	memory            *Memory
	steps             int

	// device side: ONE handle holds both networks, the solver state and the replay ring
	handle  *goldq.Handle
	pending []*Event // remembered since the last Learn, not yet on the device ring
}

// Hyperparameters for the dqn agent (reference: agent.go:44-58).
type Hyperparameters struct {
	// Gamma is the discount factor (0≤γ≤1).
	Gamma float32

	// Epsilon is the rate at which the agent should exploit vs explore.
	Epsilon common.Schedule

	// UpdateTargetSteps determines how often the target network updates its parameters.
	UpdateTargetSteps int

	// BuferSize is the buffer size of the memory.
	BufferSize int
}

// DefaultHyperparameters are the default hyperparameters (reference: agent.go:60-66).
var DefaultHyperparameters = &Hyperparameters{
	Epsilon:           common.DefaultDecaySchedule(),
	Gamma:             0.95,
	UpdateTargetSteps: 100,
	BufferSize:        10e6,
}

// AgentConfig is the config for a dqn agent (reference: agent.go:68-78).
type AgentConfig struct {
	// Base for the agent.
	Base *agentv1.Base

	// Hyperparameters for the agent.
	*Hyperparameters

	// PolicyConfig for the agent.
	PolicyConfig *PolicyConfig
}

// DefaultAgentConfig is the default config for a dqn agent (reference: agent.go:80-85).
var DefaultAgentConfig = &AgentConfig{
	Hyperparameters: DefaultHyperparameters,
	PolicyConfig:    DefaultPolicyConfig,
	Base:            agentv1.NewBase("DeepQ"),
}

// NewAgent returns a new dqn agent (reference: agent.go:88-123).  Same defaults, same "environment
// cannot be nil" error; the two policies are two views of one device handle.
func NewAgent(c *AgentConfig, env *envv1.Env) (*Agent, error) {
	if c == nil {
		c = DefaultAgentConfig
	}
	if c.Base == nil {
		c.Base = DefaultAgentConfig.Base
	}
	if env == nil {
		return nil, fmt.Errorf("environment cannot be nil")
	}
	if c.Epsilon == nil {
		c.Epsilon = common.DefaultDecaySchedule()
	}
	handle, online, target, err := makePolicies(c.PolicyConfig, c.Hyperparameters, c.Base, env)
	if err != nil {
		return nil, err
	}
	c.Base.Tracker.TrackValue("epsilon", c.Epsilon.Initial())
	return &Agent{
		Base:              c.Base,
		Hyperparameters:   c.Hyperparameters,
		memory:            NewMemory(),
		Policy:            online,
		TargetPolicy:      target,
		Epsilon:           c.Epsilon,
		epsilon:           c.Epsilon.Initial(),
		env:               env,
		updateTargetSteps: c.UpdateTargetSteps,
		batchSize:         c.PolicyConfig.BatchSize,
		handle:            handle,
	}, nil
}

// makePolicies stands in for the two MakePolicy calls of the reference's NewAgent (agent.go:101-109):
// the same inputs, layers and goro options, but one device handle for the online and the target
// network (independent GlorotN draws, as two reference models would have).
func makePolicies(config *PolicyConfig, hp *Hyperparameters, base *agentv1.Base, env *envv1.Env) (*goldq.Handle, *Sequential, *Sequential, error) {
	x := model.NewInput("state", env.ObservationSpaceShape())
	x.EnsureBatch()
	y := model.NewInput("actionPotentials", envv1.PotentialsShape(env.ActionSpace))
	y.EnsureBatch()

	var models [2]*Sequential
	for i, name := range []string{"online", "target"} {
		m, err := NewSequential(name)
		if err != nil {
			return nil, nil, nil, err
		}
		m.AddLayers(config.LayerBuilder(x, y)...)
		models[i] = m
	}
	opts := policyOpts(config, base)
	c, err := readOpts("online", opts)
	if err != nil {
		return nil, nil, nil, err
	}
	cfg, err := models[0].deviceConfig(c, x.Shape())
	if err != nil {
		return nil, nil, nil, err
	}
	cfg.ReplayCapacity, cfg.Gamma = int64(hp.BufferSize), hp.Gamma
	handle, err := goldq.New(cfg)
	if err != nil {
		return nil, nil, nil, err
	}
	for i, net := range []int{goldq.NetOnline, goldq.NetTarget} {
		if err := handle.SetParams(net, glorot(cfg)); err != nil {
			return nil, nil, nil, err
		}
		models[i].attach(handle, net)
		if err := models[i].Compile(x, y, opts...); err != nil {
			return nil, nil, nil, err
		}
	}
	return handle, models[0], models[1], nil
}

// flush sends the events remembered since the last Learn to the device replay ring, oldest first
// (each becomes logical index 0, like PushFront).
func (a *Agent) flush() error {
	n := len(a.pending)
	if n == 0 {
		return nil
	}
	var s, s2, r []float32
	var act []int32
	var done []uint8
	for _, e := range a.pending {
		s = append(s, e.State.Data().([]float32)...)
		s2 = append(s2, e.Observation.Data().([]float32)...)
		act = append(act, int32(e.Action))
		r = append(r, e.Reward)
		if e.Done {
			done = append(done, 1)
		} else {
			done = append(done, 0)
		}
	}
	a.pending = a.pending[:0]
	return a.handle.ReplayPush(n, s, act, r, s2, done)
}

// Learn the agent (reference: agent.go:126-178).  Sample, the two forwards per event, the TD label,
// FitBatch and the solver step are ONE device call; the "too few events" no-op, the epsilon decay
// and the target-sync cadence stay here.
func (a *Agent) Learn() error {
	if a.memory.Len() < a.batchSize {
		return nil
	}
	if err := a.flush(); err != nil {
		return err
	}
	switch err := a.handle.Learn(nil, rand.Uint64()); err {
	case nil:
	case goldq.ErrNotReady:
		return nil
	default:
		return err
	}
	a.epsilon = a.Epsilon.Value()

	return a.updateTarget()
}

// LearnOn is Learn with the batch drawn on the host by Memory.Sample (the reference's path,
// memory.go:54-69): the sampled logical indices are what the device step trains on.
func (a *Agent) LearnOn(batch []*Event) error {
	if len(batch) != a.batchSize {
		return fmt.Errorf("batch of %d events, batch size is %d", len(batch), a.batchSize)
	}
	if err := a.flush(); err != nil {
		return err
	}
	if err := a.handle.Learn(indices(batch), 0); err != nil {
		return err
	}
	a.epsilon = a.Epsilon.Value()
	return a.updateTarget()
}

// updateTarget copies the weights from the online network to the target network on the provided
// interval (reference: agent.go:181-190); keyed on Action calls, checked after every Learn.
func (a *Agent) updateTarget() error {
	if a.steps%a.updateTargetSteps == 0 {
		log.Debugf("updating target model - current steps %v target update %v", a.steps, a.updateTargetSteps)
		return a.Policy.(*Sequential).CloneLearnablesTo(a.TargetPolicy.(*Sequential))
	}
	return nil
}

// Action selects the best known action for the given state (reference: agent.go:193-205).
func (a *Agent) Action(state *tensor.Dense) (action int, err error) {
	a.steps++
	a.Tracker.TrackValue("epsilon", a.epsilon)
	if num.RandF32(0.0, 1.0) < a.epsilon {
		// explore
		return a.env.SampleAction()
	}
	return a.action(state)
}

// action is the greedy branch (reference: agent.go:207-221): first-argmax of the online network's
// prediction, evaluated on the device.
func (a *Agent) action(state *tensor.Dense) (action int, err error) {
	in, ok := state.Data().([]float32)
	if !ok {
		return 0, fmt.Errorf("goldq: float32 states only, got %v", state.Dtype())
	}
	return a.handle.GreedyAction(in)
}

// Remember an event (reference: agent.go:224-229): front of the host deque, oldest evicted past
// BufferSize; the device ring (capacity BufferSize) receives it at the next Learn.
func (a *Agent) Remember(event *Event) {
	a.memory.PushFront(event)
	if a.memory.Len() > a.BufferSize {
		a.memory.PopBack()
	}
	a.pending = append(a.pending, event)
}

// Close releases the device memory of the agent.
func (a *Agent) Close() {
	if a.handle != nil {
		a.handle.Close()
		a.handle = nil
	}
}
