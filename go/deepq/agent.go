// Package deepq is a drop-in for github.com/aunum/gold/pkg/v1/agent/deepq whose
// learn step runs on a B200 through libgoldq (cgo).  The exported surface — type, field
// and function names, argument meaning, error behaviour — is the reference's
// (pkg/v1/agent/deepq/agent.go, memory.go, policy.go); everything behind it is new.
//
// NOT COMPILED HERE (no Go toolchain in the build image); its executable twin is
// gold_b200/deepq.py, which tests/test_host_api.py drives through the same C ABI.
package deepq

import (
	"errors"
	"math/rand"

	agentv1 "github.com/aunum/gold/pkg/v1/agent"
	"github.com/aunum/gold/pkg/v1/common"
	"github.com/aunum/gold/pkg/v1/common/num"
	envv1 "github.com/aunum/gold/pkg/v1/env"
	"github.com/aunum/goro/pkg/v1/model"
	"gorgonia.org/tensor"

	"github.com/aunum/gold-b200/go/goldq"
)

// Hyperparameters keeps the reference's exported fields (agent.go:44-58).
type Hyperparameters struct {
	Gamma             float32         // discount of the TD target
	Epsilon           common.Schedule // exploration rate, advanced once per Learn
	UpdateTargetSteps int             // target <- online every this many Action calls
	BufferSize        int             // replay capacity (device ring)
}

// AgentConfig keeps the reference's exported fields (agent.go:68-78).
type AgentConfig struct {
	Base *agentv1.Base
	*Hyperparameters
	PolicyConfig *PolicyConfig
}

// Defaults: same values as the reference (agent.go:60-66, 80-85).
var (
	DefaultHyperparameters = &Hyperparameters{Gamma: 0.95, Epsilon: common.DefaultDecaySchedule(), UpdateTargetSteps: 100, BufferSize: 10e6}
	DefaultAgentConfig     = &AgentConfig{Base: agentv1.NewBase("DeepQ"), Hyperparameters: DefaultHyperparameters, PolicyConfig: DefaultPolicyConfig}
)

// device is everything the agent keeps on the B200 side plus the counters that drive it.
type device struct {
	handle     *goldq.Handle
	replay     *Memory
	env        *envv1.Env
	eps        float32 // current exploration rate
	syncEvery  int     // = UpdateTargetSteps
	batch      int     // = PolicyConfig.BatchSize
	actionsRun int     // Action calls so far; the target sync is keyed on it, as in the reference
}

// Agent is a dqn agent with the reference's exported fields (agent.go:19-42).  Policy and
// TargetPolicy are two views of ONE device handle.
type Agent struct {
	*agentv1.Base
	*Hyperparameters
	Policy, TargetPolicy model.Model
	Epsilon              common.Schedule
	device
}

// resolved returns c with the reference's fall-backs applied (nil config, nil Base, nil schedule).
func resolved(c *AgentConfig) *AgentConfig {
	if c == nil {
		return DefaultAgentConfig
	}
	if c.Base == nil {
		c.Base = DefaultAgentConfig.Base
	}
	if c.Epsilon == nil {
		c.Epsilon = common.DefaultDecaySchedule()
	}
	return c
}

// NewAgent builds the device handle (both networks, replay ring, solver state) and wraps it.
// Errors as the reference does for a missing environment (agent.go:88-123).
func NewAgent(c *AgentConfig, env *envv1.Env) (*Agent, error) {
	if env == nil {
		return nil, errors.New("environment cannot be nil")
	}
	cfg := resolved(c)
	handle, online, target, err := makePolicies(cfg.PolicyConfig, cfg.Hyperparameters, env)
	if err != nil {
		return nil, err
	}
	start := cfg.Epsilon.Initial()
	cfg.Base.Tracker.TrackValue("epsilon", start)
	agent := &Agent{Base: cfg.Base, Hyperparameters: cfg.Hyperparameters, Policy: online, TargetPolicy: target, Epsilon: cfg.Epsilon}
	agent.device = device{handle: handle, replay: NewMemory(handle), env: env, eps: start,
		syncEvery: cfg.UpdateTargetSteps, batch: cfg.PolicyConfig.BatchSize}
	return agent, nil
}

// Learn runs one minibatch update.  The reference's body (sample, two forwards per event, label,
// concat, FitBatch, Adam: agent.go:126-178) is one goldq_learn call; what stays in Go is the
// "too few events yet" no-op, the epsilon decay and the target-sync cadence.
func (a *Agent) Learn() error {
	if err := a.replay.flush(); err != nil { // events remembered since the last Learn go to the device ring
		return err
	}
	if a.replay.Len() < a.batch {
		return nil
	}
	switch err := a.handle.Learn(nil, rand.Uint64()); err {
	case nil:
	case goldq.ErrNotReady:
		return nil
	default:
		return err
	}
	a.eps = a.Epsilon.Value()
	if a.actionsRun%a.syncEvery != 0 { // agent.go:181-190: keyed on Action calls, checked after every Learn
		return nil
	}
	return a.handle.SyncTarget()
}

// Action is epsilon-greedy; the greedy branch is the first-argmax of the online network's Q
// (agent.go:193-221), evaluated on the device.
func (a *Agent) Action(state *tensor.Dense) (action int, err error) {
	a.actionsRun++
	a.Tracker.TrackValue("epsilon", a.eps)
	explore := num.RandF32(0.0, 1.0) < a.eps
	if explore {
		return a.env.SampleAction()
	}
	return a.handle.GreedyAction(state.Data().([]float32))
}

// Remember queues an event for the device replay ring (agent.go:224-229).
func (a *Agent) Remember(event *Event) { a.replay.PushFront(event) }
