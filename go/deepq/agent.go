// Package deepq is a drop-in for github.com/aunum/gold/pkg/v1/agent/deepq whose
// learn step runs on a B200 through libgoldq (cgo).  Exported names, argument
// meaning and error behaviour follow the reference file by file
// (pkg/v1/agent/deepq/agent.go, memory.go, policy.go).
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
	"gorgonia.org/tensor"

	"github.com/aunum/gold-b200/go/goldq"
)

// Agent is a dqn agent (reference: agent.go:19-42).
type Agent struct {
	*agentv1.Base
	*Hyperparameters

	// Policy and TargetPolicy satisfy the slice of model.Model the agent and its
	// callers use; both are views of one device handle.
	Policy       model.Model
	TargetPolicy model.Model
	Epsilon      common.Schedule

	env               *envv1.Env
	epsilon           float32
	updateTargetSteps int
	batchSize         int
	memory            *Memory
	steps             int
	h                 *goldq.Handle
}

// Hyperparameters for the dqn agent (reference: agent.go:44-58).
type Hyperparameters struct {
	Gamma             float32
	Epsilon           common.Schedule
	UpdateTargetSteps int
	BufferSize        int
}

// DefaultHyperparameters (reference: agent.go:60-66).
var DefaultHyperparameters = &Hyperparameters{
	Epsilon:           common.DefaultDecaySchedule(),
	Gamma:             0.95,
	UpdateTargetSteps: 100,
	BufferSize:        10e6,
}

// AgentConfig (reference: agent.go:68-78).
type AgentConfig struct {
	Base *agentv1.Base
	*Hyperparameters
	PolicyConfig *PolicyConfig
}

// DefaultAgentConfig (reference: agent.go:80-85).
var DefaultAgentConfig = &AgentConfig{
	Hyperparameters: DefaultHyperparameters,
	PolicyConfig:    DefaultPolicyConfig,
	Base:            agentv1.NewBase("DeepQ"),
}

// NewAgent returns a new dqn agent (reference: agent.go:88-123).
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
	h, online, target, err := makePolicies(c.PolicyConfig, c.Hyperparameters, env)
	if err != nil {
		return nil, err
	}
	c.Base.Tracker.TrackValue("epsilon", c.Epsilon.Initial())
	return &Agent{
		Base:              c.Base,
		Hyperparameters:   c.Hyperparameters,
		memory:            NewMemory(h),
		Policy:            online,
		TargetPolicy:      target,
		Epsilon:           c.Epsilon,
		epsilon:           c.Epsilon.Initial(),
		env:               env,
		updateTargetSteps: c.UpdateTargetSteps,
		batchSize:         c.PolicyConfig.BatchSize,
		h:                 h,
	}, nil
}

// Learn the agent (reference: agent.go:126-178).  Sample, the two forwards per
// event, the label, FitBatch and the Adam step run in libgoldq.
func (a *Agent) Learn() error {
	if err := a.memory.flush(); err != nil {
		return err
	}
	if a.memory.Len() < a.batchSize {
		return nil
	}
	if err := a.h.Learn(nil, rand.Uint64()); err != nil {
		if err == goldq.ErrNotReady {
			return nil
		}
		return err
	}
	a.epsilon = a.Epsilon.Value()
	return a.updateTarget()
}

// updateTarget (reference: agent.go:181-190).
func (a *Agent) updateTarget() error {
	if a.steps%a.updateTargetSteps == 0 {
		return a.h.SyncTarget()
	}
	return nil
}

// Action selects the best known action for the given state (reference: agent.go:193-205).
func (a *Agent) Action(state *tensor.Dense) (action int, err error) {
	a.steps++
	a.Tracker.TrackValue("epsilon", a.epsilon)
	if num.RandF32(0.0, 1.0) < a.epsilon {
		return a.env.SampleAction()
	}
	return a.h.GreedyAction(state.Data().([]float32))
}

// Remember an event (reference: agent.go:224-229).
func (a *Agent) Remember(event *Event) {
	a.memory.PushFront(event)
}
