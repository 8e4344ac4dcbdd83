package deepq

import (
	"fmt"
	"math/rand"
	"time"

	envv1 "github.com/aunum/gold/pkg/v1/env"
	"github.com/aunum/log"
	"github.com/gammazero/deque"
	"gorgonia.org/tensor"
)

// Event is an event that occurred (reference: memory.go:15-25).
type Event struct {
	*envv1.Outcome

	// State by which the action was taken.
	State *tensor.Dense

	// Action that was taken.
	Action int

	i int
}

// NewEvent returns a new event (reference: memory.go:28-34).
func NewEvent(state *tensor.Dense, action int, outcome *envv1.Outcome) *Event {
	return &Event{State: state, Action: action, Outcome: outcome}
}

// Print the event (reference: memory.go:37-39).
func (e *Event) Print() {
	log.Infof("event --> \n state: %v \n action: %v \n reward: %v \n done: %v \n obv: %v\n\n",
		e.State, e.Action, e.Reward, e.Done, e.Observation)
}

// Memory for the dqn agent (reference: memory.go:41-51).  As in the reference it IS a deque of
// *Event (PushFront / PopBack / At / Len come from the embedded type), so code written against
// gold's Memory keeps compiling.  The agent additionally mirrors every remembered event into the
// device replay ring (Agent.Remember), which is what Learn samples from; the host deque only
// serves callers that want the Events back.
type Memory struct {
	*deque.Deque
}

// NewMemory returns a new Memory store (reference: memory.go:46-51).
func NewMemory() *Memory {
	return &Memory{Deque: &deque.Deque{}}
}

// Sample from the memory with the given batch size (reference: memory.go:54-69): an error when
// fewer than batchsize events are held, otherwise batchsize distinct events, uniformly, each with
// its logical index (0 = newest) recorded in the unexported field i.
func (m *Memory) Sample(batchsize int) ([]*Event, error) {
	n := m.Len()
	if n < batchsize {
		return nil, fmt.Errorf("queue size %d is less than batch size %d", n, batchsize)
	}
	rng := rand.New(rand.NewSource(time.Now().UnixNano()))
	picked := rng.Perm(n)[:batchsize]
	out := make([]*Event, 0, batchsize)
	for _, at := range picked {
		ev := m.At(at).(*Event)
		ev.i = at
		out = append(out, ev)
	}
	return out, nil
}

// indices returns the logical indices of a sampled batch (what the device learn step takes when
// the host, not the device, draws the sample).
func indices(batch []*Event) []int32 {
	out := make([]int32, len(batch))
	for k, ev := range batch {
		out[k] = int32(ev.i)
	}
	return out
}
