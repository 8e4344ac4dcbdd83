package deepq

import (
	"fmt"
	"math/rand"

	envv1 "github.com/aunum/gold/pkg/v1/env"
	"gorgonia.org/tensor"

	"github.com/aunum/gold-b200/go/goldq"
)

// Event is an event that occurred (reference: memory.go:15-25).
type Event struct {
	*envv1.Outcome
	State  *tensor.Dense
	Action int
	i      int
}

// NewEvent returns a new event (reference: memory.go:28-34).
func NewEvent(state *tensor.Dense, action int, outcome *envv1.Outcome) *Event {
	return &Event{State: state, Action: action, Outcome: outcome}
}

// Memory for the dqn agent (reference: memory.go:41-51).  The events live in the
// device replay ring; PushFront buffers them on the host until the next Learn.
type Memory struct {
	h       *goldq.Handle
	pending []*Event
}

// NewMemory returns a new Memory store.
func NewMemory(h *goldq.Handle) *Memory { return &Memory{h: h} }

// Len of the memory (deque.Len).
func (m *Memory) Len() int { return m.h.ReplayLen() + len(m.pending) }

// PushFront makes the event logical index 0 (deque.PushFront).
func (m *Memory) PushFront(e *Event) { m.pending = append(m.pending, e) }

// Sample returns batchsize distinct logical indices (reference: memory.go:54-69;
// the reference returns the events, which here live on the device).
func (m *Memory) Sample(batchsize int) ([]int32, error) {
	if m.Len() < batchsize {
		return nil, fmt.Errorf("queue size %d is less than batch size %d", m.Len(), batchsize)
	}
	perm := rand.Perm(m.Len())[:batchsize]
	out := make([]int32, batchsize)
	for i, v := range perm {
		out[i] = int32(v)
	}
	return out, nil
}

func (m *Memory) flush() error {
	n := len(m.pending)
	if n == 0 {
		return nil
	}
	var s, s2, r []float32
	var a []int32
	var d []uint8
	for _, e := range m.pending {
		s = append(s, e.State.Data().([]float32)...)
		s2 = append(s2, e.Observation.Data().([]float32)...)
		a = append(a, int32(e.Action))
		r = append(r, e.Reward)
		if e.Done {
			d = append(d, 1)
		} else {
			d = append(d, 0)
		}
	}
	m.pending = m.pending[:0]
	return m.h.ReplayPush(n, s, a, r, s2, d)
}
