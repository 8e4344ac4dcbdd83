package deepq

import (
	"fmt"
	"reflect"

	envv1 "github.com/aunum/gold/pkg/v1/env"
	"github.com/aunum/goro/pkg/v1/layer"
	modelv1 "github.com/aunum/goro/pkg/v1/model"
	g "gorgonia.org/gorgonia"
	"gorgonia.org/tensor"

	"github.com/aunum/gold-b200/go/goldq"
)

// PolicyConfig are the hyperparameters for a policy (reference: policy.go:14-29).
type PolicyConfig struct {
	Loss         modelv1.Loss
	Optimizer    g.Solver
	LayerBuilder LayerBuilder
	BatchSize    int
	Track        bool
}

// LayerBuilder builds layers (reference: policy.go:50).
type LayerBuilder func(x, y *modelv1.Input) []layer.Config

// DefaultPolicyConfig (reference: policy.go:32-38).
var DefaultPolicyConfig = &PolicyConfig{
	Loss:         modelv1.MSE,
	Optimizer:    g.NewAdamSolver(g.WithLearnRate(0.0005)),
	LayerBuilder: DefaultFCLayerBuilder,
	BatchSize:    20,
	Track:        true,
}

// DefaultFCLayerBuilder (reference: policy.go:53-59).
var DefaultFCLayerBuilder = func(x, y *modelv1.Input) []layer.Config {
	return []layer.Config{
		layer.FC{Input: x.Squeeze()[0], Output: 24},
		layer.FC{Input: 24, Output: 24},
		layer.FC{Input: 24, Output: y.Squeeze()[0], Activation: layer.Linear},
	}
}

// adamFields reads the solver's unexported hyper-parameters
// (vendor/gorgonia.org/gorgonia/solvers.go:399-415) by reflection; anything other
// than a plain *AdamSolver is rejected: there is no CPU fallback by design.
func adamFields(s g.Solver) (eta, eps, b1, b2 float64, err error) {
	a, ok := s.(*g.AdamSolver)
	if !ok {
		return 0, 0, 0, 0, fmt.Errorf("goldq implements gorgonia.AdamSolver only, got %T", s)
	}
	v := reflect.ValueOf(a).Elem()
	for _, f := range []string{"useClip", "useL1Reg", "useL2Reg"} {
		if v.FieldByName(f).Bool() {
			return 0, 0, 0, 0, fmt.Errorf("goldq: AdamSolver option %s is not implemented", f)
		}
	}
	return v.FieldByName("eta").Float(), v.FieldByName("eps").Float(),
		v.FieldByName("beta1").Float(), v.FieldByName("beta2").Float(), nil
}

// makePolicies replaces the two MakePolicy calls of NewAgent (reference:
// policy.go:74-110, agent.go:101-109): one device handle holds both networks.
func makePolicies(c *PolicyConfig, hp *Hyperparameters, env *envv1.Env) (*goldq.Handle, *Sequential, *Sequential, error) {
	x := modelv1.NewInput("state", env.ObservationSpaceShape())
	x.EnsureBatch()
	y := modelv1.NewInput("actionPotentials", envv1.PotentialsShape(env.ActionSpace))
	y.EnsureBatch()
	lossKind, delta := goldq.LossMSE, float32(1)
	switch l := c.Loss.(type) {
	case *modelv1.MSELoss:
	case *modelv1.PseudoHuberLoss: // reduced by Mean on the device; the reference's Sum(.., 1) cannot be differentiated
		lossKind, delta = goldq.LossPseudoHuber, l.Delta
	default:
		return nil, nil, nil, fmt.Errorf("goldq implements model.MSE and model.PseudoHuber, got %T", c.Loss)
	}
	layers := c.LayerBuilder(x, y)
	if len(layers) != 3 {
		return nil, nil, nil, fmt.Errorf("goldq implements FC-FC-FC policies (DefaultFCLayerBuilder)")
	}
	var fc [3]layer.FC
	for i, l := range layers {
		f, ok := l.(layer.FC)
		if !ok {
			return nil, nil, nil, fmt.Errorf("goldq: layer %d is %T, only layer.FC is implemented", i, l)
		}
		fc[i] = f
	}
	solver := goldq.SolverAdam
	var eta, eps, b1, b2 float64
	var err error
	if rms, ok := c.Optimizer.(*g.RMSPropSolver); ok { // gold's Atari config (policy.go:41-47); unexported fields by reflection
		v := reflect.ValueOf(rms).Elem()
		if v.FieldByName("useClip").Bool() || v.FieldByName("useL2Reg").Bool() {
			return nil, nil, nil, fmt.Errorf("goldq: RMSProp with clipping or L2 regularisation is not implemented")
		}
		solver, eta, eps, b2 = goldq.SolverRMSProp, v.FieldByName("eta").Float(), v.FieldByName("eps").Float(), v.FieldByName("decay").Float()
	} else if eta, eps, b1, b2, err = adamFields(c.Optimizer); err != nil {
		return nil, nil, nil, err
	}
	cfg := goldq.DefaultConfig()
	cfg.ObsDim, cfg.Hidden1, cfg.Hidden2, cfg.NActions = fc[0].Input, fc[0].Output, fc[1].Output, fc[2].Output
	cfg.BatchSize, cfg.ReplayCapacity, cfg.Gamma = c.BatchSize, int64(hp.BufferSize), hp.Gamma
	cfg.LR, cfg.Eps, cfg.Beta1, cfg.Beta2 = eta, eps, b1, b2
	cfg.Loss, cfg.HuberDelta, cfg.Solver = lossKind, delta, solver
	h, err := goldq.New(cfg)
	if err != nil {
		return nil, nil, nil, err
	}
	// GlorotN(1) initial weights, two independent draws (goro layer/fc.go:76-81)
	for _, net := range []int{goldq.NetOnline, goldq.NetTarget} {
		if err := h.SetParams(net, glorot(cfg)); err != nil {
			return nil, nil, nil, err
		}
	}
	return h, &Sequential{h: h, net: goldq.NetOnline, cfg: cfg}, &Sequential{h: h, net: goldq.NetTarget, cfg: cfg}, nil
}

func glorot(c goldq.Config) []float32 {
	var out []float32
	for _, s := range [][2]int{{c.ObsDim, c.Hidden1}, {c.Hidden1, c.Hidden2}, {c.Hidden2, c.NActions}} {
		out = append(out, g.GlorotEtAlN32(1, s[0], s[1])...)
		out = append(out, g.GlorotEtAlN32(1, 1, s[1])...)
	}
	return out
}

// Sequential is the view of one network that stands in for goro's
// *model.Sequential (vendor/github.com/aunum/goro/pkg/v1/model/model.go:17-50).
type Sequential struct {
	h   *goldq.Handle
	net int
	cfg goldq.Config
}

// Predict (reference: model.go:500-512): one (1,D) row in, a fresh (1,A) tensor out.
func (s *Sequential) Predict(x g.Value) (g.Value, error) {
	q, err := s.h.Predict(s.net, 1, x.Data().([]float32))
	if err != nil {
		return nil, err
	}
	return tensor.New(tensor.WithShape(1, s.cfg.NActions), tensor.WithBacking(q)), nil
}

// PredictBatch (reference: model.go:514-526).
func (s *Sequential) PredictBatch(x g.Value) (g.Value, error) {
	n := x.Shape()[0]
	q, err := s.h.Predict(s.net, n, x.Data().([]float32))
	if err != nil {
		return nil, err
	}
	return tensor.New(tensor.WithShape(n, s.cfg.NActions), tensor.WithBacking(q)), nil
}

// Fit / FitBatch (reference: model.go:528-573): one MSE + Adam step on explicit labels.
func (s *Sequential) Fit(x modelv1.ValueOr, y g.Value) error      { return s.fit(1, x, y) }
func (s *Sequential) FitBatch(x modelv1.ValueOr, y g.Value) error { return s.fit(s.cfg.BatchSize, x, y) }

func (s *Sequential) fit(n int, x modelv1.ValueOr, y g.Value) error {
	if s.net != goldq.NetOnline {
		return fmt.Errorf("the target network is never fitted")
	}
	xv, ok := x.(g.Value)
	if !ok {
		return fmt.Errorf("goldq: a single input tensor is expected")
	}
	return s.h.FitBatch(n, xv.Data().([]float32), y.Data().([]float32))
}

// CloneLearnablesTo (reference: model.go:606-636): online -> target.
func (s *Sequential) CloneLearnablesTo(to *Sequential) error { return s.h.SyncTarget() }

// The graph-introspection part of model.Model has no meaning on the device path:
// Graphs is empty, Visualize a no-op, ResizeBatch unsupported (recreate the agent).
func (s *Sequential) Graphs() map[string]*g.ExprGraph { return map[string]*g.ExprGraph{} }
func (s *Sequential) Visualize(name string)           {}
func (s *Sequential) ResizeBatch(n int) error {
	return fmt.Errorf("goldq: ResizeBatch is not supported; create the agent with the batch size you need")
}
