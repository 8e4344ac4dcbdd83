package deepq

import (
	"fmt"
	"reflect"
	"unsafe"

	agentv1 "github.com/aunum/gold/pkg/v1/agent"
	envv1 "github.com/aunum/gold/pkg/v1/env"
	"github.com/aunum/goro/pkg/v1/layer"
	modelv1 "github.com/aunum/goro/pkg/v1/model"
	"github.com/aunum/log"
	g "gorgonia.org/gorgonia"
	"gorgonia.org/tensor"

	"github.com/aunum/gold-b200/go/goldq"
)

// PolicyConfig are the hyperparameters for a policy (reference: policy.go:14-29).
type PolicyConfig struct {
	// Loss function to evaluate network performance.
	Loss modelv1.Loss

	// Optimizer to optimize the weights with regards to the error.
	Optimizer g.Solver

	// LayerBuilder is a builder of layer.
	LayerBuilder LayerBuilder

	// BatchSize of the updates.
	BatchSize int

	// Track is whether to track the model.
	Track bool
}

// DefaultPolicyConfig are the default hyperparameters for a policy (reference: policy.go:32-38).
var DefaultPolicyConfig = &PolicyConfig{
	Loss:         modelv1.MSE,
	Optimizer:    g.NewAdamSolver(g.WithLearnRate(0.0005)),
	LayerBuilder: DefaultFCLayerBuilder,
	BatchSize:    20,
	Track:        true,
}

// DefaultAtariPolicyConfig is the default policy config for atari environments (reference:
// policy.go:41-47).  The device path implements exactly its layer stack (GOLDQ_ARCH_ATARI_CONV, fp32).
var DefaultAtariPolicyConfig = &PolicyConfig{
	Loss:         modelv1.MSE,
	Optimizer:    g.NewRMSPropSolver(g.WithBatchSize(20)),
	LayerBuilder: DefaultAtariLayerBuilder,
	BatchSize:    20,
	Track:        true,
}

// LayerBuilder builds layers (reference: policy.go:50).
type LayerBuilder func(x, y *modelv1.Input) []layer.Config

// DefaultFCLayerBuilder is a default fully connected layer builder (reference: policy.go:53-59).
var DefaultFCLayerBuilder = func(x, y *modelv1.Input) []layer.Config {
	return []layer.Config{
		layer.FC{Input: x.Squeeze()[0], Output: 24},
		layer.FC{Input: 24, Output: 24},
		layer.FC{Input: 24, Output: y.Squeeze()[0], Activation: layer.Linear},
	}
}

// DefaultAtariLayerBuilder is the default layer builder for atari environments (reference:
// policy.go:62-71).
var DefaultAtariLayerBuilder = func(x, y *modelv1.Input) []layer.Config {
	return []layer.Config{
		layer.Conv2D{Input: 1, Output: 32, Width: 8, Height: 8, Stride: []int{4, 4}},
		layer.Conv2D{Input: 32, Output: 64, Width: 4, Height: 4, Stride: []int{2, 2}},
		layer.Conv2D{Input: 64, Output: 64, Width: 3, Height: 3, Stride: []int{1, 1}},
		layer.Flatten{},
		layer.FC{Input: 6400, Output: 512},
		layer.FC{Input: 512, Output: y.Squeeze()[0], Activation: layer.Linear},
	}
}

// MakePolicy makes a policy model (reference: policy.go:74-110).  Same arguments, same result type,
// and -- like the reference -- it panics when the model cannot be built.  The model it returns owns
// one device handle; NewAgent instead puts the online and the target network on ONE handle.
func MakePolicy(name string, config *PolicyConfig, base *agentv1.Base, env *envv1.Env) (modelv1.Model, error) {
	x := modelv1.NewInput("state", env.ObservationSpaceShape())
	x.EnsureBatch()

	y := modelv1.NewInput("actionPotentials", envv1.PotentialsShape(env.ActionSpace))
	y.EnsureBatch()

	log.Infov("x shape", x.Shape())
	log.Infov("y shape", y.Shape())

	model, err := NewSequential(name)
	if err != nil {
		panic(err)
	}
	model.AddLayers(config.LayerBuilder(x, y)...)
	if err = model.Compile(x, y, policyOpts(config, base)...); err != nil {
		panic(err)
	}
	return model, nil
}

// policyOpts are the goro options MakePolicy compiles with (reference: policy.go:88-100).
func policyOpts(config *PolicyConfig, base *agentv1.Base) []modelv1.Opt {
	opts := modelv1.NewOpts()
	opts.Add(
		modelv1.WithOptimizer(config.Optimizer),
		modelv1.WithLoss(config.Loss),
		modelv1.WithBatchSize(config.BatchSize),
		modelv1.WithMetrics(modelv1.TrainBatchLossMetric),
	)
	if config.Track && base != nil {
		opts.Add(modelv1.WithTracker(base.Tracker))
	} else {
		opts.Add(modelv1.WithoutTracker())
	}
	return opts.Values()
}

// Sequential is the device-resident counterpart of goro's *model.Sequential: it implements every
// method of model.Model (vendor/github.com/aunum/goro/pkg/v1/model/model.go:17-50) on one network
// (online or target) of a goldq handle.
type Sequential struct {
	name   string
	layers []layer.Config
	x      modelv1.InputOr
	y      *modelv1.Input

	h      *goldq.Handle // nil until Compile (or attach)
	net    int
	shared bool // the handle belongs to an Agent that holds both networks
	cfg    goldq.Config

	// host mirrors of the learnables, refreshed from the device whenever Learnables() is called
	graph  *g.ExprGraph
	nodes  g.Nodes
	mirror []float32
}

var _ modelv1.Model = (*Sequential)(nil)

// NewSequential returns a new sequential model (reference: model.go:101-108).
func NewSequential(name string) (*Sequential, error) {
	return &Sequential{name: name, net: goldq.NetOnline}, nil
}

// AddLayer adds a layer (reference: model.go:242-244).
func (s *Sequential) AddLayer(layer layer.Config) {
	s.layers = append(s.layers, layer)
}

// AddLayers adds a number of layers (reference: model.go:247-252).
func (s *Sequential) AddLayers(layers ...layer.Config) {
	for _, l := range layers {
		s.AddLayer(l)
	}
}

// unexported reads an unexported field of a struct value that is addressable: goro and gorgonia keep
// their option values (loss, optimizer, batch size; eta, eps, beta1 ...) in such fields.
func unexported(v reflect.Value, field string) reflect.Value {
	f := v.FieldByName(field)
	return reflect.NewAt(f.Type(), unsafe.Pointer(f.UnsafeAddr())).Elem()
}

// compiled collects what goro's option setters would have stored.  They type-switch on
// *model.Sequential and log.Fatal on anything else (model.go:142-239), so they are applied to a
// scratch goro Sequential and read back from it.
type compiled struct {
	loss      modelv1.Loss
	optimizer g.Solver
	batchSize int
}

func readOpts(name string, opts []modelv1.Opt) (compiled, error) {
	probe, err := modelv1.NewSequential(name)
	if err != nil {
		return compiled{}, err
	}
	for _, opt := range opts {
		opt(probe)
	}
	v := reflect.ValueOf(probe).Elem()
	c := compiled{batchSize: int(unexported(v, "batchSize").Int())}
	if l, ok := unexported(v, "loss").Interface().(modelv1.Loss); ok {
		c.loss = l
	}
	if o, ok := unexported(v, "optimizer").Interface().(g.Solver); ok {
		c.optimizer = o
	}
	return c, nil
}

// solverConfig reads the solver's unexported hyper-parameters
// (vendor/gorgonia.org/gorgonia/solvers.go:219-233, 399-415).  Anything but a plain AdamSolver or
// RMSPropSolver is an error: there is no CPU fallback by design.
func solverConfig(s g.Solver, cfg *goldq.Config) error {
	switch t := s.(type) {
	case *g.AdamSolver:
		v := reflect.ValueOf(t).Elem()
		for _, f := range []string{"useClip", "useL1Reg", "useL2Reg"} {
			if unexported(v, f).Bool() {
				return fmt.Errorf("goldq: AdamSolver option %s is not implemented", f)
			}
		}
		cfg.Solver = goldq.SolverAdam
		cfg.LR, cfg.Eps = unexported(v, "eta").Float(), unexported(v, "eps").Float()
		cfg.Beta1, cfg.Beta2 = unexported(v, "beta1").Float(), unexported(v, "beta2").Float()
	case *g.RMSPropSolver: // gold's Atari config (policy.go:41-47)
		v := reflect.ValueOf(t).Elem()
		for _, f := range []string{"useClip", "useL2Reg"} {
			if unexported(v, f).Bool() {
				return fmt.Errorf("goldq: RMSPropSolver option %s is not implemented", f)
			}
		}
		cfg.Solver = goldq.SolverRMSProp
		cfg.LR, cfg.Eps = unexported(v, "eta").Float(), unexported(v, "eps").Float()
		cfg.Beta2 = unexported(v, "decay").Float()
	default:
		return fmt.Errorf("goldq implements gorgonia.AdamSolver and RMSPropSolver, got %T", s)
	}
	return nil
}

// atariConfig recognises deepq.DefaultAtariLayerBuilder's stack (policy.go:62-71): Conv2D 32x8x8/4, 64x4x4/2,
// 64x3x3/1 with the layer's defaults (pad 1, ReLU, no bias), Flatten, FC 512 (ReLU), FC A (linear).  in is the
// input shape with its batch dimension, (1, C, H, W).
func (s *Sequential) atariConfig(cfg *goldq.Config, in tensor.Shape) error {
	bad := fmt.Errorf("goldq implements deepq.DefaultAtariLayerBuilder's convolutional stack only")
	if len(s.layers) != 6 || len(in) < 3 {
		return bad
	}
	want := [3][4]int{{32, 8, 8, 4}, {64, 4, 4, 2}, {64, 3, 3, 1}}
	for i := 0; i < 3; i++ {
		c, ok := s.layers[i].(layer.Conv2D)
		if !ok || c.Output != want[i][0] || c.Height != want[i][1] || c.Width != want[i][2] {
			return bad
		}
		if len(c.Stride) != 2 || c.Stride[0] != want[i][3] || c.Stride[1] != want[i][3] {
			return bad
		}
		if (len(c.Pad) != 0 && (c.Pad[0] != 1 || c.Pad[1] != 1)) || len(c.Dilation) != 0 {
			return bad
		}
		if _, relu := c.Activation.(*layer.ReLUActivation); c.Activation != nil && !relu {
			return bad
		}
	}
	if _, ok := s.layers[3].(layer.Flatten); !ok {
		return bad
	}
	f1, ok1 := s.layers[4].(layer.FC)
	f2, ok2 := s.layers[5].(layer.FC)
	if !ok1 || !ok2 || f1.Output != 512 || f2.Input != 512 || f1.NoBias || f2.NoBias {
		return bad
	}
	if _, lin := f2.Activation.(*layer.LinearActivation); !lin {
		return bad
	}
	n := len(in)
	cfg.Arch = goldq.ArchAtariConv
	cfg.InChannels, cfg.InHeight, cfg.InWidth = in[n-3], in[n-2], in[n-1]
	cfg.ObsDim = cfg.InChannels * cfg.InHeight * cfg.InWidth
	cfg.Hidden1, cfg.Hidden2, cfg.NActions = 512, 512, f2.Output
	return nil
}

// deviceConfig turns layers + options into the handle configuration, rejecting what the device
// path does not implement.  in is the model's input shape (used by the convolutional stack only).
func (s *Sequential) deviceConfig(c compiled, in tensor.Shape) (goldq.Config, error) {
	cfg := goldq.DefaultConfig()
	if len(s.layers) > 0 {
		if _, conv := s.layers[0].(layer.Conv2D); conv {
			if err := s.atariConfig(&cfg, in); err != nil {
				return cfg, err
			}
			return s.lossAndSolver(c, cfg)
		}
	}
	if len(s.layers) != 3 {
		return cfg, fmt.Errorf("goldq implements FC-FC-FC policies (DefaultFCLayerBuilder), got %d layers", len(s.layers))
	}
	var fc [3]layer.FC
	for i, l := range s.layers {
		f, ok := l.(layer.FC)
		if !ok {
			return cfg, fmt.Errorf("goldq: layer %d is %T, only layer.FC is implemented", i, l)
		}
		if f.NoBias {
			return cfg, fmt.Errorf("goldq: layer %d: NoBias is not implemented", i)
		}
		switch f.Activation.(type) {
		case nil, *layer.ReLUActivation:
			if i == 2 {
				return cfg, fmt.Errorf("goldq: the last layer must be linear")
			}
		case *layer.LinearActivation:
			if i != 2 {
				return cfg, fmt.Errorf("goldq: hidden layer %d must be ReLU", i)
			}
		default:
			return cfg, fmt.Errorf("goldq: layer %d: activation %T is not implemented", i, f.Activation)
		}
		fc[i] = f
	}
	if fc[0].Output != fc[1].Input || fc[1].Output != fc[2].Input {
		return cfg, fmt.Errorf("goldq: layer shapes do not chain")
	}
	cfg.ObsDim, cfg.Hidden1, cfg.Hidden2, cfg.NActions = fc[0].Input, fc[0].Output, fc[1].Output, fc[2].Output
	return s.lossAndSolver(c, cfg)
}

// lossAndSolver fills the batch size, loss and solver fields from the compiled options.
func (s *Sequential) lossAndSolver(c compiled, cfg goldq.Config) (goldq.Config, error) {
	cfg.BatchSize = c.batchSize
	switch l := c.loss.(type) {
	case nil, *modelv1.MSELoss:
		cfg.Loss = goldq.LossMSE
	case *modelv1.PseudoHuberLoss: // reduced by Mean on the device; the reference's Sum(.., 1) cannot be differentiated
		cfg.Loss, cfg.HuberDelta = goldq.LossPseudoHuber, l.Delta
	default:
		return cfg, fmt.Errorf("goldq implements model.MSE and model.PseudoHuber, got %T", c.loss)
	}
	if c.optimizer == nil {
		c.optimizer = g.NewAdamSolver() // goro's default (model.go:290-292)
	}
	if err := solverConfig(c.optimizer, &cfg); err != nil {
		return cfg, err
	}
	return cfg, nil
}

// Compile the model (reference: model.go:260-312): validates layers and options and allocates the
// network on the device.  Unsupported layers, activations, losses or solvers are errors.
func (s *Sequential) Compile(x modelv1.InputOr, y *modelv1.Input, opts ...modelv1.Opt) error {
	c, err := readOpts(s.name, opts)
	if err != nil {
		return err
	}
	cfg, err := s.deviceConfig(c, x.Input().Shape())
	if err != nil {
		return err
	}
	s.x, s.y = x, y
	if s.h != nil { // attached to an agent's handle: the configuration must agree with it
		have := s.h.Config()
		if have.ObsDim != cfg.ObsDim || have.Hidden1 != cfg.Hidden1 || have.Hidden2 != cfg.Hidden2 || have.NActions != cfg.NActions {
			return fmt.Errorf("goldq: layers do not match the agent's device handle")
		}
		s.cfg = have
		return nil
	}
	cfg.ReplayCapacity = 1 // a stand-alone model has no replay memory
	h, err := goldq.New(cfg)
	if err != nil {
		return err
	}
	if err := h.SetParams(goldq.NetOnline, glorot(cfg)); err != nil {
		return err
	}
	s.h, s.cfg = h, cfg
	return nil
}

// attach binds the model to one network of an existing handle (NewAgent: one handle, two networks).
func (s *Sequential) attach(h *goldq.Handle, net int) {
	s.h, s.net, s.shared, s.cfg = h, net, true, h.Config()
}

// glorot draws the initial parameters in learnable order: FC weights and biases GlorotN(1) (goro layer/fc.go:76-81),
// convolution filters GlorotU(1) (layer/conv2d.go:104-106).
func glorot(c goldq.Config) []float32 {
	var out []float32
	if c.Arch == goldq.ArchAtariConv {
		h1, w1 := (c.InHeight+2-8)/4+1, (c.InWidth+2-8)/4+1
		h2, w2 := (h1+2-4)/2+1, (w1+2-4)/2+1
		for _, f := range [][4]int{{32, c.InChannels, 8, 8}, {64, 32, 4, 4}, {64, 64, 3, 3}} {
			out = append(out, g.GlorotU(1)(tensor.Float32, f[0], f[1], f[2], f[3]).([]float32)...)
		}
		for _, s := range [][2]int{{64 * h2 * w2, 512}, {512, c.NActions}} {
			out = append(out, g.GlorotEtAlN32(1, s[0], s[1])...)
			out = append(out, g.GlorotEtAlN32(1, 1, s[1])...)
		}
		return out
	}
	for _, s := range [][2]int{{c.ObsDim, c.Hidden1}, {c.Hidden1, c.Hidden2}, {c.Hidden2, c.NActions}} {
		out = append(out, g.GlorotEtAlN32(1, s[0], s[1])...)
		out = append(out, g.GlorotEtAlN32(1, 1, s[1])...)
	}
	return out
}

func (s *Sequential) ready() error {
	if s.h == nil {
		return fmt.Errorf("model %q is not compiled", s.name)
	}
	return nil
}

func floats(v g.Value) ([]float32, error) {
	data, ok := v.Data().([]float32)
	if !ok {
		return nil, fmt.Errorf("goldq: float32 tensors only, got %v", v.Dtype())
	}
	return data, nil
}

// Predict x (reference: model.go:500-512): one (1,D) row in, a fresh (1,A) tensor out.
func (s *Sequential) Predict(x g.Value) (prediction g.Value, err error) {
	if err = s.ready(); err != nil {
		return nil, err
	}
	in, err := floats(x)
	if err != nil {
		return nil, err
	}
	if len(in) != s.cfg.ObsDim {
		return nil, fmt.Errorf("shape mismatch: expected (1, %d), got %v", s.cfg.ObsDim, x.Shape())
	}
	q, err := s.h.Predict(s.net, 1, in)
	if err != nil {
		return nil, err
	}
	return tensor.New(tensor.WithShape(1, s.cfg.NActions), tensor.WithBacking(q)), nil
}

// PredictBatch predicts x as a batch (reference: model.go:514-526).
func (s *Sequential) PredictBatch(x g.Value) (prediction g.Value, err error) {
	if err = s.ready(); err != nil {
		return nil, err
	}
	in, err := floats(x)
	if err != nil {
		return nil, err
	}
	n := len(in) / s.cfg.ObsDim
	if n*s.cfg.ObsDim != len(in) || n == 0 {
		return nil, fmt.Errorf("shape mismatch: expected (n, %d), got %v", s.cfg.ObsDim, x.Shape())
	}
	q, err := s.h.Predict(s.net, n, in)
	if err != nil {
		return nil, err
	}
	return tensor.New(tensor.WithShape(n, s.cfg.NActions), tensor.WithBacking(q)), nil
}

// Fit x to y (reference: model.go:528-550): one loss + solver step on a single (1,D)/(1,A) pair.
func (s *Sequential) Fit(x modelv1.ValueOr, y g.Value) error {
	return s.fit(1, x, y)
}

// FitBatch fits x to y as a batch (reference: model.go:552-573): exactly BatchSize rows, as the
// reference's compiled batch graph requires (io.go:138-153).
func (s *Sequential) FitBatch(x modelv1.ValueOr, y g.Value) error {
	return s.fit(s.h.BatchSize(), x, y)
}

func (s *Sequential) fit(n int, x modelv1.ValueOr, y g.Value) error {
	if err := s.ready(); err != nil {
		return err
	}
	if s.net != goldq.NetOnline {
		return fmt.Errorf("the target network is never fitted")
	}
	vals := modelv1.ValuesFrom(x)
	if len(vals) != 1 {
		return fmt.Errorf("goldq: a single input tensor is expected, got %d", len(vals))
	}
	in, err := floats(vals[0])
	if err != nil {
		return err
	}
	lab, err := floats(y)
	if err != nil {
		return err
	}
	if len(in) != n*s.cfg.ObsDim || len(lab) != n*s.cfg.NActions {
		return fmt.Errorf("shape mismatch: expected x (%d, %d) and y (%d, %d)", n, s.cfg.ObsDim, n, s.cfg.NActions)
	}
	return s.h.FitBatch(n, in, lab)
}

// ResizeBatch resizes the batch graphs (reference: model.go:485-497): here, the device scratch that
// is sized by the batch.  Parameters, solver state and replay are kept.
func (s *Sequential) ResizeBatch(n int) error {
	if err := s.ready(); err != nil {
		return err
	}
	if err := s.h.ResizeBatch(n); err != nil {
		return err
	}
	s.cfg.BatchSize = n
	return nil
}

// Visualize the model by graph name (reference: model.go:576-578).  The device path has no
// expression graph to draw: a logged no-op.
func (s *Sequential) Visualize(name string) {
	log.Debugf("model %q runs on the device: no %q graph to visualize", s.name, name)
}

// Graphs returns the expression graphs of the model (reference: model.go:581-588): none on the
// device path, so the map is empty.
func (s *Sequential) Graphs() map[string]*g.ExprGraph {
	return map[string]*g.ExprGraph{}
}

// X is the input to the model (reference: model.go:591-593).
func (s *Sequential) X() modelv1.InputOr {
	return s.x
}

// Y is the expected output of the model (reference: model.go:596-598).
func (s *Sequential) Y() *modelv1.Input {
	return s.y
}

// Learnables are the model learnables (reference: model.go:601-603; order as layer/chain.go:38-44
// yields it: W1, b1, W2, b2, W3, b3).  The nodes live on a private graph and are bound to host
// mirrors of the device parameters, refreshed on every call; writing to them does not change the
// device (use Fit).
func (s *Sequential) Learnables() g.Nodes {
	if s.h == nil {
		return nil
	}
	theta, err := s.h.GetParams(s.net)
	if err != nil {
		log.Errorf("goldq: reading the parameters back failed: %v", err)
		return s.nodes
	}
	c := s.cfg
	if s.nodes == nil {
		s.graph = g.NewGraph()
		s.mirror = make([]float32, len(theta))
		off := 0
		dense := [][2]int{{c.ObsDim, c.Hidden1}, {c.Hidden1, c.Hidden2}, {c.Hidden2, c.NActions}}
		if c.Arch == goldq.ArchAtariConv { // three filters (O, I, H, W), then the two FC layers
			h1, w1 := (c.InHeight+2-8)/4+1, (c.InWidth+2-8)/4+1
			h2, w2 := (h1+2-4)/2+1, (w1+2-4)/2+1
			for l, f := range [][4]int{{32, c.InChannels, 8, 8}, {64, 32, 4, 4}, {64, 64, 3, 3}} {
				n := f[0] * f[1] * f[2] * f[3]
				w := tensor.New(tensor.WithShape(f[0], f[1], f[2], f[3]), tensor.WithBacking(s.mirror[off:off+n]))
				off += n
				s.nodes = append(s.nodes, g.NewTensor(s.graph, tensor.Float32, 4, g.WithShape(f[0], f[1], f[2], f[3]),
					g.WithName(fmt.Sprintf("%s_conv%d_filter", s.name, l)), g.WithValue(w)))
			}
			dense = [][2]int{{64 * h2 * w2, 512}, {512, c.NActions}}
		}
		for l, sh := range dense {
			w := tensor.New(tensor.WithShape(sh[0], sh[1]), tensor.WithBacking(s.mirror[off:off+sh[0]*sh[1]]))
			off += sh[0] * sh[1]
			b := tensor.New(tensor.WithShape(1, sh[1]), tensor.WithBacking(s.mirror[off:off+sh[1]]))
			off += sh[1]
			s.nodes = append(s.nodes,
				g.NewMatrix(s.graph, tensor.Float32, g.WithShape(sh[0], sh[1]), g.WithName(fmt.Sprintf("%s_fc%d_w", s.name, l)), g.WithValue(w)),
				g.NewMatrix(s.graph, tensor.Float32, g.WithShape(1, sh[1]), g.WithName(fmt.Sprintf("%s_fc%d_b", s.name, l)), g.WithValue(b)))
		}
	}
	copy(s.mirror, theta)
	return s.nodes
}

// CloneLearnablesTo another model (reference: model.go:606-636).  Within an agent this is the
// device-side target <- online copy; between stand-alone models the parameters travel through the
// host.
func (s *Sequential) CloneLearnablesTo(to *Sequential) error {
	if err := s.ready(); err != nil {
		return err
	}
	if to == nil || to.h == nil {
		return fmt.Errorf("models must be identical to clone learnables")
	}
	if to.h == s.h {
		if s.net == goldq.NetOnline && to.net == goldq.NetTarget {
			return s.h.SyncTarget()
		}
		return fmt.Errorf("goldq: only online -> target within one agent")
	}
	if s.h.P != to.h.P {
		return fmt.Errorf("models must be identical to clone learnables")
	}
	theta, err := s.h.GetParams(s.net)
	if err != nil {
		return err
	}
	return to.h.SetParams(to.net, theta)
}

// Close releases the device handle of a stand-alone model (an agent's models are released by the agent).
func (s *Sequential) Close() {
	if s.h != nil && !s.shared {
		s.h.Close()
	}
	s.h = nil
}
