// Package goldq is the thin cgo binding of include/goldq.h (libgoldq.so): the
// B200-native DQN learn step behind gold's deepq agent.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain
// (see DESIGN.md).  The binding is a 1:1 wrapper of the C ABI whose behaviour
// is exercised through the identical ctypes binding (gold_b200/_capi.py).
//
// Build: CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/gold_b200/lib -lgoldq" go build ./...
package goldq

/*
#cgo LDFLAGS: -lgoldq
#include <stdlib.h>
#include "goldq.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Network selectors, precisions and paths mirror the GOLDQ_* constants.
const (
	NetOnline = int(C.GOLDQ_NET_ONLINE)
	NetTarget = int(C.GOLDQ_NET_TARGET)
	PrecFP32  = int(C.GOLDQ_PREC_FP32)
	PrecBF16  = int(C.GOLDQ_PREC_BF16)

	ArchMLP       = int(C.GOLDQ_ARCH_MLP)
	ArchAtariConv = int(C.GOLDQ_ARCH_ATARI_CONV)

	LossMSE         = int(C.GOLDQ_LOSS_MSE)
	LossPseudoHuber = int(C.GOLDQ_LOSS_PSEUDO_HUBER)
	SolverAdam      = int(C.GOLDQ_SOLVER_ADAM)
	SolverRMSProp   = int(C.GOLDQ_SOLVER_RMSPROP)
)

// ErrNotReady is returned by Learn when the replay holds fewer events than the
// batch; deepq.Agent.Learn maps it to a nil return like the reference
// (pkg/v1/agent/deepq/agent.go:127-129).
var ErrNotReady = fmt.Errorf("goldq: replay shorter than batch")

// Config is the Go view of goldq_config_t.
type Config struct {
	ObsDim, Hidden1, Hidden2, NActions int
	BatchSize, GlobalBatch             int
	ReplayCapacity                     int64
	Gamma                              float32
	LR, Beta1, Beta2, Eps              float64
	Precision, Path, Device, NReplicas int
	Loss, Solver                       int     // GOLDQ_LOSS_* / GOLDQ_SOLVER_*; RMSProp reads LR, Beta2 (= rho), Eps
	HuberDelta                         float32 // PseudoHuberLoss.Delta
	// Arch == ArchAtariConv: deepq.DefaultAtariLayerBuilder's stack on (InChannels, InHeight, InWidth) states;
	// ObsDim = InChannels*InHeight*InWidth, Hidden1 / Hidden2 ignored
	Arch, InChannels, InHeight, InWidth int
}

// DefaultConfig returns gold's defaults (CartPole shape, B=20, gamma .95, Adam 5e-4).
func DefaultConfig() Config {
	var c C.goldq_config_t
	C.goldq_default_config(&c)
	return Config{
		ObsDim: int(c.obs_dim), Hidden1: int(c.hidden1), Hidden2: int(c.hidden2), NActions: int(c.n_actions),
		BatchSize: int(c.batch_size), GlobalBatch: int(c.global_batch), ReplayCapacity: int64(c.replay_capacity),
		Gamma: float32(c.gamma), LR: float64(c.lr), Beta1: float64(c.beta1), Beta2: float64(c.beta2), Eps: float64(c.eps),
		Precision: int(c.precision), Path: int(c.path), Device: int(c.device), NReplicas: int(c.n_replicas),
		Loss: int(c.loss), Solver: int(c.solver), HuberDelta: float32(c.huber_delta),
		Arch: int(c.arch), InChannels: int(c.in_channels), InHeight: int(c.in_height), InWidth: int(c.in_width),
	}
}

// Handle owns one goldq_t (one agent, or NReplicas independent agents, on one GPU).
// Calls on a Handle are not re-entrant; the library selects its device on every
// call, so goroutine-to-thread migration is harmless.
type Handle struct {
	h *C.goldq_t
	P int
	c Config
}

// New allocates both networks, the Adam state and the replay ring on the device.
func New(cfg Config) (*Handle, error) {
	var c C.goldq_config_t
	C.goldq_default_config(&c)
	c.obs_dim, c.hidden1, c.hidden2, c.n_actions = C.int32_t(cfg.ObsDim), C.int32_t(cfg.Hidden1), C.int32_t(cfg.Hidden2), C.int32_t(cfg.NActions)
	c.batch_size, c.global_batch = C.int32_t(cfg.BatchSize), C.int32_t(cfg.GlobalBatch)
	c.replay_capacity = C.int64_t(cfg.ReplayCapacity)
	c.gamma = C.float(cfg.Gamma)
	c.lr, c.beta1, c.beta2, c.eps = C.double(cfg.LR), C.double(cfg.Beta1), C.double(cfg.Beta2), C.double(cfg.Eps)
	c.precision, c.path, c.device, c.n_replicas = C.int32_t(cfg.Precision), C.int32_t(cfg.Path), C.int32_t(cfg.Device), C.int32_t(cfg.NReplicas)
	c.loss, c.solver, c.huber_delta = C.int32_t(cfg.Loss), C.int32_t(cfg.Solver), C.float(cfg.HuberDelta)
	c.arch, c.in_channels, c.in_height, c.in_width = C.int32_t(cfg.Arch), C.int32_t(cfg.InChannels), C.int32_t(cfg.InHeight), C.int32_t(cfg.InWidth)
	h := C.goldq_create(&c)
	if h == nil {
		return nil, fmt.Errorf("goldq_create: %s", C.GoString(C.goldq_last_error(nil)))
	}
	out := &Handle{h: h, P: int(C.goldq_param_count(h)), c: cfg}
	runtime.SetFinalizer(out, func(x *Handle) { x.Close() })
	return out, nil
}

// Close releases all device memory.
func (x *Handle) Close() {
	if x.h != nil {
		C.goldq_destroy(x.h)
		x.h = nil
	}
}

func (x *Handle) err(rc C.int) error {
	if rc == 0 {
		return nil
	}
	if rc == C.GOLDQ_ENOTREADY {
		return ErrNotReady
	}
	return fmt.Errorf("goldq error %d: %s", int(rc), C.GoString(C.goldq_last_error(x.h)))
}

func f32(p []float32) *C.float { return (*C.float)(unsafe.Pointer(&p[0])) }

// SetParams / GetParams move the flat learnables [W1,b1,W2,b2,W3,b3] of one network.
func (x *Handle) SetParams(net int, theta []float32) error {
	return x.err(C.goldq_set_params(x.h, C.int(net), f32(theta)))
}
func (x *Handle) GetParams(net int) ([]float32, error) {
	out := make([]float32, x.P*max1(x.c.NReplicas))
	return out, x.err(C.goldq_get_params(x.h, C.int(net), f32(out)))
}

// ReplayPush appends n transitions (oldest first); each becomes logical index 0.
func (x *Handle) ReplayPush(n int, s []float32, a []int32, r []float32, s2 []float32, done []uint8) error {
	return x.err(C.goldq_replay_push(x.h, C.int64_t(n), f32(s), (*C.int32_t)(unsafe.Pointer(&a[0])), f32(r), f32(s2),
		(*C.uint8_t)(unsafe.Pointer(&done[0]))))
}
func (x *Handle) ReplayLen() int { return int(C.goldq_replay_len(x.h)) }

// PackedBytes is the size of the one-buffer layout s|s2|a|r|done for n transitions (goldq.h).
func (x *Handle) PackedBytes(n int) int { return int(C.goldq_packed_bytes(x.h, C.int64_t(n))) }

// ReplayPushPacked appends n transitions from ONE contiguous buffer: one host-to-device copy instead of five.
func (x *Handle) ReplayPushPacked(n int, packed []byte) error {
	return x.err(C.goldq_replay_push_packed(x.h, C.int64_t(n), unsafe.Pointer(&packed[0])))
}

// ReplayStagePacked starts the host-to-device copy of a packed batch without touching the replay memory;
// ReplayCommit appends the oldest staged batch (see include/goldq.h).  ReplayPushPacked = both.
func (x *Handle) ReplayStagePacked(n int, packed []byte) error {
	return x.err(C.goldq_replay_stage_packed(x.h, C.int64_t(n), unsafe.Pointer(&packed[0])))
}
func (x *Handle) ReplayCommit() error { return x.err(C.goldq_replay_commit(x.h)) }

// ResizeBatch is Sequential.ResizeBatch: scratch sized by the batch is re-allocated, parameters, solver state
// and replay are kept.
func (x *Handle) ResizeBatch(n int) error {
	if err := x.err(C.goldq_resize_batch(x.h, C.int32_t(n))); err != nil {
		return err
	}
	x.c.BatchSize = n
	return nil
}
func (x *Handle) BatchSize() int { return int(C.goldq_batch_size(x.h)) }

// LearnPrepare instantiates the step graphs now instead of on the first Learn.
func (x *Handle) LearnPrepare() error { return x.err(C.goldq_learn_prepare(x.h)) }

// Config returns the configuration the handle was created with (batch size kept current).
func (x *Handle) Config() Config { return x.c }

// Learn runs one learn step; idx == nil draws the indices on the device from seed.
func (x *Handle) Learn(idx []int32, seed uint64) error {
	if idx == nil {
		return x.err(C.goldq_learn(x.h, nil, C.uint64_t(seed)))
	}
	return x.err(C.goldq_learn(x.h, (*C.int32_t)(unsafe.Pointer(&idx[0])), C.uint64_t(seed)))
}

// LearnMany enqueues n learn steps as CUDA-graph launches.
func (x *Handle) LearnMany(n int, seed uint64) error {
	return x.err(C.goldq_learn_many(x.h, C.int32_t(n), C.uint64_t(seed)))
}
func (x *Handle) SyncTarget() error { return x.err(C.goldq_sync_target(x.h)) }
func (x *Handle) LastLoss() (float32, error) {
	out := make([]float32, max1(x.c.NReplicas))
	return out[0], x.err(C.goldq_last_loss(x.h, f32(out)))
}

// Predict returns a fresh [n][A] slice owned by the caller, as goro returns a clone.
func (x *Handle) Predict(net int, n int, in []float32) ([]float32, error) {
	q := make([]float32, n*x.c.NActions)
	return q, x.err(C.goldq_predict(x.h, C.int(net), 0, C.int64_t(n), f32(in), f32(q)))
}
func (x *Handle) GreedyAction(in []float32) (int, error) {
	var a C.int32_t
	rc := C.goldq_greedy_action(x.h, 0, 1, f32(in), &a)
	return int(a), x.err(rc)
}
// EpsilonGreedyAction is Agent.Action for n states at once (vectorised environments).
func (x *Handle) EpsilonGreedyAction(n int, in []float32, epsilon float32, seed uint64) ([]int32, error) {
	a := make([]int32, n)
	rc := C.goldq_epsilon_greedy_action(x.h, 0, C.int64_t(n), f32(in), C.float(epsilon), C.uint64_t(seed), (*C.int32_t)(unsafe.Pointer(&a[0])))
	return a, x.err(rc)
}
func (x *Handle) FitBatch(n int, in, y []float32) error {
	return x.err(C.goldq_fit_batch(x.h, C.int64_t(n), f32(in), f32(y)))
}

// Data-parallel group (one Handle per GPU/process).  DpUniqueID on rank 0, distribute the 128 bytes
// by any side channel, DpInit everywhere; DpMode: 0 none, 1 ncclAllReduce, 2 peer-memory exchange.
func DpUniqueID() ([128]byte, error) {
	var id [128]byte
	if rc := C.goldq_dp_unique_id((*C.uint8_t)(unsafe.Pointer(&id[0]))); rc != 0 {
		return id, fmt.Errorf("goldq_dp_unique_id: %s", C.GoString(C.goldq_last_error(nil)))
	}
	return id, nil
}
func (x *Handle) DpInit(id [128]byte, rank, world int) error {
	return x.err(C.goldq_dp_init(x.h, (*C.uint8_t)(unsafe.Pointer(&id[0])), C.int(rank), C.int(world)))
}
func (x *Handle) DpMode() int { return int(C.goldq_dp_mode(x.h)) }

func max1(n int) int {
	if n < 1 {
		return 1
	}
	return n
}
