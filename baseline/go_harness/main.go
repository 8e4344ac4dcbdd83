// UNEXECUTED: there is no Go toolchain in the image this repository is developed in.
//
// Reference harness: runs gold's DQN learn step on REAL Gorgonia (v0.9.9, the version gold
// vendors) for the fixtures under tests/golden/ and dumps what the oracle claims to reproduce
// (Q(s), Q'(s'), TD targets, loss, gradients, post-Adam parameters), so that whoever has Go can
// pin oracle/goldq_oracle.c against the executed reference:
//
//	python tests/golden/make_golden.py --bin /tmp/golden_bin        # inputs as raw little-endian files
//	cp -r baseline/go_harness $GOLD/cmd/goldq_harness                # inside a checkout of aunum/gold
//	(cd $GOLD && go run -mod=vendor ./cmd/goldq_harness -case /tmp/golden_bin/cartpole_b20 -out /tmp/ref_dump)
//	python baseline/go_harness/compare.py tests/golden/cartpole_b20.npz /tmp/ref_dump
//
// It imports gorgonia and tensor only: gold's deepq package and goro's model package do not build
// from the vendored tree (model -> gold/pkg/v1/track -> gonum/plot, whose font blob is missing,
// .MISSING_LARGE_BLOBS).  The graphs are therefore rebuilt here exactly as goro builds them:
//   - layer.FC.Fwd (goro layer/fc.go:119-161): Mul(x, W), then Add (single row) or
//     BroadcastAdd(.., []byte{0}) (batch), then Rectify for hidden layers (activation.go:81-94);
//   - MSELoss.Compute (goro model/loss.go:28-42): Mean(Square(Sub(yHat, y)));
//   - Sequential.FitBatch (goro model/model.go:552-573): RunAll, NodesToValueGrads, solver.Step, Reset;
//   - deepq.Agent.Learn (gold pkg/v1/agent/deepq/agent.go:126-178) around them;
//   - learnables shared between graphs by value, as layer/fc.go:92-96 does.
package main

import (
	"encoding/binary"
	"flag"
	"fmt"
	"io/ioutil"
	"math"
	"os"
	"path/filepath"

	g "gorgonia.org/gorgonia"
	"gorgonia.org/tensor"
)

func must(err error) {
	if err != nil {
		panic(err)
	}
}

func readF32(path string) []float32 {
	b, err := ioutil.ReadFile(path)
	must(err)
	out := make([]float32, len(b)/4)
	for i := range out {
		out[i] = math.Float32frombits(binary.LittleEndian.Uint32(b[4*i:]))
	}
	return out
}

func readI32(path string) []int32 {
	b, err := ioutil.ReadFile(path)
	must(err)
	out := make([]int32, len(b)/4)
	for i := range out {
		out[i] = int32(binary.LittleEndian.Uint32(b[4*i:]))
	}
	return out
}

func writeF32(path string, v []float32) {
	b := make([]byte, 4*len(v))
	for i, x := range v {
		binary.LittleEndian.PutUint32(b[4*i:], math.Float32bits(x))
	}
	must(ioutil.WriteFile(path, b, 0644))
}

// one compiled graph of the 3-layer policy
type net struct {
	graph      *g.ExprGraph
	x, y       *g.Node
	learnables g.Nodes
	pred, loss *g.Node
	predVal    g.Value
	lossVal    g.Value
	vm         g.VM
}

// vals: W1 (D x H1), b1 (1 x H1), W2, b2, W3, b3 -- shared by every graph of one policy
func build(dims [4]int, rows int, vals []tensor.Tensor, train bool) *net {
	n := &net{graph: g.NewGraph()}
	n.x = g.NewMatrix(n.graph, tensor.Float32, g.WithShape(rows, dims[0]), g.WithName("state"))
	h := n.x
	for l := 0; l < 3; l++ {
		w := g.NewMatrix(n.graph, tensor.Float32, g.WithShape(dims[l], dims[l+1]), g.WithName(fmt.Sprintf("fc%d", l)), g.WithValue(vals[2*l]))
		b := g.NewMatrix(n.graph, tensor.Float32, g.WithShape(1, dims[l+1]), g.WithName(fmt.Sprintf("fc%d-bias", l)), g.WithValue(vals[2*l+1]))
		n.learnables = append(n.learnables, w, b)
		xw, err := g.Mul(h, w)
		must(err)
		var xwb *g.Node
		if rows > 1 {
			xwb, err = g.BroadcastAdd(xw, b, nil, []byte{0})
		} else {
			xwb, err = g.Add(xw, b)
		}
		must(err)
		if l < 2 {
			h, err = g.Rectify(xwb)
			must(err)
		} else {
			h = xwb
		}
	}
	n.pred = h
	g.Read(n.pred, &n.predVal)
	if !train {
		n.vm = g.NewTapeMachine(n.graph)
		return n
	}
	n.y = g.NewMatrix(n.graph, tensor.Float32, g.WithShape(rows, dims[3]), g.WithName("actionPotentials"))
	d, err := g.Sub(n.pred, n.y)
	must(err)
	sq, err := g.Square(d)
	must(err)
	n.loss, err = g.Mean(sq)
	must(err)
	g.Read(n.loss, &n.lossVal)
	_, err = g.Grad(n.loss, n.learnables...)
	must(err)
	n.vm = g.NewTapeMachine(n.graph, g.BindDualValues(n.learnables...))
	return n
}

func (n *net) predict(row []float32, d int) []float32 {
	x := tensor.New(tensor.WithShape(1, d), tensor.WithBacking(append([]float32(nil), row...)))
	must(g.Let(n.x, x))
	must(n.vm.RunAll())
	out := append([]float32(nil), n.predVal.Data().([]float32)...) // Predict's result is a clone (vm_tape.go:691-697)
	n.vm.Reset()
	return out
}

func params(vals []tensor.Tensor) []float32 {
	var out []float32
	for _, v := range vals {
		out = append(out, v.Data().([]float32)...)
	}
	return out
}

func newVals(dims [4]int, flat []float32) []tensor.Tensor {
	var vals []tensor.Tensor
	off := 0
	for l := 0; l < 3; l++ {
		nw := dims[l] * dims[l+1]
		vals = append(vals, tensor.New(tensor.WithShape(dims[l], dims[l+1]), tensor.WithBacking(append([]float32(nil), flat[off:off+nw]...))))
		off += nw
		vals = append(vals, tensor.New(tensor.WithShape(1, dims[l+1]), tensor.WithBacking(append([]float32(nil), flat[off:off+dims[l+1]]...))))
		off += dims[l+1]
	}
	return vals
}

func main() {
	dir := flag.String("case", "", "directory written by make_golden.py --bin")
	out := flag.String("out", "ref_dump", "output directory")
	flag.Parse()
	must(os.MkdirAll(*out, 0755))
	in := func(name string) string { return filepath.Join(*dir, name) }
	meta := readI32(in("meta.i32")) // D, H1, H2, A, B, N, steps
	dims := [4]int{int(meta[0]), int(meta[1]), int(meta[2]), int(meta[3])}
	B, N, steps := int(meta[4]), int(meta[5]), int(meta[6])
	D, A := dims[0], dims[3]
	gamma := readF32(in("gamma.f32"))[0]
	s, s2, r := readF32(in("s.f32")), readF32(in("s2.f32")), readF32(in("r.f32"))
	act, done := readI32(in("a.i32")), readI32(in("done.i32"))

	online := newVals(dims, readF32(in("theta_init.f32")))
	target := newVals(dims, readF32(in("theta_t_init.f32")))
	onlineNet := build(dims, 1, online, false)
	targetNet := build(dims, 1, target, false)
	trainNet := build(dims, B, online, true)
	solver := g.NewAdamSolver(g.WithLearnRate(0.0005)) // policy.go:34

	for step := 0; step < steps; step++ {
		idx := readI32(in(fmt.Sprintf("idx%d.i32", step))) // logical index i = i-th newest event (deque.At(i) after PushFront)
		states := make([]float32, 0, B*D)
		labels := make([]float32, 0, B*A)
		var qOn, qTg, td []float32
		for _, li := range idx {
			row := N - 1 - int(li) // arrays are in arrival order
			qUpdate := r[row]
			qt := make([]float32, A)
			if done[row] == 0 {
				qt = targetNet.predict(s2[row*D:(row+1)*D], D)
				qv := tensor.New(tensor.WithShape(1, A), tensor.WithBacking(append([]float32(nil), qt...)))
				am, err := qv.Argmax(1) // dense.AMaxF32 (pkg/v1/dense/op.go:42-62)
				must(err)
				j := am.Data().([]int)[0]
				qUpdate = r[row] + gamma*qt[j]
			}
			q := onlineNet.predict(s[row*D:(row+1)*D], D)
			qOn = append(qOn, q...)
			qTg = append(qTg, qt...)
			td = append(td, qUpdate)
			q[act[row]] = qUpdate
			states = append(states, s[row*D:(row+1)*D]...)
			labels = append(labels, q...)
		}
		must(g.Let(trainNet.x, tensor.New(tensor.WithShape(B, D), tensor.WithBacking(states))))
		must(g.Let(trainNet.y, tensor.New(tensor.WithShape(B, A), tensor.WithBacking(labels))))
		must(trainNet.vm.RunAll())
		var grads []float32
		for _, n := range trainNet.learnables {
			gv, err := n.Grad()
			must(err)
			grads = append(grads, gv.Data().([]float32)...)
		}
		loss := trainNet.lossVal.Data().(float32)
		solver.Step(g.NodesToValueGrads(trainNet.learnables))
		trainNet.vm.Reset()

		w := func(name string, v []float32) { writeF32(filepath.Join(*out, fmt.Sprintf("%s%d.f32", name, step)), v) }
		w("q_online", qOn)
		w("q_target", qTg) // rows of terminal events are zeros here and ignored by compare.py
		w("td", td)
		w("loss", []float32{loss})
		w("grads", grads)
		w("theta", params(online))
		if step == 1 { // CloneLearnablesTo (model.go:606-636)
			for i := range online {
				copy(target[i].Data().([]float32), online[i].Data().([]float32))
			}
		}
	}
	fmt.Println("dumped", steps, "steps to", *out)
}
