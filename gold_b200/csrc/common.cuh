// common.cuh — internal declarations shared by the libgoldq translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/goldq.h"

namespace gq {

struct Dims {
    int D, H1, H2, A;
};

// Flat learnable layout [W1,b1,W2,b2,W3,b3]; W is (in x out) row-major.
// Reference: vendor/github.com/aunum/goro/pkg/v1/layer/fc.go:98-101,164-169.
struct Offsets {
    int w1, b1, w2, b2, w3, b3, P;
};

__host__ __device__ inline Offsets make_offsets(const Dims &d)
{
    Offsets o;
    o.w1 = 0;
    o.b1 = o.w1 + d.D * d.H1;
    o.w2 = o.b1 + d.H1;
    o.b2 = o.w2 + d.H1 * d.H2;
    o.w3 = o.b2 + d.H2;
    o.b3 = o.w3 + d.H2 * d.A;
    o.P = o.b3 + d.A;
    return o;
}

// Replay ring, structure of arrays.  Per replica r the arrays start at r*cap.
// head = next physical slot to write; logical index i (i-th newest, the
// reference's deque.At(i) after PushFront) lives at (head - 1 - i) mod cap.
struct Replay {
    float *s;       // [R][cap][D]
    float *s2;      // [R][cap][D]
    int32_t *a;     // [R][cap]
    float *r;       // [R][cap]
    uint8_t *done;  // [R][cap]
    int64_t cap;
    int64_t len;
    int64_t head;
};

// Adam constants, computed on the host exactly as the reference does
// (vendor/gorgonia.org/gorgonia/solvers.go:446-447,490-503): the bias corrections
// in float64, then cast.
struct AdamConsts {
    float beta1, beta2, omb1, omb2, eps, neg_eta, c1, c2;
    int kind;       // GOLDQ_SOLVER_ADAM / GOLDQ_SOLVER_RMSPROP (then beta2 = decay, omb2 = float32(1 - decay); cache in v)
};

// One element of Solver.Step, op for op, one rounding per op and no FMA (shared by the fused narrow kernel, the
// wide update kernels and the generic path).
//   Adam     vendor/gorgonia.org/gorgonia/solvers.go:551-615, including v0.9.9's double weight update
//   RMSProp  solvers.go:273-335: Square(g); cw *= decay; g2 *= (1 - decay); cw += g2; upd = cw + eps;
//            InvSqrt(upd) = 1/sqrt (tensor generic_unary.go:492-496); g *= -eta; upd *= g; w += upd
__device__ __forceinline__ void solver_update(float &w, float g, float &m, float &v, const AdamConsts &c)
{
    if (c.kind == GOLDQ_SOLVER_RMSPROP) {
        const float cw = __fadd_rn(__fmul_rn(v, c.beta2), __fmul_rn(__fmul_rn(g, g), c.omb2));
        v = cw;
        const float inv = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(cw, c.eps)));
        w = __fadd_rn(w, __fmul_rn(inv, __fmul_rn(g, c.neg_eta)));
        return;
    }
    const float t1 = __fmul_rn(g, c.omb1);
    const float g2 = __fmul_rn(__fmul_rn(g, g), c.omb2);
    m = __fadd_rn(t1, __fmul_rn(c.beta1, m));
    v = __fadd_rn(g2, __fmul_rn(c.beta2, v));
    const float mh = __fmul_rn(m, c.c1);
    const float vh = __fmul_rn(v, c.c2);
    const float den = __fadd_rn(__fsqrt_rn(vh), c.eps);
    const float u = __fmul_rn(mh, c.neg_eta);
    if (den == 0.f) w = __int_as_float(0x7f800000);     // vecf32 DivIncr stores +Inf (incr.go:36-38)
    else w = __fadd_rn(w, __fdiv_rn(u, den));
    w = __fadd_rn(w, u);
}

// Per-step scalars that live in DEVICE memory when a step is replayed from a CUDA graph
// (goldq_learn_many): the kernel arguments of a captured step are frozen, so the sampler
// seed and Adam's bias corrections of step i are read from slot i instead.
struct StepDev {
    unsigned long long seed;
    float c1, c2;
    unsigned int epoch;     // data-parallel step counter (peer-memory all-reduce handshake)
    unsigned int pad_;
    long long head, len;    // replay ring position at this step: a Remember between two replays of a
                            // captured step moves the ring without invalidating the graph
};

// Scalars of one learn step.
struct StepScalars {
    float gamma;
    float count;      // float32(B_global * A)   (Mean's divisor)
    float inv_count;  // float32(1) / count
    int loss_kind;    // 0: MSE (goro loss.go:28-42), 1: pseudo-Huber (loss.go:96-151, reduced by Mean)
    float delta;      // pseudo-Huber delta
    float delta2;     // delta * delta
};

// One residual d = q - y: its term of the loss sum and dLoss/dq (Mean's 1/count folded in).
//   MSE           term = d^2                         dq = (d * 2) * (1/count)      (reference op order)
//   pseudo-Huber  term = delta^2 (sqrt(1+(d/delta)^2) - 1)   dq = d / sqrt(1+(d/delta)^2) * (1/count)
__device__ __forceinline__ void residual_loss(const StepScalars &sc, float d, float &term, float &dq)
{
    if (sc.loss_kind == 0) {
        term = __fmul_rn(d, d);
        dq = __fmul_rn(__fmul_rn(d, 2.f), sc.inv_count);
    } else {
        const float x = __fdiv_rn(d, sc.delta);
        const float s = __fsqrt_rn(__fadd_rn(1.f, __fmul_rn(x, x)));
        term = __fmul_rn(sc.delta2, __fsub_rn(s, 1.f));
        dq = __fmul_rn(__fdiv_rn(d, s), sc.inv_count);
    }
}

// ---- generic path (kernels_generic.cu) -----------------------------------
void launch_sample_indices(int32_t *idx, int B, int64_t len, uint64_t seed, const StepDev *step,
                           int replicas, cudaStream_t st);
void launch_gather(const Replay &rp, int D, const int32_t *idx, int B, float *bs, float *bs2,
                   int32_t *ba, float *br, uint8_t *bdone, const StepDev *step, cudaStream_t st);
void launch_fc_forward(const float *X, const float *W, const float *bias, float *Y,
                       uint8_t *mask, int n, int in, int out, bool relu, cudaStream_t st);
// TD label + dQ: dq is dense [B][A] (zero off-action), loss partials per block.
void launch_td(const float *q, const float *qt, const int32_t *a, const float *r,
               const uint8_t *done, int B, int A, StepScalars sc, float *td, float *dq,
               float *loss_part, int *n_loss_part, cudaStream_t st);
// Explicit-label variant (FitBatch): dq = 2 (q - y) / count, dense.
void launch_label_loss(const float *q, const float *y, int n, int A, StepScalars sc, float *dq,
                       float *loss_part, int *n_loss_part, cudaStream_t st);
// dW slices: gslice[z][w_off + i*out + j] = sum over the z-th row range of X[r][i]*dZ[r][j];
// gslice[z][b_off + j] = sum of dZ[r][j].
void launch_grad_w(const float *X, const float *dZ, int n, int in, int out, float *gslice,
                   int64_t slice_stride, int nsplit, int w_off, int b_off, cudaStream_t st);
// dX[n][in] = (dZ[n][out] . W[in][out]^T) * mask
void launch_grad_x(const float *dZ, const float *W, const uint8_t *mask, float *dX, int n, int in,
                   int out, cudaStream_t st);
// flat[i] = sum_z gslice[z][i] (i < P); flat[P] = (sum of loss partials) / count
void launch_reduce_slices(const float *gslice, int64_t slice_stride, int nsplit, int P,
                          const float *loss_part, int n_loss_part, float count, float *flat,
                          cudaStream_t st);
// Fused Adam incl. the reference's double update; replicas laid out [R][P].
void launch_adam(float *w, const float *g, float *m, float *v, int64_t n, AdamConsts c,
                 const StepDev *step, cudaStream_t st);
void launch_argmax_rows(const float *q, int n, int A, int32_t *out, cudaStream_t st);
void launch_eps_greedy_rows(const float *q, int n, int A, float epsilon, uint64_t seed, int32_t *out, cudaStream_t st);

// ---- Atari conv policy, fp32 (kernels_conv.cu) -------------------------------
struct ConvShape {
    int C, H, W, O, KH, KW, pad, stride, OH, OW;
};
// Learnable order = layer order (goro layer/chain.go:38-44): the three filters (O,I,H,W; no bias), then W, b of FC 512
// and FC A.
struct AtariNet {
    ConvShape c1, c2, c3;
    int F, HID, A;
    int64_t f1, f2, f3, w4, b4, w5, b5, P;
};
ConvShape make_conv_shape(int C, int H, int W, int O, int KH, int KW, int pad, int stride);
AtariNet make_atari_net(int C, int H, int W, int A);
void launch_conv_forward(const ConvShape &c, int n, const float *x, const float *f, float *y, uint8_t *mask, cudaStream_t st);
// gslice[z][f_off + e] = partial filter gradient of batch slice z (summed later with the other slices, fixed order)
void launch_conv_backward_filter(const ConvShape &c, int n, int nslice, const float *x, const float *dy, const uint8_t *mask,
                                 float *gslice, int64_t slice_stride, int64_t f_off, cudaStream_t st);
void launch_conv_backward_data(const ConvShape &c, int n, const float *dy, const uint8_t *mask, const float *f, float *dx,
                               cudaStream_t st);

// ---- narrow fused path (kernels_narrow.cu) --------------------------------
#define NARROW_GROUP 8        // CTAs summed by one level-1 leader
#define NARROW_MAX_GROUPS 64  // => at most 512 CTAs per replica
bool narrow_supported(const Dims &d);
size_t narrow_partial_floats(int ctas, int P);
struct NarrowArgs {
    Dims dims;
    Replay rp;
    const int32_t *idx;   // [R][B] or nullptr -> sample on device
    uint64_t seed;
    const StepDev *step;  // non-null: seed / c1 / c2 come from device memory (graph replay)
    int B;
    int replicas;
    float *theta;         // [R][P]
    const float *theta_t; // [R][P]
    float *m, *v;         // [R][P]
    float *partial;       // [R][ctas + groups][P+1] scratch
    unsigned int *ticket; // [R][NARROW_MAX_GROUPS+1] zero-initialised, self-resetting
    float *flat;          // [R][P+1]: reduced gradient + loss (always written)
    float *loss;          // [R]
    int fuse_adam;        // 1: apply Adam in the same kernel; 0: stop at `flat` (data-parallel)
    StepScalars sc;
    AdamConsts adam;
    // debug snapshots (nullptr to skip)
    float *dbg_q, *dbg_qt, *dbg_td;
    int32_t *dbg_idx;
    int ctas_per_replica; // filled by narrow_plan
    int warps_per_cta;
};
void narrow_plan(NarrowArgs &a, int sm_count);
void narrow_configure(const Dims &d, int warps_per_cta);
void launch_narrow_learn(const NarrowArgs &a, cudaStream_t st);

}  // namespace gq
