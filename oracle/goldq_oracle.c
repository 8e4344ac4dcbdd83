/* goldq_oracle.c — CPU restatement of gold's DQN learn step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gold_b200/, include/,
 * go/) may import, link or call this file.  Permitted users: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
 * and there only as the checker or the timed CPU arm.
 *
 * What it restates (paths relative to /root/reference; V/ = vendor/):
 *   pkg/v1/agent/deepq/agent.go:126-190      Agent.Learn, updateTarget
 *   pkg/v1/agent/deepq/memory.go:54-69       Memory.Sample (index convention only)
 *   pkg/v1/dense/op.go:42-72                 AMaxF32, Concat
 *   V/github.com/aunum/goro/pkg/v1/model/model.go:500-512,552-573   Predict, FitBatch
 *   V/github.com/aunum/goro/pkg/v1/layer/fc.go:119-161              FC forward
 *   V/github.com/aunum/goro/pkg/v1/model/loss.go:28-42              MSE
 *   V/gorgonia.org/gorgonia/nn.go:162-189                           Rectify
 *   V/gorgonia.org/gorgonia/solvers.go:440-617                      AdamSolver.Step
 *   V/gorgonia.org/tensor/internal/execution/generic_argmethods.go:211-233  ArgmaxF32
 *   V/gonum.org/v1/gonum/blas/gonum/sgemm.go:258-300                Sgemm kernels
 * Third-party code it restates is vendored in the reference tree at the pinned
 * versions gorgonia v0.9.9, tensor v0.9.4, gonum v0.7.0, goro 6fd962e92fb9,
 * vecf32 v0.9.0 (go.mod:5-25).
 *
 * PARITY PINNING.  The reference is Go; no Go toolchain exists in this image or
 * on the GPU boxes, and the reference's own tests pin no numeric result of the
 * learn step (pkg/v1/agent/deepq/policy_test.go asserts nothing numeric).  The
 * only known-answer material the reference holds for this path is
 * pkg/v1/dense/op_test.go:12-18 (AMax{5,8,3,2}=8), :20-31 (Concat),
 * pkg/v1/agent/deepq/memory_test.go:28 (Sample length) — all checked in
 * tests/test_oracle_kat.py.  Q-values, loss, gradients and post-Adam weights are
 * therefore PARITY UNPINNED against executed reference output; they are pinned
 * only by this line-by-line restatement, by closed-form known answers and by an
 * independent autograd cross-check (tests/test_oracle_crosscheck.py).
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { int32_t D, H1, H2, A; } oq_dims_t;

typedef struct {
    oq_dims_t dims;
    int32_t B;
    float gamma;
    double lr, beta1, beta2, eps; /* AdamSolver keeps float64 fields (solvers.go:399-415) */
    int32_t loss;                 /* 0: MSE (loss.go:28-42); 1: pseudo-Huber (loss.go:96-151, reduced by Mean) */
    int32_t solver;               /* 0: AdamSolver; 1: RMSPropSolver (lr = eta, beta2 = rho, eps) */
    float delta;                  /* PseudoHuberLoss.Delta, a float32 (loss.go:98) */
} oq_cfg_t;

typedef struct { int64_t w1, b1, w2, b2, w3, b3, total; } oq_offsets_t;

static int oq_threads = 1;

/* Learnable order [W1,b1,W2,b2,W3,b3] (V/github.com/aunum/goro/pkg/v1/layer/chain.go:38-44,
 * fc.go:164-169); W is (in×out) row-major (fc.go:98), bias (1×out). */
static oq_offsets_t oq_offsets(const oq_dims_t *d)
{
    oq_offsets_t o;
    o.w1 = 0;
    o.b1 = o.w1 + (int64_t)d->D * d->H1;
    o.w2 = o.b1 + d->H1;
    o.b2 = o.w2 + (int64_t)d->H1 * d->H2;
    o.w3 = o.b2 + d->H2;
    o.b3 = o.w3 + (int64_t)d->H2 * d->A;
    o.total = o.b3 + d->A;
    return o;
}
static int64_t oq_param_count(const oq_dims_t *d) { return oq_offsets(d).total; }

/* ArgmaxF32 (generic_argmethods.go:211-233): first maximum under strict '>',
 * except that a NaN or +Inf at index >= 1 wins immediately. */
static int argmax_f32(const float *a, int n)
{
    int set = 0, mx = 0;
    float f = 0;
    for (int i = 0; i < n; i++) {
        float v = a[i];
        if (!set) { f = v; mx = i; set = 1; continue; }
        if (isnan(v) || (isinf(v) && v > 0)) return i;
        if (v > f) { mx = i; f = v; }
    }
    return mx;
}
static int argmax_f64(const double *a, int n)
{
    int set = 0, mx = 0;
    double f = 0;
    for (int i = 0; i < n; i++) {
        double v = a[i];
        if (!set) { f = v; mx = i; set = 1; continue; }
        if (isnan(v) || (isinf(v) && v > 0)) return i;
        if (v > f) { mx = i; f = v; }
    }
    return mx;
}

/* Sum over axis 0 (bias gradient: repeatOp.SymDiff → Sum(grad, 0),
 * V/gorgonia.org/gorgonia/op_tensor.go:281-289): rows added in order. */
static void colsum_f32(int n, int c, const float *x, float *out)
{
    for (int j = 0; j < c; j++) out[j] = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < c; j++) out[j] = out[j] + x[(size_t)i * c + j];
}
static void colsum_f64(int n, int c, const double *x, double *out)
{
    for (int j = 0; j < c; j++) out[j] = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < c; j++) out[j] = out[j] + x[(size_t)i * c + j];
}

#define REAL float
#define FN(name) name##_f32
#define SQRT(x) sqrtf(x)
#include "learn_core.inc"
#include "conv_core.inc"
#undef REAL
#undef FN
#undef SQRT

#define REAL double
#define FN(name) name##_f64
#define SQRT(x) sqrt(x)
#include "learn_core.inc"
#include "conv_core.inc"
#undef REAL
#undef FN
#undef SQRT

/* ======================= exported C entry points ======================== */

void oq_set_threads(int n)
{
    oq_threads = n < 1 ? 1 : n;
}
int oq_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int64_t oq_nparams(int D, int H1, int H2, int A)
{
    oq_dims_t d = {D, H1, H2, A};
    return oq_param_count(&d);
}

int oq_argmax_f32(const float *a, int n) { return argmax_f32(a, n); }

/* dense.AMaxF32 (pkg/v1/dense/op.go:42-62). */
float oq_amax_f32(const float *a, int n) { return a[argmax_f32(a, n)]; }

/* Greedy action = Argmax(1) of the online prediction (agent.go:207-221). */
void oq_predict_f32(int D, int H1, int H2, int A, const float *theta, int n,
                    const float *x, float *q, int32_t *greedy /* nullable */)
{
    oq_dims_t d = {D, H1, H2, A};
    float *h1 = (float *)malloc(sizeof(float) * (size_t)n * (H1 + H2));
    float *h2 = h1 + (size_t)n * H1;
    forward_f32(&d, theta, n, x, h1, NULL, h2, NULL, q);
    if (greedy)
        for (int i = 0; i < n; i++) greedy[i] = argmax_f32(q + (size_t)i * A, A);
    free(h1);
}

void oq_predict_f64(int D, int H1, int H2, int A, const double *theta, int n,
                    const double *x, double *q)
{
    oq_dims_t d = {D, H1, H2, A};
    double *h1 = (double *)malloc(sizeof(double) * (size_t)n * (H1 + H2));
    double *h2 = h1 + (size_t)n * H1;
    forward_f64(&d, theta, n, x, h1, NULL, h2, NULL, q);
    free(h1);
}

void oq_adam_f32(int64_t n, float *w, float *g, float *m, float *v, int64_t iter,
                 double lr, double beta1, double beta2, double eps)
{
    adam_f32(n, w, g, m, v, iter, lr, beta1, beta2, eps);
}

/* One Agent.Learn on a sampled batch.  iter is the AdamSolver's counter BEFORE
 * the step; it is incremented first (solvers.go:445).  theta/m/v are updated in
 * place. */
int oq_learn_f32(const oq_cfg_t *cfg, float *theta, const float *theta_t, float *m,
                 float *v, int64_t *iter, const float *s, const int32_t *a,
                 const float *r, const float *s2, const uint8_t *done, float *q_online,
                 float *q_target, float *td, float *loss, float *grads)
{
    *iter += 1;
    return learn_f32(cfg, theta, theta_t, m, v, *iter, s, a, r, s2, done, q_online,
                     q_target, td, loss, grads);
}

int oq_learn_f64(const oq_cfg_t *cfg, double *theta, const double *theta_t, double *m,
                 double *v, int64_t *iter, const double *s, const int32_t *a,
                 const double *r, const double *s2, const uint8_t *done,
                 double *q_online, double *q_target, double *td, double *loss,
                 double *grads)
{
    *iter += 1;
    return learn_f64(cfg, theta, theta_t, m, v, *iter, s, a, r, s2, done, q_online,
                     q_target, td, loss, grads);
}

/* Sequential.FitBatch (n = batch) / Fit (n = 1) on explicit labels. */
int oq_fit_f32(const oq_cfg_t *cfg, int n, float *theta, float *m, float *v, int64_t *iter,
               const float *x, const float *y, float *loss, float *grads)
{
    *iter += 1;
    return fit_f32(cfg, n, theta, m, v, *iter, x, y, loss, grads);
}

int oq_fit_f64(const oq_cfg_t *cfg, int n, double *theta, double *m, double *v, int64_t *iter,
               const double *x, const double *y, double *loss, double *grads)
{
    *iter += 1;
    return fit_f64(cfg, n, theta, m, v, *iter, x, y, loss, grads);
}

/* Replay indexing convention of deepq.Memory: Remember = PushFront
 * (agent.go:224-229), Sample reads At(i) (memory.go:62) and At(0) is the front
 * (V/github.com/gammazero/deque/deque.go:112-118), so logical index i is the
 * i-th NEWEST event.  `order` holds the transitions in arrival order (oldest
 * first), `len` of them; gathers rows idx[0..B) into batch-major arrays
 * (= dense.Concat(0, …), pkg/v1/dense/op.go:65-72). */
void oq_gather_f32(int D, int64_t len, const float *order_s, const int32_t *order_a,
                   const float *order_r, const float *order_s2, const uint8_t *order_done,
                   int B, const int32_t *idx, float *s, int32_t *a, float *r, float *s2,
                   uint8_t *done)
{
    for (int b = 0; b < B; b++) {
        int64_t src = len - 1 - idx[b];
        memcpy(s + (size_t)b * D, order_s + (size_t)src * D, sizeof(float) * D);
        memcpy(s2 + (size_t)b * D, order_s2 + (size_t)src * D, sizeof(float) * D);
        a[b] = order_a[src];
        r[b] = order_r[src];
        done[b] = order_done[src];
    }
}

/* ---- the Atari (convolutional) policy: pkg/v1/agent/deepq/policy.go:41-47,62-71 (conv_core.inc) -------------- */
int64_t oq_atari_nparams(int C, int H, int W, int A) { return atari_make_f32(C, H, W, A).total; }
int oq_atari_features(int C, int H, int W) { return atari_make_f32(C, H, W, 2).F; }

void oq_atari_predict_f32(int C, int H, int W, int A, const float *theta, int n, const float *x, float *q)
{
    atari_t_f32 t = atari_make_f32(C, H, W, A);
    atari_acts_f32 a = atari_alloc_f32(&t, n);
    atari_forward_f32(&t, theta, n, x, &a, q);
    atari_free_f32(&a);
}
void oq_atari_predict_f64(int C, int H, int W, int A, const double *theta, int n, const double *x, double *q)
{
    atari_t_f64 t = atari_make_f64(C, H, W, A);
    atari_acts_f64 a = atari_alloc_f64(&t, n);
    atari_forward_f64(&t, theta, n, x, &a, q);
    atari_free_f64(&a);
}
int oq_atari_learn_f32(const oq_cfg_t *cfg, int C, int H, int W, float *theta, const float *theta_t, float *m, float *v,
                       int64_t *iter, const float *s, const int32_t *act, const float *rew, const float *s2,
                       const uint8_t *done, float *q_online, float *q_target, float *td, float *loss, float *grads)
{
    *iter += 1;
    return atari_learn_f32(cfg, C, H, W, theta, theta_t, m, v, *iter, s, act, rew, s2, done, q_online, q_target, td, loss, grads);
}
int oq_atari_learn_f64(const oq_cfg_t *cfg, int C, int H, int W, double *theta, const double *theta_t, double *m, double *v,
                       int64_t *iter, const double *s, const int32_t *act, const double *rew, const double *s2,
                       const uint8_t *done, double *q_online, double *q_target, double *td, double *loss, double *grads)
{
    *iter += 1;
    return atari_learn_f64(cfg, C, H, W, theta, theta_t, m, v, *iter, s, act, rew, s2, done, q_online, q_target, td, loss, grads);
}
