/* goldq.h — C ABI of the B200-native DQN learn step (libgoldq.so).
 *
 * This is the drop-in boundary for gold's `pkg/v1/agent/deepq` hot path: a Go
 * package (go/deepq, cgo) or any other FFI binds exactly these entry points.
 * Plain pointers and sizes only; no C++/torch types.  Every function that returns
 * `int` returns 0 on success and a negative GOLDQ_E* code on failure; the message
 * is available from goldq_last_error().  The library never retains a caller
 * pointer: host buffers are copied (or the copy is enqueued from a staging area)
 * before the call returns, which is the rule cgo needs.
 *
 * Calls on one handle are not re-entrant (the reference is single-goroutine:
 * global solver / math/rand singletons, pkg/v1/agent/deepq/policy.go:32-38,
 * memory.go:59).  Every entry point selects the handle's device first, because a
 * cgo call may arrive on any OS thread.
 *
 * Reference paths below are relative to the aunum/gold tree; V/ = vendor/.
 */
#ifndef GOLDQ_H
#define GOLDQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOLDQ_ABI_VERSION 2

/* error codes */
#define GOLDQ_OK 0
#define GOLDQ_EINVAL (-1)   /* bad argument / shape mismatch (goro io.go:138-153 returns an error) */
#define GOLDQ_ECUDA (-2)    /* CUDA runtime error */
#define GOLDQ_ENOTREADY (-3)/* replay shorter than the batch (agent.go:127 returns nil: caller maps to no-op) */
#define GOLDQ_EUNSUPPORTED (-4) /* configuration the CUDA path does not implement (no CPU fallback by design) */
#define GOLDQ_ENCCL (-5)

/* which network */
#define GOLDQ_NET_ONLINE 0  /* Agent.Policy        (agent.go:27) */
#define GOLDQ_NET_TARGET 1  /* Agent.TargetPolicy  (agent.go:30) */

/* loss (V/github.com/aunum/goro/pkg/v1/model/loss.go) */
#define GOLDQ_LOSS_MSE 0    /* model.MSE, loss.go:22-42 */
#define GOLDQ_LOSS_PSEUDO_HUBER 1 /* model.PseudoHuber, loss.go:90-151: delta^2 (sqrt(1 + ((q-y)/delta)^2) - 1), reduced
                                   * by Mean like MSE.  The reference version ends in Sum(.., 1), a vector, which
                                   * gorgonia.Grad rejects (the "!blocked" note at loss.go:94), so there is no
                                   * executable reference behaviour to match: this is the loss as it is meant. */

/* solver (V/gorgonia.org/gorgonia/solvers.go) */
#define GOLDQ_SOLVER_ADAM 0    /* AdamSolver.Step incl. v0.9.9's double weight update, solvers.go:440-617 */
#define GOLDQ_SOLVER_RMSPROP 1 /* RMSPropSolver.Step, solvers.go:249-395 (gold's Atari config, policy.go:41-47):
                                * c = c*rho + (1-rho)*g^2; w += (1/sqrt(c+eps)) * (g * -eta).  Hyper-parameters:
                                * lr = eta, beta2 = rho, eps; beta1 unused.  On every path (fused narrow / wide update
                                * kernels included). */

/* arithmetic of the forward/backward pass */
#define GOLDQ_PREC_FP32 0   /* reference arithmetic: fp32, multiply then add, k-sequential */
#define GOLDQ_PREC_BF16 1   /* bf16 operands, fp32 accumulate on tcgen05 tensor cores; fp32 master weights + Adam */

/* kernel path selection (GOLDQ_PATH_AUTO picks by shape and precision) */
#define GOLDQ_PATH_AUTO 0
#define GOLDQ_PATH_GENERIC 1 /* tiled CUDA-core kernels, any shape */
#define GOLDQ_PATH_NARROW 2  /* one fused warp-per-row kernel, classic-control shapes (H<=32) */
#define GOLDQ_PATH_WIDE 3    /* tcgen05 GEMM tiles (requires GOLDQ_PREC_BF16) */
#define GOLDQ_PATH_CONV 4    /* the Atari conv policy (arch = GOLDQ_ARCH_ATARI_CONV), fp32 CUDA-core kernels */

/* network architecture */
#define GOLDQ_ARCH_MLP 0         /* FC-FC-FC, deepq.DefaultFCLayerBuilder (policy.go:53-59) */
#define GOLDQ_ARCH_ATARI_CONV 1  /* deepq.DefaultAtariLayerBuilder (policy.go:62-71): Conv2D 32x8x8/4, 64x4x4/2, 64x3x3/1
                                  * (pad 1, ReLU, no bias), Flatten, FC 512, FC A.  obs_dim = in_channels*in_height*in_width,
                                  * states are NCHW; hidden1 / hidden2 are ignored */

/* goldq_debug_fetch selectors: state of the LAST learn step (parity tests) */
#define GOLDQ_DBG_Q_ONLINE 0  /* float[B*A]  online Q(s)            */
#define GOLDQ_DBG_Q_TARGET 1  /* float[B*A]  target Q(s')           */
#define GOLDQ_DBG_TD 2        /* float[B]    TD targets             */
#define GOLDQ_DBG_GRADS 3     /* float[P]    dL/dtheta before Adam  */
#define GOLDQ_DBG_INDICES 4   /* int32[B]    sampled logical indices */
#define GOLDQ_DBG_BATCH_S 5   /* float[B*D]  gathered states        */
#define GOLDQ_DBG_BATCH_A 6   /* int32[B]    gathered actions       */
#define GOLDQ_DBG_TRACE 8     /* uint64[4 + 2^17 * 12]: block log of the wide step's kernels (GOLDQ_TRACE=1 at create; tools/trace_step.py) */
#define GOLDQ_DBG_FF_STAMPS 7 /* int64[B/128*2][96]: clock64 stamps inside the fused forward (GOLDQ_FF_STAMPS=1 at create) */

typedef struct goldq goldq_t;

/* Flat POD mirror of deepq.AgentConfig / PolicyConfig / AdamSolver
 * (agent.go:44-85, policy.go:14-38, V/gorgonia.org/gorgonia/solvers.go:399-438). */
typedef struct goldq_config {
    int32_t abi_version;      /* must be GOLDQ_ABI_VERSION */
    int32_t obs_dim;          /* D: env.ObservationSpaceShape()[0]            (policy.go:75) */
    int32_t hidden1;          /* layer.FC{Output: 24}                         (policy.go:55) */
    int32_t hidden2;          /* layer.FC{Input: 24, Output: 24}              (policy.go:56) */
    int32_t n_actions;        /* A: PotentialsShape(env.ActionSpace)[0]       (policy.go:78) */
    int32_t batch_size;       /* PolicyConfig.BatchSize (rows THIS handle trains per step) */
    int32_t global_batch;     /* rows of the whole data-parallel job; 0 = batch_size */
    int64_t replay_capacity;  /* Hyperparameters.BufferSize                   (agent.go:57,64) */
    float gamma;              /* Hyperparameters.Gamma, a float32             (agent.go:48) */
    float huber_delta;        /* PseudoHuberLoss.Delta (loss.go:98); <= 0 means 1.0 (fills what used to be padding) */
    double lr;                /* AdamSolver.eta                               (solvers.go:400) */
    double beta1, beta2, eps; /* AdamSolver.beta1/beta2/eps                   (solvers.go:401-403) */
    int32_t loss;             /* GOLDQ_LOSS_* */
    int32_t precision;        /* GOLDQ_PREC_* */
    int32_t path;             /* GOLDQ_PATH_* */
    int32_t device;           /* CUDA device ordinal */
    int32_t n_replicas;       /* >= 1: independent agents (own weights, solver state, replay) in one grid */
    int32_t solver;           /* GOLDQ_SOLVER_* (0 = Adam, as before this field had a meaning) */
    void *stream;             /* cudaStream_t to run on, or NULL to create one */
    int32_t arch;             /* GOLDQ_ARCH_* (ABI 2) */
    int32_t in_channels, in_height, in_width;   /* GOLDQ_ARCH_ATARI_CONV: state shape (1, 84, 84 for gold's Atari wrapper) */
} goldq_config_t;

/* ---- lifetime ---------------------------------------------------------- */
/* Replaces deepq.NewAgent → MakePolicy ×2 → Sequential.Compile (agent.go:88-123,
 * policy.go:74-110, goro model.go:260-312): allocates both networks, the Adam
 * state and the replay ring on the device.  Parameters start at zero; the host
 * layer initialises them (GlorotN) through goldq_set_params. */
int goldq_default_config(goldq_config_t *cfg);               /* gold's defaults: CartPole shape, B=20, γ=.95, Adam 5e-4 */
goldq_t *goldq_create(const goldq_config_t *cfg);            /* NULL on failure: see goldq_last_error(NULL) */
void goldq_destroy(goldq_t *h);
const char *goldq_last_error(const goldq_t *h);              /* h may be NULL (creation errors) */
int goldq_abi_version(void);
int64_t goldq_param_count(const goldq_t *h);                 /* P per replica */
int goldq_selected_path(const goldq_t *h);                   /* GOLDQ_PATH_* actually in use */

/* ---- parameters and solver state --------------------------------------- */
/* Learnable order [W1,b1,W2,b2,W3,b3], W (in×out) row-major, bias (out)
 * (goro layer/chain.go:38-44, fc.go:98-101,164-169).  Arrays hold
 * n_replicas × P floats.  Replaces reading/writing Sequential.Learnables()
 * (model.go:17-50). */
int goldq_set_params(goldq_t *h, int net, const float *theta);
int goldq_get_params(goldq_t *h, int net, float *theta);
/* AdamSolver's cache (means in .Value, variances in .d) and iter
 * (solvers.go:412-414, 441-445). */
int goldq_set_adam_state(goldq_t *h, const float *m, const float *v, int64_t iter);
int goldq_get_adam_state(goldq_t *h, float *m, float *v, int64_t *iter);

/* ---- replay memory ------------------------------------------------------ */
/* deepq.Memory + Agent.Remember (memory.go:41-51, agent.go:224-229): append n
 * transitions, given oldest first; each becomes the new front (logical index 0 =
 * newest), the oldest are evicted past replay_capacity.  Layout: s, s2 are
 * [n_replicas][n][D] floats; a, r, done are [n_replicas][n]. */
int goldq_replay_push(goldq_t *h, int64_t n, const float *s, const int32_t *a,
                      const float *r, const float *s2, const uint8_t *done);
/* Actions must lie in [0, n_actions): the learn step indexes Q rows and W3 columns with them, so an
 * out-of-range action is rejected here with GOLDQ_EINVAL and nothing is pushed (the reference
 * panics later, in qValues.Set(event.Action, ...), agent.go:155). */
/* Same, from ONE contiguous host buffer (ONE host-to-device copy instead of five):
 *   s [R][n][D] f32 | s2 [R][n][D] f32 | a [R][n] i32 | r [R][n] f32 | done [R][n] u8
 * with R = n_replicas and no padding; goldq_packed_bytes(h, n) is its size.  The Go layer packs
 * Events into such a pinned buffer as they are Remembered. */
int goldq_replay_push_packed(goldq_t *h, int64_t n, const void *packed);
int64_t goldq_packed_bytes(const goldq_t *h, int64_t n);
/* The two halves of goldq_replay_push_packed, for callers that produce transitions ahead of the learner
 * (an actor thread filling the next pinned buffer while Agent.Learn runs, agent.go:224-229 + 126):
 *   goldq_replay_stage_packed  validates the batch and STARTS its host-to-device copy on the copy stream;
 *                              the replay memory does not change.  At most two batches may be staged.
 *   goldq_replay_commit        appends the OLDEST staged batch to the replay memory, ordered after every learn
 *                              step enqueued so far (exactly where goldq_replay_push_packed would have put it).
 * Staging batch t+2 before waiting for learn step t keeps a copy queued behind the one in flight, so the
 * link never idles while the host is between calls.  A later goldq_replay_push* commits anything still
 * staged first (order is kept).  GOLDQ_EINVAL: nothing staged / two batches staged already. */
int goldq_replay_stage_packed(goldq_t *h, int64_t n, const void *packed);
int goldq_replay_commit(goldq_t *h);
int64_t goldq_replay_len(const goldq_t *h);                  /* Memory.Len() */
/* Pageable host buffers are consumed before goldq_replay_push returns.  Buffers
 * from goldq_host_alloc (pinned) are copied asynchronously: do not overwrite them
 * until goldq_replay_push_wait (or goldq_sync) returns. */
int goldq_replay_push_wait(goldq_t *h);
void *goldq_host_alloc(int64_t bytes);                       /* pinned host memory (cudaHostAlloc) */
void goldq_host_free(void *p);

/* ---- the learn step ----------------------------------------------------- */
/* Agent.Learn (agent.go:126-178) minus the ε decay and the steps%N test, which
 * stay in the host layer: sample → two forwards per event → label → FitBatch
 * (goro model.go:552-573) → AdamSolver.Step.  Asynchronous on the handle's stream.
 *   idx != NULL: the sampled logical indices, [n_replicas][batch_size], host
 *                memory (Memory.Sample's output, memory.go:54-69; injected so that
 *                runs are reproducible — the reference reseeds from the clock).
 *   idx == NULL: indices are drawn on the device: the first batch_size entries of
 *                a keyed pseudo-random permutation of [0, Len) — distinct and
 *                uniform like rand.Perm(Len)[:B] — from `seed`.
 * Returns GOLDQ_ENOTREADY if Len < batch_size (agent.go:127). */
int goldq_learn(goldq_t *h, const int32_t *idx, uint64_t seed);
/* Same, with the index list already in device memory (no host traffic). */
int goldq_learn_device_indices(goldq_t *h, const int32_t *idx_dev);
/* n consecutive learn steps with device-drawn indices (step i uses seed + i), replayed from CUDA
 * graphs of 16, 8, 4, 2 and 1 steps (largest first): the same work as n calls of
 * goldq_learn(h, NULL, seed + i), without n times the launch overhead.  No target sync
 * happens inside; the host layer keeps calling it between its goldq_sync_target calls
 * (agent.go:173,181-190).  The graphs read the replay ring's position from per-step device
 * slots, so goldq_replay_push between calls does not invalidate them. */
int goldq_learn_many(goldq_t *h, int32_t n, uint64_t seed);
/* Sequential.ResizeBatch (goro model.go:485-497): change the batch size of an existing agent.  Scratch sized by
 * the batch is re-allocated; parameters, solver state and replay are kept.  Not inside a data-parallel group. */
int goldq_resize_batch(goldq_t *h, int32_t n);
int32_t goldq_batch_size(const goldq_t *h);
/* Instantiate every step graph now (16/8/4/2/1-step and the host-index step of goldq_learn)
 * instead of on first use; executes nothing.  Needs Len >= batch_size. */
int goldq_learn_prepare(goldq_t *h);
/* Diagnostic: graphs instantiated so far (a Remember/Learn loop must not grow it). */
int64_t goldq_graph_builds(const goldq_t *h);
/* Host mirror of the device sampler: the indices goldq_learn(h, NULL, seed) draws
 * for `replica` when the replay holds `len` events (lets the host layer fill
 * Event.i as Memory.Sample does, memory.go:63).  Pure CPU integer code. */
int goldq_sampler_indices(int64_t len, int32_t batch, uint64_t seed, int32_t replica, int32_t *out);
/* Agent.updateTarget's copy (agent.go:181-190) = Sequential.CloneLearnablesTo
 * (goro model.go:606-636): target ← online.  Solver state untouched. */
int goldq_sync_target(goldq_t *h);
/* Loss of the last learn step (the value gold's tracker reads from the
 * `<name>_train_batch_loss` node, goro model.go:404-408).  Synchronises.
 * out: float[n_replicas]. */
int goldq_last_loss(goldq_t *h, float *out);

/* ---- goro model.Model surface ------------------------------------------ */
/* Sequential.Predict / PredictBatch (model.go:500-512, 514-526): q = net(x) for n
 * rows of replica `replica`; x is [n][D], q is [n][A] (host).  Synchronous: the
 * returned q is a fresh copy owned by the caller, as the reference returns a
 * clone (V/gorgonia.org/gorgonia/vm_tape.go:691-697). */
int goldq_predict(goldq_t *h, int net, int replica, int64_t n, const float *x, float *q);
/* Agent.action (agent.go:207-221): first-argmax of the online prediction. */
int goldq_greedy_action(goldq_t *h, int replica, int64_t n, const float *x, int32_t *action);
/* Agent.Action for n states at once (vectorised environments): row i explores when
 * its own uniform draw is below epsilon (agent.go:196: num.RandF32(0,1) < a.epsilon) and then takes a
 * uniform random action (the reference asks the environment, env.SampleAction, agent.go:197), else
 * the first-argmax of the online network (agent.go:207-221).  Draws are counter based:
 * goldq_explore_plan returns, on the host, which rows explore and what they take for a seed. */
int goldq_epsilon_greedy_action(goldq_t *h, int replica, int64_t n, const float *x, float epsilon, uint64_t seed,
                                int32_t *action);
int goldq_explore_plan(int64_t n, int32_t n_actions, float epsilon, uint64_t seed, uint8_t *explore,
                       int32_t *random_action);
/* Sequential.FitBatch / Fit (model.go:528-573) on explicit labels: x [n][D],
 * y [n][A]; one MSE + Adam step on the ONLINE net of replica 0.  n must equal
 * batch_size (FitBatch) or 1 (Fit) — goro rejects other shapes (io.go:138-153). */
int goldq_fit_batch(goldq_t *h, int64_t n, const float *x, const float *y);

/* ---- data-parallel group (new capability) ------------------- */
/* One handle per GPU/process.  Each rank trains batch_size rows of a
 * global_batch-row step; gradients are summed with one all-reduce per step so
 * every rank applies the identical Adam update (loss = mean over the GLOBAL batch,
 * loss.go:28-42).  The 128-byte id comes from goldq_dp_unique_id on rank 0 and is
 * distributed by the caller (torch.distributed / any side channel). */
int goldq_dp_unique_id(uint8_t id[128]);
int goldq_dp_init(goldq_t *h, const uint8_t id[128], int rank, int world);
/* How the group sums gradients: 0 = no group, 1 = ncclAllReduce + update kernel,
 * 2 = exchanged over NVLink peer memory inside the update kernel (wide path; windows shared
 * through CUDA IPC at goldq_dp_init; GOLDQ_NO_P2P=1 in the environment selects 1). */
int goldq_dp_mode(const goldq_t *h);

/* ---- plumbing ----------------------------------------------------------- */
int goldq_sync(goldq_t *h);                                  /* wait for the handle's stream */
void *goldq_stream(goldq_t *h);                              /* the cudaStream_t in use */
int64_t goldq_launch_count(const goldq_t *h);                /* kernels launched so far by this handle */
int goldq_debug_enable(goldq_t *h, int on);                  /* keep Q/TD/grad snapshots of each step */
int goldq_debug_fetch(goldq_t *h, int what, void *out, int64_t out_bytes);
/* Diagnostic: average launch duration (us) of the wide step's dominant tensor-core kernel, the fused forward,
 * over `reps` back-to-back launches between two CUDA events on the handle's stream. */
int goldq_time_forward(goldq_t *h, int32_t reps, float *us_per_launch);
/* Diagnostic: device time per node of a 16-node graph of empty kernels replayed `reps` times -- the
 * launch-latency floor that the latency-bound CartPole configurations are reported against. */
int goldq_launch_floor(goldq_t *h, int32_t reps, float *us_per_launch);
/* CUDA-event timer on the handle's stream (bench.py): start; ...; stop → ms. */
int goldq_timer_start(goldq_t *h);
int goldq_timer_stop(goldq_t *h, float *ms);
/* Diagnostic: ONE eager learn step (device-drawn indices) with a CUDA event after every
 * kernel launch, serialised on the handle's stream; ms[i] = duration of launch i,
 * names = '\n'-joined kernel labels.  It is a real step (parameters are updated). */
int goldq_step_breakdown(goldq_t *h, uint64_t seed, int32_t max_kernels, float *ms, int32_t *n_kernels,
                         char *names, int32_t names_bytes);
/* Diagnostic: run one stand-alone tcgen05 GEMM on device 0 and return fp32 C[M][N].
 * A, B are bf16 bit patterns.  mode 0: C = A[M][K] . B[N][K]^T (both K-major);
 * mode 1: C = A[K][M]^T . B[K][N] (both MN-major); mode 2: mode 0 with the 32-wide
 * N tile.  nsplit > 1 exercises the split-K slices.  Used by tests/test_gpu_umma.py. */
int goldq_selftest_umma(int mode, int M, int N, int K, int nsplit, const uint16_t *A, const uint16_t *B, float *C);

#ifdef __cplusplus
}
#endif
#endif /* GOLDQ_H */
