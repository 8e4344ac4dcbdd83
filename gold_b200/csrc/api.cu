// api.cu — C ABI (include/goldq.h) and host runtime of libgoldq.
//
// Owns the device state of one deepq agent (or n_replicas of them): online and
// target parameters, Adam moments, the SoA replay ring, per-step scratch; one
// CUDA stream per handle.  No CPU compute path exists here by design: every
// entry point that computes launches CUDA kernels or fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "sampler.cuh"
#include "wide.cuh"

using namespace gq;

// ---------------------------------------------------------------------------
// NCCL, resolved at run time so that libgoldq.so loads without it (and shares
// the copy torch already loaded when it runs inside a torch process).
// ---------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err)
    {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("dlopen libnccl failed: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
            err = "libnccl is missing required symbols";
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;
std::string g_create_error;
}  // namespace

struct goldq {
    goldq_config_t cfg;
    Dims dims;
    Offsets off;
    int P = 0, R = 1, B = 0, Bglobal = 0;
    int dev = 0, sm_count = 148;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    int path = GOLDQ_PATH_GENERIC;

    float *theta = nullptr, *theta_t = nullptr, *m = nullptr, *v = nullptr;  // [R][P]
    int64_t iter = 0;
    Replay rp{};
    int32_t *d_idx = nullptr;  // [R][B]

    // push staging (device) + scatter
    // double-buffered device staging, filled on a dedicated copy stream so that the H2D copy of
    // push i+1 overlaps the learn step enqueued after push i
    struct Staging {
        uint8_t *base = nullptr;        // one allocation, packed as goldq_replay_push_packed documents
        cudaEvent_t copied = nullptr, scattered = nullptr;
        int64_t n = 0;                  // items per replica of the staged, not yet committed batch
    } stg[3];
    int stg_cur = 0;                    // next set to fill
    int stg_pending = 0;                // staged batches not yet scattered into the ring (oldest = stg_cur - stg_pending)
    int64_t stg_cap = 0;
    cudaStream_t copy_st = nullptr;
    cudaEvent_t ev_push = nullptr;      // = copied event of the latest push
    // goldq_last_loss reads the scalar on its own stream, ordered after the latest learn step only: the
    // readback does not wait for a replay push enqueued behind that step
    cudaStream_t rb_st = nullptr;
    cudaEvent_t ev_learned = nullptr;

    // reduced gradient + loss: [R][P+1]; loss per replica
    float *flat = nullptr, *loss = nullptr;
    // narrow path
    float *partial = nullptr;
    unsigned int *ticket = nullptr;
    int n_ctas = 1, n_warps = 1;
    // generic path scratch (rows = gen_rows)
    int gen_rows = 0;
    float *bs = nullptr, *bs2 = nullptr, *br = nullptr;
    int32_t *ba = nullptr;
    uint8_t *bdone = nullptr;
    float *h1 = nullptr, *h2 = nullptr, *q = nullptr, *t1 = nullptr, *t2 = nullptr, *qt = nullptr;
    float *td = nullptr, *dq = nullptr, *dh2 = nullptr, *dh1 = nullptr, *ylab = nullptr;
    uint8_t *mk1 = nullptr, *mk2 = nullptr;
    float *gslice = nullptr, *loss_part = nullptr;
    int nsplit = 1;
    // predict scratch
    int64_t pred_rows = 0;
    float *px = nullptr, *ph1 = nullptr, *ph2 = nullptr, *pq = nullptr;
    int32_t *pact = nullptr;
    // wide (tcgen05) path
    WideState *wide = nullptr;
    // Atari conv policy (GOLDQ_PATH_CONV): activations [rows][O][OH][OW] of the online (y*, masks mk*) and the target
    // (u*) network, FC-512 activations, and the gradients flowing back through the stack
    AtariNet atari{};
    int conv_rows = 0;
    float *cy1 = nullptr, *cy2 = nullptr, *cy3 = nullptr, *ch4 = nullptr, *cu1 = nullptr, *cu2 = nullptr, *cu3 = nullptr, *cu4 = nullptr;
    uint8_t *cm1 = nullptr, *cm2 = nullptr, *cm3 = nullptr, *cm4 = nullptr;
    float *cdy1 = nullptr, *cdy2 = nullptr, *cdy3 = nullptr, *cdh4 = nullptr;

    bool debug = false;
    float *dbg_q = nullptr, *dbg_qt = nullptr, *dbg_td = nullptr;
    int32_t *dbg_idx = nullptr;

    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    std::string err;

    // CUDA-graph replay of learn steps (goldq_learn_many)
    static constexpr int GRAPH_MAX = 32;
    static constexpr int GRAPH_CHUNK = 32;       // largest number of steps per graph launch; sizes: multiples of 4, then 2, 1
    static constexpr int STEP_RING = 8;          // pinned staging buffers for the per-step scalars
    int pf_set = 0, pf_have = 0;                 // batch double buffering while capturing (build_graph)
    const StepDev *pf_next = nullptr;
    StepDev *d_steps = nullptr;                  // [GRAPH_MAX] per-step scalars read by captured kernels
    StepDev *h_steps[STEP_RING] = {};            // pinned (mapped) staging ring
    int32_t *h_idx[STEP_RING] = {};              // pinned (mapped) copies of goldq_learn's host index lists, [R][B] each
    cudaEvent_t ev_steps[STEP_RING] = {};
    int steps_buf = 0;
    // graph_exec[c]: c consecutive steps with device-drawn indices; graph_exec[0]: ONE step that reads the
    // index list in d_idx (goldq_learn with host indices)
    cudaGraphExec_t graph_exec[GRAPH_MAX + 1] = {};
    int64_t graph_launches[GRAPH_MAX + 1] = {};
    int64_t graph_builds = 0;                    // cudaGraphInstantiate calls so far (goldq_graph_builds)
    bool graph_debug = false;

    cudaGraphExec_t dp_graph = nullptr;          // phase 1 of one data-parallel wide step
    int64_t dp_graph_launches = 0;

    // data-parallel tail over NVLink peer memory (wide path): see wide_p2p_apply
    struct P2P {
        bool on = false;
        int chunk_f = 0;                         // floats owned per rank
        float *base = nullptr;                   // local exported allocation: {value, epoch} windows
        void *peer_base[P2P_MAX_WORLD] = {};     // every rank's allocation as mapped here (own = base)
        uint32_t epoch = 0;                      // data-parallel steps taken (host mirror)
        int *err_host = nullptr, *err_dev = nullptr;
    } p2p;

    // goldq_step_breakdown
    cudaEvent_t *prof_evs = nullptr;
    int prof_n = 0;
};

namespace {

void p2p_teardown(goldq_t *h);

int fail(goldq_t *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    else g_create_error = msg;
    return code;
}

#define CK(h, expr)                                                                      \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(h, GOLDQ_ECUDA,                                                  \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));             \
    } while (0)

template <typename T>
cudaError_t dalloc(T **p, size_t n)
{
    if (n == 0) n = 1;
    return cudaMalloc((void **)p, n * sizeof(T));
}

int check_launch(goldq_t *h, const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, GOLDQ_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return GOLDQ_OK;
}

AdamConsts adam_consts(const goldq_config_t &c, int64_t iter)
{
    // vendor/gorgonia.org/gorgonia/solvers.go:446-447 (float64), :490-503 (casts)
    double correction1 = 1.0 - pow(c.beta1, (double)iter);
    double correction2 = 1.0 - pow(c.beta2, (double)iter);
    AdamConsts a;
    a.kind = c.solver;
    a.beta1 = (float)c.beta1;
    a.beta2 = (float)c.beta2;
    a.omb1 = 1.0f - (float)c.beta1;
    a.omb2 = 1.0f - (float)c.beta2;
    a.eps = (float)c.eps;
    a.neg_eta = (float)(-c.lr);
    a.c1 = 1.0f / (float)correction1;
    a.c2 = 1.0f / (float)correction2;
    if (c.solver == GOLDQ_SOLVER_RMSPROP) {
        // float32(s.decay), float32(1.0 - s.decay), float32(s.eps), float32(-s.eta)  (solvers.go:286-293);
        // the cache lives in `v` (goldq_get/set_adam_state: m unused, v = cache); no bias corrections
        a.omb2 = (float)(1.0 - c.beta2);
        a.c1 = a.c2 = 1.0f;
    }
    return a;
}

StepScalars step_scalars(const goldq_t *h, int rows_global)
{
    StepScalars s;
    s.gamma = h->cfg.gamma;
    s.count = (float)((int64_t)rows_global * h->dims.A);
    s.inv_count = 1.0f / s.count;
    s.loss_kind = h->cfg.loss == GOLDQ_LOSS_PSEUDO_HUBER ? 1 : 0;
    s.delta = h->cfg.huber_delta > 0.f ? h->cfg.huber_delta : 1.0f;
    s.delta2 = s.delta * s.delta;      // Square of the float32 scalar (loss.go:139)
    return s;
}

// generic path: one solver step on the flat gradient (h->iter already counts this step)
void generic_solver_step(goldq_t *h, const StepDev *step)
{
    // (Adam or RMSProp: solver_update in common.cuh switches on AdamConsts::kind)
    launch_adam(h->theta, h->flat, h->m, h->v, h->P, adam_consts(h->cfg, h->iter), step, h->st);
}

int ensure_generic(goldq_t *h, int rows)
{
    if (h->gen_rows >= rows) return GOLDQ_OK;
    const Dims &d = h->dims;
    float **fp[] = {&h->bs, &h->bs2, &h->br, &h->h1, &h->h2, &h->q, &h->t1, &h->t2, &h->qt,
                    &h->td, &h->dq, &h->dh2, &h->dh1, &h->ylab, &h->gslice, &h->loss_part};
    for (auto p : fp) { if (*p) cudaFree(*p); *p = nullptr; }
    if (h->ba) cudaFree(h->ba);
    if (h->bdone) cudaFree(h->bdone);
    if (h->mk1) cudaFree(h->mk1);
    if (h->mk2) cudaFree(h->mk2);
    size_t n = rows;
    CK(h, dalloc(&h->bs, n * d.D));
    CK(h, dalloc(&h->bs2, n * d.D));
    CK(h, dalloc(&h->br, n));
    CK(h, dalloc(&h->ba, n));
    CK(h, dalloc(&h->bdone, n));
    CK(h, dalloc(&h->h1, n * d.H1));
    CK(h, dalloc(&h->h2, n * d.H2));
    CK(h, dalloc(&h->q, n * d.A));
    CK(h, dalloc(&h->t1, n * d.H1));
    CK(h, dalloc(&h->t2, n * d.H2));
    CK(h, dalloc(&h->qt, n * d.A));
    CK(h, dalloc(&h->td, n));
    CK(h, dalloc(&h->dq, n * d.A));
    CK(h, dalloc(&h->dh2, n * d.H2));
    CK(h, dalloc(&h->dh1, n * d.H1));
    CK(h, dalloc(&h->ylab, n * d.A));
    CK(h, dalloc(&h->mk1, n * d.H1));
    CK(h, dalloc(&h->mk2, n * d.H2));
    h->nsplit = rows / 512;
    if (h->nsplit < 1) h->nsplit = 1;
    if (h->nsplit > 32) h->nsplit = 32;
    CK(h, dalloc(&h->gslice, (size_t)h->nsplit * h->P));
    CK(h, dalloc(&h->loss_part, (size_t)(rows * d.A + 255) / 256 + 1));
    h->gen_rows = rows;
    return GOLDQ_OK;
}

// forward of one net over n rows with the generic kernels
void generic_forward(goldq_t *h, const float *theta, const float *x, int n, float *h1, uint8_t *mk1,
                     float *h2, uint8_t *mk2, float *q)
{
    const Dims &d = h->dims;
    const Offsets &o = h->off;
    launch_fc_forward(x, theta + o.w1, theta + o.b1, h1, mk1, n, d.D, d.H1, true, h->st);
    launch_fc_forward(h1, theta + o.w2, theta + o.b2, h2, mk2, n, d.H1, d.H2, true, h->st);
    launch_fc_forward(h2, theta + o.w3, theta + o.b3, q, nullptr, n, d.H2, d.A, false, h->st);
    h->launches += 3;
}

// backward from dq through the online net (activations in h->h1/h2/mk1/mk2, input x)
void generic_backward(goldq_t *h, const float *x, int n)
{
    const Dims &d = h->dims;
    const Offsets &o = h->off;
    const float *th = h->theta;
    launch_grad_w(h->h2, h->dq, n, d.H2, d.A, h->gslice, h->P, h->nsplit, o.w3, o.b3, h->st);
    launch_grad_x(h->dq, th + o.w3, h->mk2, h->dh2, n, d.H2, d.A, h->st);
    launch_grad_w(h->h1, h->dh2, n, d.H1, d.H2, h->gslice, h->P, h->nsplit, o.w2, o.b2, h->st);
    launch_grad_x(h->dh2, th + o.w2, h->mk1, h->dh1, n, d.H1, d.H2, h->st);
    launch_grad_w(x, h->dh1, n, d.D, d.H1, h->gslice, h->P, h->nsplit, o.w1, o.b1, h->st);
    h->launches += 5;
}

int allreduce_flat(goldq_t *h)
{
    if (h->world <= 1) return GOLDQ_OK;
    ncclResult_t r = g_nccl.AllReduce(h->flat, h->flat, (size_t)h->P + 1, ncclFloat, ncclSum,
                                      h->comm, h->st);
    if (r != ncclSuccess)
        return fail(h, GOLDQ_ENCCL, std::string("ncclAllReduce: ") +
                                        (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    h->launches += 1;
    return GOLDQ_OK;
}

int learn_generic(goldq_t *h, const int32_t *idx_dev, const StepDev *step)
{
    int rc = ensure_generic(h, h->B);
    if (rc) return rc;
    const Dims &d = h->dims;
    launch_gather(h->rp, d.D, idx_dev, h->B, h->bs, h->bs2, h->ba, h->br, h->bdone, step, h->st);
    h->launches += 1;
    generic_forward(h, h->theta, h->bs, h->B, h->h1, h->mk1, h->h2, h->mk2, h->q);
    generic_forward(h, h->theta_t, h->bs2, h->B, h->t1, nullptr, h->t2, nullptr, h->qt);
    int nlp = 0;
    StepScalars sc = step_scalars(h, h->Bglobal);
    launch_td(h->q, h->qt, h->ba, h->br, h->bdone, h->B, d.A, sc, h->td, h->dq, h->loss_part, &nlp,
              h->st);
    h->launches += 1;
    generic_backward(h, h->bs, h->B);
    launch_reduce_slices(h->gslice, h->P, h->nsplit, h->P, h->loss_part, nlp, sc.count, h->flat,
                         h->st);
    h->launches += 1;
    rc = allreduce_flat(h);
    if (rc) return rc;
    h->iter += 1;
    generic_solver_step(h, step);
    h->launches += 1;
    CK(h, cudaMemcpyAsync(h->loss, h->flat + h->P, sizeof(float), cudaMemcpyDefault, h->st));
    return check_launch(h, "learn_generic");
}

int learn_narrow(goldq_t *h, const int32_t *idx_dev, uint64_t seed, const StepDev *step)
{
    NarrowArgs a{};
    a.dims = h->dims;
    a.rp = h->rp;
    a.idx = idx_dev;
    a.seed = seed;
    a.step = step;
    a.B = h->B;
    a.replicas = h->R;
    a.theta = h->theta;
    a.theta_t = h->theta_t;
    a.m = h->m;
    a.v = h->v;
    a.partial = h->partial;
    a.ticket = h->ticket;
    a.flat = h->flat;
    a.loss = h->loss;
    a.fuse_adam = h->world <= 1 ? 1 : 0;
    a.sc = step_scalars(h, h->Bglobal);
    h->iter += 1;
    a.adam = adam_consts(h->cfg, h->iter);
    if (h->debug) {
        a.dbg_q = h->dbg_q;
        a.dbg_qt = h->dbg_qt;
        a.dbg_td = h->dbg_td;
        a.dbg_idx = h->dbg_idx;
    }
    a.ctas_per_replica = h->n_ctas;
    a.warps_per_cta = h->n_warps;
    launch_narrow_learn(a, h->st);
    h->launches += 1;
    int rc = check_launch(h, "narrow_learn_kernel");
    if (rc) return rc;
    if (h->world > 1) {
        rc = allreduce_flat(h);
        if (rc) return rc;
        launch_adam(h->theta, h->flat, h->m, h->v, h->P, a.adam, step, h->st);
        h->launches += 1;
        CK(h, cudaMemcpyAsync(h->loss, h->flat + h->P, sizeof(float), cudaMemcpyDefault,
                              h->st));
    }
    return check_launch(h, "learn_narrow");
}

// phases: 1 = forward/backward up to the reduced gradient, 2 = all-reduce + update (data-parallel
// tail), 3 = both.  Single-GPU steps fuse the update into phase 1.
int learn_wide(goldq_t *h, const int32_t *idx_dev, uint64_t seed, const StepDev *step, int phases = 3)
{
    WideStep ws{};
    ws.rp = h->rp;
    ws.idx = idx_dev;
    ws.seed = seed;
    ws.step = step;
    ws.idx_out = h->d_idx;
    ws.B = h->B;
    ws.theta = h->theta;
    ws.theta_t = h->theta_t;
    ws.flat = h->flat;
    ws.loss_out = h->loss;
    ws.sc = step_scalars(h, h->Bglobal);
    ws.fuse_update = (h->world <= 1 || h->p2p.on) ? 1 : 0;
    ws.theta_rw = h->theta;
    ws.m = h->m;
    ws.v = h->v;
    if (phases & 1) h->iter += 1;
    ws.adam = adam_consts(h->cfg, h->iter);
    ws.prof_evs = h->prof_evs;
    ws.prof_n = &h->prof_n;
    ws.set = h->pf_set;
    ws.have = h->pf_have;
    ws.next_step = h->pf_next;
    std::string err;
    int n = 0;
    PeerSet ps{};
    if (h->p2p.on) {
        // every float of the gradient travels as a {value, epoch} word, exchanged inside the update kernel
        if (phases & 1) h->p2p.epoch += 1;
        ps.world = h->world;
        ps.rank = h->rank;
        ps.chunk_f = h->p2p.chunk_f;
        for (int r = 0; r < h->world; r++) {
            ps.rs[r] = (unsigned long long *)h->p2p.peer_base[r];
            ps.ag[r] = ps.rs[r] + (size_t)2 * h->world * h->p2p.chunk_f;
        }
        ps.err = h->p2p.err_dev;
        ws.push = &ps;
        ws.epoch = h->p2p.epoch;
    }
    if (phases & 1) {
        n = wide_step(h->wide, ws, h->st, err);
        if (n < 0) return fail(h, GOLDQ_ECUDA, err);
        h->launches += n;
    }
    if (h->p2p.on) {
        // (the exchange happened inside the update kernel of wide_step)
    } else if (h->world > 1 && (phases & 2)) {
        int rc = allreduce_flat(h);
        if (rc) return rc;
        if (h->prof_evs) cudaEventRecord(h->prof_evs[h->prof_n++], h->st);
        n = wide_apply_update(h->wide, h->flat, h->theta, h->m, h->v, ws.adam, step, h->st, err);   // Adam + bf16 shadow
        if (n < 0) return fail(h, GOLDQ_ECUDA, err);
        h->launches += n;
        if (h->prof_evs) cudaEventRecord(h->prof_evs[h->prof_n++], h->st);
        CK(h, cudaMemcpyAsync(h->loss, h->flat + h->P, sizeof(float), cudaMemcpyDefault, h->st));
    }
    return check_launch(h, "learn_wide");
}

// ---- Atari conv policy (fp32, kernels_conv.cu) ------------------------------------------------------------------
int ensure_conv(goldq_t *h, int rows)
{
    if (h->conv_rows >= rows) return GOLDQ_OK;
    const AtariNet &t = h->atari;
    void *old[] = {h->cy1, h->cy2, h->cy3, h->ch4, h->cu1, h->cu2, h->cu3, h->cu4, h->cm1, h->cm2, h->cm3, h->cm4, h->cdy1, h->cdy2,
                   h->cdy3, h->cdh4, h->bs, h->bs2, h->br, h->ba, h->bdone, h->q, h->qt, h->td, h->dq, h->gslice, h->loss_part};
    for (void *p : old)
        if (p) cudaFree(p);
    const size_t n = rows, s1 = (size_t)32 * t.c1.OH * t.c1.OW, s2 = (size_t)64 * t.c2.OH * t.c2.OW, s3 = t.F;
    CK(h, dalloc(&h->cy1, n * s1)); CK(h, dalloc(&h->cu1, n * s1)); CK(h, dalloc(&h->cm1, n * s1)); CK(h, dalloc(&h->cdy1, n * s1));
    CK(h, dalloc(&h->cy2, n * s2)); CK(h, dalloc(&h->cu2, n * s2)); CK(h, dalloc(&h->cm2, n * s2)); CK(h, dalloc(&h->cdy2, n * s2));
    CK(h, dalloc(&h->cy3, n * s3)); CK(h, dalloc(&h->cu3, n * s3)); CK(h, dalloc(&h->cm3, n * s3)); CK(h, dalloc(&h->cdy3, n * s3));
    CK(h, dalloc(&h->ch4, n * t.HID)); CK(h, dalloc(&h->cu4, n * t.HID)); CK(h, dalloc(&h->cm4, n * t.HID)); CK(h, dalloc(&h->cdh4, n * t.HID));
    CK(h, dalloc(&h->bs, n * h->dims.D)); CK(h, dalloc(&h->bs2, n * h->dims.D));
    CK(h, dalloc(&h->br, n)); CK(h, dalloc(&h->ba, n)); CK(h, dalloc(&h->bdone, n));
    CK(h, dalloc(&h->q, n * t.A)); CK(h, dalloc(&h->qt, n * t.A)); CK(h, dalloc(&h->td, n)); CK(h, dalloc(&h->dq, n * t.A));
    h->nsplit = rows / 512;
    if (h->nsplit < 1) h->nsplit = 1;
    if (h->nsplit > 4) h->nsplit = 4;          // P is ~3.3 M floats for 84 x 84 inputs: keep the slice buffer small
    CK(h, dalloc(&h->gslice, (size_t)h->nsplit * h->P));
    CK(h, dalloc(&h->loss_part, (size_t)(rows * t.A + 255) / 256 + 1));
    h->conv_rows = rows;
    return GOLDQ_OK;
}

// forward of the conv policy over n states; masks are kept for the online pass of a learn step only
void conv_forward(goldq_t *h, const float *theta, const float *x, int n, float *y1, uint8_t *m1, float *y2, uint8_t *m2,
                  float *y3, uint8_t *m3, float *h4, uint8_t *m4, float *q)
{
    const AtariNet &t = h->atari;
    launch_conv_forward(t.c1, n, x, theta + t.f1, y1, m1, h->st);
    launch_conv_forward(t.c2, n, y1, theta + t.f2, y2, m2, h->st);
    launch_conv_forward(t.c3, n, y2, theta + t.f3, y3, m3, h->st);
    // Flatten is a reshape of the NCHW tensor (goro layer/flatten.go:44-61): y3 as stored
    launch_fc_forward(y3, theta + t.w4, theta + t.b4, h4, m4, n, t.F, t.HID, true, h->st);
    launch_fc_forward(h4, theta + t.w5, theta + t.b5, q, nullptr, n, t.HID, t.A, false, h->st);
    h->launches += 5;
}

int learn_conv(goldq_t *h, const int32_t *idx_dev, const StepDev *step)
{
    int rc = ensure_conv(h, h->B);
    if (rc) return rc;
    const AtariNet &t = h->atari;
    const int B = h->B;
    launch_gather(h->rp, h->dims.D, idx_dev, B, h->bs, h->bs2, h->ba, h->br, h->bdone, step, h->st);
    h->launches += 1;
    conv_forward(h, h->theta, h->bs, B, h->cy1, h->cm1, h->cy2, h->cm2, h->cy3, h->cm3, h->ch4, h->cm4, h->q);
    conv_forward(h, h->theta_t, h->bs2, B, h->cu1, nullptr, h->cu2, nullptr, h->cu3, nullptr, h->cu4, nullptr, h->qt);
    int nlp = 0;
    StepScalars sc = step_scalars(h, h->Bglobal);
    launch_td(h->q, h->qt, h->ba, h->br, h->bdone, B, t.A, sc, h->td, h->dq, h->loss_part, &nlp, h->st);
    const float *th = h->theta;
    const int ns = h->nsplit;
    // FC A, FC 512 (the generic dense kernels), then the three convolutions; every conv kernel applies its own
    // layer's Rectify mask to the gradient it receives
    launch_grad_w(h->ch4, h->dq, B, t.HID, t.A, h->gslice, h->P, ns, (int)t.w5, (int)t.b5, h->st);
    launch_grad_x(h->dq, th + t.w5, h->cm4, h->cdh4, B, t.HID, t.A, h->st);
    launch_grad_w(h->cy3, h->cdh4, B, t.F, t.HID, h->gslice, h->P, ns, (int)t.w4, (int)t.b4, h->st);
    launch_grad_x(h->cdh4, th + t.w4, h->cm3, h->cdy3, B, t.F, t.HID, h->st);
    launch_conv_backward_filter(t.c3, B, ns, h->cy2, h->cdy3, h->cm3, h->gslice, h->P, t.f3, h->st);
    launch_conv_backward_data(t.c3, B, h->cdy3, h->cm3, th + t.f3, h->cdy2, h->st);
    launch_conv_backward_filter(t.c2, B, ns, h->cy1, h->cdy2, h->cm2, h->gslice, h->P, t.f2, h->st);
    launch_conv_backward_data(t.c2, B, h->cdy2, h->cm2, th + t.f2, h->cdy1, h->st);
    launch_conv_backward_filter(t.c1, B, ns, h->bs, h->cdy1, h->cm1, h->gslice, h->P, t.f1, h->st);
    launch_reduce_slices(h->gslice, h->P, ns, h->P, h->loss_part, nlp, sc.count, h->flat, h->st);
    h->launches += 11;
    rc = allreduce_flat(h);
    if (rc) return rc;
    h->iter += 1;
    generic_solver_step(h, step);
    h->launches += 1;
    CK(h, cudaMemcpyAsync(h->loss, h->flat + h->P, sizeof(float), cudaMemcpyDefault, h->st));
    return check_launch(h, "learn_conv");
}

int run_learn(goldq_t *h, const int32_t *idx_dev, uint64_t seed, const StepDev *step = nullptr)
{
    if (h->rp.len < h->B) return fail(h, GOLDQ_ENOTREADY, "replay shorter than batch");
    // a peer that stopped answering was detected by an earlier step: do not queue more 5 s waits
    if (h->p2p.on && *(volatile int *)h->p2p.err_host)
        return fail(h, GOLDQ_ENCCL, "data-parallel peer did not publish its gradient within 5 s");
    if (h->path == GOLDQ_PATH_NARROW) return learn_narrow(h, idx_dev, seed, step);
    if (h->path == GOLDQ_PATH_WIDE) return learn_wide(h, idx_dev, seed, step);
    if (!idx_dev) {
        launch_sample_indices(h->d_idx, h->B, h->rp.len, seed, step, h->R, h->st);
        h->launches += 1;
        idx_dev = h->d_idx;
    }
    if (h->path == GOLDQ_PATH_CONV) return learn_conv(h, idx_dev, step);
    return learn_generic(h, idx_dev, step);
}

void drop_graphs(goldq_t *h)
{
    if (h->dp_graph) { cudaGraphExecDestroy(h->dp_graph); h->dp_graph = nullptr; }
    for (int i = 0; i <= goldq::GRAPH_MAX; i++)
        if (h->graph_exec[i]) { cudaGraphExecDestroy(h->graph_exec[i]); h->graph_exec[i] = nullptr; }
}

// graph sizes: 1, 2, then the multiples of 4 up to GRAPH_CHUNK (slot 0 = the one-step graph with host indices)
inline int graph_next_size(int c) { return c == 0 ? 1 : (c < 4 ? c * 2 : c + 4); }
inline int graph_size_for(int n)
{
    if (n >= goldq::GRAPH_CHUNK) return goldq::GRAPH_CHUNK;
    return n >= 4 ? (n & ~3) : (n >= 2 ? 2 : 1);
}

// Per-step scalars (and goldq_learn's host index list) travel to the device through THIS kernel reading mapped
// pinned host memory, not through cudaMemcpyAsync: a small host-to-device copy on the learn stream shares the
// copy engine's queue with goldq_replay_push's 8.5 MB copy on the copy stream, and because the small copy is
// stream-ordered behind the previous push's scatter the next big copy queued behind IT could not start
// (tools/e2e_probe.py: 215 us per Remember+Learn against the 156 us the copy alone needs).
__global__ void stage_steps_kernel(const uint2 *__restrict__ ssrc, uint2 *__restrict__ sdst, int nsteps16,
                                   const int32_t *__restrict__ isrc, int32_t *__restrict__ idst, int nidx)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nsteps16) sdst[t] = ssrc[t];
    if (t < nidx) idst[t] = isrc[t];
}
static_assert(sizeof(StepDev) % 8 == 0, "StepDev is staged in 8-byte words");

// Capture c consecutive learn steps (device sampler, per-step scalars from d_steps[i]) into
// one executable graph; c == 0: ONE step that reads its index list from d_idx.  Kernel arguments
// are frozen at capture time; everything that changes from step to step (sampler seed, Adam's bias
// corrections, the data-parallel epoch, the replay ring's head and length) is read from the
// step's StepDev slot instead, so a graph survives goldq_replay_push.  The cache is dropped only
// when the debug flag or the data-parallel group changes.
int build_graph(goldq_t *h, int slot)
{
    const int c = slot == 0 ? 1 : slot;
    const int64_t iter0 = h->iter, launches0 = h->launches;
    cudaGraph_t graph = nullptr;
    CK(h, cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
    int rc = GOLDQ_OK;
    static const bool no_prefetch = getenv("GOLDQ_NO_PREFETCH") != nullptr;
    const bool prefetch = h->path == GOLDQ_PATH_WIDE && (h->world <= 1 || h->p2p.on) && !no_prefetch;
    const uint32_t epoch0 = h->p2p.epoch;
    for (int i = 0; i < c && rc == GOLDQ_OK; i++) {
        if (prefetch) {   // batch i+1 is gathered on a side branch while step i computes
            h->pf_set = i & 1;
            h->pf_have = i > 0;
            h->pf_next = i + 1 < c ? h->d_steps + i + 1 : nullptr;
        }
        rc = run_learn(h, slot == 0 ? h->d_idx : nullptr, 0, h->d_steps + i);
    }
    h->pf_set = h->pf_have = 0;
    h->pf_next = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->st, &graph);
    h->graph_launches[slot] = h->launches - launches0;
    h->iter = iter0;            // capturing executes nothing
    h->launches = launches0;
    h->p2p.epoch = epoch0;
    if (rc != GOLDQ_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(h, GOLDQ_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->graph_exec[slot], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(h, GOLDQ_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    // ship the executable graph to the device now: otherwise its first launch pays for the upload
    // (hundreds of microseconds) inside whatever region that launch is timed in
    e = cudaGraphUpload(h->graph_exec[slot], h->st);
    if (e != cudaSuccess) return fail(h, GOLDQ_ECUDA, std::string("cudaGraphUpload: ") + cudaGetErrorString(e));
    h->graph_builds += 1;
    return GOLDQ_OK;
}

// Replay the graph in `slot` (see build_graph) for steps seed, seed+1, ...: stage the steps' scalars,
// copy them to the device slots in stream order, launch.
int launch_graph(goldq_t *h, int slot, uint64_t seed, const int32_t *host_idx = nullptr)
{
    const int c = slot == 0 ? 1 : slot;
    // a peer that stopped answering was detected by an earlier step: do not queue more 5 s waits
    if (h->p2p.on && *(volatile int *)h->p2p.err_host)
        return fail(h, GOLDQ_ENCCL, "data-parallel peer did not publish its gradient within 5 s");
    if (h->graph_debug != h->debug) { drop_graphs(h); h->graph_debug = h->debug; }
    if (h->path == GOLDQ_PATH_GENERIC || h->path == GOLDQ_PATH_CONV) {          // no allocation may happen under capture
        int rc = h->path == GOLDQ_PATH_CONV ? ensure_conv(h, h->B) : ensure_generic(h, h->B);
        if (rc) return rc;
    }
    if (!h->graph_exec[slot]) {
        int rc = build_graph(h, slot);
        if (rc) return rc;
    }
    const int b = h->steps_buf;
    h->steps_buf = (b + 1) % goldq::STEP_RING;
    CK(h, cudaEventSynchronize(h->ev_steps[b]));     // the staging buffer's previous copy has run
    for (int i = 0; i < c; i++) {
        AdamConsts ac = adam_consts(h->cfg, h->iter + 1 + i);
        StepDev &sd = h->h_steps[b][i];
        sd.seed = seed + (uint64_t)i;
        sd.c1 = ac.c1;
        sd.c2 = ac.c2;
        sd.epoch = h->p2p.epoch + 1 + (uint32_t)i;
        sd.pad_ = 0;
        sd.head = h->rp.head;
        sd.len = h->rp.len;
    }
    const int n16 = (int)(sizeof(StepDev) / 8) * c;
    int nidx = 0;
    if (host_idx) {
        nidx = h->R * h->B;
        memcpy(h->h_idx[b], host_idx, (size_t)nidx * 4);
    }
    const int nthr = n16 > nidx ? n16 : nidx;
    stage_steps_kernel<<<(nthr + 255) / 256, 256, 0, h->st>>>((const uint2 *)h->h_steps[b], (uint2 *)h->d_steps, n16,
                                                            h->h_idx[b], h->d_idx, nidx);
    h->launches += 1;
    CK(h, cudaEventRecord(h->ev_steps[b], h->st));
    CK(h, cudaGraphLaunch(h->graph_exec[slot], h->st));
    if (h->p2p.on) h->p2p.epoch += (uint32_t)c;
    h->iter += c;
    h->launches += h->graph_launches[slot];
    return GOLDQ_OK;
}

// steps are replayed from CUDA graphs unless the group sums gradients with ncclAllReduce (a captured
// collective replays slower than the eager call) or GOLDQ_NO_GRAPH is set (diagnostic)
bool use_graphs(const goldq_t *h)
{
    static const bool no_graph = getenv("GOLDQ_NO_GRAPH") != nullptr;
    return !no_graph && (h->world <= 1 || h->p2p.on);
}

// every public entry that ran learn steps ends here: goldq_last_loss orders its readback after this point
int learned(goldq_t *h, int rc)
{
    if (rc == GOLDQ_OK && cudaEventRecord(h->ev_learned, h->st) != cudaSuccess)
        return fail(h, GOLDQ_ECUDA, "cudaEventRecord(ev_learned) failed");
    return rc;
}

// Packed push layout (host and device staging): for R replicas and n items
//   s [R][n][D] f32 | s2 [R][n][D] f32 | a [R][n] i32 | r [R][n] f32 | done [R][n] u8
struct PackedView {
    const float *s, *s2, *r;
    const int32_t *a;
    const uint8_t *done;
};
__host__ __device__ inline size_t packed_bytes(int64_t rows, int D) { return (size_t)rows * ((size_t)D * 8 + 9); }
inline PackedView packed_view(const void *base, int64_t rows, int D)
{
    const uint8_t *p = (const uint8_t *)base;
    PackedView v;
    v.s = (const float *)p;
    v.s2 = (const float *)(p + (size_t)rows * D * 4);
    v.a = (const int32_t *)(p + (size_t)rows * D * 8);
    v.r = (const float *)(p + (size_t)rows * D * 8 + (size_t)rows * 4);
    v.done = p + (size_t)rows * D * 8 + (size_t)rows * 8;
    return v;
}

// item (rep, k), k in [skip, n): physical slot (head + k - skip) mod cap.  VEC: one thread per
// 16-byte granule of a row (D % 4 == 0 and 16-byte aligned staging), else one per float.
template <bool VEC>
__global__ void __launch_bounds__(256)
replay_scatter_kernel(Replay rp, int D, int R, int64_t n, int64_t skip, PackedView src)
{
    const int per_row = VEC ? (D >> 2) : D;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t m = n - skip;
    int64_t per = m * per_row;
    if (t >= per * R) return;
    int rep = (int)(t / per);
    int64_t e = t - (int64_t)rep * per;
    int64_t k = e / per_row;
    int c = (int)(e - k * per_row);
    int64_t slot = (rp.head + k) % rp.cap;
    int64_t from = ((int64_t)rep * n + skip + k);
    int64_t dst = (int64_t)rep * rp.cap + slot;
    if (VEC) {
        reinterpret_cast<float4 *>(rp.s + dst * D)[c] = __ldg(reinterpret_cast<const float4 *>(src.s + from * D) + c);
        reinterpret_cast<float4 *>(rp.s2 + dst * D)[c] = __ldg(reinterpret_cast<const float4 *>(src.s2 + from * D) + c);
    } else {
        rp.s[dst * D + c] = src.s[from * D + c];
        rp.s2[dst * D + c] = src.s2[from * D + c];
    }
    if (c == 0) {
        rp.a[dst] = src.a[from];
        rp.r[dst] = src.r[from];
        rp.done[dst] = src.done[from];
    }
}

}  // namespace

namespace {

void p2p_teardown(goldq_t *h)
{
    for (int r = 0; r < P2P_MAX_WORLD; r++)
        if (h->p2p.peer_base[r] && h->p2p.peer_base[r] != (void *)h->p2p.base) cudaIpcCloseMemHandle(h->p2p.peer_base[r]);
    if (h->p2p.base) cudaFree(h->p2p.base);
    if (h->p2p.err_host) cudaFreeHost(h->p2p.err_host);
    h->p2p = goldq::P2P{};
}

// Export this rank's gradient/flag allocation through CUDA IPC, map every peer's, and agree with
// all ranks (NCCL min) that it worked everywhere; otherwise the group keeps the NCCL all-reduce.
// Returns GOLDQ_OK whether or not peer memory ends up enabled; *why says what prevented it.
int p2p_setup(goldq_t *h, std::string *why)
{
    goldq::P2P &pp = h->p2p;
    const int W = h->world;
    auto nccl_ok = [&](ncclResult_t r, const char *what) {
        if (r == ncclSuccess) return true;
        *why = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        return false;
    };
    int ok = 1;
    if (W > P2P_MAX_WORLD || !g_nccl.AllGather) { ok = 0; *why = "world size above 8 or ncclAllGather missing"; }
    {
        // the exchange spin-waits inside the update kernel: its whole grid has to be co-resident
        // (GOLDQ_P2P_RESIDENT_CAP overrides the device's capacity: test hook for the fallback)
        int grid = 0, resident = 0;
        const bool fits = wide_p2p_resident(h->wide, &grid, &resident);
        const char *cap = getenv("GOLDQ_P2P_RESIDENT_CAP");
        if (cap) resident = atoi(cap);
        if (ok && (!fits || grid > resident)) {
            ok = 0;
            *why = "update grid of " + std::to_string(grid) + " CTAs is not co-resident (device holds " +
                   std::to_string(resident) + "): the in-kernel exchange would wait on CTAs that cannot start";
        }
    }
    // exported allocation: reduce-scatter windows [2][W][chunk_f] + all-gather windows [2][W * chunk_f],
    // 8-byte {value, epoch} words, zeroed (epochs start at 1)
    pp.chunk_f = wide_p2p_chunk_floats(h->wide, W > 0 ? W : 1);
    const size_t bytes = (size_t)4 * W * pp.chunk_f * sizeof(unsigned long long);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok && (cudaMalloc((void **)&pp.base, bytes) != cudaSuccess || cudaMemsetAsync(pp.base, 0, bytes, h->st) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine, pp.base) != cudaSuccess)) {
        ok = 0; *why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError());
    }
    if (ok && (cudaHostAlloc((void **)&pp.err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
               cudaHostGetDevicePointer((void **)&pp.err_dev, pp.err_host, 0) != cudaSuccess)) {
        ok = 0; *why = "mapped host flag unavailable";
    }
    if (pp.err_host) *pp.err_host = 0;
    // all-gather of the 64-byte handles (a rank that failed above still takes part)
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    unsigned char *xchg = nullptr;
    std::vector<cudaIpcMemHandle_t> all(P2P_MAX_WORLD);
    if (cudaMalloc((void **)&xchg, 64 * (size_t)(W > P2P_MAX_WORLD ? W : P2P_MAX_WORLD)) != cudaSuccess)
        return fail(h, GOLDQ_ECUDA, "cudaMalloc failed (IPC exchange)");
    cudaMemcpyAsync(xchg + 64 * h->rank, &mine, 64, cudaMemcpyHostToDevice, h->st);
    if (g_nccl.AllGather && W <= P2P_MAX_WORLD) {
        if (!nccl_ok(g_nccl.AllGather(xchg + 64 * h->rank, xchg, 64, ncclChar, h->comm, h->st), "ncclAllGather")) ok = 0;
        cudaMemcpyAsync(all.data(), xchg, 64 * (size_t)W, cudaMemcpyDeviceToHost, h->st);
    }
    cudaStreamSynchronize(h->st);
    if (ok) {
        for (int r = 0; r < W && ok; r++) {
            if (r == h->rank) { pp.peer_base[r] = pp.base; continue; }
            cudaError_t e = cudaIpcOpenMemHandle(&pp.peer_base[r], all[r], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                ok = 0; pp.peer_base[r] = nullptr;
                *why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
                cudaGetLastError();
            }
        }
    }
    // unanimous or not at all (this all-reduce also orders every rank's memset before any peer's
    // first flag store)
    int *d_ok = reinterpret_cast<int *>(xchg);
    cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->st);
    int all_ok = 0;
    if (nccl_ok(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->st), "ncclAllReduce")) {
        cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->st);
        cudaStreamSynchronize(h->st);
    }
    cudaFree(xchg);
    if (!all_ok) {
        if (why->empty()) *why = "a peer rank could not map the group's memory";
        p2p_teardown(h);
        return GOLDQ_OK;
    }
    pp.on = true;
    pp.epoch = 0;
    return GOLDQ_OK;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int goldq_abi_version(void) { return GOLDQ_ABI_VERSION; }

int goldq_default_config(goldq_config_t *c)
{
    if (!c) return GOLDQ_EINVAL;
    memset(c, 0, sizeof(*c));
    c->abi_version = GOLDQ_ABI_VERSION;
    c->obs_dim = 4;       // CartPole-v0
    c->hidden1 = 24;      // policy.go:55-56
    c->hidden2 = 24;
    c->n_actions = 2;
    c->batch_size = 20;   // policy.go:36
    c->global_batch = 0;
    c->replay_capacity = 10000000;  // agent.go:64 (10e6)
    c->gamma = 0.95f;     // agent.go:62
    c->lr = 0.0005;       // policy.go:34
    c->beta1 = 0.9;       // solvers.go:425-426
    c->beta2 = 0.999;
    c->eps = 1e-8;
    c->huber_delta = 1.0f; // loss.go:102-104
    c->loss = GOLDQ_LOSS_MSE;
    c->solver = GOLDQ_SOLVER_ADAM;
    c->precision = GOLDQ_PREC_FP32;
    c->path = GOLDQ_PATH_AUTO;
    c->device = 0;
    c->n_replicas = 1;
    c->stream = nullptr;
    c->arch = GOLDQ_ARCH_MLP;
    c->in_channels = c->in_height = c->in_width = 0;
    return GOLDQ_OK;
}

const char *goldq_last_error(const goldq_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

goldq_t *goldq_create(const goldq_config_t *cfg)
{
    if (!cfg) { fail(nullptr, GOLDQ_EINVAL, "config is NULL"); return nullptr; }
    if (cfg->abi_version != GOLDQ_ABI_VERSION) { fail(nullptr, GOLDQ_EINVAL, "abi_version mismatch"); return nullptr; }
    if (cfg->obs_dim < 1 || cfg->hidden1 < 1 || cfg->hidden2 < 1 || cfg->n_actions < 1 ||
        cfg->batch_size < 1 || cfg->replay_capacity < 1 || cfg->n_replicas < 1) {
        fail(nullptr, GOLDQ_EINVAL, "dimensions, batch_size, replay_capacity and n_replicas must be >= 1");
        return nullptr;
    }
    if (cfg->loss != GOLDQ_LOSS_MSE && cfg->loss != GOLDQ_LOSS_PSEUDO_HUBER) {
        fail(nullptr, GOLDQ_EUNSUPPORTED, "loss: only model.MSE and model.PseudoHuber are implemented");
        return nullptr;
    }
    if (cfg->solver != GOLDQ_SOLVER_ADAM && cfg->solver != GOLDQ_SOLVER_RMSPROP) {
        fail(nullptr, GOLDQ_EUNSUPPORTED, "solver: only AdamSolver and RMSPropSolver are implemented");
        return nullptr;
    }
    if (cfg->replay_capacity > 0x7fffffffLL) { fail(nullptr, GOLDQ_EINVAL, "replay_capacity must fit int32 indices"); return nullptr; }

    const bool conv = cfg->arch == GOLDQ_ARCH_ATARI_CONV;
    if (cfg->arch != GOLDQ_ARCH_MLP && !conv) { fail(nullptr, GOLDQ_EUNSUPPORTED, "arch: FC-FC-FC and the Atari conv stack are implemented"); return nullptr; }
    AtariNet atari{};
    if (conv) {
        if (cfg->in_channels < 1 || cfg->in_height < 8 || cfg->in_width < 8 ||
            (int64_t)cfg->in_channels * cfg->in_height * cfg->in_width != cfg->obs_dim) {
            fail(nullptr, GOLDQ_EINVAL, "conv policy: obs_dim must equal in_channels * in_height * in_width (each dimension >= 8)");
            return nullptr;
        }
        atari = make_atari_net(cfg->in_channels, cfg->in_height, cfg->in_width, cfg->n_actions);
        if (atari.c3.OH < 1 || atari.c3.OW < 1) { fail(nullptr, GOLDQ_EINVAL, "conv policy: input too small for the 8/4, 4/2, 3/1 stack"); return nullptr; }
        if (cfg->n_replicas != 1 || cfg->precision != GOLDQ_PREC_FP32 || (cfg->path != GOLDQ_PATH_AUTO && cfg->path != GOLDQ_PATH_CONV)) {
            fail(nullptr, GOLDQ_EUNSUPPORTED, "conv policy: one replica, GOLDQ_PREC_FP32, GOLDQ_PATH_AUTO or GOLDQ_PATH_CONV");
            return nullptr;
        }
    } else if (cfg->path == GOLDQ_PATH_CONV) {
        fail(nullptr, GOLDQ_EINVAL, "GOLDQ_PATH_CONV needs arch = GOLDQ_ARCH_ATARI_CONV");
        return nullptr;
    }

    goldq_t *h = new goldq();
    h->cfg = *cfg;
    h->dims = Dims{cfg->obs_dim, conv ? atari.HID : cfg->hidden1, conv ? atari.HID : cfg->hidden2, cfg->n_actions};
    h->off = make_offsets(h->dims);
    h->P = conv ? (int)atari.P : h->off.P;
    h->atari = atari;
    h->R = cfg->n_replicas;
    h->B = cfg->batch_size;
    h->Bglobal = cfg->global_batch > 0 ? cfg->global_batch : cfg->batch_size;
    h->dev = cfg->device;

    auto bail = [&](const std::string &msg) -> goldq_t * {
        g_create_error = msg.empty() ? h->err : msg;
        goldq_destroy(h);
        return nullptr;
    };

    cudaError_t e = cudaSetDevice(h->dev);
    if (e != cudaSuccess) return bail(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, h->dev);
    if (e != cudaSuccess) return bail(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    h->sm_count = prop.multiProcessorCount;
    if (prop.major != 10)
        return bail("libgoldq is built for sm_100a (B200) only; device is sm_" +
                    std::to_string(prop.major) + std::to_string(prop.minor));

    // path selection
    int path = cfg->path;
    if (conv) path = GOLDQ_PATH_CONV;
    if (path == GOLDQ_PATH_AUTO) {
        if (cfg->precision == GOLDQ_PREC_BF16) path = GOLDQ_PATH_WIDE;
        else if (narrow_supported(h->dims)) path = GOLDQ_PATH_NARROW;
        else path = GOLDQ_PATH_GENERIC;
    }
    if (path == GOLDQ_PATH_NARROW && !narrow_supported(h->dims))
        return bail("narrow path: shape not instantiated (24-24 hidden, (D,A) in {(4,2),(2,3),(6,3),(8,4)})");
    if (path == GOLDQ_PATH_NARROW && cfg->precision != GOLDQ_PREC_FP32)
        return bail("narrow path computes in fp32");
    if (path == GOLDQ_PATH_WIDE && cfg->precision != GOLDQ_PREC_BF16)
        return bail("wide (tcgen05) path requires GOLDQ_PREC_BF16");
    if ((path == GOLDQ_PATH_GENERIC || path == GOLDQ_PATH_CONV) && cfg->precision != GOLDQ_PREC_FP32)
        return bail("generic path computes in fp32");
    if (h->R > 1 && path != GOLDQ_PATH_NARROW)
        return bail("n_replicas > 1 is implemented on the narrow path only");
    h->path = path;

    if (cfg->stream) h->st = (cudaStream_t)cfg->stream;
    else {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, getenv("GOLDQ_PRIO") ? hi : 0);
        if (e != cudaSuccess) return bail(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
        h->own_stream = true;
    }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    cudaStreamCreateWithFlags(&h->copy_st, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&h->rb_st, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_learned, cudaEventDisableTiming);
    for (auto &g : h->stg) {
        cudaEventCreateWithFlags(&g.copied, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g.scattered, cudaEventDisableTiming);
    }
    for (int i = 0; i < goldq::STEP_RING; i++) {
        cudaEventCreateWithFlags(&h->ev_steps[i], cudaEventDisableTiming);
        if (cudaHostAlloc((void **)&h->h_idx[i], (size_t)(h->R > 0 ? h->R : 1) * h->B * 4, cudaHostAllocMapped) != cudaSuccess)
            return bail("cudaHostAlloc failed (index staging)");
        if (cudaHostAlloc((void **)&h->h_steps[i], sizeof(StepDev) * goldq::GRAPH_MAX, cudaHostAllocMapped) != cudaSuccess)
            return bail("cudaHostAlloc failed (step slots)");
    }
    if (dalloc(&h->d_steps, (size_t)goldq::GRAPH_MAX) != cudaSuccess) return bail("cudaMalloc failed (step slots)");

    size_t RP = (size_t)h->R * h->P;
    bool ok = true;
    ok &= dalloc(&h->theta, RP) == cudaSuccess;
    ok &= dalloc(&h->theta_t, RP) == cudaSuccess;
    ok &= dalloc(&h->m, RP) == cudaSuccess;
    ok &= dalloc(&h->v, RP) == cudaSuccess;
    ok &= dalloc(&h->flat, (size_t)h->R * (h->P + 1)) == cudaSuccess;
    // the loss lives in mapped pinned host memory: the kernels' (posted) 4-byte store IS the readback, so
    // goldq_last_loss is an event wait plus a host load -- no copy, no copy-engine queue, no second stream
    ok &= cudaHostAlloc((void **)&h->loss, (size_t)h->R * 4, cudaHostAllocMapped) == cudaSuccess;
    if (ok) memset(h->loss, 0, (size_t)h->R * 4);
    ok &= dalloc(&h->d_idx, (size_t)h->R * h->B) == cudaSuccess;
    size_t cap = (size_t)cfg->replay_capacity, RC = (size_t)h->R * cap;
    ok &= dalloc(&h->rp.s, RC * h->dims.D) == cudaSuccess;
    ok &= dalloc(&h->rp.s2, RC * h->dims.D) == cudaSuccess;
    ok &= dalloc(&h->rp.a, RC) == cudaSuccess;
    ok &= dalloc(&h->rp.r, RC) == cudaSuccess;
    ok &= dalloc(&h->rp.done, RC) == cudaSuccess;
    if (!ok) return bail(std::string("cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->rp.cap = (int64_t)cap;
    h->rp.len = 0;
    h->rp.head = 0;
    cudaMemsetAsync(h->theta, 0, RP * 4, h->st);
    cudaMemsetAsync(h->theta_t, 0, RP * 4, h->st);
    cudaMemsetAsync(h->m, 0, RP * 4, h->st);
    cudaMemsetAsync(h->v, 0, RP * 4, h->st);

    if (path == GOLDQ_PATH_NARROW) {
        NarrowArgs a{};
        a.B = h->B;
        a.replicas = h->R;
        narrow_plan(a, h->sm_count);
        h->n_ctas = a.ctas_per_replica;
        h->n_warps = a.warps_per_cta;
        narrow_configure(h->dims, h->n_warps);
        size_t nt = (size_t)h->R * (NARROW_MAX_GROUPS + 1);
        ok &= dalloc(&h->partial, (size_t)h->R * narrow_partial_floats(h->n_ctas, h->P)) == cudaSuccess;
        ok &= dalloc(&h->ticket, nt) == cudaSuccess;
        if (!ok) return bail("cudaMalloc failed (narrow scratch)");
        cudaMemsetAsync(h->ticket, 0, nt * sizeof(unsigned), h->st);
    } else if (path == GOLDQ_PATH_WIDE) {
        std::string err;
        h->wide = wide_create(h->dims, h->B, h->sm_count, err);
        if (!h->wide) return bail("wide path: " + err);
    }
    e = cudaStreamSynchronize(h->st);
    if (e != cudaSuccess) return bail(std::string("init sync: ") + cudaGetErrorString(e));
    return h;
}

void goldq_destroy(goldq_t *h)
{
    if (!h) return;
    cudaSetDevice(h->dev);
    if (h->st) cudaStreamSynchronize(h->st);
    drop_graphs(h);
    for (int i = 0; i < goldq::STEP_RING; i++) {
        if (h->h_steps[i]) cudaFreeHost(h->h_steps[i]);
        if (h->h_idx[i]) cudaFreeHost(h->h_idx[i]);
        if (h->ev_steps[i]) cudaEventDestroy(h->ev_steps[i]);
    }
    if (h->d_steps) cudaFree(h->d_steps);
    p2p_teardown(h);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->wide) wide_destroy(h->wide);
    if (h->loss) cudaFreeHost(h->loss);
    void *ptrs[] = {h->theta, h->theta_t, h->m, h->v, h->flat, h->d_idx, h->rp.s, h->rp.s2,
                    h->rp.a, h->rp.r, h->rp.done, h->stg[0].base, h->stg[1].base, h->stg[2].base, h->partial, h->ticket, h->bs, h->bs2, h->br, h->ba, h->bdone, h->h1,
                    h->h2, h->q, h->t1, h->t2, h->qt, h->td, h->dq, h->dh2, h->dh1, h->ylab, h->mk1,
                    h->mk2, h->gslice, h->loss_part, h->px, h->ph1, h->ph2, h->pq, h->pact,
                    h->dbg_q, h->dbg_qt, h->dbg_td, h->dbg_idx, h->cy1, h->cy2, h->cy3, h->ch4, h->cu1, h->cu2, h->cu3, h->cu4,
                    h->cm1, h->cm2, h->cm3, h->cm4, h->cdy1, h->cdy2, h->cdy3, h->cdh4};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (auto &g : h->stg) {
        if (g.copied) cudaEventDestroy(g.copied);
        if (g.scattered) cudaEventDestroy(g.scattered);
    }
    if (h->copy_st) { cudaStreamSynchronize(h->copy_st); cudaStreamDestroy(h->copy_st); }
    if (h->rb_st) { cudaStreamSynchronize(h->rb_st); cudaStreamDestroy(h->rb_st); }
    if (h->ev_learned) cudaEventDestroy(h->ev_learned);
    if (h->own_stream && h->st) cudaStreamDestroy(h->st);
    delete h;
}

int64_t goldq_param_count(const goldq_t *h) { return h ? h->P : 0; }
int goldq_selected_path(const goldq_t *h) { return h ? h->path : GOLDQ_EINVAL; }
int64_t goldq_replay_len(const goldq_t *h) { return h ? h->rp.len : 0; }
int64_t goldq_launch_count(const goldq_t *h) { return h ? h->launches : 0; }
void *goldq_stream(goldq_t *h) { return h ? (void *)h->st : nullptr; }

int goldq_sync(goldq_t *h)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    CK(h, cudaStreamSynchronize(h->st));
    if (h->p2p.on && *(volatile int *)h->p2p.err_host)
        return fail(h, GOLDQ_ENCCL, "data-parallel peer did not publish its gradient within 5 s");
    return GOLDQ_OK;
}

int goldq_set_params(goldq_t *h, int net, const float *theta)
{
    if (!h || !theta || (net != GOLDQ_NET_ONLINE && net != GOLDQ_NET_TARGET))
        return fail(h, GOLDQ_EINVAL, "set_params: bad argument");
    CK(h, cudaSetDevice(h->dev));
    float *dst = net == GOLDQ_NET_ONLINE ? h->theta : h->theta_t;
    CK(h, cudaMemcpyAsync(dst, theta, (size_t)h->R * h->P * 4, cudaMemcpyHostToDevice, h->st));
    if (h->wide) {
        std::string err;
        int n = wide_set_params(h->wide, net, dst, h->st, err);
        if (n < 0) return fail(h, GOLDQ_ECUDA, err);
        h->launches += n;
    }
    CK(h, cudaStreamSynchronize(h->st));
    return GOLDQ_OK;
}

int goldq_get_params(goldq_t *h, int net, float *theta)
{
    if (!h || !theta || (net != GOLDQ_NET_ONLINE && net != GOLDQ_NET_TARGET))
        return fail(h, GOLDQ_EINVAL, "get_params: bad argument");
    CK(h, cudaSetDevice(h->dev));
    const float *src = net == GOLDQ_NET_ONLINE ? h->theta : h->theta_t;
    CK(h, cudaMemcpyAsync(theta, src, (size_t)h->R * h->P * 4, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return GOLDQ_OK;
}

int goldq_set_adam_state(goldq_t *h, const float *m, const float *v, int64_t iter)
{
    if (!h || !m || !v || iter < 0) return fail(h, GOLDQ_EINVAL, "set_adam_state: bad argument");
    CK(h, cudaSetDevice(h->dev));
    CK(h, cudaMemcpyAsync(h->m, m, (size_t)h->R * h->P * 4, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaMemcpyAsync(h->v, v, (size_t)h->R * h->P * 4, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    h->iter = iter;
    return GOLDQ_OK;
}

int goldq_get_adam_state(goldq_t *h, float *m, float *v, int64_t *iter)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    if (m) CK(h, cudaMemcpyAsync(m, h->m, (size_t)h->R * h->P * 4, cudaMemcpyDeviceToHost, h->st));
    if (v) CK(h, cudaMemcpyAsync(v, h->v, (size_t)h->R * h->P * 4, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    if (iter) *iter = h->iter;
    return GOLDQ_OK;
}

// Two halves of a push.  stage_batch: validate, then start the H2D copy into a free staging set on the copy
// stream (`packed` != NULL: one host buffer in the packed layout, ONE copy; else the five arrays, five copies
// into the same device layout).  commit_batch: scatter the OLDEST staged batch into the ring, in stream order
// on the learn stream.  goldq_replay_push* = stage + commit; goldq_replay_stage_packed / goldq_replay_commit let a
// caller start the copy of batch t+2 before it waits for learn step t, so that the copy engine always has the
// next copy queued (tools/e2e_probe.py).
constexpr int NSTG = 3;
static int commit_batch(goldq_t *h)
{
    if (h->stg_pending < 1) return fail(h, GOLDQ_EINVAL, "replay_commit: no staged batch");
    goldq::Staging &g = h->stg[(h->stg_cur + NSTG - h->stg_pending) % NSTG];
    h->stg_pending -= 1;
    const int D = h->dims.D, R = h->R;
    const int64_t n = g.n, rows = (int64_t)R * n;
    const PackedView dv = packed_view(g.base, rows, D);
    // main stream: the ring is updated after everything enqueued so far (in stream order)
    CK(h, cudaStreamWaitEvent(h->st, g.copied, 0));
    // only the newest `cap` of the n items survive (PushFront + PopBack, agent.go:224-229)
    int64_t skip = n > h->rp.cap ? n - h->rp.cap : 0;
    int64_t mcount = n - skip;
    if ((D & 3) == 0 && (((size_t)rows * D * 4) & 15) == 0) {
        int64_t total = mcount * (D >> 2) * R;
        replay_scatter_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, h->st>>>(h->rp, D, R, n, skip, dv);
    } else {
        int64_t total = mcount * D * R;
        replay_scatter_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, h->st>>>(h->rp, D, R, n, skip, dv);
    }
    CK(h, cudaEventRecord(g.scattered, h->st));
    h->launches += 1;
    h->rp.head = (h->rp.head + mcount) % h->rp.cap;
    h->rp.len = h->rp.len + mcount > h->rp.cap ? h->rp.cap : h->rp.len + mcount;
    return check_launch(h, "replay_scatter_kernel");
}

static int stage_batch(goldq_t *h, int64_t n, const void *packed, const float *s, const int32_t *a, const float *r,
                       const float *s2, const uint8_t *done)
{
    // the learn step indexes Q rows, W3 columns and dQ with the stored action: an action outside
    // [0, n_actions) would be an out-of-bounds device access (the reference panics in
    // qValues.Set(event.Action, ...), agent.go:155); reject it here
    {
        const int64_t total = (int64_t)h->R * n;
        const uint32_t A = (uint32_t)h->dims.A;
        uint32_t worst = 0;
        for (int64_t i = 0; i < total; i++) worst = (uint32_t)a[i] > worst ? (uint32_t)a[i] : worst;
        if (worst >= A)
            for (int64_t i = 0; i < total; i++)
                if ((uint32_t)a[i] >= A)
                    return fail(h, GOLDQ_EINVAL, "replay_push: action " + std::to_string(a[i]) + " at item " +
                                                     std::to_string(i) + " is outside [0, n_actions)");
    }
    CK(h, cudaSetDevice(h->dev));
    const int D = h->dims.D, R = h->R;
    if (h->stg_pending >= NSTG - 1)
        return fail(h, GOLDQ_EINVAL, "replay_stage: two batches are staged already; goldq_replay_commit one first");
    if (h->stg_cap < n) {
        while (h->stg_pending > 0) {     // staged batches keep their order: into the ring before the sets are replaced
            int rc = commit_batch(h);
            if (rc) return rc;
        }
        CK(h, cudaStreamSynchronize(h->copy_st));
        CK(h, cudaStreamSynchronize(h->st));
        int64_t cap = n < 1024 ? 1024 : n;
        for (auto &g : h->stg) {
            if (g.base) cudaFree(g.base);
            g.base = nullptr;
            CK(h, dalloc(&g.base, packed_bytes((int64_t)R * cap, D)));
        }
        h->stg_cap = cap;
    }
    goldq::Staging &g = h->stg[h->stg_cur];
    h->stg_cur = (h->stg_cur + 1) % NSTG;
    h->stg_pending += 1;
    g.n = n;
    const int64_t rows = (int64_t)R * n;
    const PackedView dv = packed_view(g.base, rows, D);
    // copy stream: wait until the scatter that last read this staging set is done, then copy
    CK(h, cudaStreamWaitEvent(h->copy_st, g.scattered, 0));
    if (packed) {
        CK(h, cudaMemcpyAsync(g.base, packed, packed_bytes(rows, D), cudaMemcpyHostToDevice, h->copy_st));
    } else {
        CK(h, cudaMemcpyAsync((void *)dv.s, s, (size_t)rows * D * 4, cudaMemcpyHostToDevice, h->copy_st));
        CK(h, cudaMemcpyAsync((void *)dv.s2, s2, (size_t)rows * D * 4, cudaMemcpyHostToDevice, h->copy_st));
        CK(h, cudaMemcpyAsync((void *)dv.a, a, (size_t)rows * 4, cudaMemcpyHostToDevice, h->copy_st));
        CK(h, cudaMemcpyAsync((void *)dv.r, r, (size_t)rows * 4, cudaMemcpyHostToDevice, h->copy_st));
        CK(h, cudaMemcpyAsync((void *)dv.done, done, (size_t)rows, cudaMemcpyHostToDevice, h->copy_st));
    }
    CK(h, cudaEventRecord(g.copied, h->copy_st));
    h->ev_push = g.copied;
    return GOLDQ_OK;
}

static int push_common(goldq_t *h, int64_t n, const void *packed, const float *s, const int32_t *a, const float *r,
                       const float *s2, const uint8_t *done)
{
    while (h->stg_pending > 0) {         // batches staged earlier go into the ring first
        int rc = commit_batch(h);
        if (rc) return rc;
    }
    int rc = stage_batch(h, n, packed, s, a, r, s2, done);
    return rc ? rc : commit_batch(h);
}

int goldq_replay_push(goldq_t *h, int64_t n, const float *s, const int32_t *a, const float *r,
                      const float *s2, const uint8_t *done)
{
    if (!h || n < 0 || (n > 0 && (!s || !a || !r || !s2 || !done)))
        return fail(h, GOLDQ_EINVAL, "replay_push: bad argument");
    if (n == 0) return GOLDQ_OK;
    return push_common(h, n, nullptr, s, a, r, s2, done);
}

int64_t goldq_packed_bytes(const goldq_t *h, int64_t n)
{
    if (!h || n < 0) return 0;
    return (int64_t)packed_bytes((int64_t)h->R * n, h->dims.D);
}

int goldq_replay_push_packed(goldq_t *h, int64_t n, const void *packed)
{
    if (!h || n < 0 || (n > 0 && !packed)) return fail(h, GOLDQ_EINVAL, "replay_push_packed: bad argument");
    if (n == 0) return GOLDQ_OK;
    const PackedView hv = packed_view(packed, (int64_t)h->R * n, h->dims.D);
    return push_common(h, n, packed, hv.s, hv.a, hv.r, hv.s2, hv.done);
}

int goldq_replay_stage_packed(goldq_t *h, int64_t n, const void *packed)
{
    if (!h || n < 0 || (n > 0 && !packed)) return fail(h, GOLDQ_EINVAL, "replay_stage_packed: bad argument");
    if (n == 0) return GOLDQ_OK;
    const PackedView hv = packed_view(packed, (int64_t)h->R * n, h->dims.D);
    return stage_batch(h, n, packed, hv.s, hv.a, hv.r, hv.s2, hv.done);
}

int goldq_replay_commit(goldq_t *h)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    return commit_batch(h);
}

int goldq_replay_push_wait(goldq_t *h)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    if (h->ev_push) CK(h, cudaEventSynchronize(h->ev_push));
    return GOLDQ_OK;
}

void *goldq_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void goldq_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int goldq_learn(goldq_t *h, const int32_t *idx, uint64_t seed)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    if (h->rp.len < h->B) return fail(h, GOLDQ_ENOTREADY, "replay shorter than batch");
    const int32_t *idx_dev = nullptr;
    if (idx) {
        size_t n = (size_t)h->R * h->B;
        uint32_t worst = 0;      // (branch-free so that the compiler vectorises it: this loop is on the e2e path)
        for (size_t i = 0; i < n; i++) worst = (uint32_t)idx[i] > worst ? (uint32_t)idx[i] : worst;
        if ((int64_t)worst >= h->rp.len) return fail(h, GOLDQ_EINVAL, "learn: sampled index out of range");
        // one launch instead of the step's kernels, events and cross-stream waits: Agent.Learn after
        // every Remember replays the same one-step graph (the ring position and the index list travel
        // through the step's staging slot)
        if (use_graphs(h)) return learned(h, launch_graph(h, 0, seed, idx));
        CK(h, cudaMemcpyAsync(h->d_idx, idx, n * 4, cudaMemcpyHostToDevice, h->st));
        idx_dev = h->d_idx;
    }
    if (use_graphs(h)) return learned(h, launch_graph(h, 1, seed));
    return learned(h, run_learn(h, idx_dev, seed));
}

/* Sequential.ResizeBatch (goro model.go:485-497): re-allocate everything sized by the batch -- index list,
 * narrow-path partials, the wide path's activations / tensor maps / split-K plan, debug snapshots -- for a new
 * batch size.  Parameters, solver state and the replay ring are untouched; cached step graphs are dropped. */
int goldq_resize_batch(goldq_t *h, int32_t n)
{
    if (!h || n < 1) return fail(h, GOLDQ_EINVAL, "resize_batch: batch size must be >= 1");
    if (h->world > 1) return fail(h, GOLDQ_EUNSUPPORTED, "resize_batch: not inside a data-parallel group (global_batch is fixed)");
    if (n == h->B) return GOLDQ_OK;
    CK(h, cudaSetDevice(h->dev));
    CK(h, cudaStreamSynchronize(h->st));
    CK(h, cudaStreamSynchronize(h->copy_st));
    drop_graphs(h);
    const int oldB = h->B;
    WideState *nw = nullptr;
    if (h->path == GOLDQ_PATH_WIDE) {        // build the new state first: a shape error leaves the handle as it was
        std::string err;
        nw = wide_create(h->dims, n, h->sm_count, err);
        if (!nw) return fail(h, GOLDQ_EINVAL, "resize_batch (wide path): " + err);
    }
    int32_t *idx = nullptr;
    if (dalloc(&idx, (size_t)h->R * n) != cudaSuccess) { if (nw) wide_destroy(nw); return fail(h, GOLDQ_ECUDA, "cudaMalloc failed (index list)"); }
    cudaFree(h->d_idx);
    h->d_idx = idx;
    for (int i = 0; i < goldq::STEP_RING; i++) {
        if (h->h_idx[i]) cudaFreeHost(h->h_idx[i]);
        h->h_idx[i] = nullptr;
        CK(h, cudaHostAlloc((void **)&h->h_idx[i], (size_t)h->R * n * 4, cudaHostAllocMapped));
    }
    h->B = n;
    h->Bglobal = n;
    h->cfg.batch_size = n;
    h->cfg.global_batch = 0;
    if (h->path == GOLDQ_PATH_NARROW) {
        NarrowArgs a{};
        a.B = h->B;
        a.replicas = h->R;
        narrow_plan(a, h->sm_count);
        h->n_ctas = a.ctas_per_replica;
        h->n_warps = a.warps_per_cta;
        narrow_configure(h->dims, h->n_warps);
        if (h->partial) cudaFree(h->partial);
        h->partial = nullptr;
        CK(h, dalloc(&h->partial, (size_t)h->R * narrow_partial_floats(h->n_ctas, h->P)));
        CK(h, cudaMemsetAsync(h->ticket, 0, (size_t)h->R * (NARROW_MAX_GROUPS + 1) * sizeof(unsigned), h->st));
    } else if (h->path == GOLDQ_PATH_WIDE) {
        wide_destroy(h->wide);
        h->wide = nw;
        wide_set_debug(h->wide, h->debug);
        std::string err;
        for (int net = 0; net < 2; net++) {      // bf16 shadows of the (unchanged) fp32 master parameters
            int k = wide_set_params(h->wide, net, net == GOLDQ_NET_ONLINE ? h->theta : h->theta_t, h->st, err);
            if (k < 0) return fail(h, GOLDQ_ECUDA, err);
            h->launches += k;
        }
    }
    // (the generic path's scratch grows on demand: ensure_generic)
    if (h->dbg_q) {
        void *old[] = {h->dbg_q, h->dbg_qt, h->dbg_td, h->dbg_idx};
        for (void *p : old) cudaFree(p);
        h->dbg_q = h->dbg_qt = h->dbg_td = nullptr;
        h->dbg_idx = nullptr;
        size_t RB = (size_t)h->R * h->B;
        CK(h, dalloc(&h->dbg_q, RB * h->dims.A));
        CK(h, dalloc(&h->dbg_qt, RB * h->dims.A));
        CK(h, dalloc(&h->dbg_td, RB));
        CK(h, dalloc(&h->dbg_idx, RB));
    }
    (void)oldB;
    CK(h, cudaStreamSynchronize(h->st));
    return GOLDQ_OK;
}

int32_t goldq_batch_size(const goldq_t *h) { return h ? h->B : 0; }

int goldq_learn_prepare(goldq_t *h)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    if (h->rp.len < h->B) return fail(h, GOLDQ_ENOTREADY, "replay shorter than batch");
    if (!use_graphs(h)) return GOLDQ_OK;
    if (h->path == GOLDQ_PATH_GENERIC || h->path == GOLDQ_PATH_CONV) {          // no allocation may happen under capture
        int rc = h->path == GOLDQ_PATH_CONV ? ensure_conv(h, h->B) : ensure_generic(h, h->B);
        if (rc) return rc;
    }
    if (h->graph_debug != h->debug) { drop_graphs(h); h->graph_debug = h->debug; }
    for (int slot = 0; slot <= goldq::GRAPH_CHUNK; slot = graph_next_size(slot)) {
        if (h->graph_exec[slot]) continue;
        int rc = build_graph(h, slot);
        if (rc) return rc;
    }
    return GOLDQ_OK;
}

int64_t goldq_graph_builds(const goldq_t *h) { return h ? h->graph_builds : 0; }

int goldq_learn_device_indices(goldq_t *h, const int32_t *idx_dev)
{
    if (!h || !idx_dev) return fail(h, GOLDQ_EINVAL, "learn_device_indices: bad argument");
    CK(h, cudaSetDevice(h->dev));
    return learned(h, run_learn(h, idx_dev, 0));
}

static int learn_many_impl(goldq_t *h, int32_t n, uint64_t seed);
int goldq_learn_many(goldq_t *h, int32_t n, uint64_t seed)
{
    if (!h || n < 0) return fail(h, GOLDQ_EINVAL, "learn_many: bad argument");
    return learned(h, learn_many_impl(h, n, seed));
}
static int learn_many_impl(goldq_t *h, int32_t n, uint64_t seed)
{
    CK(h, cudaSetDevice(h->dev));
    if (h->rp.len < h->B) return fail(h, GOLDQ_ENOTREADY, "replay shorter than batch");
    if (h->path == GOLDQ_PATH_GENERIC || h->path == GOLDQ_PATH_CONV) {          // no allocation may happen under capture
        int rc = h->path == GOLDQ_PATH_CONV ? ensure_conv(h, h->B) : ensure_generic(h, h->B);
        if (rc) return rc;
    }
    static const bool no_graph = getenv("GOLDQ_NO_GRAPH") != nullptr;   // diagnostic: eager launches
    // Data-parallel steps: an ncclAllReduce captured into a CUDA graph replays ~80 us slower per
    // step than the eager call (measured at N=2), so only the part of the step before the
    // all-reduce is replayed from a graph; the collective and the update are launched eagerly.
    if (!no_graph && h->world > 1 && h->path == GOLDQ_PATH_WIDE && !h->p2p.on) {
        if (!h->dp_graph) {
            const int64_t iter0 = h->iter, launches0 = h->launches;
            cudaGraph_t graph = nullptr;
            CK(h, cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
            int rc = learn_wide(h, nullptr, 0, h->d_steps, 1);
            cudaError_t e = cudaStreamEndCapture(h->st, &graph);
            h->dp_graph_launches = h->launches - launches0;
            h->iter = iter0;
            h->launches = launches0;
            if (rc != GOLDQ_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(h, GOLDQ_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            e = cudaGraphInstantiate(&h->dp_graph, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return fail(h, GOLDQ_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        }
        while (n > 0) {
            const int c = n < goldq::GRAPH_MAX ? n : goldq::GRAPH_MAX;
            const int b = h->steps_buf;
            h->steps_buf = (b + 1) % goldq::STEP_RING;
            CK(h, cudaEventSynchronize(h->ev_steps[b]));
            for (int i = 0; i < c; i++) {
                AdamConsts ac = adam_consts(h->cfg, h->iter + 1 + i);
                h->h_steps[b][i].seed = seed + (uint64_t)i;
                h->h_steps[b][i].c1 = ac.c1;
                h->h_steps[b][i].c2 = ac.c2;
                h->h_steps[b][i].head = h->rp.head;
                h->h_steps[b][i].len = h->rp.len;
            }
            for (int i = 0; i < c; i++) {
                // slot 0 is rewritten in stream order, after the previous step's kernels read it
                CK(h, cudaMemcpyAsync(h->d_steps, &h->h_steps[b][i], sizeof(StepDev), cudaMemcpyHostToDevice, h->st));
                CK(h, cudaGraphLaunch(h->dp_graph, h->st));
                h->iter += 1;
                h->launches += h->dp_graph_launches;
                int rc = learn_wide(h, nullptr, 0, h->d_steps, 2);
                if (rc) return rc;
            }
            CK(h, cudaEventRecord(h->ev_steps[b], h->st));
            seed += (uint64_t)c;
            n -= c;
        }
        return GOLDQ_OK;
    }
    if (no_graph || (h->world > 1 && !h->p2p.on)) {
        for (int i = 0; i < n; i++) {
            int rc = run_learn(h, nullptr, seed + (uint64_t)i);
            if (rc) return rc;
        }
        return GOLDQ_OK;
    }
    // Graphs of 32, 28, ... 8, 4, 2 and 1 steps (each instantiated once, on first use or by goldq_learn_prepare): any
    // step count is replayed from graphs, largest first.  A graph start costs about 16 us (step scalars staged, no
    // prefetched batch for its first step), so 20 steps are one graph, not 16 + 4.
    while (n > 0) {
        const int c = graph_size_for(n);
        int rc = launch_graph(h, c, seed);
        if (rc) return rc;
        seed += (uint64_t)c;
        n -= c;
    }
    return GOLDQ_OK;
}

/* Diagnostic: run ONE eager learn step (device sampler) with a CUDA event after every
 * kernel launch, everything serialised on the handle's stream, and return the time of each
 * launch.  Used by bench.py to time the dominant kernel live (roofline.achieved). */
int goldq_step_breakdown(goldq_t *h, uint64_t seed, int32_t max_kernels, float *ms, int32_t *n_kernels,
                         char *names, int32_t names_bytes)
{
    if (!h || !ms || !n_kernels || max_kernels < 1) return fail(h, GOLDQ_EINVAL, "step_breakdown: bad argument");
    CK(h, cudaSetDevice(h->dev));
    const int cap = 64;
    std::vector<cudaEvent_t> evs(cap + 1);
    for (auto &e : evs) CK(h, cudaEventCreate(&e));
    CK(h, cudaStreamSynchronize(h->st));
    int rc;
    if (h->path == GOLDQ_PATH_WIDE) {
        h->prof_evs = evs.data() + 1;
        h->prof_n = 0;
        CK(h, cudaEventRecord(evs[0], h->st));
        rc = run_learn(h, nullptr, seed);
        h->prof_evs = nullptr;
        if (names && names_bytes > 0)
            snprintf(names, names_bytes, "%s%s", wide_step_kernel_names(h->wide).c_str(),
                     (h->world > 1 && !h->p2p.on) ? "\nncclAllReduce(grad)\nAdam + bf16 shadow (K5)" : "");
    } else {
        // single fused launch (narrow) or an unfused chain (generic): one interval for the step
        CK(h, cudaEventRecord(evs[0], h->st));
        rc = run_learn(h, nullptr, seed);
        CK(h, cudaEventRecord(evs[1], h->st));
        h->prof_n = 1;
        if (names && names_bytes > 0)
            snprintf(names, names_bytes, "%s", h->path == GOLDQ_PATH_NARROW ? "narrow_learn_kernel (whole step)" : "generic path (whole step)");
    }
    if (rc == GOLDQ_OK) {
        cudaError_t e = cudaStreamSynchronize(h->st);
        if (e != cudaSuccess) rc = fail(h, GOLDQ_ECUDA, cudaGetErrorString(e));
    }
    rc = learned(h, rc);
    int n = h->prof_n < max_kernels ? h->prof_n : max_kernels;
    if (rc == GOLDQ_OK)
        for (int i = 0; i < n; i++) cudaEventElapsedTime(&ms[i], evs[i], evs[i + 1]);
    *n_kernels = n;
    for (auto &e : evs) cudaEventDestroy(e);
    return rc;
}

int goldq_sync_target(goldq_t *h)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    CK(h, cudaMemcpyAsync(h->theta_t, h->theta, (size_t)h->R * h->P * 4, cudaMemcpyDeviceToDevice,
                          h->st));
    h->launches += 1;
    if (h->wide) {
        std::string err;
        int n = wide_sync_target(h->wide, h->st, err);
        if (n < 0) return fail(h, GOLDQ_ECUDA, err);
        h->launches += n;
    }
    return GOLDQ_OK;
}

int goldq_last_loss(goldq_t *h, float *out)
{
    if (!h || !out) return fail(h, GOLDQ_EINVAL, "last_loss: bad argument");
    CK(h, cudaSetDevice(h->dev));
    // wait for the latest learn step (ev_learned), not for whatever was enqueued behind it; the loss was
    // stored into mapped host memory by the step itself (see goldq_create)
    CK(h, cudaEventSynchronize(h->ev_learned));
    memcpy(out, h->loss, (size_t)h->R * 4);
    if (h->p2p.on && *(volatile int *)h->p2p.err_host)
        return fail(h, GOLDQ_ENCCL, "data-parallel peer did not publish its gradient within 5 s");
    return GOLDQ_OK;
}

static int ensure_predict(goldq_t *h, int64_t n)
{
    if (h->pred_rows >= n) return GOLDQ_OK;
    void *old[] = {h->px, h->ph1, h->ph2, h->pq, h->pact};
    for (void *p : old)
        if (p) cudaFree(p);
    int64_t cap = n < 256 ? 256 : n;
    const Dims &d = h->dims;
    CK(h, dalloc(&h->px, (size_t)cap * d.D));
    CK(h, dalloc(&h->ph1, (size_t)cap * d.H1));
    CK(h, dalloc(&h->ph2, (size_t)cap * d.H2));
    CK(h, dalloc(&h->pq, (size_t)cap * d.A));
    CK(h, dalloc(&h->pact, (size_t)cap));
    h->pred_rows = cap;
    return GOLDQ_OK;
}

static int predict_dev(goldq_t *h, int net, int replica, int64_t n, const float *x)
{
    if (replica < 0 || replica >= h->R) return fail(h, GOLDQ_EINVAL, "predict: replica out of range");
    if (n < 1 || n > 0x7fffffff / (h->dims.D > h->dims.H1 ? h->dims.D : h->dims.H1))
        return fail(h, GOLDQ_EINVAL, "predict: bad row count");
    int rc = ensure_predict(h, n);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(h->px, x, (size_t)n * h->dims.D * 4, cudaMemcpyHostToDevice, h->st));
    const float *theta = (net == GOLDQ_NET_ONLINE ? h->theta : h->theta_t) + (size_t)replica * h->P;
    // bf16 handles: batches of 128 rows and more (vectorised environments, PredictBatch) go through the fused
    // tcgen05 forward and the bf16 weight shadows the learn step trains with, so acting and learning see the
    // same arithmetic; a single state (Agent.Action) keeps the fp32 reference-arithmetic kernels
    if (h->wide && n >= 128) {
        std::string err;
        int k = wide_predict(h->wide, net, h->px, n, h->theta, h->theta_t, h->pq, h->st, err);
        if (k < 0) return fail(h, GOLDQ_ECUDA, err);
        if (k > 0) { h->launches += k; return check_launch(h, "predict (tcgen05)"); }
    }
    if (h->path == GOLDQ_PATH_CONV) {
        rc = ensure_conv(h, (int)n > h->B ? (int)n : h->B);
        if (rc) return rc;
        conv_forward(h, theta, h->px, (int)n, h->cu1, nullptr, h->cu2, nullptr, h->cu3, nullptr, h->cu4, nullptr, h->pq);
        return check_launch(h, "predict (conv)");
    }
    generic_forward(h, theta, h->px, (int)n, h->ph1, nullptr, h->ph2, nullptr, h->pq);
    return check_launch(h, "predict");
}

int goldq_predict(goldq_t *h, int net, int replica, int64_t n, const float *x, float *q)
{
    if (!h || !x || !q || (net != GOLDQ_NET_ONLINE && net != GOLDQ_NET_TARGET))
        return fail(h, GOLDQ_EINVAL, "predict: bad argument");
    CK(h, cudaSetDevice(h->dev));
    int rc = predict_dev(h, net, replica, n, x);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(q, h->pq, (size_t)n * h->dims.A * 4, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return GOLDQ_OK;
}

int goldq_greedy_action(goldq_t *h, int replica, int64_t n, const float *x, int32_t *action)
{
    if (!h || !x || !action) return fail(h, GOLDQ_EINVAL, "greedy_action: bad argument");
    CK(h, cudaSetDevice(h->dev));
    int rc = predict_dev(h, GOLDQ_NET_ONLINE, replica, n, x);
    if (rc) return rc;
    launch_argmax_rows(h->pq, (int)n, h->dims.A, h->pact, h->st);
    h->launches += 1;
    CK(h, cudaMemcpyAsync(action, h->pact, (size_t)n * 4, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return check_launch(h, "greedy_action");
}

int goldq_epsilon_greedy_action(goldq_t *h, int replica, int64_t n, const float *x, float epsilon, uint64_t seed,
                                int32_t *action)
{
    if (!h || !x || !action || !(epsilon >= 0.f && epsilon <= 1.f))
        return fail(h, GOLDQ_EINVAL, "epsilon_greedy_action: bad argument");
    CK(h, cudaSetDevice(h->dev));
    int rc = predict_dev(h, GOLDQ_NET_ONLINE, replica, n, x);
    if (rc) return rc;
    launch_eps_greedy_rows(h->pq, (int)n, h->dims.A, epsilon, seed, h->pact, h->st);
    h->launches += 1;
    CK(h, cudaMemcpyAsync(action, h->pact, (size_t)n * 4, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return check_launch(h, "epsilon_greedy_action");
}

int goldq_explore_plan(int64_t n, int32_t n_actions, float epsilon, uint64_t seed, uint8_t *explore, int32_t *random_action)
{
    if (n < 0 || n_actions < 1 || !explore || !random_action) return GOLDQ_EINVAL;
    for (int64_t i = 0; i < n; i++) {
        int ra;
        explore[i] = explore_draw((uint32_t)i, seed, epsilon, n_actions, &ra) ? 1 : 0;
        random_action[i] = ra;
    }
    return GOLDQ_OK;
}

int goldq_fit_batch(goldq_t *h, int64_t n, const float *x, const float *y)
{
    if (!h || !x || !y) return fail(h, GOLDQ_EINVAL, "fit_batch: bad argument");
    // goro compiles one graph for batch_size rows and one for a single row
    // (model.go:260-312); any other leading dimension is a shape error (io.go:138-153)
    if (n != h->B && n != 1)
        return fail(h, GOLDQ_EINVAL, "fit_batch: rows must equal batch_size (FitBatch) or 1 (Fit)");
    if (h->world > 1) return fail(h, GOLDQ_EUNSUPPORTED, "fit_batch is single-GPU");
    if (h->path == GOLDQ_PATH_CONV) return fail(h, GOLDQ_EUNSUPPORTED, "fit_batch on explicit labels is not implemented for the conv policy (Learn is)");
    CK(h, cudaSetDevice(h->dev));
    int rc = ensure_generic(h, h->B > 1 ? h->B : 1);
    if (rc) return rc;
    const Dims &d = h->dims;
    CK(h, cudaMemcpyAsync(h->bs, x, (size_t)n * d.D * 4, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaMemcpyAsync(h->ylab, y, (size_t)n * d.A * 4, cudaMemcpyHostToDevice, h->st));
    generic_forward(h, h->theta, h->bs, (int)n, h->h1, h->mk1, h->h2, h->mk2, h->q);
    StepScalars sc = step_scalars(h, 1);     // loss kind / delta; the divisor is this call's n * A
    sc.count = (float)(n * d.A);
    sc.inv_count = 1.0f / sc.count;
    int nlp = 0;
    launch_label_loss(h->q, h->ylab, (int)n, d.A, sc, h->dq, h->loss_part, &nlp, h->st);
    h->launches += 1;
    generic_backward(h, h->bs, (int)n);
    launch_reduce_slices(h->gslice, h->P, h->nsplit, h->P, h->loss_part, nlp, sc.count, h->flat,
                         h->st);
    h->iter += 1;
    generic_solver_step(h, nullptr);
    h->launches += 2;
    CK(h, cudaMemcpyAsync(h->loss, h->flat + h->P, sizeof(float), cudaMemcpyDefault, h->st));
    if (h->wide) {
        std::string err;
        int k = wide_after_update(h->wide, h->theta, h->st, err);
        if (k < 0) return fail(h, GOLDQ_ECUDA, err);
        h->launches += k;
    }
    return learned(h, check_launch(h, "fit_batch"));
}

int goldq_dp_unique_id(uint8_t id[128])
{
    std::string err;
    if (!g_nccl.load(err)) return fail(nullptr, GOLDQ_ENCCL, err);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    ncclResult_t r = g_nccl.GetUniqueId(&uid);
    if (r != ncclSuccess) return fail(nullptr, GOLDQ_ENCCL, "ncclGetUniqueId failed");
    memcpy(id, &uid, 128);
    return GOLDQ_OK;
}

int goldq_dp_init(goldq_t *h, const uint8_t id[128], int rank, int world)
{
    if (!h || !id || world < 1 || rank < 0 || rank >= world)
        return fail(h, GOLDQ_EINVAL, "dp_init: bad argument");
    if (h->R > 1) return fail(h, GOLDQ_EUNSUPPORTED, "dp_init: replicas are independent; no group needed");
    if (h->Bglobal != h->B * world)
        return fail(h, GOLDQ_EINVAL, "dp_init: global_batch must equal batch_size * world");
    if (world == 1) { h->world = 1; h->rank = 0; return GOLDQ_OK; }
    std::string err;
    if (!g_nccl.load(err)) return fail(h, GOLDQ_ENCCL, err);
    CK(h, cudaSetDevice(h->dev));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, uid, rank);
    if (r != ncclSuccess)
        return fail(h, GOLDQ_ENCCL, std::string("ncclCommInitRank: ") +
                                        (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    h->rank = rank;
    h->world = world;
    drop_graphs(h);
    // wide path: the gradient all-reduce and the update run as one kernel over NVLink peer memory
    // (GOLDQ_NO_P2P=1 keeps ncclAllReduce + a separate update kernel)
    if (h->path == GOLDQ_PATH_WIDE && getenv("GOLDQ_NO_P2P") == nullptr) {
        std::string why;
        int rc = p2p_setup(h, &why);
        if (rc) return rc;
        if (!h->p2p.on) h->err = "peer-memory all-reduce unavailable, using NCCL: " + why;
    }
    return GOLDQ_OK;
}

int goldq_dp_mode(const goldq_t *h)
{
    if (!h) return GOLDQ_EINVAL;
    return h->world <= 1 ? 0 : (h->p2p.on ? 2 : 1);
}

int goldq_debug_enable(goldq_t *h, int on)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    if (on && !h->dbg_q) {
        size_t RB = (size_t)h->R * h->B;
        CK(h, dalloc(&h->dbg_q, RB * h->dims.A));
        CK(h, dalloc(&h->dbg_qt, RB * h->dims.A));
        CK(h, dalloc(&h->dbg_td, RB));
        CK(h, dalloc(&h->dbg_idx, RB));
    }
    h->debug = on != 0;
    if (h->wide) wide_set_debug(h->wide, h->debug);
    return GOLDQ_OK;
}

int goldq_debug_fetch(goldq_t *h, int what, void *out, int64_t out_bytes)
{
    if (!h || !out) return fail(h, GOLDQ_EINVAL, "debug_fetch: bad argument");
    CK(h, cudaSetDevice(h->dev));
    const Dims &d = h->dims;
    size_t RB = (size_t)h->R * h->B;
    const void *src = nullptr;
    size_t bytes = 0;
    bool narrow = h->path == GOLDQ_PATH_NARROW;
    if (narrow && !h->debug && what != GOLDQ_DBG_GRADS)
        return fail(h, GOLDQ_EINVAL, "debug_fetch: call goldq_debug_enable(h,1) before the step");
    const float *wq = nullptr, *wqt = nullptr, *wtd = nullptr;
    if (h->wide) wide_debug_ptrs(h->wide, &wq, &wqt, &wtd);
    switch (what) {
    case GOLDQ_DBG_Q_ONLINE: src = narrow ? h->dbg_q : (h->wide ? wq : h->q); bytes = RB * d.A * 4; break;
    case GOLDQ_DBG_Q_TARGET: src = narrow ? h->dbg_qt : (h->wide ? wqt : h->qt); bytes = RB * d.A * 4; break;
    case GOLDQ_DBG_TD: src = narrow ? h->dbg_td : (h->wide ? wtd : h->td); bytes = RB * 4; break;
    case GOLDQ_DBG_GRADS: {
        // flat is [R][P+1]; return the P gradient entries of each replica
        if ((size_t)out_bytes < (size_t)h->R * h->P * 4) return fail(h, GOLDQ_EINVAL, "debug_fetch: buffer too small");
        CK(h, cudaMemcpy2DAsync(out, (size_t)h->P * 4, h->flat, (size_t)(h->P + 1) * 4, (size_t)h->P * 4,
                                h->R, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        return GOLDQ_OK;
    }
    case GOLDQ_DBG_INDICES: src = narrow ? h->dbg_idx : h->d_idx; bytes = RB * 4; break;
    case GOLDQ_DBG_BATCH_S: src = h->bs; bytes = RB * d.D * 4; break;
    case GOLDQ_DBG_BATCH_A: src = h->ba; bytes = RB * 4; break;
    case GOLDQ_DBG_FF_STAMPS:
        if (h->wide) src = wide_ff_stamps(h->wide, &bytes);
        break;
    case GOLDQ_DBG_TRACE:
        if (h->wide) src = wide_trace(h->wide, &bytes);
        break;
    default: return fail(h, GOLDQ_EINVAL, "debug_fetch: unknown selector");
    }
    if (!src) return fail(h, GOLDQ_EINVAL, "debug_fetch: not available on this path / before any step");
    if ((size_t)out_bytes < bytes) return fail(h, GOLDQ_EINVAL, "debug_fetch: buffer too small");
    CK(h, cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, h->st));
    if (what == GOLDQ_DBG_TRACE) wide_trace_reset(h->wide, h->st);      // the next fetch starts a new log
    CK(h, cudaStreamSynchronize(h->st));
    if (what == GOLDQ_DBG_Q_TARGET && !narrow) {
        // rows of terminal events are never evaluated by the reference (agent.go:138); report 0
        std::vector<uint8_t> dn(RB);
        const uint8_t *dsrc = h->wide ? wide_done_ptr(h->wide) : h->bdone;
        CK(h, cudaMemcpy(dn.data(), dsrc, RB, cudaMemcpyDeviceToHost));
        float *o = (float *)out;
        for (size_t b = 0; b < RB; b++)
            if (dn[b])
                for (int j = 0; j < d.A; j++) o[b * d.A + j] = 0.f;
    }
    return GOLDQ_OK;
}

/* Diagnostic (bench.py roofline): average launch duration of the dominant tensor-core kernel of the wide
 * step -- the fused forward -- over `reps` back-to-back launches on the handle's stream, CUDA events around the
 * sequence; inputs are the batch of the last learn step (run one first). */
int goldq_time_forward(goldq_t *h, int32_t reps, float *us_per_launch)
{
    if (!h || !us_per_launch || reps < 1) return fail(h, GOLDQ_EINVAL, "time_forward: bad argument");
    if (!h->wide) return fail(h, GOLDQ_EUNSUPPORTED, "time_forward: wide (tcgen05) path only");
    CK(h, cudaSetDevice(h->dev));
    std::string err;
    int n = wide_time_forward(h->wide, h->theta, h->theta_t, h->st, reps, us_per_launch, err);
    if (n < 0) return fail(h, GOLDQ_EUNSUPPORTED, err);
    h->launches += n;
    return GOLDQ_OK;
}

__global__ void empty_kernel() {}

/* Diagnostic: the launch-latency floor the latency-bound configurations are compared with.  One
 * empty kernel per node of a 16-node CUDA graph (the shape of a goldq_learn_many chunk of the
 * single-launch narrow path), replayed `reps` times on the handle's stream; *us_per_launch = device
 * time per node between CUDA events. */
int goldq_launch_floor(goldq_t *h, int32_t reps, float *us_per_launch)
{
    if (!h || !us_per_launch || reps < 1) return fail(h, GOLDQ_EINVAL, "launch_floor: bad argument");
    CK(h, cudaSetDevice(h->dev));
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CK(h, cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 16; i++) empty_kernel<<<1, 32, 0, h->st>>>();
    cudaError_t e = cudaStreamEndCapture(h->st, &graph);
    if (e != cudaSuccess) return fail(h, GOLDQ_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(h, GOLDQ_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    for (int i = 0; i < 3; i++) cudaGraphLaunch(exec, h->st);
    CK(h, cudaEventRecord(h->ev0, h->st));
    for (int i = 0; i < reps; i++) cudaGraphLaunch(exec, h->st);
    CK(h, cudaEventRecord(h->ev1, h->st));
    CK(h, cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    cudaGraphExecDestroy(exec);
    *us_per_launch = ms * 1e3f / (16.f * (float)reps);
    return GOLDQ_OK;
}

int goldq_timer_start(goldq_t *h)
{
    if (!h) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    CK(h, cudaEventRecord(h->ev0, h->st));
    return GOLDQ_OK;
}

int goldq_timer_stop(goldq_t *h, float *ms)
{
    if (!h || !ms) return GOLDQ_EINVAL;
    CK(h, cudaSetDevice(h->dev));
    CK(h, cudaEventRecord(h->ev1, h->st));
    CK(h, cudaEventSynchronize(h->ev1));
    CK(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return GOLDQ_OK;
}

int goldq_selftest_umma(int mode, int M, int N, int K, int nsplit, const uint16_t *A, const uint16_t *B, float *C)
{
    std::string err;
    int rc = wide_selftest_gemm(mode, M, N, K, nsplit, A, B, C, err);
    if (rc) return fail(nullptr, GOLDQ_ECUDA, err);
    return GOLDQ_OK;
}

/* Host-side mirror of the device sampler: the indices goldq_learn(h, NULL, seed)
 * draws for replica `replica` when the replay holds `len` events. */
int goldq_sampler_indices(int64_t len, int32_t batch, uint64_t seed, int32_t replica, int32_t *out)
{
    if (len < 1 || len > 0x7fffffffLL || batch < 0 || !out) return GOLDQ_EINVAL;
    for (int b = 0; b < batch; b++)
        out[b] = (int32_t)perm_index((uint32_t)b, (uint32_t)len,
                                     seed + 0x9E3779B97F4A7C15ull * (uint64_t)replica);
    return GOLDQ_OK;
}

}  // extern "C"
