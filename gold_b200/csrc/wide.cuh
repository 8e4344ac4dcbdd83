// wide.cuh — interface of the bf16 tcgen05 path (kernels_wide.cu).
#pragma once
#include <string>

#include "common.cuh"

namespace gq {

struct WideState;

// Data-parallel exchange over NVLink peer memory (one process per GPU, buffers opened through
// CUDA IPC).  Every float travels as one 8-byte word {value, epoch}: an 8-byte store is atomic, so
// the epoch tag IS the arrival flag -- no separate flag words, no fences, no pull round trips.
// All of it happens inside the update kernel (wide_update_kernel, dp_exchange4), per element e:
//   push   rank s writes its locally reduced gradient straight into the receive window of the rank
//          that owns e:                                        rs[owner(e)][parity][s][e - lo(owner)]
//   reduce the owner polls the `world` words, adds them in rank order (so the sum is bit-identical
//          everywhere) and writes {sum, epoch} into EVERY rank's                      ag[parity][e]
//   apply  every rank polls its own ag word and runs Adam + the bf16 shadow refresh
// A rank cannot be more than one step ahead of a peer (it needs every owner's sums of step e+1,
// which an owner produces only after every rank has pushed e+1, i.e. finished step e), so a word of
// epoch e is never overwritten before it has been consumed; tests/test_dp_protocol.py checks this
// over random interleavings.  The windows additionally alternate by epoch parity (margin only).
constexpr int P2P_MAX_WORLD = 8;
struct PeerSet {
    int world, rank;
    int chunk_f;                           // floats owned per rank (multiple of 1024: block-aligned groups)
    unsigned long long *rs[P2P_MAX_WORLD]; // every rank's reduce-scatter window [2][world][chunk_f] words
    unsigned long long *ag[P2P_MAX_WORLD]; // every rank's all-gather window     [2][world * chunk_f] words
    int *err;                              // mapped host int: set to 1 when a word never arrives (5 s)
};

struct WideStep {
    Replay rp;
    const int32_t *idx;    // [B] device, or nullptr: draw on the device from `seed`
    uint64_t seed;
    const StepDev *step;   // non-null: seed / c1 / c2 from device memory (graph replay)
    int32_t *idx_out;      // where device-drawn indices are recorded (nullable)
    int B;
    const float *theta;    // fp32 master parameters (biases are read from here)
    const float *theta_t;
    float *flat;           // [P+1] out: reduced gradient + loss
    float *loss_out;       // the handle's loss scalar (written by the fused update)
    StepScalars sc;
    int fuse_update;       // 1: Adam + bf16 shadow refresh in the reduction kernel
    float *theta_rw, *m, *v;
    AdamConsts adam;
    // batch double buffering inside a captured multi-step graph: `set` = buffer set this step
    // reads; have = 1: it was already gathered by the previous step's side branch;
    // next_step != nullptr: gather the next step's batch (seed from *next_step) into the other set
    int set, have;
    const StepDev *next_step;
    const PeerSet *push;   // non-null (data-parallel group over peer memory, fuse_update = 1): the update
    unsigned int epoch;    // kernel exchanges the gradient with the peers before applying it
    // profiling (goldq_step_breakdown): record an event after every launch, single stream
    cudaEvent_t *prof_evs;
    int *prof_n;
};

WideState *wide_create(const Dims &d, int B, int sm_count, std::string &err);
void wide_destroy(WideState *w);
// each returns the number of kernels launched, or -1 with err set
int wide_set_params(WideState *w, int net, const float *theta_dev, cudaStream_t st, std::string &err);
int wide_after_update(WideState *w, const float *theta_dev, cudaStream_t st, std::string &err);
int wide_sync_target(WideState *w, cudaStream_t st, std::string &err);
int wide_apply_update(WideState *w, const float *flat, float *theta, float *m, float *v, AdamConsts ac,
                      const StepDev *step, cudaStream_t st, std::string &err);
int wide_step(WideState *w, const WideStep &s, cudaStream_t st, std::string &err);

int wide_predict(WideState *w, int net, const float *x, long long n, const float *theta, const float *theta_t, float *q,
                 cudaStream_t st, std::string &err);
int wide_time_forward(WideState *w, const float *theta, const float *theta_t, cudaStream_t st, int reps, float *us,
                      std::string &err);
int wide_p2p_chunk_floats(const WideState *w, int world);
bool wide_p2p_resident(const WideState *w, int *grid, int *resident);
std::string wide_step_kernel_names(WideState *w);
void wide_set_debug(WideState *w, bool on);
const unsigned long long *wide_trace(WideState *w, size_t *bytes);   // nullptr unless GOLDQ_TRACE was set at create time
void wide_trace_reset(WideState *w, cudaStream_t st);
const long long *wide_ff_stamps(WideState *w, size_t *bytes);   // nullptr unless GOLDQ_FF_STAMPS was set at create time
void wide_debug_ptrs(WideState *w, const float **q, const float **qt, const float **td);
const uint8_t *wide_done_ptr(WideState *w);
int wide_selftest_gemm(int mode, int M, int N, int K, int nsplit, const uint16_t *A, const uint16_t *B, float *C,
                       std::string &err);

}  // namespace gq
