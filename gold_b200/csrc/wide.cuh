// wide.cuh — interface of the bf16 tcgen05 path (kernels_wide.cu).
#pragma once
#include <string>

#include "common.cuh"

namespace gq {

struct WideState;

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

// Data-parallel tail over NVLink peer memory (one process per GPU, buffers opened through CUDA
// IPC): every rank publishes its local gradient, then sums all ranks' gradients straight out of
// their memory in rank order and applies Adam + the bf16 shadow refresh, in ONE kernel.
constexpr int P2P_MAX_WORLD = 8;
struct PeerSet {
    int world, rank;
    const float *grad[P2P_MAX_WORLD];      // this step's gradient buffer of every rank ([P+1], loss last)
    unsigned int *flags[P2P_MAX_WORLD];    // every rank's flag array; slot r is written by rank r
    unsigned int *my_flags;                // = flags[rank], polled locally
    int *err;                              // mapped host int: set to 1 when a peer never showed up
    // two-shot variant (world > 2): rank r sums block-range r of the gradient and stores the sums
    // into every rank's `sum` buffer; flags2 says "rank r's range has landed everywhere"
    int two_shot;
    float *sum[P2P_MAX_WORLD];
    unsigned int *flags2[P2P_MAX_WORLD];
    unsigned int *my_flags2;
    unsigned int *ticket;                  // local counter: blocks of my range that have delivered
};
int wide_p2p_apply(WideState *w, const PeerSet &ps, unsigned int epoch, const StepDev *step, float *flat_out,
                   float *theta, float *m, float *v, AdamConsts ac, float *loss_out, cudaStream_t st,
                   std::string &err);
std::string wide_step_kernel_names(WideState *w);
void wide_set_debug(WideState *w, bool on);
void wide_debug_ptrs(WideState *w, const float **q, const float **qt, const float **td);
const uint8_t *wide_done_ptr(WideState *w);
int wide_selftest_gemm(int mode, int M, int N, int K, int nsplit, const uint16_t *A, const uint16_t *B, float *C,
                       std::string &err);

}  // namespace gq
