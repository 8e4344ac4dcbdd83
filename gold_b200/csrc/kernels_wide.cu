// kernels_wide.cu — bf16 tcgen05 path (placeholder until the GEMM tiles land).
#include "wide.cuh"

namespace gq {
struct WideState { int unused; };
WideState *wide_create(const Dims &, int, int, std::string &err) { err = "tcgen05 path not built yet"; return nullptr; }
void wide_destroy(WideState *w) { delete w; }
int wide_set_params(WideState *, int, const float *, cudaStream_t, std::string &) { return 0; }
int wide_after_update(WideState *, const float *, cudaStream_t, std::string &) { return 0; }
int wide_sync_target(WideState *, cudaStream_t, std::string &) { return 0; }
int wide_forward_backward(WideState *, const WideStep &, cudaStream_t, std::string &err) { err = "not built"; return -1; }
void wide_set_debug(WideState *, bool) {}
void wide_debug_ptrs(WideState *, const float **q, const float **qt, const float **td) { *q = *qt = *td = nullptr; }
const uint8_t *wide_done_ptr(WideState *) { return nullptr; }
}  // namespace gq
