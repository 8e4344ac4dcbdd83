// kernels_wide.cu — bf16 tcgen05 path of the learn step for wide MLPs
// (BASELINE configs[2]/[3]: 128->512->512->18, batch 8192 per GPU).
//
// The layers here really are dense contractions, so they run on the 5th-gen tensor
// cores: TMA (cp.async.bulk.tensor, 128B swizzle) stages bf16 tiles in shared memory,
// one elected thread issues tcgen05.mma (M=128, K=16, fp32 accumulators in TMEM),
// tcgen05.commit/mbarrier pipelines producer <-> MMA <-> epilogue, and four epilogue
// warps read TMEM with tcgen05.ld and fuse bias+ReLU / ReLU-mask / split-K partial
// stores.  fp32 master weights, Adam in fp32 (kernels_generic.cu K5); a bf16 shadow
// of the weights is refreshed after every update.
//
// GEMM shapes per step (B rows):
//   forward (both nets)  H1 = relu(X  W1 + b1)      KM mode  M=B   N=H1  K=D
//                        H2 = relu(H1 W2 + b2)      KM       M=B   N=H2  K=H1
//                        Q  =      H2 W3 + b3       KM, BN=64 M=B  N=A   K=H2 (W3 padded to 64 columns)
//   backward             dW2 = H1^T dZ2             MM mode  M=H1  N=H2  K=B (split-K)
//                        dZ1 = (dZ2 W2^T) * mask1   KK       M=B   N=H1  K=H2
//                        dW1 = X^T dZ1              MM       M=D   N=H1  K=B (split-K)
//                        dW3 = H2^T dQ              MK mode  M=H2  N=32  K=B (split-K), dQ^T [32][B]
//   dQ has one non-zero per row (the taken action), so dZ2 = dq * W3[:,a] * mask2 is an
//   elementwise kernel (wide_td_dz2_kernel, fused with the TD label), not a GEMM.
// KK: both operands K-major (A [M][K], B [N][K] row-major).  MM: both MN-major
// (A [K][M], B [K][N] row-major) — the activations are used as stored, no transposes.
// MK: A MN-major, B K-major.  KM: A K-major, B MN-major (forward: W [in][out] as stored).
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <string>
#include <vector>

#include "sampler.cuh"
#include "umma.cuh"
#include "wide.cuh"

namespace gq {

using namespace umma;

namespace {

constexpr int BM = 128;   // UMMA M (rows of the output tile; TMEM lanes)
constexpr int BK = 64;    // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UK = 16;    // K of one tcgen05.mma for 16-bit inputs
constexpr int DEFAULT_STAGES = 3;
constexpr int GEMM_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

enum { MODE_KK = 0, MODE_MM = 1, MODE_MK = 2, MODE_KM = 3 };
enum { EPI_BIAS_RELU_BF16 = 0, EPI_BIAS_F32 = 1, EPI_MASK_BF16 = 2, EPI_F32_SLICE = 3 };

// Loads of data that the PREVIOUS kernel of a programmatic-dependent-launch chain produced (activations, Q, masks, the
// parameters the update kernel has just written).  They must not be `__ldg` / `const __restrict__` loads: those compile to
// ld.global.nc, which the compiler treats as invariant and may schedule ABOVE griddepcontrol.wait, i.e. before the producer
// has finished -- seen in SASS (eight LDG.E.CONSTANT in front of ACQBULK in the TD kernel) and in wrong TD targets / losses.
// ld.global.cg is an ordinary load for the compiler (it stays behind the wait's memory clobber) and skips the L1 these
// single-use values do not need.
template <typename T>
__device__ __forceinline__ T ld_dep(const T *p) { return __ldcg(p); }

// ---- block trace (diagnostic, GOLDQ_TRACE=1 at create; tools/trace_step.py) ---------------------------------
// Thread 0 of every block of the step's kernels appends {kernel id, block, globaltimer at entry and exit} to a log:
// the poor man's timeline (no nsys in this image) of a graph-replayed step with its parallel branches.
// c_trace == nullptr (the default): one constant-bank load and a predicated-off branch per block.
constexpr int TRACE_MARKS = 8;
struct TraceRec { unsigned long long t0, t1; int id, block; unsigned long long pad_; unsigned long long mark[TRACE_MARKS]; };
constexpr int TRACE_CAP = 1 << 17;
#ifndef GOLDQ_TRACE_BUILD
struct BlockTrace {      // the product build: nothing (`make trace` builds libgoldq_trace.so with the real one)
    __device__ __forceinline__ explicit BlockTrace(int, unsigned long long * = nullptr) {}
};
#define TRACE_MARKS_DECL
#define TRACE_MARKS_PTR nullptr
#define TRACE_MARK(i) do { } while (0)
#else
__constant__ unsigned long long *c_trace = nullptr;      // [0] = record counter, records start at 32 bytes
__device__ __forceinline__ unsigned long long trace_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// optional in-kernel marks: any thread stores globaltimer into the block's shared array; thread 0 copies them out
#define TRACE_MARKS_DECL __shared__ unsigned long long tr_marks_[TRACE_MARKS];
#define TRACE_MARKS_PTR tr_marks_
#define TRACE_MARK(i) do { if (c_trace != nullptr) tr_marks_[i] = trace_now(); } while (0)
struct BlockTrace {
    unsigned long long t0;
    unsigned long long *marks;
    int id;
    __device__ __forceinline__ explicit BlockTrace(int id_, unsigned long long *marks_ = nullptr) : t0(0), marks(marks_), id(id_)
    {
        if (c_trace != nullptr && threadIdx.x == 0) {
            t0 = trace_now();
            if (marks)
                for (int i = 0; i < TRACE_MARKS; i++) marks[i] = 0;
        }
    }
    __device__ __forceinline__ ~BlockTrace()
    {
        if (t0) {
            const unsigned long long slot = atomicAdd(c_trace, 1ull);
            if (slot < (unsigned long long)TRACE_CAP) {
                TraceRec *r = reinterpret_cast<TraceRec *>(c_trace + 4) + slot;
                r->t0 = t0; r->t1 = trace_now(); r->id = id;
                r->block = (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
                r->pad_ = 0;
                for (int i = 0; i < TRACE_MARKS; i++) r->mark[i] = marks ? marks[i] : 0;
            }
        }
    }
};
#endif
enum { TR_GATHER = 1, TR_FORWARD, TR_TD, TR_DW3, TR_DW2, TR_DX, TR_COLSUM, TR_DW1, TR_UPDATE, TR_OTHER };

struct GemmArgs {
    int trace_id;         // TR_* (diagnostic)
    int M, N, K;          // logical problem (K = full reduction length)
    int k_per_split;      // multiple of BK
    int nsplit;           // blockIdx.z = problem * nsplit + split
    // epilogue operands (problem 0 / problem 1: the online and the target network share one launch)
    __nv_bfloat16 *out_bf16, *out_bf16_1;   // [M][ldo]
    float *out_f32, *out_f32_1;             // [M][ldo] (+ split * slice_stride for EPI_F32_SLICE)
    const float *bias, *bias_1;             // [N]
    const __nv_bfloat16 *mask_src;          // [M][ldo]: sign bit set <=> ReLU mask 0
    int ldo;                                // row pitch of the output in elements
    int n_store;                            // columns actually stored (<= N tile coverage)
    long long slice_stride;
    int late_trigger;                       // release the dependent grid from the epilogue instead of the prologue
};

// CM > 1: clusters of CM CTAs along M (blockIdx.y) share the B tile — each CTA loads 1/CM of
// it and multicasts the slice into every CTA of the cluster (tmB* then describe the slice box).
template <int MODE, int BN, int EPI, int CM, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1,
                 const GemmArgs g)
{
    TRACE_MARKS_DECL
    BlockTrace trace_(g.trace_id, TRACE_MARKS_PTR);
    // bf16 outputs leave through shared memory + TMA store (full 128-byte lines; rows and
    // columns outside the tensor are clipped by the hardware); tmC* = box {64 cols, 128 rows}
    constexpr bool TMA_OUT = (EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_MASK_BF16);
    static_assert(!TMA_OUT || BN % 64 == 0, "bf16 tiles are stored in 64-column boxes");
    constexpr bool A_MN = (MODE == MODE_MM || MODE == MODE_MK), B_MN = (MODE == MODE_MM || MODE == MODE_KM);
    static_assert(BN == 256 || BN == 128 || BN == 64 || BN == 32, "BN");
    static_assert(!(EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_MASK_BF16) || (BN / 64) * 16384 <= STAGES * (BM * BK * 2 + BN * BK * 2),
                  "the bf16 staging tile must fit in the operand ring");
    static_assert(!B_MN || BN % 64 == 0, "MN-major B tiles are 64-wide swizzle atoms");
    static_assert(CM == 1 || CM == 2 || CM == 4 || CM == 8, "cluster size");
    static_assert(CM == 1 || (B_MN ? (64 / CM) % 8 == 0 : (BN / CM) % 8 == 0), "slices are whole swizzle atoms");
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CM) - 1u);
    constexpr uint32_t A_BYTES = BM * BK * 2;
    constexpr uint32_t B_BYTES = BN * BK * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr uint32_t IDESC = make_idesc_bf16(BM, BN, A_MN, B_MN);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_bias[BN];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int prob = blockIdx.z / g.nsplit, split = blockIdx.z - prob * g.nsplit;
    const CUtensorMap *tmA = prob ? &tmA1 : &tmA0, *tmB = prob ? &tmB1 : &tmB0;
    const int k_begin = split * g.k_per_split;
    int k_end = k_begin + g.k_per_split;
    if (k_end > g.K) k_end = g.K;
    const int nkb = (k_end - k_begin + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(tmA);
        prefetch_tmap(tmB);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CM);     // one tcgen05.commit per CTA of the cluster
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
    if (!g.late_trigger) pdl_launch_dependents();     // the next kernel may begin its own prologue
    tc_fence_before();
    if (CM > 1) cluster_sync();  // peers' barriers are initialised before anything is multicast
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;
    pdl_wait();                  // operands / bias / mask come from the previous kernels
    if (threadIdx.x == 0) TRACE_MARK(0);

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t *sA = smem + s * STAGE_BYTES, *sB = sA + A_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                const int k = k_begin + kb * BK;
                // K-major: box {64 k, rows}.  MN-major: box {64 mn, 64 k-rows}, one 8 KB atom
                // column per 64 mn.
                if (!A_MN) tma_load_2d(sA, tmA, k, m0, &full_bar[s]);
                else
                    for (int h = 0; h < BM / 64; h++) tma_load_2d(sA + h * 8192, tmA, m0 + 64 * h, k, &full_bar[s]);
                if (CM == 1) {
                    if (!B_MN) tma_load_2d(sB, tmB, k, n0, &full_bar[s]);
                    else
                        for (int h = 0; h < BN / 64; h++) tma_load_2d(sB + h * 8192, tmB, n0 + 64 * h, k, &full_bar[s]);
                } else {
                    const int r = (int)cluster_ctarank();
                    if (!B_MN) {   // slice = rows [r*BN/CM, ...) of the [BN][64 k] tile
                        tma_load_2d_mc(sB + r * (BN / CM) * 128, tmB, k, n0 + r * (BN / CM), &full_bar[s], MC_MASK);
                    } else {       // slice = k-rows [r*64/CM, ...) of every 64-wide n group
                        for (int h = 0; h < BN / 64; h++)
                            tma_load_2d_mc(sB + h * 8192 + r * (64 / CM) * 128, tmB, n0 + 64 * h, k + r * (64 / CM),
                                           &full_bar[s], MC_MASK);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (kb == 0 && lane == 0) TRACE_MARK(1);
            if (elect_one()) {
                const uint32_t aaddr = smem_u32(smem + s * STAGE_BYTES), baddr = aaddr + A_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / UK; kk++) {
                    // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); the 16
                    // k-elements of this MMA start 32 B further inside the swizzle row.
                    // MN-major SW128: k-rows of 128 B (64 mn), 8-k-row groups 1024 B apart
                    // (SBO), 64-mn groups 8192 B apart (LBO); 16 k-rows = 2048 B per MMA.
                    const uint64_t ad = A_MN ? make_smem_desc(aaddr + kk * 2048, 8192, 1024)
                                             : make_smem_desc(aaddr + kk * 32, 16, 1024);
                    const uint64_t bd = B_MN ? make_smem_desc(baddr + kk * 2048, 8192, 1024)
                                             : make_smem_desc(baddr + kk * 32, 16, 1024);
                    mma_bf16(tmem_acc, ad, bd, IDESC, (kb | kk) != 0 ? 1u : 0u);
                }
                // smem slot free once these MMAs retire (in every CTA that multicasts into it)
                if (CM == 1) mma_commit(&empty_bar[s]);
                else mma_commit_mc(&empty_bar[s], MC_MASK);
                if (kb == nkb - 1) mma_commit(&tmem_full_bar);
            }
            __syncwarp();
        }
        if (lane == 0) TRACE_MARK(2);
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int row = m0 + q * 32 + lane;
        float *const out_f32 = prob ? g.out_f32_1 : g.out_f32;
        if (EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_BIAS_F32) {
            // stage this tile's bias once (one coalesced load), while the mainloop runs
            const float *const bias = prob ? g.bias_1 : g.bias;
            for (int t = threadIdx.x - 64; t < BN; t += 128)
                s_bias[t] = (n0 + t < g.N && n0 + t < g.n_store) ? ld_dep(bias + n0 + t) : 0.f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (nkb > 0) {
            mbar_wait(&tmem_full_bar, 0);
            tc_fence_after();
        }
        // (late trigger: a dependent released from the prologue is resident for this kernel's whole main loop and holds
        // registers next to the GEMM CTAs that have not started yet)
        if (g.late_trigger && threadIdx.x == 64) pdl_launch_dependents();
        if (threadIdx.x == 64) TRACE_MARK(3);
        // fp32 split-K slices of full tiles leave through shared memory: the tile is staged in the (idle) operand ring
        // with the 16-byte chunks of a row XOR-swizzled by the row, then every warp writes whole 512-byte rows.  (Stored
        // straight from the accumulator registers, a warp's store instruction touched 32 different rows, 16 bytes each.)
        constexpr bool SLICE_STAGED = (EPI == EPI_F32_SLICE) && BN == 128 && STAGES * STAGE_BYTES >= BM * BN * 4;
        const bool slice_staged = SLICE_STAGED && n0 + BN <= g.n_store && m0 + BM <= g.M && (g.ldo & 3) == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            if (nkb > 0) {
                tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) r[j] = 0u;
            }
            const int col0 = n0 + c0;
            if (row < g.M) {
                if (EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_MASK_BF16) {
                    uint32_t packed[16];
                    uint32_t msk[16];
                    if (EPI == EPI_MASK_BF16) {
                        const uint4 *mp = reinterpret_cast<const uint4 *>(g.mask_src + (size_t)row * g.ldo + col0);
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            uint4 t = (col0 + 8 * v < g.n_store) ? ld_dep(mp + v) : make_uint4(0, 0, 0, 0);
                            msk[4 * v] = t.x; msk[4 * v + 1] = t.y; msk[4 * v + 2] = t.z; msk[4 * v + 3] = t.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[j + 1]);
                        if (EPI == EPI_BIAS_RELU_BF16) {
                            v0 += s_bias[c0 + j];
                            v1 += s_bias[c0 + j + 1];
                            // Rectify keeps the sign of a*0: -0.0 marks a masked unit (nn.go:162-189)
                            v0 = (v0 >= 0.f) ? v0 : -0.0f;
                            v1 = (v1 >= 0.f) ? v1 : -0.0f;
                        } else {
                            const uint32_t mm = msk[j >> 1];
                            v0 = (mm & 0x8000u) ? 0.f : v0;
                            v1 = (mm & 0x80000000u) ? 0.f : v1;
                        }
                        __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
                        packed[j >> 1] = *reinterpret_cast<uint32_t *>(&p);
                    }
                    // staging tile = stage 0 of the ring (free: every MMA has retired), laid out as the
                    // 128B-swizzled boxes the store tensor map describes: [BN/64][128 rows][128 B]
                    const int rt = q * 32 + lane;
                    uint8_t *dst = smem + (c0 >> 6) * 16384 + rt * 128;
                    const int ch0 = (c0 & 63) >> 3;
#pragma unroll
                    for (int v = 0; v < 4; v++)
                        sts128(dst + (((ch0 + v) ^ (rt & 7)) << 4),
                                   make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]));
                } else if (EPI == EPI_BIAS_F32) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int c = col0 + j;
                        if (c < g.n_store) out_f32[(size_t)row * g.ldo + c] = __uint_as_float(r[j]) + s_bias[c0 + j];
                    }
                } else if (SLICE_STAGED && slice_staged) {
                    const int rt = q * 32 + lane;
                    uint8_t *dst = smem + rt * (BN * 4);
#pragma unroll
                    for (int v = 0; v < 8; v++)
                        sts128(dst + ((((c0 >> 2) + v) ^ (rt & 7)) << 4),
                                   make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]));
                } else {  // EPI_F32_SLICE
                    float *o = out_f32 + (size_t)split * g.slice_stride + (size_t)row * g.ldo + col0;
                    if (col0 + 32 <= g.n_store && (g.ldo & 3) == 0) {
#pragma unroll
                        for (int v = 0; v < 8; v++)
                            reinterpret_cast<float4 *>(o)[v] =
                                make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                            __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (col0 + j < g.n_store) o[j] = __uint_as_float(r[j]);
                    }
                }
            }
        }
        if (SLICE_STAGED && slice_staged) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // warp w of the four writes rows w, w + 4, ...: lane = 16-byte chunk of the 512-byte row
            float *o = out_f32 + (size_t)split * g.slice_stride + (size_t)m0 * g.ldo + n0;
            const int ew = warp - 2;
#pragma unroll 8
            for (int rr = ew; rr < BM; rr += 4) {
                const uint4 v = lds128(smem + rr * (BN * 4) + ((lane ^ (rr & 7)) << 4));
                *reinterpret_cast<uint4 *>(o + (size_t)rr * g.ldo + lane * 4) = v;
            }
        }
        if (threadIdx.x == 64) TRACE_MARK(4);
        if (TMA_OUT) {
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 64) {
                const CUtensorMap *tmC = prob ? &tmC1 : &tmC0;
                for (int b = 0; b < BN / 64; b++)
                    if (n0 + 64 * b < g.n_store) tma_store_2d(tmC, smem + b * 16384, n0 + 64 * b, m0);
                tma_store_commit_and_wait_read();
            }
        }
        tc_fence_before();
    }
    if (CM > 1) cluster_sync();   // nobody leaves while peers may still signal its barriers
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, TMEM_COLS);
    }
}

template <int MODE, int BN, int STAGES>
constexpr size_t gemm_smem_bytes()
{
    return (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024;
}

// ---------------------------------------------------------------------------
// Fused forward: Q = relu(relu(X W1 + b1) W2 + b2) W3 + b3 for a 128-row slab of ONE network in
// ONE CTA (blockIdx.y selects online / target).  H1 and H2 never leave the SM between layers:
// the epilogue writes them (bf16, 128B-swizzled K-major boxes) into shared memory where the
// next layer's tcgen05.mma reads them as its A operand; the online network's H1/H2 are also
// TMA-stored to global for the backward pass.  Weights stream through a 3-stage ring of
// [64 k][256 n] MN-major tiles (96 KB in flight; the X tile aliases the tail of the H tile).  Accumulators: 128 x 512 fp32 = all 512 TMEM columns, as two
// 256-column halves so that the epilogue of half 0 overlaps the MMAs of half 1.
//
// Warp roles: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..9 = epilogue (warp w reads
// TMEM lane quarter w % 4 and column group (w - 2) / 4 of every half).
// Shapes: H1 == H2 == 512, D in {64, 128}, A <= 32 (W3 padded to 64 columns).
// ---------------------------------------------------------------------------
constexpr int FF_H = 512;
constexpr int FF_THREADS = 320;
constexpr int FF_EPI_THREADS = 256;
constexpr int FF_WSTAGES = 3;
constexpr uint32_t FF_BOX = 16384;            // [128 rows][64 bf16] swizzled box
constexpr uint32_t FF_WSTAGE_BYTES = 32768;   // [4 n-groups][64 k-rows][128 B]

constexpr int FF_NSTAMPS = 96;
struct FFArgs {
    int B, D, A;
    const float *bias[2][3];     // [net][layer]
    float *q[2];                 // [net]: Q / Q' fp32 [B][A]
    __nv_bfloat16 *h1, *h2;      // (unused)
    int net0;                    // network of blockIdx.y == 0 (predict launches one network: grid.y = 1)
    int x_ready;                 // the batch tile may be loaded before griddepcontrol.wait (ff_forward2_kernel)
    int keep_h;                  // store the online network's H1 / H2 for the backward pass (learn step) or not (predict)
    int late_trigger;            // release the dependent grid from the epilogue instead of the prologue
    uint32_t *maskbits[2];       // [layer]: Rectify masks of the online network, one bit per unit, [B][FF_H / 32] words; word w
                                 // covers units 32 w .. 32 w + 31: bit m = unit 32 w + 2 m, bit 16 + m = unit 32 w + 2 m + 1
                                 // (even units in the low half, odd units in the high half).  The backward pass reads these instead of the sign bits of H.
    long long *stamps;           // diagnostic (GOLDQ_FF_STAMPS): [CTAs][FF_NSTAMPS] clock64 values, else nullptr
};

template <int KD>   // KD = D / 64 k-blocks of layer 1
__global__ void __launch_bounds__(FF_THREADS, 1)   // 168 registers: 3 warps per SM sub-partition
ff_forward_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                  const __grid_constant__ CUtensorMap tmW1_0, const __grid_constant__ CUtensorMap tmW1_1,
                  const __grid_constant__ CUtensorMap tmW2_0, const __grid_constant__ CUtensorMap tmW2_1,
                  const __grid_constant__ CUtensorMap tmW3_0, const __grid_constant__ CUtensorMap tmW3_1,
                  const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmH2,
                  const FFArgs g)
{
    BlockTrace trace_(TR_FORWARD);
    constexpr uint32_t X_BYTES = KD * FF_BOX;
    constexpr uint32_t IDESC_256 = make_idesc_bf16(BM, 256, false, true);
    constexpr uint32_t IDESC_64 = make_idesc_bf16(BM, 64, false, true);
    constexpr int KH = FF_H / 64;                       // 8 k-blocks of layers 2 and 3
    static_assert(KH % 4 == 0, "layer 3 loads four k-blocks per ring stage");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sH = smem;                                 // 8 boxes: H1, then H2
    // X aliases the LAST boxes of the H tile: they belong to accumulator half 1, whose epilogue
    // starts only after every layer-1 MMA (the readers of X) has retired
    uint8_t *sX = sH + (KH - KD) * FF_BOX;
    uint8_t *sW = sH + KH * FF_BOX;                     // weight ring
    __shared__ uint64_t full_w[FF_WSTAGES], empty_w[FF_WSTAGES], x_full, acc_full[2], acc_free[2], h_ready, h_lo;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int net = blockIdx.y + g.net0;
    const int m0 = blockIdx.x * BM;
    const CUtensorMap *tmX = net ? &tmX1 : &tmX0;
    const CUtensorMap *tmW1 = net ? &tmW1_1 : &tmW1_0;
    const CUtensorMap *tmW2 = net ? &tmW2_1 : &tmW2_0;
    const CUtensorMap *tmW3 = net ? &tmW3_1 : &tmW3_0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(tmX); prefetch_tmap(tmW1); prefetch_tmap(tmW2); prefetch_tmap(tmW3);
        for (int s = 0; s < FF_WSTAGES; s++) { mbar_init(&full_w[s], 1); mbar_init(&empty_w[s], 1); }
        mbar_init(&x_full, 1);
        mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
        mbar_init(&acc_free[0], FF_EPI_THREADS / 32); mbar_init(&acc_free[1], FF_EPI_THREADS / 32);
        mbar_init(&h_ready, 1);
        mbar_init(&h_lo, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, 512);
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer: X, then the weight tiles in consumption order =====
        if (elect_one()) {
            mbar_expect_tx(&x_full, X_BYTES);
            for (int kb = 0; kb < KD; kb++) tma_load_2d(sX + kb * FF_BOX, tmX, kb * 64, m0, &x_full);
            int it = 0;
            // one ring stage = 4 boxes of [64 k][64 n]: four n-groups of one k-block (layers 1, 2) or
            // four consecutive k-blocks of the 64-column W3 tile (layer 3)
            auto load_w = [&](const CUtensorMap *tm, int n_base, int k, int dn, int dk) {
                const int s = it % FF_WSTAGES;
                const uint32_t ph = (it / FF_WSTAGES) & 1;
                mbar_wait(&empty_w[s], ph ^ 1);
                mbar_expect_tx(&full_w[s], FF_WSTAGE_BYTES);
                for (int h = 0; h < 4; h++)
                    tma_load_2d(sW + s * FF_WSTAGE_BYTES + h * 8192, tm, n_base + dn * h, k + dk * h, &full_w[s]);
                it++;
            };
            for (int half = 0; half < 2; half++)
                for (int kb = 0; kb < KD; kb++) load_w(tmW1, half * 256, kb * 64, 64, 0);
            for (int half = 0; half < 2; half++)
                for (int kb = 0; kb < KH; kb++) load_w(tmW2, half * 256, kb * 64, 64, 0);
            // all of W3 goes out as soon as two stages free up, i.e. before the H2 store queues
            // 128 KB on the same TMA unit
            for (int kq = 0; kq < KH / 4; kq++) load_w(tmW3, 0, kq * 256, 0, 64);
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int it = 0;
        auto mma_block = [&](uint32_t a_box_addr, uint32_t d_tmem, bool first) {
            const int s = it % FF_WSTAGES;
            const uint32_t ph = (it / FF_WSTAGES) & 1;
            mbar_wait(&full_w[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t baddr = smem_u32(sW + s * FF_WSTAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / UK; kk++)
                    mma_bf16(d_tmem, make_smem_desc(a_box_addr + kk * 32, 16, 1024),
                             make_smem_desc(baddr + kk * 2048, 8192, 1024), IDESC_256, (!first || kk) ? 1u : 0u);
                mma_commit(&empty_w[s]);
            }
            __syncwarp();
            it++;
        };
        // layer 1: A = X
        mbar_wait(&x_full, 0);
        for (int half = 0; half < 2; half++) {
            for (int kb = 0; kb < KD; kb++) mma_block(smem_u32(sX + kb * FF_BOX), tmem + half * 256, kb == 0);
            if (elect_one()) mma_commit(&acc_full[half]);
            __syncwarp();
        }
        // layer 2: A = H1 (written by the epilogue).  Its first four k-blocks are the columns the
        // epilogue of accumulator half 0 produced: they start while half 1 is still being drained.
        mbar_wait(&h_lo, 0);
        mbar_wait(&acc_free[0], 0);
        tc_fence_after();
        for (int half = 0; half < 2; half++) {
            for (int kb = 0; kb < KH; kb++) {
                if (half == 0 && kb == KH / 2) {
                    mbar_wait(&h_ready, 0);
                    mbar_wait(&acc_free[1], 0);
                    tc_fence_after();
                }
                mma_block(smem_u32(sH + kb * FF_BOX), tmem + half * 256, kb == 0);
            }
            if (elect_one()) mma_commit(&acc_full[half]);
            __syncwarp();
        }
        // layer 3: A = H2, N = 64 (W3 padded); a ring stage holds four k-blocks.  The first stage needs
        // only the first half of H2, which the epilogue writes out before it drains half 1.
        mbar_wait(&h_lo, 1);
        mbar_wait(&acc_free[0], 1);
        tc_fence_after();
        for (int kq = 0; kq < KH / 4; kq++) {
            if (kq == 1) {
                mbar_wait(&h_ready, 1);
                mbar_wait(&acc_free[1], 1);
                tc_fence_after();
            }
            const int s = it % FF_WSTAGES;
            const uint32_t ph = (it / FF_WSTAGES) & 1;
            mbar_wait(&full_w[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t baddr = smem_u32(sW + s * FF_WSTAGE_BYTES);
#pragma unroll
                for (int h = 0; h < 4; h++)
#pragma unroll
                    for (int kk = 0; kk < BK / UK; kk++)
                        mma_bf16(tmem, make_smem_desc(smem_u32(sH + (kq * 4 + h) * FF_BOX) + kk * 32, 16, 1024),
                                 make_smem_desc(baddr + h * 8192 + kk * 2048, 8192, 1024), IDESC_64,
                                 (kq || h || kk) ? 1u : 0u);
                mma_commit(&empty_w[s]);
            }
            __syncwarp();
            it++;
        }
        if (elect_one()) mma_commit(&acc_full[0]);
        __syncwarp();
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3;                 // TMEM lane quarter
        const int cg = (warp - 2) >> 2;         // column group within a half: chunks cg*4 .. cg*4+3
        const int rt = q * 32 + lane;           // row within the slab
        const int et = threadIdx.x - 64;        // 0..255
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        // Bias: lane l keeps element l of each 32-column chunk this warp will process and the values
        // are broadcast by shuffle.  All loads are issued here, long before they are needed (loads
        // inside the chunk loop are predicated on nothing the compiler can hoist and cost a full L2
        // round trip per chunk).
        float bl[2][8];
#pragma unroll
        for (int l = 0; l < 2; l++)
#pragma unroll
            for (int i = 0; i < 8; i++) bl[l][i] = ld_dep(g.bias[net][l] + (i >> 2) * 256 + (cg * 4 + (i & 3)) * 32 + lane);
        const float b3 = (lane < g.A) ? ld_dep(g.bias[net][2] + lane) : 0.f;

        auto pack_chunk = [&](const uint32_t(&r)[32], float bme, uint32_t(&packed)[16]) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float v0 = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bme, j);
                float v1 = __uint_as_float(r[j + 1]) + __shfl_sync(0xffffffffu, bme, j + 1);
                v0 = (v0 >= 0.f) ? v0 : -0.0f;       // Rectify, sign bit = mask (nn.go:162-189)
                v1 = (v1 >= 0.f) ? v1 : -0.0f;
                __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
                packed[j >> 1] = *reinterpret_cast<uint32_t *>(&pk);
            }
        };
        auto store_chunk = [&](int c0, const uint32_t(&packed)[16]) {
            uint8_t *dst = sH + (c0 >> 6) * FF_BOX + rt * 128;
            const int ch0 = (c0 & 63) >> 3;
#pragma unroll
            for (int v = 0; v < 4; v++)
                sts128(dst + (((ch0 + v) ^ (rt & 7)) << 4),
                           make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]));
        };
        auto arrive = [&](uint64_t *bar) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        };
        auto col0 = [&](int half, int c) { return half * 256 + (cg * 4 + c) * 32; };

        // ---- layer 1 -> H1.  The TMEM load of chunk c+1 is in flight while chunk c is packed.
#pragma unroll
        for (int half = 0; half < 2; half++) {
            mbar_wait(&acc_full[half], 0);
            tc_fence_after();
            uint32_t r[2][32];
            tmem_ld_32x32(tq + (uint32_t)col0(half, 0), r[0]);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                tmem_ld_wait();
                if (c < 3) tmem_ld_32x32(tq + (uint32_t)col0(half, c + 1), r[(c + 1) & 1]);
                uint32_t packed[16];
                pack_chunk(r[c & 1], bl[0][half * 4 + c], packed);
                store_chunk(col0(half, c), packed);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive(&acc_free[half]);
            if (half == 0) {           // boxes 0..3 of H1 are complete: layer 2 may start on them
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (et == 0) arrive(&h_lo);
            }
        }
        fence_proxy_async_smem();      // H1 complete: visible to the async proxy (layer-2 MMA, TMA store)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
            if (net == 0 && g.keep_h) {            // the backward pass needs the online network's activations
                for (int b = 0; b < KH; b++) tma_store_2d(&tmH1, sH + b * FF_BOX, 64 * b, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            arrive(&h_ready);
        }

        // ---- layer 2 -> H2, which overwrites H1 in place.  Accumulator half 0 is drained into
        // registers while the MMAs of half 1 still read H1; it is written out once they retire.
        {
            uint32_t hold[4][16];
            uint32_t r[32];
            mbar_wait(&acc_full[0], 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; c++) {
                tmem_ld_32x32(tq + (uint32_t)col0(0, c), r);
                tmem_ld_wait();
                pack_chunk(r, bl[1][c], hold[c]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive(&acc_free[0]);
            // every layer-2 MMA has read H1 once acc_full[1] fires; the H1 TMA store must also have
            // finished reading the tile
            mbar_wait(&acc_full[1], 1);
            tc_fence_after();
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            tmem_ld_32x32(tq + (uint32_t)col0(1, 0), r);
#pragma unroll
            for (int c = 0; c < 4; c++) store_chunk(col0(0, c), hold[c]);
            // boxes 0..3 of H2 are complete: layer 3 may start on them
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) arrive(&h_lo);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (c) tmem_ld_32x32(tq + (uint32_t)col0(1, c), r);
                tmem_ld_wait();
                uint32_t packed[16];
                pack_chunk(r, bl[1][4 + c], packed);
                store_chunk(col0(1, c), packed);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive(&acc_free[1]);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
            if (net == 0 && g.keep_h) {
                for (int b = 0; b < KH; b++) tma_store_2d(&tmH2, sH + b * FF_BOX, 64 * b, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            arrive(&h_ready);
        }

        // ---- layer 3: Q = acc + b3 -> fp32 [B][A] (A <= 32: one chunk).  Column group 0 warps stage
        // the slab's rows in the (now idle) weight ring and copy them out coalesced: the rows of a
        // slab are contiguous in Q.
        if (cg == 0) {
            mbar_wait(&acc_full[0], 0);          // third completion of acc_full[0]: parity 0 again
            tc_fence_after();
            uint32_t r[32];
            tmem_ld_32x32(tq, r);
            tmem_ld_wait();
            tc_fence_before();
            float *sQ = reinterpret_cast<float *>(sW);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const float v = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, b3, j);
                if (j < g.A) sts32(sQ + rt * g.A + j, v);
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            const int rows = (g.B - m0 < BM) ? g.B - m0 : BM;
            float *dst = g.q[net] + (size_t)m0 * g.A;
            for (int i = et; i < rows * g.A; i += 128) dst[i] = lds32(sQ + i);
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int KD>
constexpr size_t ff_smem_bytes()
{
    return (size_t)8 * FF_BOX + FF_WSTAGES * FF_WSTAGE_BYTES + 1024;
}

// ---------------------------------------------------------------------------
// Fused forward on CTA PAIRS (cta_group::2).  Same dataflow as ff_forward_kernel, but two CTAs of a
// cluster -- two SMs of one TPC -- share every tcgen05.mma: M = 256 (each CTA's 128-row slab is the
// A operand for its own half of the rows and its own TMEM holds its rows of D), and the B operand's
// N columns are SPLIT between the two CTAs' shared memory.  Each SM therefore ingests HALF of every
// weight tile (16 KB instead of 32 KB per k-block): the measured limiter of the single-CTA kernel
// was the SM's L2 ingest of the re-streamed weights (98.8 MB L2->SM per launch).
//
// Protocol (leader = cluster rank 0):
//   TMA       both CTAs load their X tile and their half of each weight tile; all completion bytes
//             are counted on the LEADER's full barriers (cp.async.bulk.tensor .cta_group::2)
//   MMA       one thread of the leader issues tcgen05.mma.cta_group::2; tcgen05.commit multicasts
//             "stage free" / "accumulator ready" to the barriers of BOTH CTAs
//   epilogue  8 warps per CTA drain their own TMEM, write H (bf16, K-major SW128 boxes) into their
//             own shared memory and arrive (release.cluster) on the LEADER's h_lo / h_box /
//             acc_free barriers, which the leader's MMA thread acquires at cluster scope
// Epilogue arithmetic per element pair: two fp32 bias adds (bias staged in shared memory, read as
// broadcast float4), one cvt.rn.bf16x2, one max.bf16x2 against -0.0 (Rectify keeps the sign of a*0,
// nn.go:162-189: the sign bit is the mask) -- about half the instructions of the single-CTA version.
// Layer 3 uses W3 padded to 128 columns (N = 128: 64 from each CTA; only columns < A are read back).
// ---------------------------------------------------------------------------
constexpr int FF2_WSTAGES = 5;
// order in which layer 3 consumes the k-blocks (= boxes of H2): the two column groups of the epilogue complete boxes
// 4 and 6 first, then 5 and 7
__host__ __device__ constexpr int ff2_kperm(int kb) { return (kb == 5 || kb == 6) ? (kb ^ 3) : kb; }
constexpr uint32_t FF2_WSTAGE_BYTES = 16384;   // this CTA's half of a [64 k][256 n] tile: [2 n-groups][64 k-rows][128 B]

// CODE SIZE MATTERS HERE.  The first version of this kernel was fully unrolled (6 000 SASS instructions,
// 97 KB, most of them executed exactly once): in-kernel clock stamps (tools/ff_stamps.py) showed the
// single-thread roles -- TMA producer, MMA issuer -- losing 1-2.5 k cycles at every phase change to cold
// instruction fetches (the first k-block after each gate took 2.5 k cycles, the following ones 450).  Every
// role is therefore ONE rolled loop over a schedule computed arithmetically from the iteration index, so
// that its instructions are fetched once and re-executed.
template <int KD, int EW>   // EW = epilogue warps per CTA: 8 (two column groups of 128 per accumulator half) or 16 (four of 64)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
ff_forward2_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                   const __grid_constant__ CUtensorMap tmW1_0, const __grid_constant__ CUtensorMap tmW1_1,
                   const __grid_constant__ CUtensorMap tmW2_0, const __grid_constant__ CUtensorMap tmW2_1,
                   const __grid_constant__ CUtensorMap tmW3_0, const __grid_constant__ CUtensorMap tmW3_1,
                   const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmH2,
                   const FFArgs g)
{
    BlockTrace trace_(TR_FORWARD);
    constexpr uint32_t X_BYTES = KD * FF_BOX;
    constexpr uint32_t IDESC_256 = make_idesc_bf16(256, 256, false, true);
    constexpr uint32_t IDESC_128 = make_idesc_bf16(256, 128, false, true);
    constexpr int KH = FF_H / 64;                       // 8 k-blocks of layers 2 and 3
    constexpr int NB1 = 2 * KD, NB2 = 2 * KH, NB3 = KH / 2;   // ring stages consumed by layers 1, 2, 3
    constexpr int NB = NB1 + NB2 + NB3;
    constexpr uint16_t BOTH = 3;
    constexpr int EPI_THREADS = 32 * EW, NCG = EW / 4, NC = 8 / NCG;   // column groups per half, 32-column chunks per thread and phase
    static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sH = smem;                                 // 8 boxes: H1, then H2
    uint8_t *sX = sH + (KH - KD) * FF_BOX;              // aliases the last boxes of H (see ff_forward_kernel)
    uint8_t *sW = sH + KH * FF_BOX;                     // weight ring
    __shared__ uint64_t full_w[FF2_WSTAGES], empty_w[FF2_WSTAGES], x_full, acc_full[2], acc_free[2], h_box[4], h_lo, lo_free;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_bias[2 * FF_H + 32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int net = blockIdx.y + g.net0;
    const int m0 = blockIdx.x * BM;
    const uint32_t rank = cluster_ctarank();            // 0 = leader
    long long *const stp = g.stamps ? g.stamps + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * FF_NSTAMPS : nullptr;
#define FF_STAMP(i) do { if (stp) stp[i] = clock64(); } while (0)
    if (threadIdx.x == 0) {
        FF_STAMP(0);
        if (stp) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); stp[30] = (long long)t; }
    }
    const CUtensorMap *tmX = net ? &tmX1 : &tmX0;
    const CUtensorMap *tmW1 = net ? &tmW1_1 : &tmW1_0;
    const CUtensorMap *tmW2 = net ? &tmW2_1 : &tmW2_0;
    const CUtensorMap *tmW3 = net ? &tmW3_1 : &tmW3_0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(tmX); prefetch_tmap(tmW1); prefetch_tmap(tmW2); prefetch_tmap(tmW3);
        for (int s = 0; s < FF2_WSTAGES; s++) { mbar_init(&full_w[s], 1); mbar_init(&empty_w[s], 1); }
        mbar_init(&x_full, 1);
        mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
        mbar_init(&lo_free, 1);
        // the leader's copies collect the arrivals of both CTAs
        mbar_init(&acc_free[0], 2 * EW); mbar_init(&acc_free[1], 2 * EW);
        for (int b = 0; b < 4; b++) mbar_init(&h_box[b], 2);     // boxes 4..7 of H: one arrival per CTA
        mbar_init(&h_lo, 2);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(&tmem_base_slot, 512);
    // (the dependent grid -- TD/dZ2 -- is released late, from the epilogue: released here, its blocks would sit on the
    // idle SMs from the start, and the blocks that do not fit stall the block scheduler's queue, so that the NEXT
    // step's gather on the side branch cannot run under this kernel: tools/trace_step.py)
    if (!g.late_trigger) pdl_launch_dependents();
    tc_fence_before();
    cluster_sync();          // both CTAs' barriers are initialised and both TMEM allocations exist
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    // the gathered batch does not come from the kernel this one is launched behind (x_ready: it was gathered on the
    // previous step's side branch, a full dependency): its tile is requested before the wait
    if (warp == 0 && g.x_ready) {
        if (elect_one()) {
            const uint32_t x_full_leader = mapa_u32(&x_full, 0);
            if (rank == 0) mbar_expect_tx(&x_full, 2 * X_BYTES);
#pragma unroll
            for (int kb = 0; kb < KD; kb++) tma_load_2d_pair(sX + kb * FF_BOX, tmX, kb * 64, m0, x_full_leader);
        }
        __syncwarp();
    }
    pdl_wait();
    if (threadIdx.x == 0) FF_STAMP(1);

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own X tile, own half of every weight tile =====
        if (elect_one()) {
            const uint32_t x_full_leader = mapa_u32(&x_full, 0);
            if (!g.x_ready) {
                if (rank == 0) mbar_expect_tx(&x_full, 2 * X_BYTES);
#pragma unroll
                for (int kb = 0; kb < KD; kb++) tma_load_2d_pair(sX + kb * FF_BOX, tmX, kb * 64, m0, x_full_leader);
            }
            const uint32_t full_leader0 = mapa_u32(&full_w[0], 0);
            const int nr = (int)rank * 128;                        // this CTA's half of a 256-column tile
            // ring stage `it` = two boxes of [64 k][64 n]: layers 1-2 -> two n-groups of one k-block,
            // layer 3 (N = 128, this CTA's 64 columns) -> two consecutive k-blocks
#pragma unroll 1
            for (int it = 0; it < NB; it++) {
                const CUtensorMap *tm;
                int n0, k0, n1, k1;
                if (it < NB1 + NB2) {
                    const bool l1 = it < NB1;
                    const int j = l1 ? it : it - NB1;
                    // layer 1: half 0, half 1.  Layer 2 in four groups of KH/2 stages (see the MMA issuer):
                    // (half 0, k-blocks 0..3), (half 1, 0..3), (half 0, 4..7), (half 1, 4..7)
                    const int grp = j / (KH / 2);
                    const int half = l1 ? (j >= KD ? 1 : 0) : (grp & 1);
                    const int kb = l1 ? j - half * KD : (grp >> 1) * (KH / 2) + j % (KH / 2);
                    tm = l1 ? tmW1 : tmW2;
                    n0 = half * 256 + nr; k0 = kb * 64; n1 = n0 + 64; k1 = k0;
                } else {
                    const int j = it - NB1 - NB2;
                    tm = tmW3;
                    n0 = (int)rank * 64; k0 = ff2_kperm(2 * j) * 64; n1 = n0; k1 = ff2_kperm(2 * j + 1) * 64;
                }
                const int s = it % FF2_WSTAGES;
                const uint32_t ph = (it / FF2_WSTAGES) & 1;
                mbar_wait_bounded<false>(&empty_w[s], ph ^ 1);     // own ring slot free (the leader's commit reaches both CTAs)
                if (80 + it < FF_NSTAMPS) FF_STAMP(80 + it);
                if (rank == 0) mbar_expect_tx(&full_w[s], 2 * FF2_WSTAGE_BYTES);
                const uint32_t bar = full_leader0 + (uint32_t)s * 8u;
                tma_load_2d_pair(sW + s * FF2_WSTAGE_BYTES, tm, n0, k0, bar);
                tma_load_2d_pair(sW + s * FF2_WSTAGE_BYTES + 8192, tm, n1, k1, bar);
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===== MMA issuer: the leader's warp 1, three loop nests (one per layer) =====
            // All waits are CTA-scope acquires: operands arrive through the async proxy (TMA), and the peer's
            // epilogue publishes H with fence.proxy.async + bar.sync before it arrives on the leader's barrier.
            //
            // THIS THREAD HAS NO SLACK.  tcgen05.mma issue blocks until the previous MMA has started, so of a stage's
            // 4 x 139 cycles only the last ~280 are left for everything else the thread does before the next stage's
            // first MMA must be in the queue; whatever exceeds that is added to EVERY stage (tools/ff_stamps.py: the
            // single rolled loop that derived layer / half / k-block / ring slot from its index ran at 566 cycles per
            // stage, 636 with three more selects in it and 856 with one more barrier test).  Hence: no index
            // arithmetic (nested loops, ring slot and phase carried), descriptors advanced by adding to their low word,
            // gates only where a layer really has to wait.  The WHOLE WARP runs the loops and one elected lane issues:
            // tcgen05.mma takes its operands from uniform registers, and inside an `if (lane == 0)` region the compiler
            // has to move every descriptor there with an ELECT / R2UR.BROADCAST loop (671 cycles per stage).  Layer 2
            // waits once for the second half of H1 (the last box of each column group), layer 3 -- stages twice as
            // long -- box by box, in the order the epilogue completes them (ff2_kperm).
            {
                mbar_wait_bounded<false>(&x_full, 0);
                if (lane == 0) FF_STAMP(2);
                constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | ((uint32_t)SW128 << 29);   // SBO 1024, bit 46, SWIZZLE_128B
                const uint32_t ax_lo = ((smem_u32(sX) & 0x3FFFFu) >> 4) | ((16u >> 4) << 16);       // A: LBO 16
                const uint32_t ah_lo = ((smem_u32(sH) & 0x3FFFFu) >> 4) | ((16u >> 4) << 16);
                const uint32_t bw_lo = ((smem_u32(sW) & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16);     // B: LBO 8192
                uint32_t slot = 0, ph = 0;
                int nst = 0;                                    // (stage counter, stamps only)
                // one ring stage: wait for its weights, 4 MMAs per k-block (nkb = 1 or 2 boxes of A), free the slot
                auto stage = [&](uint32_t a0_lo, uint32_t a1_lo, int nkb, uint32_t d_tmem, uint32_t idesc, uint32_t acc) {
                    mbar_wait_bounded<false>(&full_w[slot], ph);
                    tc_fence_after();
                    if (lane == 0 && 32 + 2 * nst + 1 < 80) FF_STAMP(32 + 2 * nst);
                    if (elect_one()) {
                        const uint32_t b_lo = bw_lo + slot * (FF2_WSTAGE_BYTES >> 4);
#pragma unroll
                        for (int kk = 0; kk < BK / UK; kk++) {
                            mma_bf16_pair_lohi(d_tmem, a0_lo + kk * (32 >> 4), DESC_HI, b_lo + kk * (2048 >> 4), DESC_HI, idesc, acc);
                            acc = 1u;
                        }
                        if (nkb == 2) {
#pragma unroll
                            for (int kk = 0; kk < BK / UK; kk++)
                                mma_bf16_pair_lohi(d_tmem, a1_lo + kk * (32 >> 4), DESC_HI, b_lo + (8192 >> 4) + kk * (2048 >> 4), DESC_HI, idesc, 1u);
                        }
                        mma_commit_pair(&empty_w[slot], BOTH);
                    }
                    __syncwarp();
                    if (++slot == FF2_WSTAGES) { slot = 0; ph ^= 1u; }
                };
                auto commit = [&](uint64_t *bar) {
                    if (elect_one()) mma_commit_pair(bar, BOTH);
                    __syncwarp();
                };
                // ---- layer 1: A = X
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
#pragma unroll 1
                    for (int kb = 0; kb < KD; kb++) {
                        stage(ax_lo + kb * (FF_BOX >> 4), 0u, 1, tmem + half * 256u, IDESC_256, kb ? 1u : 0u);
                        if (lane == 0 && 32 + 2 * nst + 1 < 80) FF_STAMP(32 + 2 * nst + 1);
                        nst++;
                    }
                    commit(&acc_full[half]);
                }
                // ---- layer 2: A = H1, in four groups of KH/2 stages:
                //   (half 0, k-blocks 0..3) after h_lo | (half 1, 0..3) once accumulator half 1 is drained | (half 0, 4..7)
                //   after the last H1 boxes of both column groups | (half 1, 4..7).
                // H2's first half is written IN PLACE over H1 boxes 0..3: in this order they have been read by both
                // halves after 8 of the 16 stages (lo_free), so the epilogue of half 0 starts when its accumulator is
                // complete (12 stages) instead of waiting for stage 12 of the old order (half 0 first, then half 1).
#pragma unroll 1
                for (int grp = 0; grp < 4; grp++) {
                    const int half = grp & 1, kb0 = (grp >> 1) * (KH / 2);
                    if (grp == 0) { mbar_wait_bounded<false>(&h_lo, 0); mbar_wait_bounded<false>(&acc_free[0], 0); }
                    else if (grp == 1) mbar_wait_bounded<false>(&acc_free[1], 0);
                    else if (grp == 2) {
                        mbar_wait_bounded<false>(&h_box[1], 0); mbar_wait_bounded<false>(&h_box[3], 0);
                        if (NCG == 4) { mbar_wait_bounded<false>(&h_box[0], 0); mbar_wait_bounded<false>(&h_box[2], 0); }
                    }
                    tc_fence_after();
#pragma unroll 1
                    for (int kb = kb0; kb < kb0 + KH / 2; kb++) {
                        stage(ah_lo + kb * (FF_BOX >> 4), 0u, 1, tmem + half * 256u, IDESC_256, kb ? 1u : 0u);
                        if (lane == 0 && 32 + 2 * nst + 1 < 80) FF_STAMP(32 + 2 * nst + 1);
                        nst++;
                    }
                    if (grp == 1) commit(&lo_free);      // the first four H1 boxes are no longer read: H2 may land on them
                    if (grp >= 2) commit(&acc_full[half]);
                }
                // ---- layer 3: A = H2, two boxes per stage, N = 128
                mbar_wait_bounded<false>(&h_lo, 1); mbar_wait_bounded<false>(&acc_free[0], 1);
                tc_fence_after();
#pragma unroll 1
                for (int j = 0; j < NB3; j++) {
                    const int b0 = ff2_kperm(2 * j), b1 = ff2_kperm(2 * j + 1);
                    if (j >= NB3 / 2) {
                        mbar_wait_bounded<false>(&h_box[b0 - KH / 2], 1); mbar_wait_bounded<false>(&h_box[b1 - KH / 2], 1);
                        tc_fence_after();
                    }
                    stage(ah_lo + b0 * (FF_BOX >> 4), ah_lo + b1 * (FF_BOX >> 4), 2, tmem, IDESC_128, j ? 1u : 0u);
                    if (lane == 0 && 32 + 2 * nst + 1 < 80) FF_STAMP(32 + 2 * nst + 1);
                    nst++;
                }
                commit(&acc_full[0]);
                if (lane == 0) FF_STAMP(9);
                mbar_wait_bounded<false>(&acc_free[1], 1);     // (every remote arrival is consumed before the pair may exit)
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue warps (both CTAs, own rows) =====
        const int q = warp & 3;                 // TMEM lane quarter
        const int cg = (warp - 2) >> 2;         // column group within a half: chunks cg*NC .. cg*NC+NC-1
        const int rt = q * 32 + lane;           // row within the slab
        const int et = threadIdx.x - 64;        // 0..EPI_THREADS-1
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t acc_free_leader0 = mapa_u32(&acc_free[0], 0);
        const uint32_t h_lo_leader = mapa_u32(&h_lo, 0), h_box_leader0 = mapa_u32(&h_box[0], 0);
        // biases of the three layers -> shared memory (read back as broadcast float4)
        for (int i = et; i < 2 * FF_H + 32; i += EPI_THREADS) {
            float v = 0.f;
            if (i < FF_H) v = ld_dep(g.bias[net][0] + i);
            else if (i < 2 * FF_H) v = ld_dep(g.bias[net][1] + i - FF_H);
            else if (i - 2 * FF_H < g.A) v = ld_dep(g.bias[net][2] + i - 2 * FF_H);
            s_bias[i] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");

        const __nv_bfloat162 neg0 = __floats2bfloat162_rn(-0.0f, -0.0f), zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
        const bool store_early = net == 0 && g.keep_h == 1;
        // phases: (layer 1, half 0), (layer 1, half 1), (layer 2, half 0), (layer 2, half 1) -- one loop body.
        // H2 overwrites H1 in place: its first half may be written once the layer-2 MMAs of accumulator half 1
        // have consumed H1 boxes 0..3 (lo_free), its second half once every layer-2 MMA has retired.
#pragma unroll 1
        for (int p = 0; p < 4; p++) {
            const int layer = p >> 1, half = p & 1;
            mbar_wait_bounded<false>(&acc_full[half], (uint32_t)layer);
            if (p == 2) {
                mbar_wait_bounded<false>(&lo_free, 0);
                // the TMA stores of H1 must have finished reading the tile before H2 lands on it
                if ((et & 127) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }
            if (p == 3) mbar_wait_bounded<false>(&acc_full[1], 1);   // (same barrier as above; kept for symmetry of the listing)
            tc_fence_after();
            if (et == 0) FF_STAMP(10 + p);
            const int cbase = half * 256 + cg * (32 * NC);           // first column of this warp's NC chunks
            const float *bias = s_bias + layer * FF_H + cbase;
            uint32_t r[2][32];
            uint32_t mw[NC];
            tmem_ld_32x32(tq + (uint32_t)cbase, r[0]);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                tmem_ld_wait();
                if (c < NC - 1) tmem_ld_32x32(tq + (uint32_t)(cbase + 32 * (c + 1)), r[(c + 1) & 1]);
                uint32_t packed[16];
                uint32_t mwc = 0u;
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4) {
                    const float4 bv = *reinterpret_cast<const float4 *>(bias + 32 * c + jj);
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(__uint_as_float(r[c & 1][jj]) + bv.x, __uint_as_float(r[c & 1][jj + 1]) + bv.y);
                    __nv_bfloat162 p1 = __floats2bfloat162_rn(__uint_as_float(r[c & 1][jj + 2]) + bv.z, __uint_as_float(r[c & 1][jj + 3]) + bv.w);
                    // Rectify mask of the four units = "pre-activation < 0 or NaN" (a * (a >= 0), nn.go:162-189): one packed
                    // compare per pair (0xffff per true half) and one and-or into the chunk's mask word
                    mwc |= __hltu2_mask(p0, zero2) & (0x00010001u << (jj >> 1));
                    mwc |= __hltu2_mask(p1, zero2) & (0x00010001u << ((jj >> 1) + 1));
                    p0 = __hmax2(p0, neg0);         // Rectify: negative (and NaN) -> -0.0
                    p1 = __hmax2(p1, neg0);
                    packed[jj >> 1] = *reinterpret_cast<uint32_t *>(&p0);
                    packed[(jj >> 1) + 1] = *reinterpret_cast<uint32_t *>(&p1);
                }
                mw[c] = mwc;                                   // bits 0..15: units 0, 2, .. 30; bits 16..31: units 1, 3, .. 31
                const int c0 = cbase + 32 * c;
                uint8_t *dst = sH + (c0 >> 6) * FF_BOX + rt * 128;
                const int ch0 = (c0 & 63) >> 3;
#pragma unroll
                for (int v = 0; v < 4; v++)
                    sts128(dst + (((ch0 + v) ^ (rt & 7)) << 4),
                               make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]));
                // The online network's activations leave box by box, as soon as the four warps of this column group
                // have written one: a single 128 KB burst at the end of a layer kept the TMA unit and the shared-memory
                // port busy for ~2.6 k cycles, during which the weight ring ran dry (tools/ff_stamps.py; 17.8 -> 16.3 us).
                // The same point releases the boxes of the second half to the MMA thread one by one.
                if ((c & 1) && (store_early || half == 1)) {
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + cg) : "memory");
                    if ((et & 127) == 0) {
                        const int box = c0 >> 6;
                        if (half == 1) mbar_arrive_cluster(h_box_leader0 + (uint32_t)(box - KH / 2) * 8u);
                        if (store_early) {
                            tma_store_2d(layer == 0 ? &tmH1 : &tmH2, sH + box * FF_BOX, 64 * box, m0);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                }
            }
            if (net == 0 && g.keep_h && m0 + rt < g.B)
            {
                uint32_t *mdst = g.maskbits[layer] + (size_t)(m0 + rt) * (FF_H / 32) + (cbase >> 5);
                if (NC == 4) *reinterpret_cast<uint4 *>(mdst) = make_uint4(mw[0], mw[1], mw[NC - 2], mw[NC - 1]);
                else *reinterpret_cast<uint2 *>(mdst) = make_uint2(mw[0], mw[1]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_free_leader0 + (uint32_t)half * 8u);
            if (half == 0) {
                // the first half of H is complete in this CTA: visible to the async proxy (MMA of the next layer)
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                if (et == 0) mbar_arrive_cluster(h_lo_leader);
            }
            if (et == 0) FF_STAMP(14 + p);
        }

        // ---- layer 3: Q = acc + b3 -> fp32 [B][A] (A <= 32: one chunk), staged in the idle weight ring
        // and copied out coalesced (the rows of a slab are contiguous in Q)
        if (g.late_trigger && et == 0) pdl_launch_dependents();
        if (cg == 0) {
            mbar_wait_bounded<false>(&acc_full[0], 0);          // third completion of acc_full[0]: parity 0 again
            tc_fence_after();
            if (et == 0) FF_STAMP(18);
            // (rolled loops: this code runs once per CTA, and straight-line code executed once costs its own
            // instruction fetch -- the 32-column unrolled version took 1 900 cycles from "accumulator ready" to "staged")
            float *sQ = reinterpret_cast<float *>(sW);
#pragma unroll 1
            for (int j0 = 0; j0 < g.A; j0 += 8) {
                uint32_t r8[8];
                tmem_ld_32x32_x8(tq + (uint32_t)j0, r8);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (j0 + u < g.A) sts32(sQ + rt * g.A + j0 + u, __uint_as_float(r8[u]) + s_bias[2 * FF_H + j0 + u]);
            }
            tc_fence_before();
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (et == 0) FF_STAMP(20);
            const int nq = ((g.B - m0 < BM) ? g.B - m0 : BM) * g.A;
            float *dst = g.q[net] + (size_t)m0 * g.A;
#pragma unroll 1
            for (int i0 = et; i0 < nq; i0 += 4 * 128) {
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = (i0 + 128 * u < nq) ? lds32(sQ + i0 + 128 * u) : 0.f;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (i0 + 128 * u < nq) dst[i0 + 128 * u] = v[u];
            }
            if (et == 0) FF_STAMP(21);
        }
        if ((et & 127) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (et == 0) FF_STAMP(19);
    }
#undef FF_STAMP
    tc_fence_before();
    cluster_sync_exit();     // nobody leaves while the pair's MMAs, commits or remote arrives may still touch it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem, 512);
    }
    if (threadIdx.x == 32 && stp) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); stp[31] = (long long)t; }
}

template <int KD>
constexpr size_t ff2_smem_bytes()
{
    return (size_t)8 * FF_BOX + FF2_WSTAGES * FF2_WSTAGE_BYTES + 1024;
}

// ---------------------------------------------------------------------------
// GEMM on CTA pairs (cta_group::2) for the backward pass: one 256 x 256 output tile per pair, each CTA owning
// 128 rows of A / D and loading 128 of the tile's 256 B columns, so a k-block costs each SM 32 KB of L2->SM
// traffic instead of 48 KB (single-CTA 128 x 256 tile) or 2 x 32 KB (two 128 x 128 tiles).  The measured
// L2->SM rate is 33-35 B/clk per SM whatever the access pattern (tools/ingest_probe.py); at this problem size
// every backward GEMM is bound by it, not by the tensor pipe.
//   PG_DX   dZ1 = (dZ2 W2^T) * mask1 : A = dZ2 [B][H2] K-major, B = W2 [H1][H2] K-major; epilogue = ReLU mask
//           from H1's sign bits -> bf16 -> TMA store, plus the tile's column sums (db1 partials) read back
//           from the staged tile -- the separate column-sum kernel and its 8 MB re-read are gone
//   PG_DW   dW = A^T-form split-K : A [K][M] MN-major, B [K][N] MN-major (activations as stored);
//           epilogue = fp32 partial tile -> slice `split`
// All roles are rolled loops (see ff_forward2_kernel on why code size matters).
// ---------------------------------------------------------------------------
enum { PG_DX = 0, PG_DW = 1 };
constexpr int PG_STAGES = 3;                  // 3 x 32 KB: two CTAs (of different kernels / tiles) share an SM and interleave their phases
constexpr uint32_t PG_STAGE_BYTES = 32768;      // A [128 x 64] + this CTA's half of B [128 x 64], bf16
constexpr int PG_THREADS = 320;                 // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 epilogue

struct PairGemmArgs {
    int trace_id;                 // TR_* (diagnostic)
    int late_trigger;             // release the dependent grid after the main loop instead of in the prologue
    int M, N, K;                  // logical problem; K = full reduction length
    int k_per_split;              // multiple of 64
    // PG_DX
    __nv_bfloat16 *out_bf16;      // (unused: the tile leaves by TMA store through tmC)
    const __nv_bfloat16 *mask_src;   // [M][ldo], sign bit set <=> masked (used when mask_bits is null)
    const uint32_t *mask_bits;       // [M][N / 32] mask words of the fused forward (FFArgs::maskbits)
    float *colsum;                // [M / 128][N] column sums of the stored (rounded) tile rows
    // PG_DW
    float *out_f32;               // [nsplit][M][ldo]
    long long slice_stride;
    int ldo;
};

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PG_THREADS, 2)
pair_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const PairGemmArgs g)
{
    TRACE_MARKS_DECL
    BlockTrace trace_(g.trace_id, TRACE_MARKS_PTR);
    constexpr bool MN = (KIND == PG_DW);          // both operands MN-major (A [K][M], B [K][N]) or both K-major
    constexpr uint32_t IDESC = make_idesc_bf16(256, 256, MN, MN);
    constexpr uint16_t BOTH = 3;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[PG_STAGES], empty_bar[PG_STAGES], acc_full;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int m0 = blockIdx.x * BM;                       // this CTA's 128 rows of the 256-row tile
    const int n0 = blockIdx.y * 256;                      // the pair's 256 columns
    const int nb = n0 + (int)rank * 128;                  // ... of which this CTA loads 128
    const int split = blockIdx.z;
    const int k_begin = split * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int nkb = (k_end - k_begin + BK - 1) / BK;      // same in both CTAs of the pair

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int s = 0; s < PG_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(&tmem_base_slot, 256);
    if (!g.late_trigger) pdl_launch_dependents();
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    pdl_wait();
    if (threadIdx.x == 0) TRACE_MARK(0);

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (elect_one()) {
            const uint32_t full_leader0 = mapa_u32(&full_bar[0], 0);
#pragma unroll 1
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % PG_STAGES;
                const uint32_t ph = (kb / PG_STAGES) & 1;
                mbar_wait_bounded<false>(&empty_bar[s], ph ^ 1);
                if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * PG_STAGE_BYTES);
                const uint32_t bar = full_leader0 + (uint32_t)s * 8u;
                uint8_t *sA = smem + s * PG_STAGE_BYTES, *sB = sA + 16384;
                const int k = k_begin + kb * BK;
                if (!MN) {      // K-major: box {64 k, 128 rows}
                    tma_load_2d_pair(sA, &tmA, k, m0, bar);
                    tma_load_2d_pair(sB, &tmB, k, nb, bar);
                } else {        // MN-major: boxes {64 mn, 64 k-rows}, 8 KB each
                    tma_load_2d_pair(sA, &tmA, m0, k, bar);
                    tma_load_2d_pair(sA + 8192, &tmA, m0 + 64, k, bar);
                    tma_load_2d_pair(sB, &tmB, nb, k, bar);
                    tma_load_2d_pair(sB + 8192, &tmB, nb + 64, k, bar);
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===== MMA issuer (leader) =====
#pragma unroll 1
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % PG_STAGES;
                const uint32_t ph = (kb / PG_STAGES) & 1;
                mbar_wait_bounded<false>(&full_bar[s], ph);
                tc_fence_after();
                if (kb == 0 && lane == 0) TRACE_MARK(1);
                if (elect_one()) {
                    const uint32_t aaddr = smem_u32(smem + s * PG_STAGE_BYTES), baddr = aaddr + 16384;
#pragma unroll
                    for (int kk = 0; kk < BK / UK; kk++) {
                        const uint64_t ad = MN ? make_smem_desc(aaddr + kk * 2048, 8192, 1024) : make_smem_desc(aaddr + kk * 32, 16, 1024);
                        const uint64_t bd = MN ? make_smem_desc(baddr + kk * 2048, 8192, 1024) : make_smem_desc(baddr + kk * 32, 16, 1024);
                        mma_bf16_pair(tmem, ad, bd, IDESC, (kb | kk) != 0 ? 1u : 0u);
                    }
                    mma_commit_pair(&empty_bar[s], BOTH);
                    if (kb == nkb - 1) mma_commit_pair(&acc_full, BOTH);
                }
                __syncwarp();
            }
            if (lane == 0) TRACE_MARK(2);
        }
    } else {
        // ===== epilogue (both CTAs, own 128 rows x 256 columns) =====
        const int q = warp & 3, cg = (warp - 2) >> 2;
        const int rt = q * 32 + lane;
        const int row = m0 + rt;
        const int et = threadIdx.x - 64;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        const int cbase = cg * 128;                        // this warp's four 32-column chunks
        if (KIND == PG_DX) {
            // the ReLU mask (sign bits of H1) of this thread's 128 columns, compressed to one bit per column and loaded
            // WHILE THE MAIN LOOP RUNS (these warps are idle until the accumulator is ready).  The first version
            // fetched the mask one 32-column chunk ahead of its use: every chunk then waited a whole L2 round trip
            // (in-kernel marks, tools/trace_step.py: 5.7 us for the four chunks -- more than the main loop).
            uint32_t mbits[4] = {0u, 0u, 0u, 0u};
            if (g.mask_bits) {       // one 16-byte load: the four mask words of this thread's 128 columns
                if (row < g.M) {
                    const uint4 t = ld_dep(reinterpret_cast<const uint4 *>(g.mask_bits + (size_t)row * (g.N >> 5) + ((n0 + cbase) >> 5)));
                    mbits[0] = t.x; mbits[1] = t.y; mbits[2] = t.z; mbits[3] = t.w;
                }
            } else {
                const uint4 *mp = reinterpret_cast<const uint4 *>(g.mask_src + (size_t)row * g.ldo + n0 + cbase);
#pragma unroll
                for (int hb = 0; hb < 2; hb++) {             // two batches of eight 16-byte loads
                    uint4 t[8];
#pragma unroll
                    for (int v = 0; v < 8; v++) t[v] = (row < g.M) ? ld_dep(mp + 8 * hb + v) : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int v = 0; v < 8; v++) {            // same word layout as the forward's (FFArgs::maskbits)
                        const uint32_t ev = ((t[v].x >> 15) & 1u) | (((t[v].y >> 15) & 1u) << 1) | (((t[v].z >> 15) & 1u) << 2) |
                                            (((t[v].w >> 15) & 1u) << 3);
                        const uint32_t od = (t[v].x >> 31) | ((t[v].y >> 31) << 1) | ((t[v].z >> 31) << 2) | ((t[v].w >> 31) << 3);
                        const int sh = ((8 * hb + v) & 3) * 4;
                        mbits[(8 * hb + v) >> 2] |= (ev << sh) | (od << (16 + sh));
                    }
                }
            }
            if (nkb > 0) mbar_wait_bounded<false>(&acc_full, 0);
            tc_fence_after();
            // the dependent grid (dW1) is released when the main loop is over: released in the prologue, its CTAs held
            // their shared memory idle for the whole of this kernel while dW2's CTAs waited for a slot
            if (g.late_trigger && et == 0) pdl_launch_dependents();
            if (et == 0) TRACE_MARK(3);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t r[32];
                tmem_ld_32x32(tq + (uint32_t)(cbase + 32 * c), r);
                tmem_ld_wait();
                uint32_t packed[16];
                const uint32_t mb = mbits[c];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float v0 = ((mb >> (j >> 1)) & 1u) ? 0.f : __uint_as_float(r[j]);
                    const float v1 = ((mb >> (16 + (j >> 1))) & 1u) ? 0.f : __uint_as_float(r[j + 1]);
                    __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
                    packed[j >> 1] = *reinterpret_cast<uint32_t *>(&pk);
                }
                // staged tile = the first 64 KB of the (now idle) ring: [4 boxes of 64 columns][128 rows][128 B], SW128
                const int c0 = cbase + 32 * c;
                uint8_t *dst = smem + (c0 >> 6) * 16384 + rt * 128;
                const int ch0 = (c0 & 63) >> 3;
#pragma unroll
                for (int v = 0; v < 4; v++)
                    sts128(dst + (((ch0 + v) ^ (rt & 7)) << 4),
                               make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]));
            }
            tc_fence_before();
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) {
                TRACE_MARK(4);
                for (int b = 0; b < 4; b++) tma_store_2d(&tmC, smem + b * 16384, n0 + 64 * b, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            // db1 partials of this 128-row slab, all eight warps: warp = (box of 64 columns, half of the rows); a lane
            // owns one 16-byte chunk (8 columns) of every fourth row of its half, so that a warp reads 512 contiguous
            // bytes per step; the four row phases are folded by shuffles, the two halves through shared memory --
            // every sum in a fixed order.
            {
                const int ew = et >> 5, bx = ew & 3, hf = ew >> 2;
                const uint8_t *box = smem + bx * 16384;
                const int ro = lane >> 3, ch = lane & 7;
                float acc[8];
#pragma unroll
                for (int u = 0; u < 8; u++) acc[u] = 0.f;
#pragma unroll
                for (int i0 = 0; i0 < 16; i0 += 8) {          // eight independent 16-byte shared loads in flight
                    uint4 wv[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int rr = 64 * hf + 4 * (i0 + i) + ro;
                        wv[i] = lds128(box + rr * 128 + ((ch ^ (rr & 7)) << 4));
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int rr = 64 * hf + 4 * (i0 + i) + ro;
                        if (m0 + rr < g.M) {
                            acc[0] += __uint_as_float(wv[i].x << 16); acc[1] += __uint_as_float(wv[i].x & 0xffff0000u);
                            acc[2] += __uint_as_float(wv[i].y << 16); acc[3] += __uint_as_float(wv[i].y & 0xffff0000u);
                            acc[4] += __uint_as_float(wv[i].z << 16); acc[5] += __uint_as_float(wv[i].z & 0xffff0000u);
                            acc[6] += __uint_as_float(wv[i].w << 16); acc[7] += __uint_as_float(wv[i].w & 0xffff0000u);
                        }
                    }
                }
                if (et == 0) TRACE_MARK(7);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], 8);
                    acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], 16);
                }
                float *fold = reinterpret_cast<float *>(smem + 4 * 16384);     // [4 boxes][64 columns], behind the staged tile
                if (hf == 1 && ro == 0) {
#pragma unroll
                    for (int u = 0; u < 8; u++) sts32(fold + bx * 64 + ch * 8 + u, acc[u]);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (hf == 0 && ro == 0) {
                    float *cs = g.colsum + (size_t)blockIdx.x * g.N + n0 + bx * 64 + ch * 8;
                    const float *fo = fold + bx * 64 + ch * 8;
                    const uint4 f0 = lds128(fo), f1 = lds128(fo + 4);
                    *reinterpret_cast<float4 *>(cs) = make_float4(acc[0] + __uint_as_float(f0.x), acc[1] + __uint_as_float(f0.y),
                                                                  acc[2] + __uint_as_float(f0.z), acc[3] + __uint_as_float(f0.w));
                    *reinterpret_cast<float4 *>(cs + 4) = make_float4(acc[4] + __uint_as_float(f1.x), acc[5] + __uint_as_float(f1.y),
                                                                      acc[6] + __uint_as_float(f1.z), acc[7] + __uint_as_float(f1.w));
                }
            }
            if (et == 0) TRACE_MARK(5);
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (et == 0) TRACE_MARK(6);
        } else {
            if (nkb > 0) mbar_wait_bounded<false>(&acc_full, 0);
            tc_fence_after();
            float *o = g.out_f32 + (size_t)split * g.slice_stride + (size_t)row * g.ldo + n0 + cbase;
#pragma unroll 1
            for (int c = 0; c < 4; c++) {
                uint32_t r[32];
                if (nkb > 0) { tmem_ld_32x32(tq + (uint32_t)(cbase + 32 * c), r); tmem_ld_wait(); }
                else {
#pragma unroll
                    for (int j = 0; j < 32; j++) r[j] = 0u;
                }
                if (row < g.M) {
#pragma unroll
                    for (int v = 0; v < 8; v++)
                        reinterpret_cast<float4 *>(o + 32 * c)[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                                                              __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    cluster_sync_exit();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem, 256);
    }
}

constexpr size_t pair_gemm_smem_bytes() { return (size_t)PG_STAGES * PG_STAGE_BYTES + 1024; }

// ---------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

bool load_encode(std::string &err)
{
    if (g_encode) return true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
        err = "cuTensorMapEncodeTiled entry point unavailable";
        return false;
    }
    g_encode = (EncodeTiledFn)fn;
    return true;
}

// row-major bf16 matrix [rows][cols] (cols contiguous, pitch ld elements); box {64 cols, box_rows}
bool make_tmap(CUtensorMap *m, const void *base, int rows, int cols, int ld, int box_rows, std::string &err)
{
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") rows=" + std::to_string(rows) +
              " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld);
        return false;
    }
    return true;
}

// one problem (ta1/tb1 = nullptr) or two problems sharing the launch
// Launch with (optionally) programmatic stream serialization: the kernel may start while its
// same-stream predecessor drains; see pdl_wait() in umma.cuh.
template <typename... KArgs, typename... Args>
cudaError_t launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                      int cluster_y, Args &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    if (cluster_y > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = 1;
        at[n].val.clusterDim.y = (unsigned)cluster_y;
        at[n].val.clusterDim.z = 1;
        n++;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                     Args &&...args)
{
    return launch_kc(kernel, grid, block, smem, st, pdl, 1, static_cast<Args &&>(args)...);
}

// opt in to the dynamic shared memory of one instantiation (called at create time, never
// under stream capture)
template <int MODE, int BN, int EPI, int CM = 1, int STAGES = DEFAULT_STAGES>
cudaError_t configure_gemm()
{
    return cudaFuncSetAttribute(umma_gemm_kernel<MODE, BN, EPI, CM, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)gemm_smem_bytes<MODE, BN, STAGES>());
}

cudaError_t configure_all_gemms()
{
    cudaError_t e = configure_gemm<MODE_KM, 128, EPI_BIAS_RELU_BF16>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KM, 64, EPI_BIAS_F32>();
    if (e == cudaSuccess) e = configure_gemm<MODE_MK, 32, EPI_F32_SLICE>();
    if (e == cudaSuccess) e = configure_gemm<MODE_MM, 128, EPI_F32_SLICE>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KK, 128, EPI_MASK_BF16>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KK, 128, EPI_F32_SLICE>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KK, 32, EPI_F32_SLICE>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KM, 128, EPI_F32_SLICE>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KM, 64, EPI_F32_SLICE>();
    // multicast variants
    if (e == cudaSuccess) e = configure_gemm<MODE_KM, 128, EPI_BIAS_RELU_BF16, 4>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KK, 128, EPI_MASK_BF16, 4>();
    if (e == cudaSuccess) e = configure_gemm<MODE_MM, 128, EPI_F32_SLICE, 4>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KK, 128, EPI_F32_SLICE, 4>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KM, 128, EPI_F32_SLICE, 4>();
    // 128x256 tiles, two stages (two CTAs per SM: 2 x 97 KB smem, 2 x 256 TMEM columns)
    if (e == cudaSuccess) e = configure_gemm<MODE_KM, 256, EPI_BIAS_RELU_BF16, 1, 2>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KK, 256, EPI_MASK_BF16, 1, 2>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KK, 256, EPI_MASK_BF16, 1, 4>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KM, 256, EPI_F32_SLICE, 1, 2>();
    if (e == cudaSuccess) e = configure_gemm<MODE_KK, 256, EPI_F32_SLICE, 1, 2>();
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ff_forward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff_smem_bytes<1>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ff_forward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff_smem_bytes<2>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pair_gemm_kernel<PG_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_gemm_smem_bytes());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pair_gemm_kernel<PG_DW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_gemm_smem_bytes());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ff_forward2_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff2_smem_bytes<1>());

    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ff_forward2_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff2_smem_bytes<2>());

    return e;
}

template <int MODE, int BN, int EPI, int CM = 1, int STAGES = DEFAULT_STAGES>
cudaError_t launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, GemmArgs g, int nsplit, cudaStream_t st,
                        const CUtensorMap *ta1 = nullptr, const CUtensorMap *tb1 = nullptr, bool pdl = false,
                        const CUtensorMap *tc0 = nullptr, const CUtensorMap *tc1 = nullptr)
{
    constexpr size_t smem = gemm_smem_bytes<MODE, BN, STAGES>();
    g.nsplit = nsplit;
    const int nprob = ta1 ? 2 : 1;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, nsplit * nprob);
    if (CM > 1 && grid.y % CM != 0) return cudaErrorInvalidConfiguration;
    if ((EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_MASK_BF16) && !tc0) return cudaErrorInvalidValue;
    return launch_kc(umma_gemm_kernel<MODE, BN, EPI, CM, STAGES>, grid, dim3(GEMM_THREADS), smem, st, pdl, CM, ta, tb,
                     ta1 ? *ta1 : ta, tb1 ? *tb1 : tb, tc0 ? *tc0 : ta, tc1 ? *tc1 : (tc0 ? *tc0 : ta), g);
}

// ---------------------------------------------------------------------------
// CUDA-core helpers of the wide path
// ---------------------------------------------------------------------------
// K1: sample (optional) + gather + fp32 -> bf16.  The replay stays fp32 in HBM; index
// generation and the cast are fused into the gather.  One thread per 16-byte granule.  (Four rows per thread,
// eight loads in flight each, measured SLOWER -- 19 vs 13 us: random 512-byte rows want more threads in
// flight, not fewer.)
__global__ void __launch_bounds__(256)
wide_gather_kernel(Replay rp, int D, const int32_t *__restrict__ idx, uint64_t seed,
                   const StepDev *__restrict__ step, int32_t *__restrict__ idx_out, int B, __nv_bfloat16 *__restrict__ x, __nv_bfloat16 *__restrict__ x2, int32_t *__restrict__ ba,
                   float *__restrict__ br, uint8_t *__restrict__ bdone)
{
    BlockTrace trace_(TR_GATHER);
    const int chunks = D >> 2;  // float4 granules per row (D % 8 == 0 is required by the path)
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();                 // x/x2/ba/... are still being read by the previous step's kernels
    if (t >= B * chunks) return;
    int b = t / chunks, c = t - b * chunks;
    if (step) { seed = step->seed; rp.head = step->head; rp.len = step->len; }
    const int32_t li = idx ? idx[b] : (int32_t)perm_index((uint32_t)b, (uint32_t)rp.len, seed);
    int64_t phys = rp.head - 1 - (int64_t)li;
    if (phys < 0) phys += rp.cap;
    float4 v = __ldg(reinterpret_cast<const float4 *>(rp.s + phys * D) + c);
    float4 w = __ldg(reinterpret_cast<const float4 *>(rp.s2 + phys * D) + c);
    __nv_bfloat162 a0 = __floats2bfloat162_rn(v.x, v.y), a1 = __floats2bfloat162_rn(v.z, v.w);
    __nv_bfloat162 b0 = __floats2bfloat162_rn(w.x, w.y), b1 = __floats2bfloat162_rn(w.z, w.w);
    uint2 pa = make_uint2(*reinterpret_cast<uint32_t *>(&a0), *reinterpret_cast<uint32_t *>(&a1));
    uint2 pb = make_uint2(*reinterpret_cast<uint32_t *>(&b0), *reinterpret_cast<uint32_t *>(&b1));
    reinterpret_cast<uint2 *>(x + (int64_t)b * D)[c] = pa;
    reinterpret_cast<uint2 *>(x2 + (int64_t)b * D)[c] = pb;
    if (c == 0) {
        ba[b] = rp.a[phys];
        br[b] = rp.r[phys];
        bdone[b] = rp.done[phys];
        if (idx_out && !idx) idx_out[b] = li;
    }
}

// bf16 shadow element writes for parameter i of value w (shared by the shadow kernel and
// the fused update kernel).  The weights keep the reference's (in x out) layout: W1 [D][H1]
// and W2 [H1][H2] are MN-major B operands of the forward GEMMs as stored (W2 is also the
// K-major B operand of dX); W3 is padded to 64 columns for the forward tile, and kept
// transposed [32][H2] for the dZ2 kernel's contiguous column reads.
constexpr int W3_PAD = 128;   // columns of the padded W3 shadow (the pair kernel runs layer 3 at N = 128)
struct ShadowPtrs {
    __nv_bfloat16 *w1, *w2, *w3p, *w3t;
};
__device__ __forceinline__ void shadow_store(const Dims &d, const Offsets &o, const ShadowPtrs &sp, int i, float w)
{
    const __nv_bfloat16 v = __float2bfloat16_rn(w);
    if (i < o.b1) {
        sp.w1[i] = v;
    } else if (i >= o.w2 && i < o.b2) {
        sp.w2[i - o.w2] = v;
    } else if (i >= o.w3 && i < o.b3) {
        const int u = i - o.w3;
        const int k = u / d.A, a = u - k * d.A;            // W3[k][a]
        sp.w3p[(size_t)k * W3_PAD + a] = v;
        if (sp.w3t) sp.w3t[(size_t)a * d.H2 + k] = v;
    }
}

__global__ void __launch_bounds__(256)
wide_shadow_kernel(const float *__restrict__ theta, Dims d, Offsets o, ShadowPtrs sp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < o.P) shadow_store(d, o, sp, i, theta[i]);
    // the padding of W3 stays zero from allocation time (cudaMemset in wide_create)
}

// K3 + layer-3 backward, one kernel.  Block = DZ_ROWS = 32 rows = 8 warps x 4 rows; everything about a row
// happens inside ONE warp, so the only block-level synchronisation is the final sum of the eight warps' partials
// (round 1's version computed the TD rows in warp 0 while the other seven waited at a barrier: ncu showed it
// stalled on that barrier, 6.1 per issue).
//   lanes 0..3 of a warp, one row each: first-max over Q'(s') (ArgmaxF32 rules), TD target (agent.go:149),
//     d = q[a] - t, dqv = (2 d) / count, loss term; broadcast by shuffle
//   all lanes: dQ^T [32][B] bf16 (one non-zero per column b; the K-major B operand of the dW3 GEMM), lane = action
//     row, four consecutive b per 8-byte store
//   all lanes: dZ2[b][j] = dqv[b] * W3[j][a_b] * mask2[b][j] (bf16 out, operand of dW2 / dX): lane l takes the
//     16-byte column chunks l, l + 32, ...; every load of the warp's four rows is issued before the first use
//   partials per block in a fixed order: db2 (column sums of the ROUNDED dZ2), db3 by action, loss
constexpr int DZ_ROWS = 32;
constexpr int DZ_MAXCH = 2;      // 16-byte chunks per lane and row: H2 <= 32 * 8 * DZ_MAXCH = 512 (more: chunk loop)
__global__ void __launch_bounds__(256, 2)   // two blocks per SM: all 256 blocks of the headline shape in one wave
wide_td_dz2_kernel(const float *__restrict__ q, const float *__restrict__ qt, const int32_t *__restrict__ act,
                   const float *__restrict__ rew, const uint8_t *__restrict__ done, int B, int H2, int A, StepScalars sc,
                   const __nv_bfloat16 *__restrict__ h2, const uint32_t *__restrict__ m2bits /* [B][H2/32] or null */,
                   const __nv_bfloat16 *__restrict__ w3t /* [32][H2] */, float *__restrict__ td_out, __nv_bfloat16 *__restrict__ dqT, __nv_bfloat16 *__restrict__ dz2,
                   float *__restrict__ b2_part /* [blocks][H2] */, float *__restrict__ row_part /* [blocks][A+1] */, int early_ok)
{
    TRACE_MARKS_DECL
    BlockTrace trace_(TR_TD, TRACE_MARKS_PTR);
    extern __shared__ float sm[];            // [8 warps][H2] column sums, then [8][33] row partials
    float *sRow = sm + 8 * H2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rw0 = blockIdx.x * DZ_ROWS + warp * 4;      // this warp's four rows
    // ---- Ahead of the wait: what does NOT come from the forward this kernel is launched behind.  Actions, rewards and
    // done flags were written by the gather (a full dependency of the forward, complete before its first CTA ran), W3^T by
    // the previous update.  Loading them here -- and the W3 rows the actions pick, a dependent round trip -- leaves ONE L2
    // round trip behind the wait (Q, Q', mask words).  These are plain cached loads issued in program order in front of an
    // `asm volatile` with a memory clobber: nothing the compiler may move; tests/test_sass_invariants.py holds this kernel
    // to exactly these early loads and to none of the invariant (.CONSTANT) kind.
    // early_ok = the host's promise that the gather HAS completed when this grid starts: true behind the fused forward,
    // which releases its dependent only after its own wait (on the gather, or on a step whose batch was prefetched).  The
    // per-layer GEMM forward releases dependents from its prologue, down the whole chain: there this kernel can start
    // while the gather still runs, and loads everything behind the wait.
    pdl_launch_dependents();
    int a = 0, dn = 0;
    float rwd = 0.f;
    const int nch = H2 >> 3;
    uint4 wv[4][DZ_MAXCH];
    int a4[4];
    auto load_batch = [&]() {
        if (lane < 4 && rw0 + lane < B) {
            a = __ldca(act + rw0 + lane);
            rwd = __ldca(rew + rw0 + lane);
            dn = __ldca(done + rw0 + lane);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) a4[k] = __shfl_sync(0xffffffffu, a, k);
    };
    if (early_ok) load_batch();
    auto load_w = [&](int c0) {              // the W3 rows picked by the actions (only A rows are ever read: through L1)
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int u = 0; u < DZ_MAXCH; u++) {
                const int c = c0 + lane + 32 * u;
                const bool live = c < nch && rw0 + k < B;
                wv[k][u] = live ? __ldca(reinterpret_cast<const uint4 *>(w3t + (int64_t)a4[k] * H2) + c) : make_uint4(0, 0, 0, 0);
            }
    };
    if (early_ok) load_w(0);
    pdl_wait();
    if (threadIdx.x == 0) TRACE_MARK(0);
    if (!early_ok) { load_batch(); load_w(0); }
    // ---- behind the wait: Q'(s'), Q(s) and the mask words of the four rows, all issued before the first use
    float qtv[4], qv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool ok = rw0 + k < B && lane < A;
        qtv[k] = ok ? ld_dep(qt + (int64_t)(rw0 + k) * A + lane) : -INFINITY;
        qv[k] = ok ? ld_dep(q + (int64_t)(rw0 + k) * A + lane) : 0.f;
    }
    uint32_t mk[4][DZ_MAXCH];      // masks of the chunk's 8 units: bit e = unit 2e masked, bit 4 + e = unit 2e + 1 masked
    // (all loads first, decoding afterwards: decoded inside the load loop, every mask word's shift waited for its own
    // load -- eight serial L2 round trips, 2.8 us in the block trace)
    auto load_masks = [&](int c0) {
        if (m2bits) {            // the forward's mask words (FFArgs::maskbits): 4 bytes instead of the 16 bytes of H2
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int u = 0; u < DZ_MAXCH; u++) {
                    const int c = c0 + lane + 32 * u;
                    mk[k][u] = (c < nch && rw0 + k < B) ? ld_dep(m2bits + (int64_t)(rw0 + k) * (H2 >> 5) + (c >> 2)) : 0u;
                }
            // (decoded where it is used, below: nothing waits for these loads here)
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint4 hv[DZ_MAXCH];
#pragma unroll
                for (int u = 0; u < DZ_MAXCH; u++) {
                    const int c = c0 + lane + 32 * u;
                    hv[u] = (c < nch && rw0 + k < B) ? ld_dep(reinterpret_cast<const uint4 *>(h2 + (int64_t)(rw0 + k) * H2) + c) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < DZ_MAXCH; u++)
                    mk[k][u] = ((hv[u].x >> 15) & 1u) | (((hv[u].y >> 15) & 1u) << 1) | (((hv[u].z >> 15) & 1u) << 2) |
                               (((hv[u].w >> 15) & 1u) << 3) | ((hv[u].x >> 31) << 4) | ((hv[u].y >> 31) << 5) |
                               ((hv[u].z >> 31) << 6) | ((hv[u].w >> 31) << 7);
            }
        }
    };
    load_masks(0);
#ifdef GOLDQ_TRACE_BUILD
    if (threadIdx.x == 0 && a4[0] >= 0) TRACE_MARK(5);
#endif
    // ---- TD label: first-max over Q'(s') with ArgmaxF32's rules, lane = action, one row at a time; lane k keeps row k's.
    //   sequential rule (dense/op.go:42-62 over gonum floats.MaxIdx): f = v[0]; for i >= 1: the first NaN or +Inf ends
    //   the scan with f = v[i]; otherwise f = v[i] if v[i] > f.  In parallel: the lowest special lane if any, else a
    //   shuffle tree that takes the higher lane's value only if it is strictly greater (so the first maximum wins and a
    //   NaN in v[0] survives, as `v > NaN` is false).
    float fmine = 0.f, qamine = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float v = qtv[k];
        const bool special = lane >= 1 && lane < A && (isnan(v) || (isinf(v) && v > 0.f));
        const unsigned bal = __ballot_sync(0xffffffffu, special);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float o = __shfl_down_sync(0xffffffffu, v, off);
            v = (lane + off < 32 && o > v) ? o : v;
        }
        const float fmax_ = __shfl_sync(0xffffffffu, v, 0);
        const float fsp = __shfl_sync(0xffffffffu, qtv[k], bal ? (__ffs(bal) - 1) : 0);
        const float qa = __shfl_sync(0xffffffffu, qv[k], a4[k]);
        if (lane == k) { fmine = bal ? fsp : fmax_; qamine = qa; }
    }
#ifdef GOLDQ_TRACE_BUILD
    if (threadIdx.x == 0 && fmine == fmine) TRACE_MARK(6);
#endif
    float dqv = 0.f, sq = 0.f;
    if (lane < 4 && rw0 + lane < B) {
        const float t = dn ? rwd : __fadd_rn(rwd, __fmul_rn(sc.gamma, fmine));
        td_out[rw0 + lane] = t;
        residual_loss(sc, __fsub_rn(qamine, t), sq, dqv);
    }
    if (threadIdx.x == 0) TRACE_MARK(1);
    float dq4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) dq4[k] = __shfl_sync(0xffffffffu, dqv, k);
    // ---- dQ^T: lane = action row i, columns rw0 .. rw0 + 3 in one 8-byte store (B % 8 == 0, rw0 % 4 == 0) ----
    {
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(a4[0] == lane ? dq4[0] : 0.f, a4[1] == lane ? dq4[1] : 0.f);
        const __nv_bfloat162 p1 = __floats2bfloat162_rn(a4[2] == lane ? dq4[2] : 0.f, a4[3] == lane ? dq4[3] : 0.f);
        if (rw0 + 3 < B)
            *reinterpret_cast<uint2 *>(dqT + (int64_t)lane * B + rw0) =
                make_uint2(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1));
        else {
            const float v[4] = {a4[0] == lane ? dq4[0] : 0.f, a4[1] == lane ? dq4[1] : 0.f, a4[2] == lane ? dq4[2] : 0.f,
                                a4[3] == lane ? dq4[3] : 0.f};
            for (int k = 0; k < 4; k++)
                if (rw0 + k < B) dqT[(int64_t)lane * B + rw0 + k] = __float2bfloat16_rn(v[k]);
        }
    }
    // ---- row partials of this warp: db3[i] (lane i), loss (lane A) --------------------------------------
    {
        float v = 0.f, ls = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            ls += __shfl_sync(0xffffffffu, sq, k);
            if (lane < A) v += (a4[k] == lane && rw0 + k < B) ? dq4[k] : 0.f;
        }
        if (lane < A) sRow[warp * 33 + lane] = v;
        if (lane == 0) sRow[warp * 33 + A] = ls;      // (slot A <= 32: also right for 32 actions)
    }
    if (threadIdx.x == 0) TRACE_MARK(2);
    // ---- dZ2 of the four rows + column sums ----------------------------------------------------------------
    for (int c0 = 0; c0 < nch; c0 += 32 * DZ_MAXCH) {
        if (c0 > 0) { load_masks(c0); load_w(c0); }
        if (m2bits) {
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int u = 0; u < DZ_MAXCH; u++) {
                    const int sh = ((c0 + lane + 32 * u) & 3) * 4;
                    mk[k][u] = ((mk[k][u] >> sh) & 0xfu) | (((mk[k][u] >> (16 + sh)) & 0xfu) << 4);
                }
        }
#pragma unroll
        for (int u = 0; u < DZ_MAXCH; u++) {
            const int c = c0 + lane + 32 * u;
            float cs[8];
#pragma unroll
            for (int e = 0; e < 8; e++) cs[e] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t ww[4] = {wv[k][u].x, wv[k][u].y, wv[k][u].z, wv[k][u].w};
                uint32_t ow[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float w0 = __uint_as_float(ww[e] << 16), w1 = __uint_as_float(ww[e] & 0xffff0000u);
                    const float z0 = ((mk[k][u] >> e) & 1u) ? 0.f : dq4[k] * w0;
                    const float z1 = ((mk[k][u] >> (4 + e)) & 1u) ? 0.f : dq4[k] * w1;
                    __nv_bfloat162 pk = __floats2bfloat162_rn(z0, z1);
                    ow[e] = *reinterpret_cast<uint32_t *>(&pk);
                    // sum the ROUNDED values so that db2 matches the bf16 operand of dW2 / dX
                    cs[2 * e] += __uint_as_float(ow[e] << 16);
                    cs[2 * e + 1] += __uint_as_float(ow[e] & 0xffff0000u);
                }
                if (c < nch && rw0 + k < B) reinterpret_cast<uint4 *>(dz2 + (int64_t)(rw0 + k) * H2)[c] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
            if (c < nch) {
                float4 *dst = reinterpret_cast<float4 *>(sm + warp * H2 + c * 8);
                dst[0] = make_float4(cs[0], cs[1], cs[2], cs[3]);
                dst[1] = make_float4(cs[4], cs[5], cs[6], cs[7]);
            }
        }
    }
    if (threadIdx.x == 0) TRACE_MARK(3);
    __syncthreads();
    if (threadIdx.x == 0) TRACE_MARK(4);
    // ---- block partials, warps summed in order ------------------------------------------------------------------
    for (int j = threadIdx.x; j < H2; j += blockDim.x) {
        float v = sm[j];
#pragma unroll
        for (int w = 1; w < 8; w++) v += sm[w * H2 + j];
        b2_part[(int64_t)blockIdx.x * H2 + j] = v;
    }
    if (threadIdx.x <= A) {
        float v = sRow[threadIdx.x];
#pragma unroll
        for (int w = 1; w < 8; w++) v += sRow[w * 33 + threadIdx.x];
        row_part[(int64_t)blockIdx.x * (A + 1) + threadIdx.x] = v;
    }
}

// db1 partials: column sums of dZ1 (bf16 [B][H1]) over CS_ROWS-row ranges.  A warp covers
// 64 columns (one bfloat162 per lane), the 8 warps of a block take interleaved rows with
// 8 independent loads in flight each; warps are then summed in order through smem.
constexpr int CS_ROWS = 256;
__global__ void __launch_bounds__(256)
wide_colsum_kernel(const __nv_bfloat16 *__restrict__ dz, int B, int H, float *__restrict__ part /* [blocks.y][H] */)
{
    BlockTrace trace_(TR_COLSUM);
    __shared__ float2 sh[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.x * 64 + lane * 2;
    const int r0 = blockIdx.y * CS_ROWS, r1 = min(B, r0 + CS_ROWS);
    float2 s = make_float2(0.f, 0.f);
    if (j < H) {
        for (int b = r0 + warp; b < r1; b += 64) {
            __nv_bfloat162 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int bb = b + 8 * u;
                v[u] = (bb < r1) ? ld_dep(reinterpret_cast<const __nv_bfloat162 *>(dz + (int64_t)bb * H + j))
                                 : __floats2bfloat162_rn(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                float2 f = __bfloat1622float2(v[u]);
                s.x += f.x;
                s.y += f.y;
            }
        }
    }
    sh[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && j < H) {
        float2 t = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; w++) { t.x += sh[w][lane].x; t.y += sh[w][lane].y; }
        part[(int64_t)blockIdx.y * H + j] = t.x;
        if (j + 1 < H) part[(int64_t)blockIdx.y * H + j + 1] = t.y;
    }
}

// Fixed-order sums of all partial buffers -> flat[P+1] (flat[P] = loss) and, when FUSE,
// the Adam update (kernels_generic.cu K5 semantics) plus the bf16 shadow refresh in the
// same pass.
struct PartialSrc {
    const float *w1_slices; int n1;     // [n1][D*H1]
    const float *b1_part;   int nb1;    // [nb1][H1]
    const float *w2_slices; int n2;     // [n2][H1*H2]
    const float *b2_part;   int nb2;    // [nb2][H2]
    const float *w3_slices; int n3;     // [n3][H2*A]
    const float *row_part;  int nrow;   // [nrow][A+1]: db3, loss
};

// Work decomposition of the update kernel (one launch, three block ranges):
//   A  thread per float4 group of W1 / W2 / W3: few partials (split-K slices), batches of 8
//      independent 16-byte loads;
//   B  warp per float4 group of b1 / b2: many partials (one per row block): lanes stride
//      over the partials, fixed-order butterfly;
//   C  warp per scalar of b3 and the loss.
struct UpdatePlan {
    int nA;        // float4 groups in W1, W2, W3 (in parameter order)
    int nB;        // float4 groups in b1, b2 (this launch's: gB0 .. gB0 + nB - 1)
    int gB0;       // first bias group of this launch (0, or H1 / 4 when only b2 is updated)
    int nC;        // A + 1 scalars
    int blocksA, blocksB, blocksC;
    int lp[3];     // lanes per float4 group of W1 / W2 / W3 (1, 2 or 4): a range with many split-K slices gives each group
                   // several lanes, so that every lane sums its share of the slices in ONE round of loads
    int bA[3];     // blocks of the three ranges (blocksA = their sum)
};

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ---- {value, epoch} words (data-parallel exchange, see PeerSet in wide.cuh) ---------------------
__device__ __forceinline__ void ll_store(unsigned long long *p, float val, unsigned int epoch)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(val)), "r"(epoch) : "memory");
}
// two adjacent words in one 16-byte access (each 8-byte half stays atomic; halves carry their own tag)
__device__ __forceinline__ void ll_store2(unsigned long long *p, float v0, float v1, unsigned int epoch)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(v0)), "r"(epoch),
                 "r"(__float_as_uint(v1)), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint4 ll_load2(const unsigned long long *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// local gradient element e of this rank -> the receive window of the rank that owns e
__device__ __forceinline__ void ll_push(const PeerSet &L, unsigned int epoch, int e, float val)
{
    const int owner = e / L.chunk_f;
    unsigned long long *dst = L.rs[owner] + ((size_t)(epoch & 1u) * L.world + L.rank) * L.chunk_f + (e - owner * L.chunk_f);
    ll_store(dst, val, epoch);
}

// a whole float4 group (i multiple of 4: one owner, 32-byte aligned destination)
__device__ __forceinline__ void ll_push4(const PeerSet &L, unsigned int epoch, int i, float4 g)
{
    const int owner = i / L.chunk_f;
    unsigned long long *dst = L.rs[owner] + ((size_t)(epoch & 1u) * L.world + L.rank) * L.chunk_f + (i - owner * L.chunk_f);
    ll_store2(dst, g.x, g.y, epoch);
    ll_store2(dst + 2, g.z, g.w, epoch);
}

// The whole exchange for one float4 group, called by the SAME thread (same launch geometry) on
// every rank: push the local value to the owner; on the owner wait for the `world` copies, add
// them in rank order and write the sums into every rank's all-gather window; wait for the sums.
// A thread only ever waits on words produced by threads that wait on nothing but pushes, and the
// whole grid is resident (wide_p2p_resident, checked by goldq_dp_init: otherwise the group keeps
// ncclAllReduce), so no wait can starve its producer.  A word that never arrives (5 s) sets the
// host-visible error flag and the thread leaves parameters, moments and shadows untouched.
__device__ __forceinline__ bool ll_timed_out(unsigned long long &t0, int *err)
{
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (!t0) t0 = t1;
    if (t1 - t0 > 5000000000ull) { *err = 1; return true; }      // 5 s: a peer is gone; do not hang the GPU
    return false;
}
__device__ __forceinline__ float4 dp_exchange4(const PeerSet &L, unsigned int epoch, int i, float4 g, bool &ok)
{
    ll_push4(L, epoch, i, g);
    const size_t par = epoch & 1u;
    const int owner = i / L.chunk_f;
    if (owner == L.rank) {
        const unsigned long long *my_rs = L.rs[L.rank] + par * (size_t)L.world * L.chunk_f + (i - owner * L.chunk_f);
        uint4 a[P2P_MAX_WORLD], b[P2P_MAX_WORLD];
        unsigned long long t0 = 0;
        for (;;) {
            bool all = true;
#pragma unroll
            for (int r = 0; r < P2P_MAX_WORLD; r++) {
                if (r < L.world) {
                    a[r] = ll_load2(my_rs + (size_t)r * L.chunk_f);
                    b[r] = ll_load2(my_rs + (size_t)r * L.chunk_f + 2);
                }
            }
#pragma unroll
            for (int r = 0; r < P2P_MAX_WORLD; r++)
                if (r < L.world) all = all && a[r].y == epoch && a[r].w == epoch && b[r].y == epoch && b[r].w == epoch;
            if (all) break;
            if (ll_timed_out(t0, L.err)) { ok = false; break; }
        }
        float4 acc = make_float4(__uint_as_float(a[0].x), __uint_as_float(a[0].z), __uint_as_float(b[0].x), __uint_as_float(b[0].z));
#pragma unroll
        for (int r = 1; r < P2P_MAX_WORLD; r++)
            if (r < L.world)
                acc = f4add(acc, make_float4(__uint_as_float(a[r].x), __uint_as_float(a[r].z), __uint_as_float(b[r].x),
                                             __uint_as_float(b[r].z)));
#pragma unroll
        for (int r = 0; r < P2P_MAX_WORLD; r++) {
            if (r < L.world) {
                unsigned long long *dst = L.ag[r] + par * (size_t)L.world * L.chunk_f + i;
                ll_store2(dst, acc.x, acc.y, epoch);
                ll_store2(dst + 2, acc.z, acc.w, epoch);
            }
        }
    }
    const unsigned long long *my_ag = L.ag[L.rank] + par * (size_t)L.world * L.chunk_f + i;
    uint4 a, b;
    unsigned long long t0 = 0;
    for (;;) {
        a = ll_load2(my_ag);
        b = ll_load2(my_ag + 2);
        if (a.y == epoch && a.w == epoch && b.y == epoch && b.w == epoch) break;
        if (ll_timed_out(t0, L.err)) { ok = false; break; }
    }
    return make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
}
// same for a single element (b3 and the loss)
__device__ __forceinline__ float dp_exchange1(const PeerSet &L, unsigned int epoch, int e, float g, bool &ok)
{
    ll_push(L, epoch, e, g);
    const size_t par = epoch & 1u;
    const int owner = e / L.chunk_f;
    unsigned long long t0 = 0;
    if (owner == L.rank) {
        const unsigned long long *my_rs = L.rs[L.rank] + par * (size_t)L.world * L.chunk_f + (e - owner * L.chunk_f);
        unsigned int val[P2P_MAX_WORLD], tag[P2P_MAX_WORLD];
        for (;;) {
            bool all = true;
#pragma unroll
            for (int r = 0; r < P2P_MAX_WORLD; r++)
                if (r < L.world)
                    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(val[r]), "=r"(tag[r]) : "l"(my_rs + (size_t)r * L.chunk_f) : "memory");
#pragma unroll
            for (int r = 0; r < P2P_MAX_WORLD; r++)
                if (r < L.world) all = all && tag[r] == epoch;
            if (all) break;
            if (ll_timed_out(t0, L.err)) { ok = false; break; }
        }
        float acc = __uint_as_float(val[0]);
#pragma unroll
        for (int r = 1; r < P2P_MAX_WORLD; r++)
            if (r < L.world) acc += __uint_as_float(val[r]);
        for (int r = 0; r < L.world; r++) ll_store(L.ag[r] + par * (size_t)L.world * L.chunk_f + e, acc, epoch);
    }
    const unsigned long long *my_ag = L.ag[L.rank] + par * (size_t)L.world * L.chunk_f + e;
    unsigned int val, tag;
    t0 = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(val), "=r"(tag) : "l"(my_ag) : "memory");
        if (tag == epoch) break;
        if (ll_timed_out(t0, L.err)) { ok = false; break; }
    }
    return __uint_as_float(val);
}

template <bool FUSE>
__device__ __forceinline__ void finish4(const Dims &d, const Offsets &o, int i, float4 g, float *flat, float *theta,
                                        float *m, float *v, float4 w4, float4 m4, float4 v4, const AdamConsts &ac,
                                        const ShadowPtrs &sp)
{
    if (flat) *reinterpret_cast<float4 *>(flat + i) = g;
    if (!FUSE) return;
    solver_update(w4.x, g.x, m4.x, v4.x, ac);
    solver_update(w4.y, g.y, m4.y, v4.y, ac);
    solver_update(w4.z, g.z, m4.z, v4.z, ac);
    solver_update(w4.w, g.w, m4.w, v4.w, ac);
    *reinterpret_cast<float4 *>(theta + i) = w4;
    *reinterpret_cast<float4 *>(m + i) = m4;
    *reinterpret_cast<float4 *>(v + i) = v4;
    if (i < o.b1 || (i >= o.w2 && i < o.b2)) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(w4.x, w4.y), p1 = __floats2bfloat162_rn(w4.z, w4.w);
        uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&p0), *reinterpret_cast<uint32_t *>(&p1));
        __nv_bfloat16 *dst = (i < o.b1) ? sp.w1 + i : sp.w2 + (i - o.w2);
        *reinterpret_cast<uint2 *>(dst) = pk;
    } else if (i >= o.w3 && i < o.b3) {
        shadow_store(d, o, sp, i, w4.x);
        shadow_store(d, o, sp, i + 1, w4.y);
        shadow_store(d, o, sp, i + 2, w4.z);
        shadow_store(d, o, sp, i + 3, w4.w);
    }
}

// (DP: at least 5 CTAs per SM, so that the whole grid -- 727 CTAs on the wide configuration -- is
// resident while its threads wait on their peers)
// LPG = lanes per float4 group in part A.  With one lane per group a group with 32 split-K slices needs four
// dependent rounds of eight loads (four L2 round trips: the tail of the whole step); with four lanes each lane
// sums a quarter of the slices in ONE round and the quarters are combined in a fixed shuffle tree (so the sum
// stays deterministic).  The data-parallel instantiation keeps one lane per group: its grid has to be
// co-resident (wide_p2p_resident).
template <bool FUSE, bool DP = false>
__global__ void __launch_bounds__(128, DP ? 5 : 0)
wide_update_kernel(Dims d, Offsets o, PartialSrc ps, UpdatePlan pl, float count, float *__restrict__ flat,
                   float *__restrict__ theta, float *__restrict__ m, float *__restrict__ v, AdamConsts ac,
                   const StepDev *__restrict__ step, ShadowPtrs sp, float *__restrict__ loss_out, PeerSet push,
                   unsigned int epoch)
{
    TRACE_MARKS_DECL
    BlockTrace trace_(TR_UPDATE, TRACE_MARKS_PTR);
    // push.world > 0 (data-parallel group over peer memory, FUSE only): the locally reduced gradient is
    // exchanged (dp_exchange*) and the update uses the global sum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) TRACE_MARK(0);
    if (FUSE && step) { ac.c1 = step->c1; ac.c2 = step->c2; }
    if (step) epoch = step->epoch;
    if (threadIdx.x == 0 && epoch != 0xffffffffu) TRACE_MARK(1);
    constexpr bool dp = FUSE && DP;      // (a separate instantiation: the exchange costs registers)
    if ((int)blockIdx.x < pl.blocksA) {
        // ---- A: weights ------------------------------------------------------------
        const int n1g = (d.D * d.H1) >> 2, n2g = (d.H1 * d.H2) >> 2, n3g = (d.H2 * d.A) >> 2;
        int bi = blockIdx.x, range = 0;
        if (bi >= pl.bA[0]) { bi -= pl.bA[0]; range = 1; }
        if (range == 1 && bi >= pl.bA[1]) { bi -= pl.bA[1]; range = 2; }
        const int lpg = pl.lp[range];
        int gl = (bi * 128 + threadIdx.x) / lpg;              // group within its range
        const int sub = threadIdx.x & (lpg - 1);
        const bool live = gl < (range == 0 ? n1g : (range == 1 ? n2g : n3g));   // (every lane stays for the shuffles below)
        if (!live) gl = 0;
        int i, n;
        const float *base;
        int64_t stride;
        if (range == 0) { i = o.w1 + gl * 4; base = ps.w1_slices + gl * 4; n = ps.n1; stride = (int64_t)d.D * d.H1; }
        else if (range == 1) { i = o.w2 + gl * 4; base = ps.w2_slices + gl * 4; n = ps.n2; stride = (int64_t)d.H1 * d.H2; }
        else { i = o.w3 + gl * 4; base = ps.w3_slices + gl * 4; n = ps.n3; stride = (int64_t)d.H2 * d.A; }
        float4 w4 = make_float4(0, 0, 0, 0), m4 = w4, v4 = w4;
        if (!live) n = 0;
        if (FUSE && sub == 0 && live) {
            w4 = *reinterpret_cast<const float4 *>(theta + i);
            m4 = *reinterpret_cast<const float4 *>(m + i);
            v4 = *reinterpret_cast<const float4 *>(v + i);
        }
        float4 g = make_float4(0, 0, 0, 0);
        for (int z = sub; z < n; z += 8 * lpg) {     // lane `sub` takes slices sub, sub + lpg, ...
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
                t[u] = (z + u * lpg < n) ? __ldg(reinterpret_cast<const float4 *>(base + (int64_t)(z + u * lpg) * stride))
                                         : make_float4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 8; u++) g = f4add(g, t[u]);
        }
        if (threadIdx.x == 0 && g.x == g.x) TRACE_MARK(2);
        for (int off = 1; off < lpg; off <<= 1) {    // fixed tree: the sum does not depend on timing
            g.x += __shfl_xor_sync(0xffffffffu, g.x, off);
            g.y += __shfl_xor_sync(0xffffffffu, g.y, off);
            g.z += __shfl_xor_sync(0xffffffffu, g.z, off);
            g.w += __shfl_xor_sync(0xffffffffu, g.w, off);
        }
        if (sub != 0 || !live) return;
        bool ok = true;     // false: a peer's word never arrived -- leave theta / m / v / shadows untouched
        if (dp) g = dp_exchange4(push, epoch, i, g, ok);
        if (dp && threadIdx.x == 0 && g.x == g.x) TRACE_MARK(4);
        if (ok) finish4<FUSE>(d, o, i, g, flat, theta, m, v, w4, m4, v4, ac, sp);
        if (threadIdx.x == 0) TRACE_MARK(3);
    } else if ((int)blockIdx.x < pl.blocksA + pl.blocksB) {
        // ---- B: biases of the hidden layers, warp per float4 group -------------------
        const int gidx = pl.gB0 + (blockIdx.x - pl.blocksA) * 4 + warp;
        if (gidx >= pl.gB0 + pl.nB) return;
        const int nb1g = d.H1 >> 2;
        int i, n;
        const float *base;
        int64_t stride;
        if (gidx < nb1g) { i = o.b1 + gidx * 4; base = ps.b1_part + gidx * 4; n = ps.nb1; stride = d.H1; }
        else { const int u = (gidx - nb1g) * 4; i = o.b2 + u; base = ps.b2_part + u; n = ps.nb2; stride = d.H2; }
        float4 g = make_float4(0, 0, 0, 0);
        for (int z = lane; z < n; z += 32 * 4) {
            float4 t[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                t[u] = (z + 32 * u < n) ? __ldg(reinterpret_cast<const float4 *>(base + (int64_t)(z + 32 * u) * stride))
                                        : make_float4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 4; u++) g = f4add(g, t[u]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            g.x += __shfl_xor_sync(0xffffffffu, g.x, off);
            g.y += __shfl_xor_sync(0xffffffffu, g.y, off);
            g.z += __shfl_xor_sync(0xffffffffu, g.z, off);
            g.w += __shfl_xor_sync(0xffffffffu, g.w, off);
        }
        if (lane == 0) {
            float4 w4 = make_float4(0, 0, 0, 0), m4 = w4, v4 = w4;
            if (FUSE) {
                w4 = *reinterpret_cast<const float4 *>(theta + i);
                m4 = *reinterpret_cast<const float4 *>(m + i);
                v4 = *reinterpret_cast<const float4 *>(v + i);
            }
            bool ok = true;
            if (dp) g = dp_exchange4(push, epoch, i, g, ok);
            if (ok) finish4<FUSE>(d, o, i, g, flat, theta, m, v, w4, m4, v4, ac, sp);
        }
    } else {
        // ---- C: b3 and the loss, warp per scalar ---------------------------------------
        const int e = (blockIdx.x - pl.blocksA - pl.blocksB) * 4 + warp;
        if (e >= pl.nC) return;
        const int i = o.b3 + e;
        const float *base = ps.row_part + e;
        float g = 0.f;
        for (int z = lane; z < ps.nrow; z += 32) g += __ldg(base + (int64_t)z * (d.A + 1));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) g += __shfl_xor_sync(0xffffffffu, g, off);
        if (lane == 0) {
            if (i == o.P) g = __fdiv_rn(g, count);
            bool ok = true;
            if (dp) g = dp_exchange1(push, epoch, i, g, ok);
            if (i == o.P && loss_out) *loss_out = g;
            flat[i] = g;
            if (FUSE && i < o.P && ok) {
                float wi = theta[i], mi = m[i], vi = v[i];
                solver_update(wi, g, mi, vi, ac);
                theta[i] = wi;
                m[i] = mi;
                v[i] = vi;
            }
        }
    }
}

template <typename T>
bool dmalloc(T **p, size_t n, std::string &err)
{
    cudaError_t e = cudaMalloc((void **)p, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) { err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return false; }
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------
// WideState
// ---------------------------------------------------------------------------
struct WideNet {
    __nv_bfloat16 *w1 = nullptr, *w2 = nullptr, *w3p = nullptr, *w3t = nullptr;
    CUtensorMap tm_w1_mn, tm_w2_mn, tm_w3p_mn, tm_w2_k;
    CUtensorMap tm_w1_mn_s4, tm_w2_mn_s4, tm_w2_k_s4;   // multicast slice boxes (clusters of 4)
    CUtensorMap tm_w2_k256;                              // K-major B operand, 256-row box
};

struct WideState {
    Dims d;
    Offsets o;
    int B = 0, sm_count = 148;
    WideNet net[2];
    // gathered batch, double buffered: inside a multi-step graph the gather of step i+1 runs on
    // a side branch while step i computes (WideStep::set / have / next_step)
    struct BatchSet {
        __nv_bfloat16 *x = nullptr, *x2 = nullptr;
        float *br = nullptr;
        int32_t *ba = nullptr;
        uint8_t *bdone = nullptr;
        CUtensorMap tm_x_a, tm_x2_a, tm_x_mm;
    } bs[2];
    int last_set = 0;
    BatchSet pred;                         // predict: bf16 states [B][D] (tm_x_a, tm_x2_a both describe it)
    float *pred_q = nullptr;               // predict: Q [B][A]
    __nv_bfloat16 *h1 = nullptr, *h2 = nullptr, *t1 = nullptr, *t2 = nullptr;
    __nv_bfloat16 *dz2 = nullptr, *dz1 = nullptr, *dqT = nullptr;
    uint32_t *m1bits = nullptr, *m2bits = nullptr;      // Rectify masks of the online network as bits (fused pair forward only)
    float *q = nullptr, *qt = nullptr, *td = nullptr, *dqv = nullptr;
    float *w1_slices = nullptr, *w2_slices = nullptr, *w3_slices = nullptr, *b1_part = nullptr, *b2_part = nullptr,
          *row_part = nullptr;
    int ns1 = 1, ns2 = 1, ns3 = 1, kps1 = 0, kps2 = 0, kps3 = 0, nb1 = 1, nb2 = 1, nrow = 1;
    // activation tensor maps: KK A-operands (box 128 rows) and MN-major operands (box 64 k-rows)
    CUtensorMap tm_h1_a, tm_t1_a, tm_h2_a, tm_t2_a, tm_dz2_a, tm_dz1_a;
    CUtensorMap tm_h1_mm, tm_h2_mm, tm_dz2_mm, tm_dz1_mm, tm_dqT_b, tm_dz2_mm_s4;
    // side streams: independent backward GEMMs run concurrently (fork/join by events; under
    // stream capture these become parallel branches of the step graph)
    cudaStream_t s1 = nullptr, s2 = nullptr, s3 = nullptr;
    cudaEvent_t ev_dz2 = nullptr, ev_dx = nullptr, ev_s1 = nullptr, ev_s2 = nullptr, ev_fork = nullptr, ev_gather = nullptr;
    bool pdl = true;
    bool ff = false;                       // fused three-layer forward kernel usable
    bool pair_dx = false, pair_dw2 = false;   // backward GEMMs on CTA pairs (pair_gemm_kernel); pair_dx also produces the db1 partials
    int skip = 0;                          // GOLDQ_SKIP bit mask (diagnostic): launches left out of the step, see wide_step
    long long *ff_stamps = nullptr;        // GOLDQ_FF_STAMPS: in-kernel clock stamps of the fused forward (diagnostic)
    unsigned long long *trace = nullptr;   // GOLDQ_TRACE: block log of the step's kernels (BlockTrace, diagnostic)
    bool ff2 = false;                      // ... on CTA pairs (cta_group::2): needs an even number of 128-row slabs
    bool bn256 = false;                    // 128x256 output tiles for the N = H GEMMs
    std::vector<const char *> prof_names;  // labels of the launches of the last profiled step
    bool mc_rows = false, mc_h1 = false;   // B-tile multicast usable: row tiles / H1 tiles divide by 4
    bool debug = false;
};

static int split_for(int tiles, int K, int sm_count, int *kps)
{
    // enough CTAs to cover the SMs once, k ranges in multiples of BK
    int want = (sm_count + tiles - 1) / tiles;
    int kblocks = (K + BK - 1) / BK;
    if (want > kblocks) want = kblocks;
    if (want < 1) want = 1;
    int per = (kblocks + want - 1) / want;
    *kps = per * BK;
    return (kblocks + per - 1) / per;
}

WideState *wide_create(const Dims &d, int B, int sm_count, std::string &err)
{
    if ((d.D % 8) || (d.H1 % 8) || (d.H2 % 8)) { err = "bf16 path needs D, H1, H2 to be multiples of 8 (16-byte TMA pitch)"; return nullptr; }
    if (B % 8) { err = "bf16 path needs batch_size to be a multiple of 8 (16-byte pitch of dQ^T)"; return nullptr; }
    if (d.A > 32) { err = "bf16 path supports up to 32 actions"; return nullptr; }
    if (!load_encode(err)) return nullptr;
    {
        cudaError_t e = configure_all_gemms();
        // The small kernels INSIDE the step (TD/dZ2, gather) ask for the shared-memory carve-out of the GEMMs around them.
        // An SM re-partitions L1 / shared memory only when it is idle: with the default carve-out the TD blocks that did not
        // fit next to the forward's CTAs started a full wave late (block trace: TD ended 8.9 us after the forward instead of
        // 6.0).  NOT the update kernel: with the same carve-out the next forward's CTAs become resident next to its last
        // blocks and take registers from the data-parallel instantiation, whose grid must stay co-resident for the exchange
        // (N=2: 56.9 vs 53.9 us/step).  GOLDQ_CARVEOUT=0: none; =2: the update kernels too.
        const int cv = getenv("GOLDQ_CARVEOUT") ? atoi(getenv("GOLDQ_CARVEOUT")) : 1;
        const int mx = (int)cudaSharedmemCarveoutMaxShared;
        if (cv >= 1) {
            if (e == cudaSuccess) e = cudaFuncSetAttribute(wide_td_dz2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(wide_gather_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        }
        if (cv >= 2) {
            if (e == cudaSuccess) e = cudaFuncSetAttribute(wide_update_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(wide_update_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(wide_update_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        }
        if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return nullptr; }
    }
    WideState *w = new WideState();
    w->d = d;
    w->o = make_offsets(d);
    w->B = B;
    w->sm_count = sm_count;
    bool ok = true;
    for (int n = 0; n < 2; n++) {
        ok = ok && dmalloc(&w->net[n].w1, (size_t)d.D * d.H1, err) && dmalloc(&w->net[n].w2, (size_t)d.H1 * d.H2, err) &&
             dmalloc(&w->net[n].w3p, (size_t)d.H2 * W3_PAD, err);
        if (ok) cudaMemset(w->net[n].w3p, 0, (size_t)d.H2 * W3_PAD * 2);
    }
    ok = ok && dmalloc(&w->net[0].w3t, (size_t)32 * d.H2, err);
    if (ok) cudaMemset(w->net[0].w3t, 0, (size_t)32 * d.H2 * 2);
    size_t n = B;
    for (int k = 0; k < 2; k++)
        ok = ok && dmalloc(&w->bs[k].x, n * d.D, err) && dmalloc(&w->bs[k].x2, n * d.D, err) && dmalloc(&w->bs[k].br, n, err) &&
             dmalloc(&w->bs[k].ba, n, err) && dmalloc(&w->bs[k].bdone, n, err);
    ok = ok && dmalloc(&w->pred.x, n * d.D, err) && dmalloc(&w->pred_q, n * d.A, err);
    if (ok) cudaMemset(w->pred.x, 0, n * d.D * 2);
    ok = ok && dmalloc(&w->h1, n * d.H1, err) &&
         dmalloc(&w->h2, n * d.H2, err) && dmalloc(&w->t1, n * d.H1, err) && dmalloc(&w->t2, n * d.H2, err) &&
         dmalloc(&w->dz2, n * d.H2, err) && dmalloc(&w->dz1, n * d.H1, err) && dmalloc(&w->dqT, n * 32, err) &&
         dmalloc(&w->m1bits, n * (d.H1 / 32 + 1), err) && dmalloc(&w->m2bits, n * (d.H2 / 32 + 1), err) &&
         dmalloc(&w->q, n * d.A, err) && dmalloc(&w->qt, n * d.A, err) && dmalloc(&w->td, n, err) &&
         dmalloc(&w->dqv, n, err);
    if (!ok) { wide_destroy(w); return nullptr; }
    int tiles1 = ((d.D + BM - 1) / BM) * ((d.H1 + 127) / 128);
    int tiles2 = ((d.H1 + BM - 1) / BM) * ((d.H2 + 127) / 128);
    int tiles3 = (d.H2 + BM - 1) / BM;
    // CTAs each split-K GEMM aims for (GOLDQ_CTAS1/2/3: tuning knobs).  dW2 (the big one): one per SM.  dW1 and dW3 are
    // small GEMMs with few output tiles: a CTA per SM meant 32 split-K slices each, i.e. four dependent rounds of loads
    // in the update kernel's one-lane (data-parallel) form; 64 / 32 CTAs (16 / 8 slices on the wide configuration) cost
    // the GEMMs nothing measurable (49.8 vs 50.6 us/step) and halve / quarter what the update kernel reads.
    auto ctas_for = [&](const char *name, int dflt) { const char *e = getenv(name); return e && atoi(e) > 0 ? atoi(e) : dflt; };
    w->ns1 = split_for(tiles1, B, ctas_for("GOLDQ_CTAS1", sm_count * 7 / 16), &w->kps1);
    w->ns2 = split_for(tiles2, B, ctas_for("GOLDQ_CTAS2", sm_count), &w->kps2);
    w->ns3 = split_for(tiles3, B, ctas_for("GOLDQ_CTAS3", sm_count * 7 / 32), &w->kps3);
    w->nb1 = (B + CS_ROWS - 1) / CS_ROWS;
    const bool no_pair = getenv("GOLDQ_NO_PAIR_GEMM") != nullptr;
    // dX on CTA pairs: 256 x 256 tiles need whole 256-row / 256-column tiles; it writes one db1 partial per 128-row slab
    w->pair_dx = !no_pair && getenv("GOLDQ_NO_PAIR_DX") == nullptr && B % 256 == 0 && d.H1 % 256 == 0 && d.H2 % 64 == 0;
    if (w->pair_dx) w->nb1 = B / BM;
    // dW2 on CTA pairs: (H1/256) x (H2/256) pair tiles, split-K so that about one CTA lands on every SM
    // (opt-in: measured no faster inside the step -- 64.3 vs 63.4 us -- and its 16 split-K slices are 16 MB against 10 MB)
    w->pair_dw2 = !no_pair && getenv("GOLDQ_PAIR_DW2") != nullptr && d.H1 % 256 == 0 && d.H2 % 256 == 0 && B % 64 == 0;
    if (w->pair_dw2) {
        // never more CTAs than SMs (one CTA per SM: a few CTAs left for a second wave would double the time)
        const int tiles = (d.H1 / BM) * (d.H2 / 256), kblocks = B / BK;
        int want = sm_count / tiles;
        if (want < 1) want = 1;
        if (want > kblocks) want = kblocks;
        const int per = (kblocks + want - 1) / want;
        w->kps2 = per * BK;
        w->ns2 = (kblocks + per - 1) / per;
    }
    w->nb2 = (B + DZ_ROWS - 1) / DZ_ROWS;
    w->nrow = w->nb2;   // db3 / loss partials come from the same blocks as db2
    ok = ok && dmalloc(&w->w1_slices, (size_t)w->ns1 * d.D * d.H1, err) &&
         dmalloc(&w->w2_slices, (size_t)w->ns2 * d.H1 * d.H2, err) &&
         dmalloc(&w->w3_slices, (size_t)w->ns3 * d.H2 * d.A, err) && dmalloc(&w->b1_part, (size_t)w->nb1 * d.H1, err) &&
         dmalloc(&w->b2_part, (size_t)w->nb2 * d.H2, err) && dmalloc(&w->row_part, (size_t)w->nrow * (d.A + 1), err);
    if (!ok) { wide_destroy(w); return nullptr; }
    // weight maps.  MN-major B operands: [K rows][N cols], box 64 k-rows x 64 n
    for (int k = 0; k < 2; k++) {
        WideNet &nn = w->net[k];
        ok = ok && make_tmap(&nn.tm_w1_mn, nn.w1, d.D, d.H1, d.H1, 64, err) &&
             make_tmap(&nn.tm_w2_mn, nn.w2, d.H1, d.H2, d.H2, 64, err) &&
             make_tmap(&nn.tm_w3p_mn, nn.w3p, d.H2, W3_PAD, W3_PAD, 64, err);
    }
    // W2 as the K-major B operand of dX: [N = H1 rows][K = H2 cols], box 128 rows
    ok = ok && make_tmap(&w->net[0].tm_w2_k, w->net[0].w2, d.H1, d.H2, d.H2, 128, err);
    // slice boxes for B-tile multicast over clusters of 4 CTAs
    for (int k = 0; k < 2; k++) {
        WideNet &nn = w->net[k];
        ok = ok && make_tmap(&nn.tm_w1_mn_s4, nn.w1, d.D, d.H1, d.H1, 16, err) &&
             make_tmap(&nn.tm_w2_mn_s4, nn.w2, d.H1, d.H2, d.H2, 16, err);
    }
    ok = ok && make_tmap(&w->net[0].tm_w2_k_s4, w->net[0].w2, d.H1, d.H2, d.H2, 32, err);
    ok = ok && make_tmap(&w->net[0].tm_w2_k256, w->net[0].w2, d.H1, d.H2, d.H2, 256, err);
#ifdef GOLDQ_TRACE_BUILD
    if (getenv("GOLDQ_TRACE")) {
        const size_t n = 4 + (size_t)TRACE_CAP * (sizeof(TraceRec) / 8);
        if (dmalloc(&w->trace, n, err)) {
            cudaMemset(w->trace, 0, n * 8);
            cudaMemcpyToSymbol(c_trace, &w->trace, sizeof(w->trace));
        }
    }
#endif
    w->ff = getenv("GOLDQ_NO_FF") == nullptr && d.H1 == FF_H && d.H2 == FF_H && (d.D == 64 || d.D == 128) && d.A <= W3_PAD;
    w->ff2 = w->ff && getenv("GOLDQ_NO_FF2") == nullptr && B % (2 * BM) == 0;
    // diagnostic only (results are garbage): leave launches out of the step to read their share of the
    // critical path off the step time.  1 gather, 2 forward, 4 TD/dZ2, 8 dW3, 16 dW2, 32 dX, 64 db1, 128 dW1, 256 update
    if (const char *sk = getenv("GOLDQ_SKIP")) w->skip = atoi(sk);
    if (w->ff2 && getenv("GOLDQ_FF_STAMPS")) {
        const size_t n = (size_t)(B / BM) * 2 * FF_NSTAMPS;
        if (dmalloc(&w->ff_stamps, n, err)) cudaMemset(w->ff_stamps, 0, n * sizeof(long long));
    }
    w->bn256 = getenv("GOLDQ_NO_BN256") == nullptr && d.H1 % 256 == 0 && d.H2 % 256 == 0;   // 82.3 -> 79.7 us/step
    // measured on B200 (round 1): with 128x128 tiles the cluster's lock-step costs more than the
    // saved L2->SM traffic (95 vs 87 us/step), so B-tile multicast is opt-in
    w->mc_rows = getenv("GOLDQ_MULTICAST") != nullptr && ((B + BM - 1) / BM) % 4 == 0;
    w->mc_h1 = getenv("GOLDQ_MULTICAST") != nullptr && ((d.H1 + BM - 1) / BM) % 4 == 0;
    // activations as KK A operands: [B rows][K cols], box 128 rows
    for (int k = 0; k < 2; k++)
        ok = ok && make_tmap(&w->bs[k].tm_x_a, w->bs[k].x, B, d.D, d.D, BM, err) &&
             make_tmap(&w->bs[k].tm_x2_a, w->bs[k].x2, B, d.D, d.D, BM, err) &&
             make_tmap(&w->bs[k].tm_x_mm, w->bs[k].x, B, d.D, d.D, 64, err);
    ok = ok && make_tmap(&w->pred.tm_x_a, w->pred.x, B, d.D, d.D, BM, err) && make_tmap(&w->pred.tm_x2_a, w->pred.x, B, d.D, d.D, BM, err);
    ok = ok && make_tmap(&w->tm_h1_a, w->h1, B, d.H1, d.H1, BM, err) && make_tmap(&w->tm_t1_a, w->t1, B, d.H1, d.H1, BM, err) &&
         make_tmap(&w->tm_h2_a, w->h2, B, d.H2, d.H2, BM, err) && make_tmap(&w->tm_t2_a, w->t2, B, d.H2, d.H2, BM, err) &&
         make_tmap(&w->tm_dz2_a, w->dz2, B, d.H2, d.H2, BM, err) && make_tmap(&w->tm_dz1_a, w->dz1, B, d.H1, d.H1, BM, err);
    // activations as MN-major operands: [K = B rows][MN cols], box 64 k-rows x 64 mn
    ok = ok && make_tmap(&w->tm_h1_mm, w->h1, B, d.H1, d.H1, 64, err) &&
         make_tmap(&w->tm_h2_mm, w->h2, B, d.H2, d.H2, 64, err) && make_tmap(&w->tm_dz2_mm, w->dz2, B, d.H2, d.H2, 64, err) &&
         make_tmap(&w->tm_dz1_mm, w->dz1, B, d.H1, d.H1, 64, err) &&
         make_tmap(&w->tm_dz2_mm_s4, w->dz2, B, d.H2, d.H2, 16, err);
    // dQ^T [32 rows][B cols]: K-major B operand of the dW3 GEMM, box 32 rows
    ok = ok && make_tmap(&w->tm_dqT_b, w->dqT, 32, B, B, 32, err);
    if (!ok) { wide_destroy(w); return nullptr; }
    w->pdl = getenv("GOLDQ_NO_PDL") == nullptr;
    {
        // GOLDQ_PRIO (experiment): the side branches (dW2 | dW3 | next step's gather) at the lowest priority, so that the
        // block scheduler prefers the critical chain dX -> dW1 on the learn stream when slots free up
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const int side = getenv("GOLDQ_PRIO") ? lo : 0;
        cudaStreamCreateWithPriority(&w->s1, cudaStreamNonBlocking, side);
        cudaStreamCreateWithPriority(&w->s2, cudaStreamNonBlocking, side);
        cudaStreamCreateWithPriority(&w->s3, cudaStreamNonBlocking, side);
    }
    cudaEvent_t *evs[] = {&w->ev_dz2, &w->ev_dx, &w->ev_s1, &w->ev_s2, &w->ev_fork, &w->ev_gather};
    for (cudaEvent_t *e : evs) cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    return w;
}

void wide_destroy(WideState *w)
{
    if (!w) return;
    for (int n = 0; n < 2; n++) {
        void *p[] = {w->net[n].w1, w->net[n].w2, w->net[n].w3p, w->net[n].w3t};
        for (void *q : p)
            if (q) cudaFree(q);
    }
    for (int k = 0; k < 2; k++) {
        void *p[] = {w->bs[k].x, w->bs[k].x2, w->bs[k].br, w->bs[k].ba, w->bs[k].bdone};
        for (void *q : p)
            if (q) cudaFree(q);
    }
    void *p[] = {w->pred.x, w->pred_q, w->ff_stamps, w->m1bits, w->m2bits, w->trace, w->h1, w->h2, w->t1, w->t2, w->dz2, w->dz1, w->dqT, w->q, w->qt, w->td, w->dqv, w->w1_slices, w->w2_slices, w->w3_slices, w->b1_part, w->b2_part, w->row_part};
    for (void *q : p)
        if (q) cudaFree(q);
    if (w->s1) cudaStreamDestroy(w->s1);
    if (w->s2) cudaStreamDestroy(w->s2);
    if (w->s3) cudaStreamDestroy(w->s3);
    cudaEvent_t evs[] = {w->ev_dz2, w->ev_dx, w->ev_s1, w->ev_s2, w->ev_fork, w->ev_gather};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    delete w;
}

static ShadowPtrs shadow_ptrs(WideState *w, int net)
{
    WideNet &n = w->net[net];
    return ShadowPtrs{n.w1, n.w2, n.w3p, n.w3t};
}

static int shadow(WideState *w, int net, const float *theta, cudaStream_t st, std::string &err)
{
    wide_shadow_kernel<<<(w->o.P + 255) / 256, 256, 0, st>>>(theta, w->d, w->o, shadow_ptrs(w, net));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("wide_shadow_kernel: ") + cudaGetErrorString(e); return -1; }
    return 1;
}

int wide_set_params(WideState *w, int net, const float *theta_dev, cudaStream_t st, std::string &err)
{
    return shadow(w, net, theta_dev, st, err);
}

int wide_after_update(WideState *w, const float *theta_dev, cudaStream_t st, std::string &err)
{
    return shadow(w, 0, theta_dev, st, err);
}

// data-parallel tail: Adam on the all-reduced flat gradient + bf16 shadow refresh, one pass
__global__ void __launch_bounds__(256)
wide_apply_kernel(Dims d, Offsets o, const float *__restrict__ flat, float *__restrict__ theta, float *__restrict__ m,
                  float *__restrict__ v, AdamConsts ac, const StepDev *__restrict__ step, ShadowPtrs sp)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (step) { ac.c1 = step->c1; ac.c2 = step->c2; }
    if (i + 3 < o.P && (i + 3 < o.b1 || (i >= o.b1 && i + 3 < o.w2) || (i >= o.w2 && i + 3 < o.b2) ||
                        (i >= o.b2 && i + 3 < o.w3) || (i >= o.w3 && i + 3 < o.b3))) {
        const float4 g = *reinterpret_cast<const float4 *>(flat + i);
        float4 w4 = *reinterpret_cast<const float4 *>(theta + i);
        float4 m4 = *reinterpret_cast<const float4 *>(m + i);
        float4 v4 = *reinterpret_cast<const float4 *>(v + i);
        finish4<true>(d, o, i, g, nullptr, theta, m, v, w4, m4, v4, ac, sp);
    } else {
        for (int e = i; e < i + 4 && e < o.P; e++) {
            float wi = theta[e], mi = m[e], vi = v[e];
            solver_update(wi, flat[e], mi, vi, ac);
            theta[e] = wi; m[e] = mi; v[e] = vi;
            shadow_store(d, o, sp, e, wi);
        }
    }
}

int wide_apply_update(WideState *w, const float *flat, float *theta, float *m, float *v, AdamConsts ac,
                      const StepDev *step, cudaStream_t st, std::string &err)
{
    const int groups = (w->o.P + 3) / 4;
    wide_apply_kernel<<<(groups + 255) / 256, 256, 0, st>>>(w->d, w->o, flat, theta, m, v, ac, step, shadow_ptrs(w, 0));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("wide_apply_kernel: ") + cudaGetErrorString(e); return -1; }
    return 1;
}

// many_lanes: ranges with more than 8 split-K slices get 2 or 4 lanes per group (one round of loads per lane instead of
// up to four dependent rounds: the block trace showed the 32-slice groups of W1 and W3 finishing 5 us after the 10-slice
// groups of W2 -- the tail of the whole step).  The data-parallel instantiation keeps one lane per group: its grid has
// to stay co-resident (wide_p2p_resident).
// part 0: every parameter in one launch.  The single-GPU step splits the update in two launches that join different
// branches of the backward pass: part 1 = W1, b1 (behind dW1 and dX), part 2 = W2, b2, W3, b3 and the loss (behind dW2,
// dW3 and TD/dZ2) -- a range with no blocks simply drops out of the kernel's block -> range arithmetic.
static UpdatePlan make_update_plan(const Dims &d, int n1 = 1, int n2 = 1, int n3 = 1, bool many_lanes = false, int part = 0)
{
    UpdatePlan pl;
    const int ng[3] = {(d.D * d.H1) >> 2, (d.H1 * d.H2) >> 2, (d.H2 * d.A) >> 2};
    const int ns[3] = {n1, n2, n3};
    pl.nA = ng[0] + ng[1] + ng[2];
    pl.nB = (part == 1 ? d.H1 : (part == 2 ? d.H2 : d.H1 + d.H2)) >> 2;
    pl.gB0 = part == 2 ? (d.H1 >> 2) : 0;
    pl.nC = part == 1 ? 0 : d.A + 1;
    pl.blocksA = 0;
    for (int r = 0; r < 3; r++) {
        pl.lp[r] = !many_lanes ? 1 : (ns[r] > 16 ? 4 : (ns[r] > 10 ? 2 : 1));
        const bool in = part == 0 || (part == 1) == (r == 0);
        pl.bA[r] = in ? (ng[r] * pl.lp[r] + 127) / 128 : 0;
        pl.blocksA += pl.bA[r];
    }
    pl.blocksB = (pl.nB + 3) / 4;
    pl.blocksC = (pl.nC + 3) / 4;
    return pl;
}

// The in-kernel exchange spin-waits on words its peers' CTAs of the same index produce, so every
// CTA of the update grid must be resident at once.  *grid = CTAs the update kernel launches for
// this shape, *resident = CTAs the device holds concurrently (occupancy x SMs).
bool wide_p2p_resident(const WideState *w, int *grid, int *resident)
{
    const UpdatePlan pl = make_update_plan(w->d);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wide_update_kernel<true, true>, 128, 0) != cudaSuccess) {
        cudaGetLastError();
        per_sm = 0;
    }
    *grid = pl.blocksA + pl.blocksB + pl.blocksC;
    *resident = per_sm * w->sm_count;
    return *grid <= *resident;
}

// floats owned per rank: an equal split of [P + 1] rounded up to whole 256-group blocks
int wide_p2p_chunk_floats(const WideState *w, int world)
{
    const int per = (w->o.P + 1 + world - 1) / world;
    return (per + 1023) / 1024 * 1024;
}

int wide_sync_target(WideState *w, cudaStream_t st, std::string &err)
{
    // target <- online: copy the bf16 shadows (W3^T is online-only)
    const Dims &d = w->d;
    cudaError_t e = cudaMemcpyAsync(w->net[1].w1, w->net[0].w1, (size_t)d.D * d.H1 * 2, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->net[1].w2, w->net[0].w2, (size_t)d.H1 * d.H2 * 2, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->net[1].w3p, w->net[0].w3p, (size_t)d.H2 * W3_PAD * 2, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { err = std::string("wide_sync_target: ") + cudaGetErrorString(e); return -1; }
    return 3;
}

#define WCK(expr)                                                                    \
    do {                                                                             \
        cudaError_t e_ = (expr);                                                     \
        if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return -1; } \
    } while (0)

template <int KIND>
static cudaError_t launch_pair_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &tc, const PairGemmArgs &g,
                                    int nsplit, cudaStream_t st, bool pdl)
{
    dim3 grid((g.M + BM - 1) / BM, (g.N + 255) / 256, nsplit);
    if (grid.x & 1) return cudaErrorInvalidConfiguration;       // pairs along M
    return launch_k(pair_gemm_kernel<KIND>, grid, dim3(PG_THREADS), pair_gemm_smem_bytes(), st, pdl, ta, tb, tc, g);
}

// the fused forward of both networks on batch set `bt` (pair kernel when the slab count is even)
static cudaError_t launch_fused_forward(WideState *w, WideState::BatchSet &bt, const float *theta, const float *theta_t,
                                        cudaStream_t st, bool pdl, int only_net = -1, float *q_out = nullptr, bool x_ready = false)
{
    // only_net < 0: the learn step's launch (both networks, online activations kept); else one network of a
    // predict launch (bt.tm_x_a = the states, Q -> q_out)
    const Dims &d = w->d;
    const Offsets &o = w->o;
    const int B = w->B;
    WideNet &on = w->net[0], &tg = w->net[1];
    FFArgs fa{};
    fa.B = B; fa.D = d.D; fa.A = d.A;
    fa.bias[0][0] = theta + o.b1; fa.bias[0][1] = theta + o.b2; fa.bias[0][2] = theta + o.b3;
    fa.bias[1][0] = theta_t + o.b1; fa.bias[1][1] = theta_t + o.b2; fa.bias[1][2] = theta_t + o.b3;
    fa.q[0] = w->q; fa.q[1] = w->qt;
    fa.stamps = only_net < 0 ? w->ff_stamps : nullptr;
    static const bool early = getenv("GOLDQ_EARLY_TRIGGER") != nullptr;      // diagnostic: round-2's first placement
    fa.late_trigger = early ? 0 : 1;
    fa.net0 = only_net < 0 ? 0 : only_net;
    // (GOLDQ_FF_KEEPH=0, diagnostic: the activations are not stored -- results garbage)
    static const int keeph = getenv("GOLDQ_FF_KEEPH") ? atoi(getenv("GOLDQ_FF_KEEPH")) : 1;
    fa.keep_h = only_net < 0 ? keeph : 0;
    static const bool no_xpre = getenv("GOLDQ_NO_XPREFETCH") != nullptr;
    fa.x_ready = (x_ready && !no_xpre) ? 1 : 0;
    fa.maskbits[0] = w->m1bits; fa.maskbits[1] = w->m2bits;
    if (only_net >= 0) fa.q[only_net] = q_out;
    dim3 grid((B + BM - 1) / BM, only_net < 0 ? 2 : 1);
    // (8 epilogue warps.  16 -- four column groups of 64, 96 registers -- measured SLOWER, 15.9 vs 14.7 us: a phase drains
    // 128 KB of accumulator through the 64 B/clk TMEM read path, 2 048 cycles whatever the number of warps)
    if (w->ff2 && d.D == 128)
        return launch_k(ff_forward2_kernel<2, 8>, grid, dim3(64 + 32 * 8), ff2_smem_bytes<2>(), st, pdl, bt.tm_x_a, bt.tm_x2_a,
                        on.tm_w1_mn, tg.tm_w1_mn, on.tm_w2_mn, tg.tm_w2_mn, on.tm_w3p_mn, tg.tm_w3p_mn, w->tm_h1_a, w->tm_h2_a, fa);
    if (w->ff2)
        return launch_k(ff_forward2_kernel<1, 8>, grid, dim3(64 + 32 * 8), ff2_smem_bytes<1>(), st, pdl, bt.tm_x_a, bt.tm_x2_a,
                        on.tm_w1_mn, tg.tm_w1_mn, on.tm_w2_mn, tg.tm_w2_mn, on.tm_w3p_mn, tg.tm_w3p_mn, w->tm_h1_a, w->tm_h2_a, fa);
    if (d.D == 128)
        return launch_k(ff_forward_kernel<2>, grid, dim3(FF_THREADS), ff_smem_bytes<2>(), st, pdl, bt.tm_x_a, bt.tm_x2_a,
                        on.tm_w1_mn, tg.tm_w1_mn, on.tm_w2_mn, tg.tm_w2_mn, on.tm_w3p_mn, tg.tm_w3p_mn, w->tm_h1_a, w->tm_h2_a, fa);
    return launch_k(ff_forward_kernel<1>, grid, dim3(FF_THREADS), ff_smem_bytes<1>(), st, pdl, bt.tm_x_a, bt.tm_x2_a,
                    on.tm_w1_mn, tg.tm_w1_mn, on.tm_w2_mn, tg.tm_w2_mn, on.tm_w3p_mn, tg.tm_w3p_mn, w->tm_h1_a, w->tm_h2_a, fa);
}

__global__ void __launch_bounds__(256)
wide_cast_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ out, long long n4)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2 *>(out)[i] = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
}

// Sequential.PredictBatch / batched acting on the tensor cores: Q = net(x) for n rows of fp32 states in device
// memory, through the SAME fused bf16 forward kernel (and bf16 weight shadows) the learn step trains with, in
// chunks of the handle's batch size.  Returns launches, or -1 (err set), or 0 when the fused forward does not
// exist for this shape (the caller then uses the generic fp32 kernels).
int wide_predict(WideState *w, int net, const float *x, long long n, const float *theta, const float *theta_t, float *q,
                 cudaStream_t st, std::string &err)
{
    if (!w->ff) return 0;
    const Dims &d = w->d;
    int launches = 0;
    for (long long r0 = 0; r0 < n; r0 += w->B) {
        const long long rows = (n - r0 < w->B) ? n - r0 : w->B;
        const long long n4 = rows * d.D / 4;
        wide_cast_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x + r0 * d.D, w->pred.x, n4);
        WCK(cudaGetLastError());
        WCK(launch_fused_forward(w, w->pred, theta, theta_t, st, false, net, w->pred_q));
        WCK(cudaMemcpyAsync(q + r0 * d.A, w->pred_q, (size_t)rows * d.A * sizeof(float), cudaMemcpyDeviceToDevice, st));
        launches += 2;
    }
    return launches;
}

// Diagnostic (bench.py roofline): average duration of the fused forward over `reps` back-to-back launches on
// `st` (inputs = the batch of the last step), CUDA events around the whole sequence.  Returns -1 when the fused
// forward is not in use for this shape.
int wide_time_forward(WideState *w, const float *theta, const float *theta_t, cudaStream_t st, int reps, float *us,
                      std::string &err)
{
    if (!w->ff) { err = "the fused forward kernel is not used for this shape"; return -1; }
    cudaEvent_t e0, e1;
    WCK(cudaEventCreate(&e0));
    WCK(cudaEventCreate(&e1));
    WideState::BatchSet &bt = w->bs[w->last_set];
    // (the batch was gathered long before: as in a graph-replayed step, its tile may be requested ahead of the wait)
    for (int i = 0; i < 3; i++) WCK(launch_fused_forward(w, bt, theta, theta_t, st, w->pdl, -1, nullptr, true));
    WCK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; i++) WCK(launch_fused_forward(w, bt, theta, theta_t, st, w->pdl, -1, nullptr, true));
    WCK(cudaEventRecord(e1, st));
    WCK(cudaEventSynchronize(e1));
    float ms = 0.f;
    WCK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *us = ms * 1e3f / (float)reps;
    return reps + 3;
}

int wide_step(WideState *w, const WideStep &s, cudaStream_t st, std::string &err)
{
    const Dims &d = w->d;
    const Offsets &o = w->o;
    const int B = w->B;
    int launches = 0;
    const bool prof = s.prof_evs != nullptr;
    const bool pdl = w->pdl && !prof;   // programmatic dependent launch along the st chain
    static const bool serial = getenv("GOLDQ_SERIAL") != nullptr;   // diagnostic: no parallel backward branches
    const bool one_stream = prof || serial;
    cudaStream_t s1 = one_stream ? st : w->s1, s2 = one_stream ? st : w->s2;
    if (prof) w->prof_names.clear();
    auto mark = [&](const char *name) {
        if (prof) { cudaEventRecord(s.prof_evs[(*s.prof_n)++], st); w->prof_names.push_back(name); }
    };
    WideState::BatchSet &bt = w->bs[s.set & 1];
    w->last_set = s.set & 1;
    const int gather_threads = B * (d.D >> 2);
    if (!s.have && !(w->skip & 1)) {
        WCK(launch_k(wide_gather_kernel, dim3((gather_threads + 255) / 256), dim3(256), 0, st, pdl, s.rp, d.D, s.idx, s.seed,
                     s.step, s.idx_out, B, bt.x, bt.x2, bt.ba, bt.br, bt.bdone));
        launches++;
        mark("gather+cast (K1)");
    } else if (s.have && !(w->skip & 1)) {
        WCK(cudaStreamWaitEvent(st, w->ev_gather, 0));   // this step's batch was gathered by the previous step's side branch
    }
    if (s.next_step && !(w->skip & 1)) {
        // gather the NEXT step's batch into the other buffer set while this step computes.  The fork
        // follows everything the previous step launched, i.e. every reader of that buffer set.
        WideState::BatchSet &nx = w->bs[(s.set & 1) ^ 1];
        WCK(cudaEventRecord(w->ev_fork, st));
        WCK(cudaStreamWaitEvent(w->s3, w->ev_fork, 0));
        // (a few big blocks that ask for 100 KB of shared memory, so that they land only on the SMs the forward leaves
        // idle, did take the gather off the forward's SMs -- forward 19.3 -> 18.3 us -- but needed 35-53 us themselves:
        // random 512-byte rows want hundreds of thousands of threads in flight, not 40 k.  Measured 51.9 / 62.0 vs 49.5 us/step.)
        wide_gather_kernel<<<(gather_threads + 255) / 256, 256, 0, w->s3>>>(s.rp, d.D, nullptr, 0, s.next_step, s.idx_out, B,
                                                                            nx.x, nx.x2, nx.ba, nx.br, nx.bdone);
        WCK(cudaGetLastError());
        WCK(cudaEventRecord(w->ev_gather, w->s3));
        launches++;
    }
    WideNet &on = w->net[0], &tg = w->net[1];
    GemmArgs g{};
    g.trace_id = TR_FORWARD;
    if (w->skip & 2) {
    } else if (w->ff) {
        // forward of both networks, all three layers, one launch
        WCK(launch_fused_forward(w, bt, s.theta, s.theta_t, st, pdl, -1, nullptr, s.have && !(w->skip & 1)));
        launches += 1;
        mark(w->ff2 ? "fused forward L1-L3 online+target, CTA pairs (tcgen05 cta_group::2)"
                    : "fused forward L1-L3 online+target (tcgen05)");
    } else {
    // forward, online (problem 0) and target (problem 1) in one launch per layer
    g.M = B; g.N = d.H1; g.K = d.D; g.k_per_split = ((d.D + BK - 1) / BK) * BK;
    g.out_bf16 = w->h1; g.out_bf16_1 = w->t1; g.bias = s.theta + o.b1; g.bias_1 = s.theta_t + o.b1;
    g.ldo = d.H1; g.n_store = d.H1;
    if (w->bn256)
        WCK((launch_gemm<MODE_KM, 256, EPI_BIAS_RELU_BF16, 1, 2>(bt.tm_x_a, on.tm_w1_mn, g, 1, st, &bt.tm_x2_a, &tg.tm_w1_mn, pdl, &w->tm_h1_a, &w->tm_t1_a)));
    else if (w->mc_rows)
        WCK((launch_gemm<MODE_KM, 128, EPI_BIAS_RELU_BF16, 4>(bt.tm_x_a, on.tm_w1_mn_s4, g, 1, st, &bt.tm_x2_a, &tg.tm_w1_mn_s4, pdl, &w->tm_h1_a, &w->tm_t1_a)));
    else
        WCK((launch_gemm<MODE_KM, 128, EPI_BIAS_RELU_BF16>(bt.tm_x_a, on.tm_w1_mn, g, 1, st, &bt.tm_x2_a, &tg.tm_w1_mn, pdl, &w->tm_h1_a, &w->tm_t1_a)));
    mark("fwd L1 online+target (tcgen05)");
    g.N = d.H2; g.K = d.H1; g.k_per_split = ((d.H1 + BK - 1) / BK) * BK;
    g.out_bf16 = w->h2; g.out_bf16_1 = w->t2; g.bias = s.theta + o.b2; g.bias_1 = s.theta_t + o.b2;
    g.ldo = d.H2; g.n_store = d.H2;
    if (w->bn256)
        WCK((launch_gemm<MODE_KM, 256, EPI_BIAS_RELU_BF16, 1, 2>(w->tm_h1_a, on.tm_w2_mn, g, 1, st, &w->tm_t1_a, &tg.tm_w2_mn, pdl, &w->tm_h2_a, &w->tm_t2_a)));
    else if (w->mc_rows)
        WCK((launch_gemm<MODE_KM, 128, EPI_BIAS_RELU_BF16, 4>(w->tm_h1_a, on.tm_w2_mn_s4, g, 1, st, &w->tm_t1_a, &tg.tm_w2_mn_s4, pdl, &w->tm_h2_a, &w->tm_t2_a)));
    else
        WCK((launch_gemm<MODE_KM, 128, EPI_BIAS_RELU_BF16>(w->tm_h1_a, on.tm_w2_mn, g, 1, st, &w->tm_t1_a, &tg.tm_w2_mn, pdl, &w->tm_h2_a, &w->tm_t2_a)));
    mark("fwd L2 online+target (tcgen05)");
    g.N = 64; g.K = d.H2; g.k_per_split = ((d.H2 + BK - 1) / BK) * BK;     // the first 64 of W3's padded columns
    g.out_bf16 = g.out_bf16_1 = nullptr; g.out_f32 = w->q; g.out_f32_1 = w->qt;
    g.bias = s.theta + o.b3; g.bias_1 = s.theta_t + o.b3; g.ldo = d.A; g.n_store = d.A;
    WCK((launch_gemm<MODE_KM, 64, EPI_BIAS_F32>(w->tm_h2_a, on.tm_w3p_mn, g, 1, st, &w->tm_t2_a, &tg.tm_w3p_mn, pdl)));
    mark("fwd L3 online+target (tcgen05)");
    launches += 3;
    }

    // the pair forward leaves the online network's Rectify masks as bit words: the backward pass reads those instead of
    // the sign bits of H1 / H2 (8 MB each)
    const bool bits = w->ff && w->ff2 && !(w->skip & 2) && getenv("GOLDQ_NO_MASKBITS") == nullptr;
    // TD/dZ2 may read the batch's actions / rewards / done flags ahead of its wait only behind the fused forward with its
    // late trigger (see the kernel)
    static const bool early_trig = getenv("GOLDQ_EARLY_TRIGGER") != nullptr, td_noearly = getenv("GOLDQ_TD_NO_EARLY") != nullptr;
    const bool td_early = w->ff && w->ff2 && !(w->skip & 2) && !early_trig && !td_noearly;     // (ff_forward2_kernel only)
    // (Tried: dW2 / dW3 as PROGRAMMATIC dependents of TD/dZ2 across streams, through an event recorded by TD's launch
    // (cudaLaunchAttributeProgrammaticEvent, triggerAtBlockStart) -- it captures, but the step went from 42.2 to 49.6 us.)
    static const bool dx_first = getenv("GOLDQ_DX_FIRST") != nullptr;
    if (!(w->skip & 4))
    // (W3^T staged in shared memory ahead of the wait was tried first: block 4.8 -> 3.9 us, but 35 KB blocks made the step
    // longer; loading the rows themselves ahead of the wait needs no shared memory)
    WCK(launch_k(wide_td_dz2_kernel, dim3(w->nb2), dim3(256), sizeof(float) * (8 * d.H2 + 8 * 33), st, pdl, w->q, w->qt, bt.ba, bt.br,
                 bt.bdone, B, d.H2, d.A, s.sc, w->h2, bits ? w->m2bits : nullptr, on.w3t, w->td, w->dqT, w->dz2, w->b2_part, w->row_part,
                 td_early ? 1 : 0));
    mark("TD + dZ2 + db2/db3 (K3)");
    launches += 1;

    // backward: three independent chains after dZ2 —  s1: dW2 | s2: dW3, db1 | st: dX -> dW1
    // (GOLDQ_DX_FIRST, experiment: dW2 and dW3 start when dX is done instead of next to it)
    if (!one_stream && !dx_first) {
        WCK(cudaEventRecord(w->ev_dz2, st));
        WCK(cudaStreamWaitEvent(s1, w->ev_dz2, 0));
        WCK(cudaStreamWaitEvent(s2, w->ev_dz2, 0));
    }
    // GOLDQ_UPDATE_PDL: 0 = both parts of the update are plain launches (2 us behind the GEMM they follow), 1 = part 1 is
    // launched programmatically behind dW1, 2 (default) = part 2 behind dW2 as well.  dW1 / dW2 then release their
    // dependent from the EPILOGUE: released from the prologue, the update's blocks sat on the SMs for the whole GEMM and
    // the step did not get shorter (42.8 vs 42.6 us); released late, 42.1.
    static const int upd_pdl_mode = getenv("GOLDQ_UPDATE_PDL") ? atoi(getenv("GOLDQ_UPDATE_PDL")) : 2;
    auto side_gemms = [&]() -> int {
        // dW3 = H2^T dQ : M = H2, N = 32 (A stored), K = B ; A MN-major (h2 as stored), B = dQ^T K-major
        g = GemmArgs{};
        g.trace_id = TR_DW3;
        g.M = d.H2; g.N = 32; g.K = B; g.k_per_split = w->kps3;
        g.out_f32 = w->w3_slices; g.ldo = d.A; g.n_store = d.A; g.slice_stride = (long long)d.H2 * d.A;
        if (!(w->skip & 8))
        WCK((launch_gemm<MODE_MK, 32, EPI_F32_SLICE>(w->tm_h2_mm, w->tm_dqT_b, g, w->ns3, s2)));
        mark("dW3 (tcgen05)");
        // dW2 = H1^T dZ2 : M = H1, N = H2, K = B
        g = GemmArgs{};
        g.trace_id = TR_DW2;
        g.M = d.H1; g.N = d.H2; g.K = B; g.k_per_split = w->kps2;
        g.out_f32 = w->w2_slices; g.ldo = d.H2; g.n_store = d.H2; g.slice_stride = (long long)d.H1 * d.H2;
        g.late_trigger = upd_pdl_mode >= 2 ? 1 : 0;
        if (w->skip & 16) {
        } else if (w->pair_dw2) {
            PairGemmArgs pg{};
            pg.trace_id = TR_DW2;
            pg.M = d.H1; pg.N = d.H2; pg.K = B; pg.k_per_split = w->kps2;
            pg.out_f32 = w->w2_slices; pg.ldo = d.H2; pg.slice_stride = (long long)d.H1 * d.H2;
            WCK(launch_pair_gemm<PG_DW>(w->tm_h1_mm, w->tm_dz2_mm, w->tm_dz2_mm, pg, w->ns2, s1, false));
        } else if (w->mc_h1)
            WCK((launch_gemm<MODE_MM, 128, EPI_F32_SLICE, 4>(w->tm_h1_mm, w->tm_dz2_mm_s4, g, w->ns2, s1)));
        else
            WCK((launch_gemm<MODE_MM, 128, EPI_F32_SLICE>(w->tm_h1_mm, w->tm_dz2_mm, g, w->ns2, s1)));
        mark(w->pair_dw2 ? "dW2 split-K, CTA pairs (tcgen05 cta_group::2)" : "dW2 split-K (tcgen05)");
        if (!one_stream) WCK(cudaEventRecord(w->ev_s1, s1));
        return 0;
    };
    if (!dx_first && side_gemms() < 0) return -1;
    // dZ1 = (dZ2 W2^T) * mask1 : M = B, N = H1, K = H2 ; B operand = W2 [H1 rows][H2 cols]
    g = GemmArgs{};
    g.trace_id = TR_DX;
    g.M = B; g.N = d.H1; g.K = d.H2; g.k_per_split = ((d.H2 + BK - 1) / BK) * BK;
    g.out_bf16 = w->dz1; g.mask_src = w->h1; g.ldo = d.H1; g.n_store = d.H1;
    if (w->skip & 32) {
    } else if (w->pair_dx) {
        PairGemmArgs pg{};
        pg.trace_id = TR_DX;
        pg.M = B; pg.N = d.H1; pg.K = d.H2; pg.k_per_split = ((d.H2 + BK - 1) / BK) * BK;
        pg.mask_src = w->h1; pg.ldo = d.H1; pg.colsum = w->b1_part;
        pg.mask_bits = bits ? w->m1bits : nullptr;
        static const bool early = getenv("GOLDQ_EARLY_TRIGGER") != nullptr;
        pg.late_trigger = early ? 0 : 1;
        WCK(launch_pair_gemm<PG_DX>(w->tm_dz2_a, on.tm_w2_k, w->tm_dz1_a, pg, 1, st, pdl));
    } else if (w->bn256)
        // one CTA per SM here (B/128 x H1/256 tiles <= SMs): a 4-stage ring (192 KB) instead of 2
        if (((B + BM - 1) / BM) * ((d.H1 + 255) / 256) <= w->sm_count)
            WCK((launch_gemm<MODE_KK, 256, EPI_MASK_BF16, 1, 4>(w->tm_dz2_a, on.tm_w2_k256, g, 1, st, nullptr, nullptr, pdl, &w->tm_dz1_a)));
        else
            WCK((launch_gemm<MODE_KK, 256, EPI_MASK_BF16, 1, 2>(w->tm_dz2_a, on.tm_w2_k256, g, 1, st, nullptr, nullptr, pdl, &w->tm_dz1_a)));
    else if (w->mc_rows)
        WCK((launch_gemm<MODE_KK, 128, EPI_MASK_BF16, 4>(w->tm_dz2_a, on.tm_w2_k_s4, g, 1, st, nullptr, nullptr, pdl, &w->tm_dz1_a)));
    else
        WCK((launch_gemm<MODE_KK, 128, EPI_MASK_BF16>(w->tm_dz2_a, on.tm_w2_k, g, 1, st, nullptr, nullptr, pdl, &w->tm_dz1_a)));
    mark(w->pair_dx ? "dX -> dZ1 + db1, CTA pairs (tcgen05 cta_group::2)" : "dX -> dZ1 (tcgen05)");
    if (!one_stream) {
        WCK(cudaEventRecord(w->ev_dx, st));
        // (s2 continues with the db1 column sums of dZ1 -- or, with the pair dX that produces them itself, with nothing:
        // then dW3's branch does not have to wait for dX)
        if (!w->pair_dx || dx_first) WCK(cudaStreamWaitEvent(s2, w->ev_dx, 0));
    }
    if (dx_first) {
        if (!one_stream) WCK(cudaStreamWaitEvent(s1, w->ev_dx, 0));      // (s2 waits for ev_dx above)
        if (side_gemms() < 0) return -1;
    }
    if (!(w->skip & 64) && !w->pair_dx) {
        dim3 grid((d.H1 + 63) / 64, w->nb1);
        wide_colsum_kernel<<<grid, 256, 0, s2>>>(w->dz1, B, d.H1, w->b1_part);
        WCK(cudaGetLastError());
        mark("db1 colsum");
        launches++;
    }
    if (!one_stream) WCK(cudaEventRecord(w->ev_s2, s2));
    // dW1 = X^T dZ1 : M = D, N = H1, K = B
    g = GemmArgs{};
    g.trace_id = TR_DW1;
    g.M = d.D; g.N = d.H1; g.K = B; g.k_per_split = w->kps1;
    g.out_f32 = w->w1_slices; g.ldo = d.H1; g.n_store = d.H1; g.slice_stride = (long long)d.D * d.H1;
    g.late_trigger = upd_pdl_mode >= 1 ? 1 : 0;
    // (GOLDQ_DW1_NOPDL, experiment: launched programmatically, dW1's CTAs become resident as soon as dX's CTAs have
    // started and then hold their shared memory idle until dX is done, while dW2's CTAs wait for a slot)
    static const bool dw1_nopdl = getenv("GOLDQ_DW1_NOPDL") != nullptr;
    if (!(w->skip & 128))
    WCK((launch_gemm<MODE_MM, 128, EPI_F32_SLICE>(bt.tm_x_mm, w->tm_dz1_mm, g, w->ns1, st, nullptr, nullptr, pdl && !dw1_nopdl)));
    mark("dW1 split-K (tcgen05)");
    launches += 4;

    PartialSrc ps{w->w1_slices, w->ns1, w->b1_part, w->nb1, w->w2_slices, w->ns2,
                  w->b2_part,   w->nb2, w->w3_slices, w->ns3, w->row_part, w->nrow};
    static const bool one_lane = getenv("GOLDQ_UPDATE_ONE_LANE") != nullptr;      // diagnostic: round-1 decomposition
    static const bool one_update = getenv("GOLDQ_UPDATE_ONE") != nullptr;         // diagnostic: the update as one launch
    const bool dp_launch = s.fuse_update && s.push;
    float *nop = nullptr;
    // The update as TWO launches on the single-GPU step, each joining only the branches it needs:
    //   part 2 (W2, b2, W3, b3, loss) on dW2's branch, behind dW2, dW3 and TD/dZ2 -- next to dW1;
    //   part 1 (W1, b1) on the main chain behind dW1: 66 k of the 338 k parameters, 4 of the 18 MB of split-K slices.
    // As one launch behind the join of all three branches it was the last 6 us of the step (L2 bandwidth: 18.5 MB).
    // The data-parallel instantiation stays one launch (its exchange pairs blocks of equal index across ranks).
    // (GOLDQ_DP_SPLIT: the data-parallel step split the same way -- blocks of equal index still pair across ranks, launch by
    // launch, and the two grids together are the one grid whose co-residency wide_p2p_resident checks; 80 % of the
    // gradient is then exchanged under dW1)
    static const bool dp_split = getenv("GOLDQ_DP_SPLIT") != nullptr;
    const bool split = (!dp_launch || dp_split) && !one_stream && !one_update && !(w->skip & 256);
    auto launch_update = [&](int part, cudaStream_t us, bool upd_pdl = false) -> cudaError_t {
        const UpdatePlan pl = make_update_plan(d, w->ns1, w->ns2, w->ns3, !dp_launch && !one_lane, part);
        const int blocks = pl.blocksA + pl.blocksB + pl.blocksC;
        if (dp_launch)
            return launch_k(wide_update_kernel<true, true>, dim3(blocks), dim3(128), 0, us, upd_pdl, d, o, ps, pl, s.sc.count, s.flat,
                            s.theta_rw, s.m, s.v, s.adam, s.step, shadow_ptrs(w, 0), s.loss_out, *s.push, s.epoch);
        if (s.fuse_update)
            return launch_k(wide_update_kernel<true, false>, dim3(blocks), dim3(128), 0, us, upd_pdl, d, o, ps, pl, s.sc.count, s.flat,
                            s.theta_rw, s.m, s.v, s.adam, s.step, shadow_ptrs(w, 0), s.loss_out, PeerSet{}, 0u);
        return launch_k(wide_update_kernel<false, false>, dim3(blocks), dim3(128), 0, us, upd_pdl, d, o, ps, pl, s.sc.count, s.flat,
                        nop, nop, nop, s.adam, s.step, shadow_ptrs(w, 0), nop, PeerSet{}, 0u);
    };
    if (split) {
        WCK(cudaStreamWaitEvent(s1, w->ev_s2, 0));          // dW3 (and, without the pair dX, the db1 column sums)
        // ... and dX: part 2 rewrites the bf16 shadow of W2, which dX reads as its B operand.  (dW2 normally ends 6 us after
        // dX, so the wait costs nothing -- without it the result depended on that timing: the variants test caught it.)
        WCK(cudaStreamWaitEvent(s1, w->ev_dx, 0));
        WCK(launch_update(2, s1, pdl && upd_pdl_mode >= 2 && !dp_launch));
        WCK(cudaEventRecord(w->ev_s1, s1));                  // (re-recorded: the branch now ends with its half of the update)
        if (!w->pair_dx) WCK(cudaStreamWaitEvent(st, w->ev_s2, 0));
        WCK(launch_update(1, st, pdl && upd_pdl_mode >= 1 && !dp_launch));
        WCK(cudaStreamWaitEvent(st, w->ev_s1, 0));
        launches++;
    } else {
        if (!one_stream) {
            WCK(cudaStreamWaitEvent(st, w->ev_s1, 0));
            WCK(cudaStreamWaitEvent(st, w->ev_s2, 0));
        }
        if (!(w->skip & 256)) WCK(launch_update(0, st));
    }
    mark("reduce + Adam + bf16 shadow (K5)");
    launches++;
    return launches;
}

std::string wide_step_kernel_names(WideState *w)
{
    std::string out;
    for (size_t i = 0; i < w->prof_names.size(); i++) {
        if (i) out += "\n";
        out += w->prof_names[i];
    }
    return out;
}

void wide_set_debug(WideState *w, bool on) { w->debug = on; }
// the block log since the last call: [0] = number of records, records from word 4 on (TraceRec = 4 words)
const unsigned long long *wide_trace(WideState *w, size_t *bytes)
{
    *bytes = w->trace ? (4 + (size_t)TRACE_CAP * (sizeof(TraceRec) / 8)) * 8 : 0;
    return w->trace;
}
void wide_trace_reset(WideState *w, cudaStream_t st)
{
    if (w->trace) cudaMemsetAsync(w->trace, 0, 8, st);
}

const long long *wide_ff_stamps(WideState *w, size_t *bytes)
{
    *bytes = w->ff_stamps ? (size_t)(w->B / BM) * 2 * FF_NSTAMPS * sizeof(long long) : 0;
    return w->ff_stamps;
}
void wide_debug_ptrs(WideState *w, const float **q, const float **qt, const float **td)
{
    *q = w->q; *qt = w->qt; *td = w->td;
}
const uint8_t *wide_done_ptr(WideState *w) { return w->bs[w->last_set].bdone; }


// ---------------------------------------------------------------------------
// Diagnostic: issue / completion rate of back-to-back tcgen05.mma (selftest modes 30, 31).
// One thread issues 4*iters MMAs (M = 128, K = 16) on resident shared-memory operands.
// Measured on B200: 139.4 cycles per MMA for N = 64, 128 and 256 alike, K- or MN-major B, one or
// two accumulators: the instruction streams the 128 rows of A at one row per cycle, so only
// N = 256 reaches the tensor peak and narrower tiles cost the same time for less work.
// ---------------------------------------------------------------------------
template <int N, bool B_MN>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int naccs, float *out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N, false, B_MN);
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
                mma_bf16(tmem + ((i * 4 + kk) % naccs) * N, make_smem_desc(a + kk * 32, 16, 1024),
                         B_MN ? make_smem_desc(b + kk * 2048, 8192, 1024) : make_smem_desc(b + kk * 32, 16, 1024), idesc, 1u);
        }
        const long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[blockIdx.x * 2] = float(t1 - t0) / (4.f * iters);
        out[blockIdx.x * 2 + 1] = float(t2 - t0) / (4.f * iters);
    }
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------
// Diagnostic: how fast can ONE thread of a CTA stream L2-resident data into shared memory (selftest modes
// 40..43)?  A ring of `SLOTS` 16 KB slots; every slot is refilled as soon as its previous load has landed.
//   kind 0: two TMA 2D boxes of [64 rows][128 B] per slot out of a row-major matrix with a 1 KB pitch, the
//           SAME boxes in every CTA (the weight stream of the fused forward)
//   kind 1: the same boxes, but every CTA reads its own rows (activation tiles)
//   kind 2: one contiguous 16 KB bulk copy per slot (cp.async.bulk), the same bytes in every CTA
//   kind 3: contiguous 16 KB bulk copies, own region per CTA
// out[2*cta] = bytes per SM clock, out[2*cta+1] = cycles from first issue to last completion.
// ---------------------------------------------------------------------------
constexpr int IG_SLOTS = 6;
__global__ void __launch_bounds__(128, 1)
ingest_probe_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t *flat, int kind, int iters, int rows_total, float *out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[IG_SLOTS];
    if (threadIdx.x == 0) {
        for (int s = 0; s < IG_SLOTS; s++) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int own = (kind & 1) ? (int)blockIdx.x : 0;
        const long long t0 = clock64();
        for (int i = 0; i < iters + IG_SLOTS; i++) {
            const int s = i % IG_SLOTS;
            if (i >= IG_SLOTS) mbar_wait_bounded<false>(&full[s], ((i / IG_SLOTS) - 1) & 1);
            if (i < iters) {
                mbar_expect_tx(&full[s], 16384);
                if (kind < 2) {
                    // matrix [rows_total][512 bf16]: slot i = rows [64 * (i % 8) + 512 * own, +64), columns 128 * (i / 8 % 4) ...
                    const int r0 = (64 * (i & 7) + 512 * own) % rows_total, c0 = 128 * ((i >> 3) & 3);
                    tma_load_2d(smem + s * 16384, &tm, c0, r0, &full[s]);
                    tma_load_2d(smem + s * 16384 + 8192, &tm, c0 + 64, r0, &full[s]);
                } else {
                    const size_t off = ((size_t)(i & 31) * 16384 + (size_t)own * 524288) % ((size_t)rows_total * 1024);
                    bulk_load_1d(smem + s * 16384, flat + off, 16384, &full[s]);
                }
            }
        }
        const long long t1 = clock64();
        out[2 * blockIdx.x] = (float)((double)iters * 16384.0 / (double)(t1 - t0));
        out[2 * blockIdx.x + 1] = (float)(t1 - t0);
    }
}

template <int N, bool B_MN>
static cudaError_t run_mma_rate(int ctas, int iters, int naccs, float *dout)
{
    cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel<N, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
    if (e != cudaSuccess) return e;
    mma_rate_kernel<N, B_MN><<<ctas, 128, 16384 + 32768 + 1024>>>(iters, naccs, dout);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// stand-alone tcgen05 GEMM self-test (diagnostic ABI: goldq_selftest_umma)
//   mode 0 (KK): C[M][N] = sum_k A[m][k] B[n][k]      A [M][K], B [N][K] bf16 (as uint16)
//   mode 1 (MM): C[M][N] = sum_k A[k][m] B[k][n]      A [K][M], B [K][N]
// ---------------------------------------------------------------------------
int wide_selftest_gemm(int mode, int M, int N, int K, int nsplit, const uint16_t *A, const uint16_t *Bm, float *C,
                       std::string &err)
{
    if (!load_encode(err)) return -1;
    WCK(configure_all_gemms());
    if (nsplit < 1) nsplit = 1;
    if (mode >= 40 && mode <= 43) {
        // ingest probe: M/128 CTAs, K = 16 KB slots per CTA; C[2*cta] = bytes per clock, C[2*cta+1] = cycles
        const int ctas = M / 128, rows_total = 512 * ctas;
        if (ctas < 1 || K < 1) { err = "selftest: ingest probe needs M >= 128, K >= 1"; return -1; }
        __nv_bfloat16 *dsrc = nullptr;
        float *dout = nullptr;
        if (!dmalloc(&dsrc, (size_t)rows_total * 512, err) || !dmalloc(&dout, (size_t)ctas * 2, err)) return -1;
        WCK(cudaMemset(dsrc, 0, (size_t)rows_total * 1024));
        CUtensorMap tm;
        if (!make_tmap(&tm, dsrc, rows_total, 512, 512, 64, err)) return -1;
        WCK(cudaFuncSetAttribute(ingest_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SLOTS * 16384 + 1024));
        for (int rep = 0; rep < 2; rep++) {      // the second run reads from a warm L2
            ingest_probe_kernel<<<ctas, 128, IG_SLOTS * 16384 + 1024>>>(tm, reinterpret_cast<const uint8_t *>(dsrc), mode - 40, K,
                                                                        rows_total, dout);
            WCK(cudaGetLastError());
            WCK(cudaDeviceSynchronize());
        }
        WCK(cudaMemcpy(C, dout, (size_t)ctas * 2 * sizeof(float), cudaMemcpyDeviceToHost));
        cudaFree(dsrc);
        cudaFree(dout);
        return 0;
    }
    if (mode == 30 || mode == 31) {
        // MMA rate probe: M/128 CTAs, N = MMA width, K = iterations, nsplit = accumulators cycled
        // through; mode 30: B MN-major, 31: B K-major.  C[0] = issue cycles per MMA, C[1] = issue +
        // completion cycles per MMA (CTA 0).
        float *dout = nullptr;
        const int ctas = M / 128;
        if (ctas < 1 || K < 1) { err = "selftest: rate probe needs M >= 128, K >= 1"; return -1; }
        if (!dmalloc(&dout, (size_t)ctas * 2, err)) return -1;
        const int naccs = nsplit * N <= 512 ? nsplit : 512 / N;
        cudaError_t e = cudaErrorInvalidValue;
        if (mode == 30 && N == 256) e = run_mma_rate<256, true>(ctas, K, naccs, dout);
        else if (mode == 30 && N == 128) e = run_mma_rate<128, true>(ctas, K, naccs, dout);
        else if (mode == 30 && N == 64) e = run_mma_rate<64, true>(ctas, K, naccs, dout);
        else if (mode == 31 && N == 256) e = run_mma_rate<256, false>(ctas, K, naccs, dout);
        else if (mode == 31 && N == 128) e = run_mma_rate<128, false>(ctas, K, naccs, dout);
        else if (mode == 31 && N == 64) e = run_mma_rate<64, false>(ctas, K, naccs, dout);
        WCK(e);
        WCK(cudaDeviceSynchronize());
        WCK(cudaMemcpy(C, dout, 2 * sizeof(float), cudaMemcpyDeviceToHost));
        cudaFree(dout);
        return 0;
    }
    __nv_bfloat16 *dA = nullptr, *dB = nullptr;
    float *dC = nullptr;
    size_t na = (size_t)M * K, nb = (size_t)N * K;
    if (!dmalloc(&dA, na, err) || !dmalloc(&dB, nb, err) || !dmalloc(&dC, (size_t)nsplit * M * N, err)) return -1;
    WCK(cudaMemcpy(dA, A, na * 2, cudaMemcpyHostToDevice));
    WCK(cudaMemcpy(dB, Bm, nb * 2, cudaMemcpyHostToDevice));
    WCK(cudaMemset(dC, 0xff, (size_t)nsplit * M * N * 4));
    CUtensorMap ta, tb;
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K;
    int kblocks = (K + BK - 1) / BK, per = (kblocks + nsplit - 1) / nsplit;
    g.k_per_split = per * BK;
    g.out_f32 = dC; g.ldo = N; g.n_store = N; g.slice_stride = (long long)M * N;
    if (mode == 0) {
        if (!make_tmap(&ta, dA, M, K, K, BM, err) || !make_tmap(&tb, dB, N, K, K, 128, err)) return -1;
        WCK((launch_gemm<MODE_KK, 128, EPI_F32_SLICE>(ta, tb, g, nsplit, 0)));
    } else if (mode == 1) {
        if (!make_tmap(&ta, dA, K, M, M, 64, err) || !make_tmap(&tb, dB, K, N, N, 64, err)) return -1;
        WCK((launch_gemm<MODE_MM, 128, EPI_F32_SLICE>(ta, tb, g, nsplit, 0)));
    } else if (mode == 2) {   // KK with the 32-wide N tile
        if (!make_tmap(&ta, dA, M, K, K, BM, err) || !make_tmap(&tb, dB, N, K, K, 32, err)) return -1;
        WCK((launch_gemm<MODE_KK, 32, EPI_F32_SLICE>(ta, tb, g, nsplit, 0)));
    } else if (mode == 3) {   // MK: A MN-major [K][M], B K-major [N=32][K]
        if (!make_tmap(&ta, dA, K, M, M, 64, err) || !make_tmap(&tb, dB, N, K, K, 32, err)) return -1;
        WCK((launch_gemm<MODE_MK, 32, EPI_F32_SLICE>(ta, tb, g, nsplit, 0)));
    } else if (mode == 4) {   // KM: A K-major [M][K], B MN-major [K][N]
        if (!make_tmap(&ta, dA, M, K, K, BM, err) || !make_tmap(&tb, dB, K, N, N, 64, err)) return -1;
        WCK((launch_gemm<MODE_KM, 128, EPI_F32_SLICE>(ta, tb, g, nsplit, 0)));
    } else if (mode == 5) {   // KM with the 64-wide N tile
        if (!make_tmap(&ta, dA, M, K, K, BM, err) || !make_tmap(&tb, dB, K, N, N, 64, err)) return -1;
        WCK((launch_gemm<MODE_KM, 64, EPI_F32_SLICE>(ta, tb, g, nsplit, 0)));
    } else if (mode == 10) {  // KK, B multicast over clusters of 4 along M
        if (!make_tmap(&ta, dA, M, K, K, BM, err) || !make_tmap(&tb, dB, N, K, K, 32, err)) return -1;
        WCK((launch_gemm<MODE_KK, 128, EPI_F32_SLICE, 4>(ta, tb, g, nsplit, 0)));
    } else if (mode == 11) {  // MM, B multicast
        if (!make_tmap(&ta, dA, K, M, M, 64, err) || !make_tmap(&tb, dB, K, N, N, 16, err)) return -1;
        WCK((launch_gemm<MODE_MM, 128, EPI_F32_SLICE, 4>(ta, tb, g, nsplit, 0)));
    } else if (mode == 14) {  // KM, B multicast
        if (!make_tmap(&ta, dA, M, K, K, BM, err) || !make_tmap(&tb, dB, K, N, N, 16, err)) return -1;
        WCK((launch_gemm<MODE_KM, 128, EPI_F32_SLICE, 4>(ta, tb, g, nsplit, 0)));
    } else if (mode == 20) {  // KK, 128x256 tiles, 2 stages
        if (!make_tmap(&ta, dA, M, K, K, BM, err) || !make_tmap(&tb, dB, N, K, K, 256, err)) return -1;
        WCK((launch_gemm<MODE_KK, 256, EPI_F32_SLICE, 1, 2>(ta, tb, g, nsplit, 0)));
    } else if (mode == 24) {  // KM, 128x256 tiles, 2 stages
        if (!make_tmap(&ta, dA, M, K, K, BM, err) || !make_tmap(&tb, dB, K, N, N, 64, err)) return -1;
        WCK((launch_gemm<MODE_KM, 256, EPI_F32_SLICE, 1, 2>(ta, tb, g, nsplit, 0)));
    } else {
        err = "selftest: unknown mode";
        return -1;
    }
    WCK(cudaDeviceSynchronize());
    std::vector<float> tmp((size_t)nsplit * M * N);
    WCK(cudaMemcpy(tmp.data(), dC, tmp.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < (size_t)M * N; i++) {
        float s = 0.f;
        for (int z = 0; z < nsplit; z++) s += tmp[(size_t)z * M * N + i];
        C[i] = s;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return 0;
}

}  // namespace gq
