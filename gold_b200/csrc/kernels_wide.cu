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
//   forward (both nets)  H1 = relu(X  W1 + b1)      KK mode  M=B   N=H1  K=D
//                        H2 = relu(H1 W2 + b2)      KK       M=B   N=H2  K=H1
//                        Q  =      H2 W3 + b3       KK, BN=32 M=B  N=A   K=H2
//   backward             dW2 = H1^T dZ2             MM mode  M=H1  N=H2  K=B (split-K)
//                        dZ1 = (dZ2 W2^T) * mask1   KK       M=B   N=H1  K=H2
//                        dW1 = X^T dZ1              MM       M=D   N=H1  K=B (split-K)
//   layer 3's backward is rank-1 per row (dq has one non-zero per row: the taken
//   action), so dW3/db3/dZ2 are computed on CUDA cores in wide_td_kernel.
// KK: both operands K-major (A [M][K], B [N][K] row-major).  MM: both MN-major
// (A [K][M], B [K][N] row-major) — the activations are used as stored, no transposes.
#include <cuda.h>
#include <cuda_bf16.h>

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
constexpr int STAGES = 3;
constexpr int GEMM_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

enum { MODE_KK = 0, MODE_MM = 1 };
enum { EPI_BIAS_RELU_BF16 = 0, EPI_BIAS_F32 = 1, EPI_MASK_BF16 = 2, EPI_F32_SLICE = 3 };

struct GemmArgs {
    int M, N, K;          // logical problem (K = full reduction length)
    int k_per_split;      // multiple of BK; blockIdx.z selects the range
    // epilogue operands
    __nv_bfloat16 *out_bf16;      // [M][ldo]
    float *out_f32;               // [M][ldo] (+ blockIdx.z * slice_stride for EPI_F32_SLICE)
    const float *bias;            // [N]
    const __nv_bfloat16 *mask_src;// [M][ldo]: sign bit set <=> ReLU mask 0
    int ldo;                      // row pitch of the output in elements
    int n_store;                  // columns actually stored (<= N tile coverage)
    long long slice_stride;
};

__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }

template <int MODE, int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmArgs g)
{
    static_assert(BN == 128 || BN == 64 || BN == 32, "BN");
    static_assert(MODE == MODE_KK || BN % 64 == 0, "MN-major B tiles are 64-wide swizzle atoms");
    constexpr uint32_t A_BYTES = BM * BK * 2;
    constexpr uint32_t B_BYTES = BN * BK * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr uint32_t IDESC = make_idesc_bf16(BM, BN, MODE == MODE_MM, MODE == MODE_MM);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int k_begin = blockIdx.z * g.k_per_split;
    int k_end = k_begin + g.k_per_split;
    if (k_end > g.K) k_end = g.K;
    const int nkb = (k_end - k_begin + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

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
                if (MODE == MODE_KK) {
                    tma_load_2d(sA, &tmA, k, m0, &full_bar[s]);       // box {64 k, 128 rows}
                    tma_load_2d(sB, &tmB, k, n0, &full_bar[s]);       // box {64 k, BN rows}
                } else {
                    // MN-major: box {64 mn, 64 k-rows}; one 8 KB atom column per 64 mn
                    for (int h = 0; h < BM / 64; h++)
                        tma_load_2d(sA + h * 8192, &tmA, m0 + 64 * h, k, &full_bar[s]);
                    for (int h = 0; h < BN / 64; h++)
                        tma_load_2d(sB + h * 8192, &tmB, n0 + 64 * h, k, &full_bar[s]);
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
            if (elect_one()) {
                const uint32_t aaddr = smem_u32(smem + s * STAGE_BYTES), baddr = aaddr + A_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / UK; kk++) {
                    uint64_t ad, bd;
                    if (MODE == MODE_KK) {
                        // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); the 16
                        // k-elements of this MMA start 32 B further inside the swizzle row.
                        ad = make_smem_desc(aaddr + kk * 32, 16, 1024);
                        bd = make_smem_desc(baddr + kk * 32, 16, 1024);
                    } else {
                        // MN-major SW128: k-rows of 128 B (64 mn), 8-k-row groups 1024 B apart
                        // (SBO), 64-mn groups 8192 B apart (LBO); 16 k-rows = 2048 B per MMA.
                        ad = make_smem_desc(aaddr + kk * 2048, 8192, 1024);
                        bd = make_smem_desc(baddr + kk * 2048, 8192, 1024);
                    }
                    mma_bf16(tmem_acc, ad, bd, IDESC, (kb | kk) != 0 ? 1u : 0u);
                }
                mma_commit(&empty_bar[s]);                 // smem slot free once these MMAs retire
                if (kb == nkb - 1) mma_commit(&tmem_full_bar);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int row = m0 + q * 32 + lane;
        if (nkb > 0) {
            mbar_wait(&tmem_full_bar, 0);
            tc_fence_after();
        }
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
                            uint4 t = (col0 + 8 * v < g.n_store) ? __ldg(mp + v) : make_uint4(0, 0, 0, 0);
                            msk[4 * v] = t.x; msk[4 * v + 1] = t.y; msk[4 * v + 2] = t.z; msk[4 * v + 3] = t.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[j + 1]);
                        if (EPI == EPI_BIAS_RELU_BF16) {
                            const int c = col0 + j;
                            v0 += (c < g.N) ? __ldg(g.bias + c) : 0.f;
                            v1 += (c + 1 < g.N) ? __ldg(g.bias + c + 1) : 0.f;
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
                    uint4 *op = reinterpret_cast<uint4 *>(g.out_bf16 + (size_t)row * g.ldo + col0);
#pragma unroll
                    for (int v = 0; v < 4; v++)
                        if (col0 + 8 * v < g.n_store)
                            op[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
                } else if (EPI == EPI_BIAS_F32) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int c = col0 + j;
                        if (c < g.n_store) g.out_f32[(size_t)row * g.ldo + c] = __uint_as_float(r[j]) + __ldg(g.bias + c);
                    }
                } else {  // EPI_F32_SLICE
                    float *o = g.out_f32 + (size_t)blockIdx.z * g.slice_stride + (size_t)row * g.ldo + col0;
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
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, TMEM_COLS);
    }
}

template <int MODE, int BN>
constexpr size_t gemm_smem_bytes()
{
    return (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024;
}

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

template <int MODE, int BN, int EPI>
cudaError_t launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const GemmArgs &g, int nsplit, cudaStream_t st)
{
    constexpr size_t smem = gemm_smem_bytes<MODE, BN>();
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<MODE, BN, EPI>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, nsplit);
    umma_gemm_kernel<MODE, BN, EPI><<<grid, GEMM_THREADS, smem, st>>>(ta, tb, g);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// CUDA-core helpers of the wide path
// ---------------------------------------------------------------------------
// K1: gather + fp32 -> bf16 (replay stays fp32 in HBM; the cast is fused into the gather)
__global__ void __launch_bounds__(256)
wide_gather_kernel(Replay rp, int D, const int32_t *__restrict__ idx, int B, __nv_bfloat16 *__restrict__ x,
                   __nv_bfloat16 *__restrict__ x2, int32_t *__restrict__ ba, float *__restrict__ br,
                   uint8_t *__restrict__ bdone)
{
    const int chunks = D >> 2;  // float4 granules per row (D % 8 == 0 is required by the path)
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * chunks) return;
    int b = t / chunks, c = t - b * chunks;
    int64_t phys = rp.head - 1 - (int64_t)idx[b];
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
    }
}

// bf16 shadow of one network's weights: W1^T [H1][D], W2^T [H2][H1], W2 [H1][H2] (online
// only, for dX), W3^T padded [32][H2].  K-major B operands of the KK GEMMs.
__global__ void __launch_bounds__(256)
wide_shadow_kernel(const float *__restrict__ theta, Dims d, Offsets o, __nv_bfloat16 *__restrict__ w1t,
                   __nv_bfloat16 *__restrict__ w2t, __nv_bfloat16 *__restrict__ w2, __nv_bfloat16 *__restrict__ w3t)
{
    const int n1 = d.D * d.H1, n2 = d.H1 * d.H2, n3 = 32 * d.H2;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n1 + n2 + n3; t += gridDim.x * blockDim.x) {
        if (t < n1) {
            int j = t / d.D, k = t - j * d.D;                 // w1t[j][k] = W1[k][j]
            w1t[t] = __float2bfloat16_rn(theta[o.w1 + k * d.H1 + j]);
        } else if (t < n1 + n2) {
            int u = t - n1;
            int j = u / d.H1, k = u - j * d.H1;               // w2t[j][k] = W2[k][j]
            w2t[u] = __float2bfloat16_rn(theta[o.w2 + k * d.H2 + j]);
            if (w2) w2[u] = __float2bfloat16_rn(theta[o.w2 + u]);
        } else {
            int u = t - n1 - n2;
            int a = u / d.H2, k = u - a * d.H2;               // w3t[a][k] = W3[k][a], rows >= A are zero
            w3t[u] = (a < d.A) ? __float2bfloat16_rn(theta[o.w3 + k * d.A + a]) : __float2bfloat16_rn(0.f);
        }
    }
}

// K3 + layer-3 backward.  One block per TD_ROWS rows, one thread per hidden unit j.
//   nextMax (first-max rules), td, d = q[a]-td, loss partial, dqv = 2 d / count
//   dZ2[b][j] = dqv * W3[j][a] * mask2[b][j]              (bf16 out, operand of dW2 / dX)
//   partials: dW3[j][a] += h2[b][j]*dqv ; db3[a] += dqv ; db2[j] += dZ2[b][j]
constexpr int TD_ROWS = 64;
__global__ void __launch_bounds__(512)
wide_td_kernel(const float *__restrict__ q, const float *__restrict__ qt, const int32_t *__restrict__ act,
               const float *__restrict__ rew, const uint8_t *__restrict__ done, int B, int H2, int A,
               const float *__restrict__ w3 /* fp32 master [H2][A] */, const __nv_bfloat16 *__restrict__ h2,
               StepScalars sc, float *__restrict__ td_out, __nv_bfloat16 *__restrict__ dz2,
               float *__restrict__ part /* [blocks][H2*A + H2 + A + 1] */)
{
    extern __shared__ float sm[];
    float *sW3 = sm;                 // [H2*A]
    float *sG3 = sW3 + H2 * A;       // [A][H2] dW3 accumulators (thread j owns column j)
    float *sRow = sG3 + H2 * A;      // [TD_ROWS] dqv
    int *sAct = reinterpret_cast<int *>(sRow + TD_ROWS);  // [TD_ROWS]
    __shared__ float sSq[TD_ROWS];
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * TD_ROWS;
    const int nrows = min(TD_ROWS, B - r0);
    for (int i = tid; i < H2 * A; i += blockDim.x) { sW3[i] = w3[i]; sG3[i] = 0.f; }
    __syncthreads();
    if (tid < nrows) {
        const int b = r0 + tid;
        float t = rew[b];
        if (!done[b]) {
            const float *row = qt + (int64_t)b * A;
            float f = row[0];
            for (int i = 1; i < A; i++) {
                float v = row[i];
                if (isnan(v) || (isinf(v) && v > 0.f)) { f = v; break; }
                if (v > f) f = v;
            }
            t = __fadd_rn(rew[b], __fmul_rn(sc.gamma, f));
        }
        td_out[b] = t;
        const int a = act[b];
        const float dd = __fsub_rn(q[(int64_t)b * A + a], t);
        sRow[tid] = __fmul_rn(__fmul_rn(dd, 2.f), sc.inv_count);
        sAct[tid] = a;
        sSq[tid] = __fmul_rn(dd, dd);
    }
    __syncthreads();
    for (int j = tid; j < H2; j += blockDim.x) {
        float db2 = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < nrows; rr++) {
            const int b = r0 + rr;
            const float dqv = sRow[rr];
            const int a = sAct[rr];
            const __nv_bfloat16 hb = h2[(int64_t)b * H2 + j];
            const bool masked = (__bfloat16_as_ushort(hb) & 0x8000u) != 0;
            const float hv = __bfloat162float(hb);
            float dz = masked ? 0.f : dqv * sW3[j * A + a];
            dz2[(int64_t)b * H2 + j] = __float2bfloat16_rn(dz);
            // accumulate the ROUNDED dz so that db2 matches the bf16 operand of dW2 / dX
            db2 += __bfloat162float(__float2bfloat16_rn(dz));
            sG3[a * H2 + j] += hv * dqv;
        }
        float *p = part + (int64_t)blockIdx.x * (H2 * A + H2 + A + 1);
        for (int a = 0; a < A; a++) p[j * A + a] = sG3[a * H2 + j];
        p[H2 * A + j] = db2;
    }
    // db3 and the loss partial, rows in order (deterministic)
    float *p = part + (int64_t)blockIdx.x * (H2 * A + H2 + A + 1);
    if (tid < A) {
        float sb = 0.f;
        for (int rr = 0; rr < nrows; rr++) sb += (sAct[rr] == tid) ? sRow[rr] : 0.f;
        p[H2 * A + H2 + tid] = sb;
    }
    if (tid == A) {
        float sl = 0.f;
        for (int rr = 0; rr < nrows; rr++) sl += sSq[rr];
        p[H2 * A + H2 + A] = sl;
    }
}

// db1 partials: column sums of dZ1 (bf16 [B][H1]) over CS_ROWS-row ranges.  A warp covers
// 64 columns (one bfloat162 per lane), the 8 warps of a block take interleaved rows with
// 8 independent loads in flight each; warps are then summed in order through smem.
constexpr int CS_ROWS = 256;
__global__ void __launch_bounds__(256)
wide_colsum_kernel(const __nv_bfloat16 *__restrict__ dz, int B, int H, float *__restrict__ part /* [blocks.y][H] */)
{
    __shared__ float2 sh[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * 64 + lane * 2;
    const int r0 = blockIdx.y * CS_ROWS, r1 = min(B, r0 + CS_ROWS);
    float2 s = make_float2(0.f, 0.f);
    if (j < H) {
        for (int b = r0 + warp; b < r1; b += 64) {
            __nv_bfloat162 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int bb = b + 8 * u;
                v[u] = (bb < r1) ? *reinterpret_cast<const __nv_bfloat162 *>(dz + (int64_t)bb * H + j)
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

// flat[P+1] = fixed-order sums of all partial buffers; flat[P] = loss.
// Block = 32 elements x 8 partial-chunks: thread (x, y) sums partials y, y+8, ... with the
// loads issued in batches of 4, then the 8 chunk sums are added in order.
__global__ void __launch_bounds__(256)
wide_reduce_kernel(Dims d, Offsets o, const float *__restrict__ w1_slices, int n1, const float *__restrict__ w2_slices,
                   int n2, const float *__restrict__ b1_part, int nb1, const float *__restrict__ td_part, int ntd,
                   float count, float *__restrict__ flat)
{
    __shared__ float sh[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + x;
    const int tdw = d.H2 * d.A + d.H2 + d.A + 1;
    const float *base = nullptr;
    int n = 0;
    int64_t stride = 0;
    if (i < o.b1) { base = w1_slices + i; n = n1; stride = (int64_t)d.D * d.H1; }
    else if (i < o.w2) { base = b1_part + (i - o.b1); n = nb1; stride = d.H1; }
    else if (i < o.b2) { base = w2_slices + (i - o.w2); n = n2; stride = (int64_t)d.H1 * d.H2; }
    else if (i < o.w3) { base = td_part + d.H2 * d.A + (i - o.b2); n = ntd; stride = tdw; }
    else if (i < o.b3) { base = td_part + (i - o.w3); n = ntd; stride = tdw; }
    else if (i < o.P) { base = td_part + d.H2 * d.A + d.H2 + (i - o.b3); n = ntd; stride = tdw; }
    else if (i == o.P) { base = td_part + d.H2 * d.A + d.H2 + d.A; n = ntd; stride = tdw; }
    float s = 0.f;
    for (int z = y; z < n; z += 32) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = (z + 8 * u < n) ? __ldg(base + (int64_t)(z + 8 * u) * stride) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; u++) s += v[u];
    }
    sh[y][x] = s;
    __syncthreads();
    if (y == 0 && i <= o.P) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += sh[w][x];
        if (i == o.P) t = __fdiv_rn(t, count);
        flat[i] = t;
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
    __nv_bfloat16 *w1t = nullptr, *w2t = nullptr, *w2 = nullptr, *w3t = nullptr;
    CUtensorMap tm_w1t, tm_w2t, tm_w2, tm_w3t;
};

struct WideState {
    Dims d;
    Offsets o;
    int B = 0, sm_count = 148;
    WideNet net[2];
    __nv_bfloat16 *x = nullptr, *x2 = nullptr, *h1 = nullptr, *h2 = nullptr, *t1 = nullptr, *t2 = nullptr;
    __nv_bfloat16 *dz2 = nullptr, *dz1 = nullptr;
    float *q = nullptr, *qt = nullptr, *td = nullptr, *br = nullptr;
    int32_t *ba = nullptr;
    uint8_t *bdone = nullptr;
    float *w1_slices = nullptr, *w2_slices = nullptr, *b1_part = nullptr, *td_part = nullptr;
    int ns1 = 1, ns2 = 1, kps1 = 0, kps2 = 0, nb1 = 1, ntd = 1;
    // activation tensor maps: KK A-operands (box 128 rows) and MM operands (box 64 k-rows)
    CUtensorMap tm_x_a, tm_x2_a, tm_h1_a, tm_t1_a, tm_h2_a, tm_t2_a, tm_dz2_a;
    CUtensorMap tm_x_mm, tm_h1_mm, tm_dz2_mm, tm_dz1_mm;
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
    if (d.A > 32) { err = "bf16 path supports up to 32 actions"; return nullptr; }
    if ((size_t)2 * d.H2 * d.A * 4 > 200 * 1024) { err = "bf16 path: hidden2 * n_actions too large for the layer-3 backward kernel"; return nullptr; }
    if (!load_encode(err)) return nullptr;
    WideState *w = new WideState();
    w->d = d;
    w->o = make_offsets(d);
    w->B = B;
    w->sm_count = sm_count;
    bool ok = true;
    for (int n = 0; n < 2; n++) {
        ok = ok && dmalloc(&w->net[n].w1t, (size_t)d.H1 * d.D, err) && dmalloc(&w->net[n].w2t, (size_t)d.H2 * d.H1, err) &&
             dmalloc(&w->net[n].w3t, (size_t)32 * d.H2, err);
    }
    ok = ok && dmalloc(&w->net[0].w2, (size_t)d.H1 * d.H2, err);
    size_t n = B;
    ok = ok && dmalloc(&w->x, n * d.D, err) && dmalloc(&w->x2, n * d.D, err) && dmalloc(&w->h1, n * d.H1, err) &&
         dmalloc(&w->h2, n * d.H2, err) && dmalloc(&w->t1, n * d.H1, err) && dmalloc(&w->t2, n * d.H2, err) &&
         dmalloc(&w->dz2, n * d.H2, err) && dmalloc(&w->dz1, n * d.H1, err) && dmalloc(&w->q, n * d.A, err) &&
         dmalloc(&w->qt, n * d.A, err) && dmalloc(&w->td, n, err) && dmalloc(&w->br, n, err) &&
         dmalloc(&w->ba, n, err) && dmalloc(&w->bdone, n, err);
    if (!ok) { wide_destroy(w); return nullptr; }
    int tiles1 = ((d.D + BM - 1) / BM) * ((d.H1 + 127) / 128);
    int tiles2 = ((d.H1 + BM - 1) / BM) * ((d.H2 + 127) / 128);
    w->ns1 = split_for(tiles1, B, sm_count, &w->kps1);
    w->ns2 = split_for(tiles2, B, sm_count, &w->kps2);
    w->nb1 = (B + CS_ROWS - 1) / CS_ROWS;
    w->ntd = (B + TD_ROWS - 1) / TD_ROWS;
    ok = ok && dmalloc(&w->w1_slices, (size_t)w->ns1 * d.D * d.H1, err) &&
         dmalloc(&w->w2_slices, (size_t)w->ns2 * d.H1 * d.H2, err) && dmalloc(&w->b1_part, (size_t)w->nb1 * d.H1, err) &&
         dmalloc(&w->td_part, (size_t)w->ntd * (d.H2 * d.A + d.H2 + d.A + 1), err);
    if (!ok) { wide_destroy(w); return nullptr; }
    // weight maps: [N rows][K cols]
    for (int k = 0; k < 2; k++) {
        WideNet &nn = w->net[k];
        ok = ok && make_tmap(&nn.tm_w1t, nn.w1t, d.H1, d.D, d.D, 128, err) &&
             make_tmap(&nn.tm_w2t, nn.w2t, d.H2, d.H1, d.H1, 128, err) &&
             make_tmap(&nn.tm_w3t, nn.w3t, 32, d.H2, d.H2, 32, err);
    }
    ok = ok && make_tmap(&w->net[0].tm_w2, w->net[0].w2, d.H1, d.H2, d.H2, 128, err);
    // activations as KK A operands: [B rows][K cols], box 128 rows
    ok = ok && make_tmap(&w->tm_x_a, w->x, B, d.D, d.D, BM, err) && make_tmap(&w->tm_x2_a, w->x2, B, d.D, d.D, BM, err) &&
         make_tmap(&w->tm_h1_a, w->h1, B, d.H1, d.H1, BM, err) && make_tmap(&w->tm_t1_a, w->t1, B, d.H1, d.H1, BM, err) &&
         make_tmap(&w->tm_h2_a, w->h2, B, d.H2, d.H2, BM, err) && make_tmap(&w->tm_t2_a, w->t2, B, d.H2, d.H2, BM, err) &&
         make_tmap(&w->tm_dz2_a, w->dz2, B, d.H2, d.H2, BM, err);
    // activations as MM operands: [K = B rows][MN cols], box 64 k-rows x 64 mn
    ok = ok && make_tmap(&w->tm_x_mm, w->x, B, d.D, d.D, 64, err) && make_tmap(&w->tm_h1_mm, w->h1, B, d.H1, d.H1, 64, err) &&
         make_tmap(&w->tm_dz2_mm, w->dz2, B, d.H2, d.H2, 64, err) && make_tmap(&w->tm_dz1_mm, w->dz1, B, d.H1, d.H1, 64, err);
    if (!ok) { wide_destroy(w); return nullptr; }
    return w;
}

void wide_destroy(WideState *w)
{
    if (!w) return;
    for (int n = 0; n < 2; n++) {
        void *p[] = {w->net[n].w1t, w->net[n].w2t, w->net[n].w2, w->net[n].w3t};
        for (void *q : p)
            if (q) cudaFree(q);
    }
    void *p[] = {w->x, w->x2, w->h1, w->h2, w->t1, w->t2, w->dz2, w->dz1, w->q, w->qt, w->td, w->br, w->ba, w->bdone,
                 w->w1_slices, w->w2_slices, w->b1_part, w->td_part};
    for (void *q : p)
        if (q) cudaFree(q);
    delete w;
}

static int shadow(WideState *w, int net, const float *theta, cudaStream_t st, std::string &err)
{
    WideNet &n = w->net[net];
    int total = w->d.D * w->d.H1 + w->d.H1 * w->d.H2 + 32 * w->d.H2;
    wide_shadow_kernel<<<(total + 255) / 256, 256, 0, st>>>(theta, w->d, w->o, n.w1t, n.w2t, n.w2, n.w3t);
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

int wide_sync_target(WideState *w, cudaStream_t st, std::string &err)
{
    // target <- online: copy the bf16 shadows (W2 plain is online-only)
    const Dims &d = w->d;
    cudaError_t e = cudaMemcpyAsync(w->net[1].w1t, w->net[0].w1t, (size_t)d.H1 * d.D * 2, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->net[1].w2t, w->net[0].w2t, (size_t)d.H2 * d.H1 * 2, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->net[1].w3t, w->net[0].w3t, (size_t)32 * d.H2 * 2, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { err = std::string("wide_sync_target: ") + cudaGetErrorString(e); return -1; }
    return 3;
}

#define WCK(expr)                                                                    \
    do {                                                                             \
        cudaError_t e_ = (expr);                                                     \
        if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return -1; } \
    } while (0)

static int forward_net(WideState *w, int net, const CUtensorMap &tx, const CUtensorMap &th1, const CUtensorMap &th2,
                       __nv_bfloat16 *h1, __nv_bfloat16 *h2, float *q, const float *theta, cudaStream_t st,
                       std::string &err)
{
    const Dims &d = w->d;
    const Offsets &o = w->o;
    WideNet &n = w->net[net];
    GemmArgs g{};
    g.M = w->B; g.N = d.H1; g.K = d.D; g.k_per_split = ((d.D + BK - 1) / BK) * BK;
    g.out_bf16 = h1; g.bias = theta + o.b1; g.ldo = d.H1; g.n_store = d.H1;
    WCK((launch_gemm<MODE_KK, 128, EPI_BIAS_RELU_BF16>(tx, n.tm_w1t, g, 1, st)));
    g.N = d.H2; g.K = d.H1; g.k_per_split = ((d.H1 + BK - 1) / BK) * BK;
    g.out_bf16 = h2; g.bias = theta + o.b2; g.ldo = d.H2; g.n_store = d.H2;
    WCK((launch_gemm<MODE_KK, 128, EPI_BIAS_RELU_BF16>(th1, n.tm_w2t, g, 1, st)));
    g.N = 32; g.K = d.H2; g.k_per_split = ((d.H2 + BK - 1) / BK) * BK;
    g.out_bf16 = nullptr; g.out_f32 = q; g.bias = theta + o.b3; g.ldo = d.A; g.n_store = d.A;
    WCK((launch_gemm<MODE_KK, 32, EPI_BIAS_F32>(th2, n.tm_w3t, g, 1, st)));
    return 3;
}

int wide_forward_backward(WideState *w, const WideStep &s, cudaStream_t st, std::string &err)
{
    const Dims &d = w->d;
    const Offsets &o = w->o;
    const int B = w->B;
    int launches = 0;
    {
        int n = B * (d.D >> 2);
        wide_gather_kernel<<<(n + 255) / 256, 256, 0, st>>>(s.rp, d.D, s.idx, B, w->x, w->x2, w->ba, w->br, w->bdone);
        WCK(cudaGetLastError());
        launches++;
    }
    int k = forward_net(w, 0, w->tm_x_a, w->tm_h1_a, w->tm_h2_a, w->h1, w->h2, w->q, s.theta, st, err);
    if (k < 0) return -1;
    launches += k;
    k = forward_net(w, 1, w->tm_x2_a, w->tm_t1_a, w->tm_t2_a, w->t1, w->t2, w->qt, s.theta_t, st, err);
    if (k < 0) return -1;
    launches += k;
    {
        size_t smem = sizeof(float) * ((size_t)2 * d.H2 * d.A + TD_ROWS) + sizeof(int) * TD_ROWS;
        static bool conf = false;
        if (!conf && smem > 48 * 1024) {
            WCK(cudaFuncSetAttribute(wide_td_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            conf = true;
        }
        int threads = d.H2 < 512 ? ((d.H2 + 31) / 32) * 32 : 512;
        if (threads < TD_ROWS) threads = TD_ROWS;
        wide_td_kernel<<<w->ntd, threads, smem, st>>>(w->q, w->qt, w->ba, w->br, w->bdone, B, d.H2, d.A, s.theta + o.w3,
                                                       w->h2, s.sc, w->td, w->dz2, w->td_part);
        WCK(cudaGetLastError());
        launches++;
    }
    GemmArgs g{};
    // dW2 = H1^T dZ2 : M = H1, N = H2, K = B
    g.M = d.H1; g.N = d.H2; g.K = B; g.k_per_split = w->kps2;
    g.out_f32 = w->w2_slices; g.ldo = d.H2; g.n_store = d.H2; g.slice_stride = (long long)d.H1 * d.H2;
    WCK((launch_gemm<MODE_MM, 128, EPI_F32_SLICE>(w->tm_h1_mm, w->tm_dz2_mm, g, w->ns2, st)));
    // dZ1 = (dZ2 W2^T) * mask1 : M = B, N = H1, K = H2 ; B operand = W2 [H1 rows][H2 cols]
    g = GemmArgs{};
    g.M = B; g.N = d.H1; g.K = d.H2; g.k_per_split = ((d.H2 + BK - 1) / BK) * BK;
    g.out_bf16 = w->dz1; g.mask_src = w->h1; g.ldo = d.H1; g.n_store = d.H1;
    WCK((launch_gemm<MODE_KK, 128, EPI_MASK_BF16>(w->tm_dz2_a, w->net[0].tm_w2, g, 1, st)));
    // dW1 = X^T dZ1 : M = D, N = H1, K = B
    g = GemmArgs{};
    g.M = d.D; g.N = d.H1; g.K = B; g.k_per_split = w->kps1;
    g.out_f32 = w->w1_slices; g.ldo = d.H1; g.n_store = d.H1; g.slice_stride = (long long)d.D * d.H1;
    WCK((launch_gemm<MODE_MM, 128, EPI_F32_SLICE>(w->tm_x_mm, w->tm_dz1_mm, g, w->ns1, st)));
    {
        dim3 grid((d.H1 + 63) / 64, w->nb1);
        wide_colsum_kernel<<<grid, 256, 0, st>>>(w->dz1, B, d.H1, w->b1_part);
        WCK(cudaGetLastError());
    }
    wide_reduce_kernel<<<(o.P + 1 + 31) / 32, 256, 0, st>>>(d, o, w->w1_slices, w->ns1, w->w2_slices, w->ns2, w->b1_part,
                                                           w->nb1, w->td_part, w->ntd, s.sc.count, s.flat);
    WCK(cudaGetLastError());
    launches += 5;
    return launches;
}

void wide_set_debug(WideState *w, bool on) { w->debug = on; }
void wide_debug_ptrs(WideState *w, const float **q, const float **qt, const float **td)
{
    *q = w->q; *qt = w->qt; *td = w->td;
}
const uint8_t *wide_done_ptr(WideState *w) { return w->bdone; }

// ---------------------------------------------------------------------------
// stand-alone tcgen05 GEMM self-test (diagnostic ABI: goldq_selftest_umma)
//   mode 0 (KK): C[M][N] = sum_k A[m][k] B[n][k]      A [M][K], B [N][K] bf16 (as uint16)
//   mode 1 (MM): C[M][N] = sum_k A[k][m] B[k][n]      A [K][M], B [K][N]
// ---------------------------------------------------------------------------
int wide_selftest_gemm(int mode, int M, int N, int K, int nsplit, const uint16_t *A, const uint16_t *Bm, float *C,
                       std::string &err)
{
    if (!load_encode(err)) return -1;
    if (nsplit < 1) nsplit = 1;
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
