// kernels_generic.cu — shape-agnostic fp32 CUDA-core kernels of the learn step.
//
// This is the "reference arithmetic" path: the forward pass accumulates
// k-sequentially with a separate multiply and add (__fmul_rn/__fadd_rn are never
// contracted), which reproduces gonum's Sgemm bit for bit
// (reference: vendor/gonum.org/v1/gonum/blas/gonum/sgemm.go:258-270,
// internal/asm/f32/axpyunitary_amd64.s:33-34).  Used for any shape the fused
// narrow / tcgen05 paths do not cover, for goro's FitBatch with explicit labels,
// and as the fp32 cross-check of the bf16 path.
#include "common.cuh"
#include "sampler.cuh"

namespace gq {

// ---------------------------------------------------------------------------
// K1a: device sampler — first B entries of a keyed permutation of [0, len)
// (replaces rand.Perm(Len)[:B], pkg/v1/agent/deepq/memory.go:54-69)
// ---------------------------------------------------------------------------
__global__ void sample_indices_kernel(int32_t *__restrict__ idx, int B, uint32_t len, uint64_t seed,
                                      const StepDev *__restrict__ step, int replicas)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * replicas) return;
    if (step) { seed = step->seed; len = (uint32_t)step->len; }
    int rep = t / B, b = t - rep * B;
    idx[t] = (int32_t)perm_index((uint32_t)b, len, seed + 0x9E3779B97F4A7C15ull * (uint64_t)rep);
}

void launch_sample_indices(int32_t *idx, int B, int64_t len, uint64_t seed, const StepDev *step,
                           int replicas, cudaStream_t st)
{
    int n = B * replicas;
    sample_indices_kernel<<<(n + 255) / 256, 256, 0, st>>>(idx, B, (uint32_t)len, seed, step, replicas);
}

// ---------------------------------------------------------------------------
// K1b: replay gather — rows idx[b] (i-th newest) into batch-major buffers.
// Replaces Memory.At(i) + dense.Concat (memory.go:62, pkg/v1/dense/op.go:65-72).
// One thread per 16-byte chunk of a row when D % 4 == 0.
// ---------------------------------------------------------------------------
template <bool VEC4>
__global__ void gather_kernel(Replay rp, int D, const int32_t *__restrict__ idx, int B,
                              float *__restrict__ bs, float *__restrict__ bs2,
                              int32_t *__restrict__ ba, float *__restrict__ br,
                              uint8_t *__restrict__ bdone, const StepDev *__restrict__ step)
{
    const int chunks = VEC4 ? (D >> 2) : D;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * chunks) return;
    int b = t / chunks, c = t - b * chunks;
    int64_t phys = (step ? (int64_t)step->head : rp.head) - 1 - (int64_t)idx[b];
    if (phys < 0) phys += rp.cap;
    if (VEC4) {
        const float4 *src = reinterpret_cast<const float4 *>(rp.s + phys * D);
        const float4 *src2 = reinterpret_cast<const float4 *>(rp.s2 + phys * D);
        reinterpret_cast<float4 *>(bs + (int64_t)b * D)[c] = __ldg(src + c);
        reinterpret_cast<float4 *>(bs2 + (int64_t)b * D)[c] = __ldg(src2 + c);
    } else {
        bs[(int64_t)b * D + c] = __ldg(rp.s + phys * D + c);
        bs2[(int64_t)b * D + c] = __ldg(rp.s2 + phys * D + c);
    }
    if (c == 0) {
        ba[b] = rp.a[phys];
        br[b] = rp.r[phys];
        bdone[b] = rp.done[phys];
    }
}

void launch_gather(const Replay &rp, int D, const int32_t *idx, int B, float *bs, float *bs2,
                   int32_t *ba, float *br, uint8_t *bdone, const StepDev *step, cudaStream_t st)
{
    if ((D & 3) == 0) {
        int n = B * (D >> 2);
        gather_kernel<true><<<(n + 255) / 256, 256, 0, st>>>(rp, D, idx, B, bs, bs2, ba, br, bdone, step);
    } else {
        int n = B * D;
        gather_kernel<false><<<(n + 255) / 256, 256, 0, st>>>(rp, D, idx, B, bs, bs2, ba, br, bdone, step);
    }
}

// ---------------------------------------------------------------------------
// K2: one FC layer, Y = act(X·W + b).  64x64 output tile, 16-deep k tiles,
// 256 threads with a 4x4 micro-tile each.  Accumulation order per output element
// is k = 0..in-1, multiply then add: gonum's order (see file header).
// ReLU is Rectify: mask = (a >= 0), h = a * mask
// (reference: vendor/gorgonia.org/gorgonia/nn.go:162-189).
// ---------------------------------------------------------------------------
constexpr int FC_BM = 64, FC_BN = 64, FC_BK = 16;

__global__ void __launch_bounds__(256)
fc_forward_kernel(const float *__restrict__ X, const float *__restrict__ W,
                  const float *__restrict__ bias, float *__restrict__ Y,
                  uint8_t *__restrict__ mask, int n, int in, int out, int relu)
{
    __shared__ float Xs[FC_BK][FC_BM + 4];
    __shared__ float Ws[FC_BK][FC_BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.y * FC_BM, col0 = blockIdx.x * FC_BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < in; k0 += FC_BK) {
#pragma unroll
        for (int e = tid; e < FC_BM * FC_BK; e += 256) {
            int mm = e / FC_BK, kk = e - mm * FC_BK;
            int gr = row0 + mm, gk = k0 + kk;
            Xs[kk][mm] = (gr < n && gk < in) ? X[(int64_t)gr * in + gk] : 0.f;
        }
#pragma unroll
        for (int e = tid; e < FC_BK * FC_BN; e += 256) {
            int kk = e / FC_BN, nn = e - kk * FC_BN;
            int gk = k0 + kk, gc = col0 + nn;
            Ws[kk][nn] = (gk < in && gc < out) ? W[(int64_t)gk * out + gc] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < FC_BK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = Xs[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(a[i], b[j]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int gr = row0 + ty * 4 + i;
        if (gr >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int gc = col0 + tx * 4 + j;
            if (gc >= out) continue;
            float a = __fadd_rn(acc[i][j], bias[gc]);
            if (relu) {
                float mk = (a >= 0.f) ? 1.f : 0.f;
                if (mask) mask[(int64_t)gr * out + gc] = (a >= 0.f) ? 1 : 0;
                a = __fmul_rn(a, mk);
            }
            Y[(int64_t)gr * out + gc] = a;
        }
    }
}

void launch_fc_forward(const float *X, const float *W, const float *bias, float *Y,
                       uint8_t *mask, int n, int in, int out, bool relu, cudaStream_t st)
{
    dim3 grid((out + FC_BN - 1) / FC_BN, (n + FC_BM - 1) / FC_BM);
    fc_forward_kernel<<<grid, 256, 0, st>>>(X, W, bias, Y, mask, n, in, out, relu ? 1 : 0);
}

// ---------------------------------------------------------------------------
// K3: first-max over the target row, TD target, label row and dQ.
//   nextMax = q_t[first argmax]  (ArgmaxF32, vendor/gorgonia.org/tensor/internal/
//             execution/generic_argmethods.go:211-233: strict '>', NaN/+Inf at i>=1 win)
//   t = done ? r : r + gamma*nextMax                     (agent.go:137-149)
//   y = q_online, y[a] = t  ->  d = q - y is zero off-action     (agent.go:150-155)
//   L = sum d^2 / (B*A) ; dq = (d*2) * (1/(B*A))          (goro loss.go:28-42)
// ---------------------------------------------------------------------------
__device__ __forceinline__ int first_argmax(const float *q, int A)
{
    float f = q[0];
    int mx = 0;
    for (int i = 1; i < A; i++) {
        float v = q[i];
        if (isnan(v) || (isinf(v) && v > 0.f)) return i;
        if (v > f) { mx = i; f = v; }
    }
    return mx;
}

__device__ __forceinline__ float block_sum_256(float v, float *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) r += sh[i];
    return r;
}

__global__ void __launch_bounds__(256)
td_kernel(const float *__restrict__ q, const float *__restrict__ qt,
          const int32_t *__restrict__ act, const float *__restrict__ rew,
          const uint8_t *__restrict__ done, int B, int A, StepScalars sc,
          float *__restrict__ td, float *__restrict__ dq, float *__restrict__ loss_part)
{
    __shared__ float sh[8];
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.f;
    if (b < B) {
        float t = rew[b];
        if (!done[b]) {
            const float *row = qt + (int64_t)b * A;
            float nm = row[first_argmax(row, A)];
            t = __fadd_rn(rew[b], __fmul_rn(sc.gamma, nm));
        }
        td[b] = t;
        int a = act[b];
        float d = __fsub_rn(q[(int64_t)b * A + a], t);
        float dqa;
        residual_loss(sc, d, sq, dqa);
        for (int j = 0; j < A; j++) dq[(int64_t)b * A + j] = 0.f;
        dq[(int64_t)b * A + a] = dqa;
    }
    float s = block_sum_256(sq, sh);
    if (threadIdx.x == 0) loss_part[blockIdx.x] = s;
}

void launch_td(const float *q, const float *qt, const int32_t *a, const float *r,
               const uint8_t *done, int B, int A, StepScalars sc, float *td, float *dq,
               float *loss_part, int *n_loss_part, cudaStream_t st)
{
    int blocks = (B + 255) / 256;
    *n_loss_part = blocks;
    td_kernel<<<blocks, 256, 0, st>>>(q, qt, a, r, done, B, A, sc, td, dq, loss_part);
}

__global__ void __launch_bounds__(256)
label_loss_kernel(const float *__restrict__ q, const float *__restrict__ y, int total,
                  StepScalars sc, float *__restrict__ dq, float *__restrict__ loss_part)
{
    __shared__ float sh[8];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.f;
    if (i < total) {
        float d = __fsub_rn(q[i], y[i]);
        float dqi;
        residual_loss(sc, d, sq, dqi);
        dq[i] = dqi;
    }
    float s = block_sum_256(sq, sh);
    if (threadIdx.x == 0) loss_part[blockIdx.x] = s;
}

void launch_label_loss(const float *q, const float *y, int n, int A, StepScalars sc, float *dq,
                       float *loss_part, int *n_loss_part, cudaStream_t st)
{
    int total = n * A;
    int blocks = (total + 255) / 256;
    *n_loss_part = blocks;
    label_loss_kernel<<<blocks, 256, 0, st>>>(q, y, total, sc, dq, loss_part);
}

// ---------------------------------------------------------------------------
// K4a: weight/bias gradient slices.  dW = X^T·dZ, db = column sums of dZ
// (reference: vendor/gorgonia.org/gorgonia/operatorLinAlg.go:92-101 and
// op_tensor.go:281-289).  Split over the batch into `nsplit` row ranges; each
// block writes its own slice, reduced later in a fixed order (deterministic).
// ---------------------------------------------------------------------------
constexpr int GW_BI = 64, GW_BJ = 64, GW_BR = 16;

__global__ void __launch_bounds__(256)
grad_w_kernel(const float *__restrict__ X, const float *__restrict__ dZ, int n, int in, int out,
              float *__restrict__ gslice, int64_t slice_stride, int nsplit, int w_off, int b_off)
{
    __shared__ float Xs[GW_BR][GW_BI];
    __shared__ float Zs[GW_BR][GW_BJ];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int j0 = blockIdx.x * GW_BJ, i0 = blockIdx.y * GW_BI, z = blockIdx.z;
    const int rows_per = (n + nsplit - 1) / nsplit;
    const int r_begin = z * rows_per;
    const int r_end = min(n, r_begin + rows_per);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    float bsum = 0.f;  // threads 0..63 of blockIdx.y == 0 own one bias column each

    for (int r0 = r_begin; r0 < r_end; r0 += GW_BR) {
#pragma unroll
        for (int e = tid; e < GW_BR * GW_BI; e += 256) {
            int rr = e / GW_BI, ii = e - rr * GW_BI;
            int gr = r0 + rr, gi = i0 + ii;
            Xs[rr][ii] = (gr < r_end && gi < in) ? X[(int64_t)gr * in + gi] : 0.f;
        }
#pragma unroll
        for (int e = tid; e < GW_BR * GW_BJ; e += 256) {
            int rr = e / GW_BJ, jj = e - rr * GW_BJ;
            int gr = r0 + rr, gj = j0 + jj;
            Zs[rr][jj] = (gr < r_end && gj < out) ? dZ[(int64_t)gr * out + gj] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < GW_BR; rr++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = Xs[rr][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Zs[rr][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (blockIdx.y == 0 && tid < GW_BJ) {
#pragma unroll
            for (int rr = 0; rr < GW_BR; rr++) bsum += Zs[rr][tid];
        }
        __syncthreads();
    }
    float *g = gslice + (int64_t)z * slice_stride;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int gi = i0 + ty * 4 + i;
        if (gi >= in) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int gj = j0 + tx * 4 + j;
            if (gj < out) g[w_off + (int64_t)gi * out + gj] = acc[i][j];
        }
    }
    if (blockIdx.y == 0 && tid < GW_BJ && j0 + tid < out) g[b_off + j0 + tid] = bsum;
}

void launch_grad_w(const float *X, const float *dZ, int n, int in, int out, float *gslice,
                   int64_t slice_stride, int nsplit, int w_off, int b_off, cudaStream_t st)
{
    dim3 grid((out + GW_BJ - 1) / GW_BJ, (in + GW_BI - 1) / GW_BI, nsplit);
    grad_w_kernel<<<grid, 256, 0, st>>>(X, dZ, n, in, out, gslice, slice_stride, nsplit, w_off,
                                        b_off);
}

// ---------------------------------------------------------------------------
// K4b: input gradient, dX = (dZ·W^T) ⊙ mask   (operatorLinAlg.go:88-91; the mask
// factor is HadamardProd's derivative, operatorPointwise_binary.go:809-829).
// ---------------------------------------------------------------------------
constexpr int GX_BM = 64, GX_BN = 64, GX_BK = 16;

__global__ void __launch_bounds__(256)
grad_x_kernel(const float *__restrict__ dZ, const float *__restrict__ W,
              const uint8_t *__restrict__ mask, float *__restrict__ dX, int n, int in, int out)
{
    __shared__ float Zs[GX_BK][GX_BM + 4];  // [k=j][row]
    __shared__ float Ws[GX_BK][GX_BN + 4];  // [k=j][i]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.y * GX_BM, i0 = blockIdx.x * GX_BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < out; k0 += GX_BK) {
#pragma unroll
        for (int e = tid; e < GX_BM * GX_BK; e += 256) {
            int mm = e / GX_BK, kk = e - mm * GX_BK;
            int gr = row0 + mm, gk = k0 + kk;
            Zs[kk][mm] = (gr < n && gk < out) ? dZ[(int64_t)gr * out + gk] : 0.f;
        }
#pragma unroll
        for (int e = tid; e < GX_BN * GX_BK; e += 256) {
            int ii = e / GX_BK, kk = e - ii * GX_BK;
            int gi = i0 + ii, gk = k0 + kk;
            Ws[kk][ii] = (gi < in && gk < out) ? W[(int64_t)gi * out + gk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GX_BK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = Zs[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int gr = row0 + ty * 4 + i;
        if (gr >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int gi = i0 + tx * 4 + j;
            if (gi >= in) continue;
            float mk = mask[(int64_t)gr * in + gi] ? 1.f : 0.f;
            dX[(int64_t)gr * in + gi] = __fmul_rn(acc[i][j], mk);
        }
    }
}

void launch_grad_x(const float *dZ, const float *W, const uint8_t *mask, float *dX, int n, int in,
                   int out, cudaStream_t st)
{
    dim3 grid((in + GX_BN - 1) / GX_BN, (n + GX_BM - 1) / GX_BM);
    grad_x_kernel<<<grid, 256, 0, st>>>(dZ, W, mask, dX, n, in, out);
}

// ---------------------------------------------------------------------------
// Slice reduction: flat[i] = sum_z slice[z][i] in z order; flat[P] = loss.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reduce_slices_kernel(const float *__restrict__ gslice, int64_t slice_stride, int nsplit, int P,
                     const float *__restrict__ loss_part, int n_loss_part, float count,
                     float *__restrict__ flat)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        float s = 0.f;
        for (int z = 0; z < nsplit; z++) s += gslice[(int64_t)z * slice_stride + i];
        flat[i] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        float s = 0.f;
        for (int k = threadIdx.x; k < n_loss_part; k += 32) s += loss_part[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) flat[P] = __fdiv_rn(s, count);
    }
}

void launch_reduce_slices(const float *gslice, int64_t slice_stride, int nsplit, int P,
                          const float *loss_part, int n_loss_part, float count, float *flat,
                          cudaStream_t st)
{
    reduce_slices_kernel<<<(P + 255) / 256, 256, 0, st>>>(gslice, slice_stride, nsplit, P, loss_part,
                                                         n_loss_part, count, flat);
}

// ---------------------------------------------------------------------------
// K5: fused Adam, gorgonia v0.9.9 semantics, one pass (read w,g,m,v; write w,m,v).
// Reference: vendor/gorgonia.org/gorgonia/solvers.go:551-615.  The weight receives
// BOTH u/(sqrt(vhat)+eps) (Div WithIncr(w), :604) and u (Add(w, mHats), :611),
// u = -eta*mhat.  Every operation is a separately rounded fp32 op, as in the
// reference's whole-tensor passes.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ w, const float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, int64_t n, AdamConsts c, const StepDev *__restrict__ step)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (step) { c.c1 = step->c1; c.c2 = step->c2; }
    float wi = w[i], mi = m[i], vi = v[i];
    solver_update(wi, g[i], mi, vi, c);
    w[i] = wi;
    m[i] = mi;
    v[i] = vi;
}

void launch_adam(float *w, const float *g, float *m, float *v, int64_t n, AdamConsts c,
                 const StepDev *step, cudaStream_t st)
{
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, g, m, v, n, c, step);
}

__global__ void argmax_rows_kernel(const float *__restrict__ q, int n, int A,
                                   int32_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = first_argmax(q + (int64_t)i * A, A);
}

void launch_argmax_rows(const float *q, int n, int A, int32_t *out, cudaStream_t st)
{
    argmax_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(q, n, A, out);
}

// batched epsilon-greedy: first-argmax of the online Q unless row i's draw explores
__global__ void eps_greedy_rows_kernel(const float *__restrict__ q, int n, int A, float epsilon, uint64_t seed,
                                       int32_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ra;
    out[i] = explore_draw((uint32_t)i, seed, epsilon, A, &ra) ? ra : first_argmax(q + (int64_t)i * A, A);
}

void launch_eps_greedy_rows(const float *q, int n, int A, float epsilon, uint64_t seed, int32_t *out, cudaStream_t st)
{
    eps_greedy_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(q, n, A, epsilon, seed, out);
}

}  // namespace gq
