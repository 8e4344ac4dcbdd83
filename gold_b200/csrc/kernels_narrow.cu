// kernels_narrow.cu — the whole learn step of a classic-control DQN in ONE kernel.
//
// Shapes like CartPole's 4->24->24->2 (770 parameters, 5.6 kFLOP per transition)
// are launch-latency bound, not HBM or FLOP bound: the step moves 205 KB at
// B = 4096.  So the design goal is one launch and no intermediate HBM traffic:
//
//   * one warp per sampled row, lane j = neuron j of the current layer; the
//     activations of the previous layer are broadcast with warp shuffles, both
//     networks' weights live in shared memory (6 KB), nothing but the replay row
//     is read from HBM;
//   * the forward pass keeps gonum's arithmetic (k-sequential, multiply then add,
//     reference: vendor/gonum.org/v1/gonum/blas/gonum/sgemm.go:258-270), so Q(s),
//     Q'(s'), the first-max and the TD target are bit-identical to the CPU path;
//   * each lane accumulates "its" column of every weight gradient in registers
//     over the rows its warp handles: no atomics, no cross-lane reduction;
//   * warps are reduced through shared memory, CTAs through a [ctas][P+1] scratch
//     in a fixed order (deterministic); the last CTA to arrive sums the scratch,
//     finishes the loss and applies the fused Adam update in the same launch.
//
// Replicas (independent agents) are blockIdx.y: same kernel, own
// weights / solver state / replay ring.
#include "common.cuh"
#include "sampler.cuh"

namespace gq {

namespace {

__device__ __forceinline__ void adam_update_n(float &w, float g, float &m, float &v,
                                              const AdamConsts &c)
{
    // reference: vendor/gorgonia.org/gorgonia/solvers.go:551-615 (see kernels_generic.cu K5)
    float t1 = __fmul_rn(g, c.omb1);
    float g2 = __fmul_rn(__fmul_rn(g, g), c.omb2);
    m = __fadd_rn(t1, __fmul_rn(c.beta1, m));
    v = __fadd_rn(g2, __fmul_rn(c.beta2, v));
    float mh = __fmul_rn(m, c.c1);
    float vh = __fmul_rn(v, c.c2);
    float den = __fadd_rn(__fsqrt_rn(vh), c.eps);
    float u = __fmul_rn(mh, c.neg_eta);
    if (den == 0.f) w = __int_as_float(0x7f800000);
    else w = __fadd_rn(w, __fdiv_rn(u, den));
    w = __fadd_rn(w, u);
}

constexpr unsigned FULL = 0xffffffffu;

template <int D, int H1, int H2, int A>
__global__ void __launch_bounds__(512)
narrow_learn_kernel(NarrowArgs p)
{
    static_assert(D <= 32 && H1 <= 32 && H2 <= 32 && A <= 32, "lane-per-neuron layout");
    constexpr int oW1 = 0, oB1 = oW1 + D * H1, oW2 = oB1 + H1, oB2 = oW2 + H1 * H2,
                  oW3 = oB2 + H2, oB3 = oW3 + H2 * A, P = oB3 + A;
    constexpr int P1 = P + 1;

    extern __shared__ float smem[];
    float *sOn = smem;             // online parameters, flat
    float *sTg = sOn + P;          // target parameters, flat
    float *sW2T = sTg + P;         // online W2 transposed [H2][H1] (for dh1)
    float *sPart = sW2T + H1 * H2; // [warps][P1] per-warp gradient partials
    __shared__ int sIsLast;

    const int rep = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int64_t pofs = (int64_t)rep * P;

    for (int i = tid; i < P; i += blockDim.x) {
        sOn[i] = p.theta[pofs + i];
        sTg[i] = p.theta_t[pofs + i];
    }
    for (int i = tid; i < H1 * H2; i += blockDim.x) {
        int r = i / H2, c = i - r * H2;  // W2[r][c]
        sW2T[c * H1 + r] = p.theta[pofs + oW2 + i];
    }
    __syncthreads();

    // replay views of this replica
    const float *rs = p.rp.s + (int64_t)rep * p.rp.cap * D;
    const float *rs2 = p.rp.s2 + (int64_t)rep * p.rp.cap * D;
    const int32_t *ra = p.rp.a + (int64_t)rep * p.rp.cap;
    const float *rr = p.rp.r + (int64_t)rep * p.rp.cap;
    const uint8_t *rd = p.rp.done + (int64_t)rep * p.rp.cap;

    const bool in1 = lane < H1, in2 = lane < H2, inA = lane < A;
    const int j1 = in1 ? lane : 0, j2 = in2 ? lane : 0, jA = inA ? lane : 0;

    float gW1[D], gW2[H1], gW3[A];
    float gb1 = 0.f, gb2 = 0.f, gb3 = 0.f, lossacc = 0.f;
#pragma unroll
    for (int k = 0; k < D; k++) gW1[k] = 0.f;
#pragma unroll
    for (int k = 0; k < H1; k++) gW2[k] = 0.f;
#pragma unroll
    for (int k = 0; k < A; k++) gW3[k] = 0.f;

    const int total_warps = gridDim.x * nwarps;
    for (int b = blockIdx.x * nwarps + warp; b < p.B; b += total_warps) {
        // ---- K1: sample / gather -------------------------------------------------
        int32_t li;
        if (p.idx) li = p.idx[(int64_t)rep * p.B + b];
        else li = (int32_t)perm_index((uint32_t)b, (uint32_t)p.rp.len,
                                      p.seed + 0x9E3779B97F4A7C15ull * (uint64_t)rep);
        int64_t phys = p.rp.head - 1 - (int64_t)li;
        if (phys < 0) phys += p.rp.cap;
        float x = 0.f, x2 = 0.f;
        if (lane < D) {
            x = __ldg(rs + phys * D + lane);
            x2 = __ldg(rs2 + phys * D + lane);
        }
        const int act = __ldg(ra + phys);
        const float rew = __ldg(rr + phys);
        const int dn = __ldg(rd + phys);

        // ---- K2: forward, online on s and target on s' interleaved ---------------
        float z = 0.f, zt = 0.f;
#pragma unroll
        for (int k = 0; k < D; k++) {
            float xk = __shfl_sync(FULL, x, k), xk2 = __shfl_sync(FULL, x2, k);
            z = __fadd_rn(z, __fmul_rn(xk, sOn[oW1 + k * H1 + j1]));
            zt = __fadd_rn(zt, __fmul_rn(xk2, sTg[oW1 + k * H1 + j1]));
        }
        float a1 = __fadd_rn(z, sOn[oB1 + j1]);
        float a1t = __fadd_rn(zt, sTg[oB1 + j1]);
        // Rectify: h = a * (a >= 0 ? 1 : 0)   (vendor/gorgonia.org/gorgonia/nn.go:162-189)
        const float mk1 = (in1 && a1 >= 0.f) ? 1.f : 0.f;
        float h1 = in1 ? __fmul_rn(a1, mk1) : 0.f;
        float h1t = in1 ? __fmul_rn(a1t, (a1t >= 0.f) ? 1.f : 0.f) : 0.f;

        z = 0.f; zt = 0.f;
#pragma unroll
        for (int k = 0; k < H1; k++) {
            float hk = __shfl_sync(FULL, h1, k), hkt = __shfl_sync(FULL, h1t, k);
            z = __fadd_rn(z, __fmul_rn(hk, sOn[oW2 + k * H2 + j2]));
            zt = __fadd_rn(zt, __fmul_rn(hkt, sTg[oW2 + k * H2 + j2]));
        }
        float a2 = __fadd_rn(z, sOn[oB2 + j2]);
        float a2t = __fadd_rn(zt, sTg[oB2 + j2]);
        const float mk2 = (in2 && a2 >= 0.f) ? 1.f : 0.f;
        float h2 = in2 ? __fmul_rn(a2, mk2) : 0.f;
        float h2t = in2 ? __fmul_rn(a2t, (a2t >= 0.f) ? 1.f : 0.f) : 0.f;

        z = 0.f; zt = 0.f;
#pragma unroll
        for (int k = 0; k < H2; k++) {
            float hk = __shfl_sync(FULL, h2, k), hkt = __shfl_sync(FULL, h2t, k);
            z = __fadd_rn(z, __fmul_rn(hk, sOn[oW3 + k * A + jA]));
            zt = __fadd_rn(zt, __fmul_rn(hkt, sTg[oW3 + k * A + jA]));
        }
        const float q = __fadd_rn(z, sOn[oB3 + jA]);     // lane a < A holds Q(s)[a]
        const float qt = __fadd_rn(zt, sTg[oB3 + jA]);   // lane a < A holds Q'(s')[a]

        // ---- K3: first-max (ArgmaxF32 rules), TD target, dQ ----------------------
        float f = __shfl_sync(FULL, qt, 0);
        bool decided = false;
#pragma unroll
        for (int i = 1; i < A; i++) {
            float v = __shfl_sync(FULL, qt, i);
            if (!decided) {
                if (isnan(v) || (isinf(v) && v > 0.f)) { f = v; decided = true; }
                else if (v > f) f = v;
            }
        }
        float t = rew;
        if (!dn) t = __fadd_rn(rew, __fmul_rn(p.sc.gamma, f));   // agent.go:149
        const float qa = __shfl_sync(FULL, q, act);
        const float d = __fsub_rn(qa, t);
        lossacc = __fadd_rn(lossacc, __fmul_rn(d, d));
        const float dqv = __fmul_rn(__fmul_rn(d, 2.f), p.sc.inv_count);

        if (p.dbg_q) {
            if (inA) {
                p.dbg_q[((int64_t)rep * p.B + b) * A + lane] = q;
                p.dbg_qt[((int64_t)rep * p.B + b) * A + lane] = dn ? 0.f : qt;
            }
            if (lane == 0) {
                p.dbg_td[(int64_t)rep * p.B + b] = t;
                p.dbg_idx[(int64_t)rep * p.B + b] = li;
            }
        }

        // ---- K4: backward --------------------------------------------------------
        // layer 3: dW3[j][a] = h2[j]*dq[a]; db3 = dq; dh2[j] = dq[act]*W3[j][act]
        if (lane == act) gb3 += dqv;
#pragma unroll
        for (int a = 0; a < A; a++) gW3[a] += (a == act) ? h2 * dqv : 0.f;
        const float dz2 = __fmul_rn(__fmul_rn(dqv, sOn[oW3 + j2 * A + act]), mk2);
        // layer 2: dW2[i][j] += h1[i]*dz2[j]; dh1[i] = sum_j dz2[j]*W2[i][j]
        gb2 += dz2;
        float dh1 = 0.f;
#pragma unroll
        for (int k = 0; k < H1; k++) gW2[k] = fmaf(__shfl_sync(FULL, h1, k), dz2, gW2[k]);
#pragma unroll
        for (int k = 0; k < H2; k++) dh1 = fmaf(__shfl_sync(FULL, dz2, k), sW2T[k * H1 + j1], dh1);
        const float dz1 = __fmul_rn(dh1, mk1);
        // layer 1: dW1[k][j] += x[k]*dz1[j]
        gb1 += dz1;
#pragma unroll
        for (int k = 0; k < D; k++) gW1[k] = fmaf(__shfl_sync(FULL, x, k), dz1, gW1[k]);
    }

    // ---- warp partials -> shared ---------------------------------------------------
    float *mine = sPart + warp * P1;
    if (in1) {
#pragma unroll
        for (int k = 0; k < D; k++) mine[oW1 + k * H1 + lane] = gW1[k];
        mine[oB1 + lane] = gb1;
    }
    if (in2) {
#pragma unroll
        for (int k = 0; k < H1; k++) mine[oW2 + k * H2 + lane] = gW2[k];
        mine[oB2 + lane] = gb2;
#pragma unroll
        for (int a = 0; a < A; a++) mine[oW3 + lane * A + a] = gW3[a];
    }
    if (inA) mine[oB3 + lane] = gb3;
    if (lane == 0) mine[P] = lossacc;
    __syncthreads();

    const int ctas = gridDim.x;
    float *gpart = p.partial + ((int64_t)rep * ctas + blockIdx.x) * P1;
    float *flat = p.flat + (int64_t)rep * P1;
    if (ctas > 1) {
        for (int i = tid; i < P1; i += blockDim.x) {
            float s = 0.f;
            for (int w = 0; w < nwarps; w++) s += sPart[w * P1 + i];
            gpart[i] = s;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            unsigned tk = atomicAdd(p.ticket + rep, 1u);
            sIsLast = (tk == (unsigned)(ctas - 1));
        }
        __syncthreads();
        if (!sIsLast) return;
        __threadfence();
    }

    // ---- last CTA of this replica: reduce, loss, K5 Adam ----------------------------
    // Cross-CTA sum in a fixed order, with the loads spread over the whole CTA:
    // warp w sums the partials of CTAs w, w+nwarps, ... (independent coalesced loads),
    // then the per-warp sums are added in warp order.
    if (ctas > 1) {
        const float *allpart = p.partial + (int64_t)rep * ctas * P1;
        for (int i = lane; i < P1; i += 32) {
            float g = 0.f;
            for (int c = warp; c < ctas; c += nwarps) g += __ldcg(allpart + (int64_t)c * P1 + i);
            sPart[warp * P1 + i] = g;
        }
        __syncthreads();
    }
    for (int i = tid; i < P1; i += blockDim.x) {
        float g = 0.f;
        for (int w = 0; w < nwarps; w++) g += sPart[w * P1 + i];
        if (i == P) {
            g = __fdiv_rn(g, p.sc.count);
            p.loss[rep] = g;
            flat[P] = g;
        } else {
            flat[i] = g;
            if (p.fuse_adam) {
                float w = sOn[i], m = p.m[pofs + i], v = p.v[pofs + i];
                adam_update_n(w, g, m, v, p.adam);
                p.theta[pofs + i] = w;
                p.m[pofs + i] = m;
                p.v[pofs + i] = v;
            }
        }
    }
    if (tid == 0 && ctas > 1) p.ticket[rep] = 0u;
}

template <int D, int H1, int H2, int A>
void launch_t(const NarrowArgs &a, cudaStream_t st)
{
    constexpr int P = D * H1 + H1 + H1 * H2 + H2 + H2 * A + A;
    size_t smem = sizeof(float) * (2 * P + H1 * H2 + (size_t)a.warps_per_cta * (P + 1));
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(narrow_learn_kernel<D, H1, H2, A>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(a.ctas_per_replica, a.replicas);
    narrow_learn_kernel<D, H1, H2, A><<<grid, a.warps_per_cta * 32, smem, st>>>(a);
}

}  // namespace

// Instantiated for gym's classic-control observation/action shapes behind
// deepq.DefaultFCLayerBuilder's 24-24 hidden layers (pkg/v1/agent/deepq/policy.go:53-59):
// CartPole (4,2), MountainCar (2,3), Acrobot (6,3), LunarLander (8,4).
bool narrow_supported(const Dims &d)
{
    if (d.H1 != 24 || d.H2 != 24) return false;
    return (d.D == 4 && d.A == 2) || (d.D == 2 && d.A == 3) || (d.D == 6 && d.A == 3) ||
           (d.D == 8 && d.A == 4);
}

void narrow_plan(NarrowArgs &a, int sm_count)
{
    int need = (a.B + 1) / 2;  // two rows per warp
    int max_ctas = (2 * sm_count) / (a.replicas > 0 ? a.replicas : 1);
    if (max_ctas < 1) max_ctas = 1;
    int ctas = (need + 15) / 16;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    int warps = (need + ctas - 1) / ctas;
    if (warps > 16) warps = 16;
    if (warps < 1) warps = 1;
    a.ctas_per_replica = ctas;
    a.warps_per_cta = warps;
}

void launch_narrow_learn(const NarrowArgs &a, cudaStream_t st)
{
    const Dims &d = a.dims;
    if (d.D == 4 && d.A == 2) launch_t<4, 24, 24, 2>(a, st);
    else if (d.D == 2 && d.A == 3) launch_t<2, 24, 24, 3>(a, st);
    else if (d.D == 6 && d.A == 3) launch_t<6, 24, 24, 3>(a, st);
    else if (d.D == 8 && d.A == 4) launch_t<8, 24, 24, 4>(a, st);
}

}  // namespace gq
