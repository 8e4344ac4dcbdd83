// kernels_narrow.cu — the whole learn step of a classic-control DQN in ONE kernel.
//
// Shapes like CartPole's 4->24->24->2 (770 parameters, 5.6 kFLOP per transition)
// are launch-latency bound, not HBM or FLOP bound: the step moves 205 KB at
// B = 4096.  So the design goal is one launch and no intermediate HBM traffic:
//
//   * one warp per sampled row (two rows in flight per warp for latency hiding),
//     lane j = neuron j of the current layer; the activations of the previous layer
//     are broadcast with warp shuffles, both networks' weights live in shared
//     memory (6 KB), nothing but the replay row is read from HBM;
//   * the forward pass keeps gonum's arithmetic (k-sequential, multiply then add,
//     reference: vendor/gonum.org/v1/gonum/blas/gonum/sgemm.go:258-270), so Q(s),
//     Q'(s'), the first-max and the TD target are bit-identical to the CPU path;
//   * each lane accumulates "its" column of every weight gradient in registers
//     over the rows its warp handles: no atomics, no cross-lane reduction;
//   * warps are reduced through shared memory; CTAs through global scratch in two
//     levels of "last arriver sums a fixed group in a fixed order" (deterministic):
//     groups of NARROW_GROUP CTAs, then the groups.  The CTA that finishes last also
//     finishes the loss and applies the fused Adam update in the same launch.
//
// Replicas (independent agents) are blockIdx.y: same kernel, own
// weights / solver state / replay ring.
#include "common.cuh"
#include "sampler.cuh"

namespace gq {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NR = 2;          // rows in flight per warp
constexpr int MIN_WARPS = 4;   // the tail keeps ELEMS sums per thread in registers

template <int D, int H1, int H2, int A>
__global__ void __launch_bounds__(512)
narrow_learn_kernel(NarrowArgs p)
{
    static_assert(D <= 32 && H1 <= 32 && H2 <= 32 && A <= 32, "lane-per-neuron layout");
    constexpr int oW1 = 0, oB1 = oW1 + D * H1, oW2 = oB1 + H1, oB2 = oW2 + H1 * H2,
                  oW3 = oB2 + H2, oB3 = oW3 + H2 * A, P = oB3 + A;
    constexpr int P1 = P + 1;
    constexpr int PE = (P + 3) & ~3;   // 16-byte aligned smem slabs
    constexpr int GROUP = NARROW_GROUP;
    constexpr int ELEMS = (P1 + MIN_WARPS * 32 - 1) / (MIN_WARPS * 32);

    extern __shared__ float smem[];
    float *sOn = smem;             // online parameters, flat          (P rounded up to even)
    float *sTg = sOn + PE;         // target parameters, flat
    float *sM = sTg + PE;          // Adam first moments  (prefetched for the CTA that finishes last)
    float *sV = sM + PE;           // Adam second moments
    float *sW2T = sV + PE;         // online W2 transposed [H2][H1] (for dh1)
    float *sPart = sW2T + H1 * H2; // [warps][P1] per-warp gradient partials
    __shared__ int sFlag;

    if (p.step) {   // graph replay: per-step scalars live in device memory
        p.seed = p.step->seed;
        p.adam.c1 = p.step->c1;
        p.adam.c2 = p.step->c2;
        p.rp.head = p.step->head;
        p.rp.len = p.step->len;
    }
    const int rep = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    // shuffled from lane 0 so that the compiler knows it is warp-uniform (no
    // convergence barriers around the shuffles in the row loop)
    const int warp = __shfl_sync(FULL, tid >> 5, 0);
    const int nwarps = blockDim.x >> 5;
    const int64_t pofs = (int64_t)rep * P;

    // Stage both networks and the Adam moments with cp.async (4-byte granules: P is odd for
    // some shapes, so a replica's slab is only 4-byte aligned); the copies fly while the warps
    // compute their sample indices and fetch their replay rows.
    {
        const float *srcs[4] = {p.theta + pofs, p.theta_t + pofs, p.m + pofs, p.v + pofs};
        float *dsts[4] = {sOn, sTg, sM, sV};
#pragma unroll
        for (int a = 0; a < 4; a++)
            for (int i = tid; i < P; i += blockDim.x)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dsts[a] + i)),
                             "l"(srcs[a] + i)
                             : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

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
    const int gw = blockIdx.x * nwarps + warp;

    // ---- K1: sample / gather (one call fetches the NR rows of an iteration) ---------------
    float x[NR], x2[NR], rew[NR];
    int act[NR], dn[NR], li[NR];
    bool live[NR];
    auto gather = [&](int b0) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const int b = b0 + r * total_warps;
            live[r] = b < p.B;
            const int bb = live[r] ? b : (b0 < p.B ? b0 : 0);
            if (p.idx) li[r] = p.idx[(int64_t)rep * p.B + bb];
            else li[r] = (int32_t)perm_index((uint32_t)bb, (uint32_t)p.rp.len,
                                             p.seed + 0x9E3779B97F4A7C15ull * (uint64_t)rep);
        }
#pragma unroll
        for (int r = 0; r < NR; r++) {
            int64_t phys = p.rp.head - 1 - (int64_t)li[r];
            if (phys < 0) phys += p.rp.cap;
            x[r] = 0.f; x2[r] = 0.f;
            if (lane < D) {
                x[r] = __ldg(rs + phys * D + lane);
                x2[r] = __ldg(rs2 + phys * D + lane);
            }
            act[r] = __ldg(ra + phys);
            rew[r] = __ldg(rr + phys);
            dn[r] = __ldg(rd + phys);
        }
    };
    if (gw < p.B) gather(gw);     // replay rows are in flight while the parameters land

    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int i = tid; i < H1 * H2; i += blockDim.x) {
        int r = i / H2, c = i - r * H2;  // W2[r][c]
        sW2T[c * H1 + r] = sOn[oW2 + i];
    }
    __syncthreads();

    // warp `gw` owns rows gw, gw + total_warps, ...; NR of them are processed together
    for (int b0 = gw; b0 < p.B; b0 += NR * total_warps) {
        if (b0 != gw) gather(b0);
        // ---- K2: forward, online on s and target on s' (2*NR independent chains) ----
        float z[NR], zt[NR], h1[NR], h1t[NR], mk1[NR], h2[NR], h2t[NR], mk2[NR], q[NR], qt[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) { z[r] = 0.f; zt[r] = 0.f; }
#pragma unroll
        for (int k = 0; k < D; k++) {
            const float w = sOn[oW1 + k * H1 + j1], wt = sTg[oW1 + k * H1 + j1];
#pragma unroll
            for (int r = 0; r < NR; r++) {
                z[r] = __fadd_rn(z[r], __fmul_rn(__shfl_sync(FULL, x[r], k), w));
                zt[r] = __fadd_rn(zt[r], __fmul_rn(__shfl_sync(FULL, x2[r], k), wt));
            }
        }
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const float a1 = __fadd_rn(z[r], sOn[oB1 + j1]);
            const float a1t = __fadd_rn(zt[r], sTg[oB1 + j1]);
            // Rectify: h = a * (a >= 0 ? 1 : 0)   (vendor/gorgonia.org/gorgonia/nn.go:162-189)
            mk1[r] = (in1 && a1 >= 0.f) ? 1.f : 0.f;
            h1[r] = in1 ? __fmul_rn(a1, mk1[r]) : 0.f;
            h1t[r] = in1 ? __fmul_rn(a1t, (a1t >= 0.f) ? 1.f : 0.f) : 0.f;
            z[r] = 0.f; zt[r] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < H1; k++) {
            const float w = sOn[oW2 + k * H2 + j2], wt = sTg[oW2 + k * H2 + j2];
#pragma unroll
            for (int r = 0; r < NR; r++) {
                z[r] = __fadd_rn(z[r], __fmul_rn(__shfl_sync(FULL, h1[r], k), w));
                zt[r] = __fadd_rn(zt[r], __fmul_rn(__shfl_sync(FULL, h1t[r], k), wt));
            }
        }
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const float a2 = __fadd_rn(z[r], sOn[oB2 + j2]);
            const float a2t = __fadd_rn(zt[r], sTg[oB2 + j2]);
            mk2[r] = (in2 && a2 >= 0.f) ? 1.f : 0.f;
            h2[r] = in2 ? __fmul_rn(a2, mk2[r]) : 0.f;
            h2t[r] = in2 ? __fmul_rn(a2t, (a2t >= 0.f) ? 1.f : 0.f) : 0.f;
            z[r] = 0.f; zt[r] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < H2; k++) {
            const float w = sOn[oW3 + k * A + jA], wt = sTg[oW3 + k * A + jA];
#pragma unroll
            for (int r = 0; r < NR; r++) {
                z[r] = __fadd_rn(z[r], __fmul_rn(__shfl_sync(FULL, h2[r], k), w));
                zt[r] = __fadd_rn(zt[r], __fmul_rn(__shfl_sync(FULL, h2t[r], k), wt));
            }
        }
#pragma unroll
        for (int r = 0; r < NR; r++) {
            q[r] = __fadd_rn(z[r], sOn[oB3 + jA]);    // lane a < A holds Q(s)[a]
            qt[r] = __fadd_rn(zt[r], sTg[oB3 + jA]);  // lane a < A holds Q'(s')[a]
        }

        // ---- K3: first-max (ArgmaxF32 rules), TD target, dQ ----------------------
        float dqv[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) {
            float f = __shfl_sync(FULL, qt[r], 0);
            bool decided = false;
#pragma unroll
            for (int i = 1; i < A; i++) {
                float v = __shfl_sync(FULL, qt[r], i);
                if (!decided) {
                    if (isnan(v) || (isinf(v) && v > 0.f)) { f = v; decided = true; }
                    else if (v > f) f = v;
                }
            }
            float t = rew[r];
            if (!dn[r]) t = __fadd_rn(rew[r], __fmul_rn(p.sc.gamma, f));   // agent.go:149
            const float qa = __shfl_sync(FULL, q[r], act[r]);
            const float d = __fsub_rn(qa, t);
            float term, dqr;
            residual_loss(p.sc, d, term, dqr);
            dqv[r] = live[r] ? dqr : 0.f;
            if (live[r]) lossacc = __fadd_rn(lossacc, term);
            if (p.dbg_q && live[r]) {
                const int b = b0 + r * total_warps;
                if (inA) {
                    p.dbg_q[((int64_t)rep * p.B + b) * A + lane] = q[r];
                    p.dbg_qt[((int64_t)rep * p.B + b) * A + lane] = dn[r] ? 0.f : qt[r];
                }
                if (lane == 0) {
                    p.dbg_td[(int64_t)rep * p.B + b] = t;
                    p.dbg_idx[(int64_t)rep * p.B + b] = li[r];
                }
            }
        }

        // ---- K4: backward (a dead row has dqv = 0 and contributes exact zeros) -----
#pragma unroll
        for (int r = 0; r < NR; r++) {
            // layer 3: dW3[j][a] = h2[j]*dq[a]; db3 = dq; dh2[j] = dq[act]*W3[j][act]
            if (lane == act[r]) gb3 += dqv[r];
#pragma unroll
            for (int a = 0; a < A; a++) gW3[a] += (a == act[r]) ? h2[r] * dqv[r] : 0.f;
            const float dz2 = __fmul_rn(__fmul_rn(dqv[r], sOn[oW3 + j2 * A + act[r]]), mk2[r]);
            // layer 2: dW2[i][j] += h1[i]*dz2[j]; dh1[i] = sum_j dz2[j]*W2[i][j]
            gb2 += dz2;
            float dh1 = 0.f;
#pragma unroll
            for (int k = 0; k < H1; k++) gW2[k] = fmaf(__shfl_sync(FULL, h1[r], k), dz2, gW2[k]);
#pragma unroll
            for (int k = 0; k < H2; k++) dh1 = fmaf(__shfl_sync(FULL, dz2, k), sW2T[k * H1 + j1], dh1);
            const float dz1 = __fmul_rn(dh1, mk1[r]);
            // layer 1: dW1[k][j] += x[k]*dz1[j]
            gb1 += dz1;
#pragma unroll
            for (int k = 0; k < D; k++) gW1[k] = fmaf(__shfl_sync(FULL, x[r], k), dz1, gW1[k]);
        }
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

    // CTA sum in warp order, kept in registers (thread t owns elements t, t+blockDim, ...)
    float mysum[ELEMS];
#pragma unroll
    for (int e = 0; e < ELEMS; e++) {
        const int i = tid + e * blockDim.x;
        float s = 0.f;
        if (i < P1)
            for (int w = 0; w < nwarps; w++) s += sPart[w * P1 + i];
        mysum[e] = s;
    }

    const int ctas = gridDim.x;
    float *flat = p.flat + (int64_t)rep * P1;
    if (ctas > 1) {
        // ---- level 1: the last CTA of each group of GROUP sums the group in order -----
        const int ngroups = (ctas + GROUP - 1) / GROUP;
        const int grp = blockIdx.x / GROUP;
        const int gfirst = grp * GROUP;
        const int gsize = min(GROUP, ctas - gfirst);
        float *part1 = p.partial + (int64_t)rep * (ctas + ngroups) * P1;   // [ctas][P1]
        float *part2 = part1 + (int64_t)ctas * P1;                         // [ngroups][P1]
        unsigned int *tick1 = p.ticket + (int64_t)rep * (NARROW_MAX_GROUPS + 1);
        unsigned int *tick2 = tick1 + NARROW_MAX_GROUPS;
#pragma unroll
        for (int e = 0; e < ELEMS; e++) {
            const int i = tid + e * blockDim.x;
            if (i < P1) part1[(int64_t)blockIdx.x * P1 + i] = mysum[e];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) sFlag = (atomicAdd(tick1 + grp, 1u) == (unsigned)(gsize - 1));
        __syncthreads();
        if (!sFlag) return;
        __threadfence();
#pragma unroll
        for (int e = 0; e < ELEMS; e++) {
            const int i = tid + e * blockDim.x;
            float v[GROUP];
#pragma unroll
            for (int c = 0; c < GROUP; c++)
                v[c] = (i < P1 && c < gsize) ? __ldcg(part1 + (int64_t)(gfirst + c) * P1 + i) : 0.f;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < GROUP; c++) s += v[c];
            mysum[e] = s;
            if (i < P1) part2[(int64_t)grp * P1 + i] = s;
        }
        if (tid == 0) tick1[grp] = 0u;
        // ---- level 2: the last group leader sums the groups in order --------------------
        if (ngroups > 1) {
            __threadfence();
            __syncthreads();
            if (tid == 0) sFlag = (atomicAdd(tick2, 1u) == (unsigned)(ngroups - 1));
            __syncthreads();
            if (!sFlag) return;
            __threadfence();
#pragma unroll
            for (int e = 0; e < ELEMS; e++) {
                const int i = tid + e * blockDim.x;
                float s = 0.f;
                for (int g0 = 0; g0 < ngroups; g0 += GROUP) {
                    float v[GROUP];
#pragma unroll
                    for (int c = 0; c < GROUP; c++)
                        v[c] = (i < P1 && g0 + c < ngroups) ? __ldcg(part2 + (int64_t)(g0 + c) * P1 + i) : 0.f;
#pragma unroll
                    for (int c = 0; c < GROUP; c++) s += v[c];
                }
                mysum[e] = s;
            }
            if (tid == 0) *tick2 = 0u;
        }
    }

    // ---- the last CTA of this replica: loss, K5 Adam --------------------------------
#pragma unroll
    for (int e = 0; e < ELEMS; e++) {
        const int i = tid + e * blockDim.x;
        if (i >= P1) continue;
        float g = mysum[e];
        if (i == P) {
            g = __fdiv_rn(g, p.sc.count);
            p.loss[rep] = g;
            flat[P] = g;
        } else {
            flat[i] = g;
            if (p.fuse_adam) {
                float w = sOn[i], m = sM[i], v = sV[i];
                solver_update(w, g, m, v, p.adam);
                p.theta[pofs + i] = w;
                p.m[pofs + i] = m;
                p.v[pofs + i] = v;
            }
        }
    }
}

template <int D, int H1, int H2, int A>
void launch_t(const NarrowArgs &a, cudaStream_t st)
{
    constexpr int P = D * H1 + H1 + H1 * H2 + H2 + H2 * A + A;
    size_t smem = sizeof(float) * (4 * ((P + 3) & ~3) + H1 * H2 + (size_t)a.warps_per_cta * (P + 1));
    dim3 grid(a.ctas_per_replica, a.replicas);
    narrow_learn_kernel<D, H1, H2, A><<<grid, a.warps_per_cta * 32, smem, st>>>(a);
}

template <int D, int H1, int H2, int A>
void configure_t(int warps)
{
    // The attribute belongs to the function, not to the handle: a later handle with fewer warps must not
    // lower it below what an earlier handle launches with, so always opt in to the 16-warp maximum.
    (void)warps;
    constexpr int P = D * H1 + H1 + H1 * H2 + H2 + H2 * A + A;
    constexpr int MAX_WARPS = 16;
    size_t smem = sizeof(float) * (4 * ((P + 3) & ~3) + H1 * H2 + (size_t)MAX_WARPS * (P + 1));
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(narrow_learn_kernel<D, H1, H2, A>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace

// opt in to > 48 KB of dynamic shared memory once, at handle creation (not at launch:
// launches may be under stream capture)
void narrow_configure(const Dims &d, int warps)
{
    if (d.D == 4 && d.A == 2) configure_t<4, 24, 24, 2>(warps);
    else if (d.D == 2 && d.A == 3) configure_t<2, 24, 24, 3>(warps);
    else if (d.D == 6 && d.A == 3) configure_t<6, 24, 24, 3>(warps);
    else if (d.D == 8 && d.A == 4) configure_t<8, 24, 24, 4>(warps);
}

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
    int need = (a.B + NR - 1) / NR;  // NR rows per warp
    int max_ctas = (2 * sm_count) / (a.replicas > 0 ? a.replicas : 1);
    if (max_ctas < 1) max_ctas = 1;
    if (max_ctas > NARROW_GROUP * NARROW_MAX_GROUPS) max_ctas = NARROW_GROUP * NARROW_MAX_GROUPS;
    int ctas = (need + 15) / 16;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    int warps = (need + ctas - 1) / ctas;
    if (warps > 16) warps = 16;
    if (warps < MIN_WARPS) warps = MIN_WARPS;
    a.ctas_per_replica = ctas;
    a.warps_per_cta = warps;
}

size_t narrow_partial_floats(int ctas, int P)
{
    int ngroups = (ctas + NARROW_GROUP - 1) / NARROW_GROUP;
    return (size_t)(ctas + ngroups) * (P + 1);
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
