// sampler.cuh — keyed pseudo-random permutation of [0, n).
//
// Memory.Sample draws the first B entries of rand.Perm(Len)
// (reference: pkg/v1/agent/deepq/memory.go:54-69): B distinct, uniform indices, at
// O(Len) cost per call.  Here entry b of the permutation is computed directly:
// a 4-round Feistel network over the smallest even-bit domain >= n, cycle-walked
// back into [0, n).  A bijection, so the B indices are distinct; O(1) per index.
#pragma once
#include <stdint.h>

namespace gq {

__host__ __device__ inline uint32_t mix32(uint32_t x, uint32_t k)
{
    x ^= k;
    x *= 0x9E3779B1u;
    x ^= x >> 15;
    x *= 0x85EBCA77u;
    x ^= x >> 13;
    x *= 0xC2B2AE3Du;
    x ^= x >> 16;
    return x;
}

__host__ __device__ inline uint32_t perm_index(uint32_t i, uint32_t n, uint64_t seed)
{
    if (n <= 1) return 0;
    int bits = 1;
    while ((1ull << bits) < (uint64_t)n) bits++;
    if (bits & 1) bits++;
    const int h = bits >> 1;
    const uint32_t mask = (1u << h) - 1u;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t x = i;
    do {
        uint32_t L = x >> h, R = x & mask;
        for (int round = 0; round < 4; round++) {
            uint32_t key = mix32(k0 + 0x632BE5ABu * (uint32_t)round, k1 ^ (uint32_t)round);
            uint32_t t = L ^ (mix32(R, key) & mask);
            L = R;
            R = t;
        }
        x = (L << h) | R;
    } while (x >= n);
    return x;
}

// Epsilon-greedy draw for row i of a batched Action call (reference: agent.go:193-205 draws
// num.RandF32(0,1) < epsilon per call and then env.SampleAction): a uniform in [0,1) with 24 bits
// and, when exploring, a uniform action.  Counter based, so host and device agree for a given seed.
__host__ __device__ inline bool explore_draw(uint32_t i, uint64_t seed, float epsilon, int n_actions, int *action)
{
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t u = mix32(mix32(i, k0), k1 ^ 0x5bd1e995u);
    const float r = (float)(u >> 8) * (1.0f / 16777216.0f);
    *action = (int)(mix32(u, k1 + 0x27d4eb2fu) % (uint32_t)n_actions);
    return r < epsilon;
}

}  // namespace gq
