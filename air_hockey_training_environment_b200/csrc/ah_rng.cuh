// Counter-based RNG for in-kernel resets and synthetic actions: Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
// "Parallel random numbers: as easy as 1, 2, 3", SC'11).  Not part of the reference, which draws the reset
// velocities from numpy's global MT19937 (environment/puck.py:115-116); a counter-based generator keyed on
// the GLOBAL env id makes the draws independent of thread mapping and of multi-GPU sharding.
//
//   counter = (env_id lo, env_id hi, index, stream)      key = (seed lo, seed hi)
//   stream 0: reset velocities (FP64 build), index = resets consumed so far by that env
//   stream 1: actions, index = frame / 64 -- one 128-bit block holds 64 two-bit actions
//
// The FP32 THROUGHPUT build draws its reset velocities from a cheaper counter-based generator, pcg3d (Jarzynski &
// Olano, "Hash Functions for GPU Rendering", JCGT 9(3), 2020: the 3-in / 3-out member of the PCG family, among the best
// quality-per-instruction hashes of that survey), keyed the same way -- (seed, GLOBAL env id, reset index) -> (dx, dy):
// 22 instructions instead of Philox's 83 INSIDE the frame loop's goal branch, which one lane of a warp enters in 23 %
// of warp-frames (measured: 81.5 -> 76.6 us per step of BASELINE config 3).  Both generators are restated on the CPU by the
// test infrastructure; tests pin the two statements to each other bit for bit (tests/test_gpu_api.py) and check known
// answers, uniformity and independence of the FP32 one (tests/test_reference_pins.py).
#pragma once
#include <stdint.h>

namespace ah {

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// U(-10, 10) pair for one reset, in the arithmetic of np.random.uniform: low + (high - low) * u  (puck.py:115-116)
// The key of an env set's reset stream: the seed itself for Philox (FP64 build), splitmix64(seed) for pcg3d (FP32 build:
// distinct seeds give unrelated streams, not shifted copies).  Computed once per launch on the host.
struct ResetKey { uint64_t seed, mix; };
inline ResetKey reset_key(uint64_t seed) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;               // splitmix64 (Steele, Lea, Flood 2014), one step
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ResetKey{seed, z};
}

__device__ __forceinline__ void reset_velocity(const ResetKey& key, uint64_t env_id, uint32_t reset_idx, double& dx, double& dy) {
    const uint64_t seed = key.seed;
    const Philox4 o = philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), reset_idx, 0u,
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    // 53-bit uniforms built like MT19937's genrand_res53: (a >> 5) * 2^26 + (b >> 6), scaled by 2^-53
    const double u0 = __dmul_rn(__dadd_rn(__dmul_rn((double)(o.x >> 5), 67108864.0), (double)(o.y >> 6)), 1.0 / 9007199254740992.0);
    const double u1 = __dmul_rn(__dadd_rn(__dmul_rn((double)(o.z >> 5), 67108864.0), (double)(o.w >> 6)), 1.0 / 9007199254740992.0);
    dx = __dadd_rn(-10.0, __dmul_rn(20.0, u0));
    dy = __dadd_rn(-10.0, __dmul_rn(20.0, u1));
}

// FP32 build: pcg3d on (env id lo ^ mix lo, reset index ^ env id hi, mix hi), mix = splitmix64(seed).  24-bit uniforms,
// one fma each.
__device__ __forceinline__ void reset_velocity(const ResetKey& key, uint64_t env_id, uint32_t reset_idx, float& dx, float& dy) {
    uint32_t x = (uint32_t)env_id ^ (uint32_t)key.mix, y = reset_idx ^ (uint32_t)(env_id >> 32), z = (uint32_t)(key.mix >> 32);
    x = x * 1664525u + 1013904223u; y = y * 1664525u + 1013904223u; z = z * 1664525u + 1013904223u;
    x += y * z; y += z * x; z += x * y;
    x ^= x >> 16; y ^= y >> 16; z ^= z >> 16;
    x += y * z; y += z * x;                                   // (pcg3d's last line, z += x * y, feeds nothing here)
    dx = fmaf((float)(x >> 8), 20.0f / 16777216.0f, -10.0f);
    dy = fmaf((float)(y >> 8), 20.0f / 16777216.0f, -10.0f);
}

// 128 bits = the 64 actions of frames [64*blk, 64*blk + 64)
__device__ __forceinline__ Philox4 action_block(uint64_t seed, uint64_t env_id, uint32_t blk) {
    return philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), blk, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
}

__device__ __forceinline__ int action_from_block(const Philox4& b, uint32_t frame_in_block) {
    const uint32_t w = frame_in_block >> 4;
    const uint32_t v = w == 0 ? b.x : w == 1 ? b.y : w == 2 ? b.z : b.w;
    return (int)((v >> (2u * (frame_in_block & 15u))) & 3u);
}

// stream 2: uniforms for sampling an action from a policy's distribution; one 128-bit block serves 4 frames
__device__ __forceinline__ Philox4 sample_block(uint64_t seed, uint64_t env_id, unsigned long long frame) {
    return philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)(frame >> 2), 2u ^ ((uint32_t)(frame >> 34) << 8),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}
// stream 3: a second, independent uniform per frame (epsilon-greedy's random action, Box-Muller's angle)
__device__ __forceinline__ Philox4 sample_block2(uint64_t seed, uint64_t env_id, unsigned long long frame) {
    return philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)(frame >> 2), 3u ^ ((uint32_t)(frame >> 34) << 8),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ float sample_uniform(const Philox4& b, unsigned long long frame) {
    const uint32_t w = (uint32_t)(frame & 3u);
    const uint32_t v = w == 0 ? b.x : w == 1 ? b.y : w == 2 ? b.z : b.w;
    return (float)(v >> 8) * (1.0f / 16777216.0f);       // 24-bit uniform in [0, 1)
}

}  // namespace ah
