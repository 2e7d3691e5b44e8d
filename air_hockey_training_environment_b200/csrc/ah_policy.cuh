// On-device action selection for ANY small dense policy (include/ah_b200.h: ah_policy): the hand-off the reference
// does on the host every frame -- Agent.get_action (lib/agents/agent.py:42-48) -> model.predict -> exploration strategy
// (lib/exploration/strategies.py, lib/exploration/random.py) -> Agent.move -- evaluated per env inside a kernel.
//
//   network   up to 4 dense layers, widths <= 32, relu / tanh / linear, BatchNorm folded by the caller.  Covers the
//             reference's nets: PPO discrete actor 8-4-6-4-4 softmax and continuous actor 8-6-6-4-2 tanh
//             (lib/agents/ppo/model.py:56-110), A2C actor and DDQN net 8-8-4 softmax (a2c/model.py:50-60,
//             ddqn/model.py:39-53), the dueling Q-net 8-32-(8|8)-(1|4) as a block-structured chain
//             (dueling/model.py:50-66).  (C51's 51-atom heads exceed the width; it goes through ah_policy_act with
//             head values computed by the caller's network.)
//   head      softmax probabilities | Q-values | dueling v + a - mean(a) | continuous mean
//   select    SoftmaxPolicy / argmax / EpsilonGreedy / BoltzmannQPolicy / MaxBoltzmannQPolicy        strategies.py:7-123
//             GaussianWhiteNoiseProcess / OrnsteinUhlenbeckProcess added to the continuous mean      random.py:11-67
//
// Evaluation: the weights (transposed, outputs padded to 4) are staged in shared memory once per block and read as
// warp-uniform LDS.128 broadcasts; a thread's activations live in its own column of a [32][T] shared array (no bank
// conflicts, no barriers: nobody else touches the column); four outputs are accumulated per pass over the inputs:
// 1 LDS.128 + 1 LDS + 4 FFMA per 4 multiply-adds.
//
// Randomness: counter-based Philox keyed (seed, global env id), counter = frame, streams 2 (u0) and 3 (u1) -- the same
// draw whether the policy runs in the step kernel (one launch per rollout) or frame by frame (ah_policy_act).
#pragma once

namespace ah {

#ifndef AH_POL_THREADS
#define AH_POL_THREADS 128
#endif
constexpr int kPolThreads = AH_POL_THREADS;            // block size of every kernel that evaluates a policy
constexpr int kPolMaxW = AH_POLICY_MAX_WIDTH;
constexpr int kPolMaxFloats = (8 * kPolMaxW + kPolMaxW) + 3 * (kPolMaxW * kPolMaxW + kPolMaxW);

// DYNAMIC shared memory of a kernel that evaluates a policy: the staged weights, then (general path only) the two
// activation arrays [kPolMaxW][blockDim.x].  Sized per launch from the network -- 1.2 KB for the reference's PPO actor,
// 46 KB for the widest net -- so that a small network does not cap the resident blocks per SM.
__host__ __device__ inline int policy_weight_floats(int n_floats) { return (n_floats + 3) & ~3; }

struct PolicyDev {                                      // by value in the kernel arguments
    int n_layers, width[5], act[4], w_off[4], b_off[4], n_floats;
    int narrow;                                         // every width <= 8: the register-resident evaluation
    int head, select, n_actions;
    float eps0, eps_delta, eps_final; int observe, max_decrements;
    float inv_tau, clip_lo, clip_hi;
    float mu, sigma0, sigma_min, sigma_m, theta, dt, sqrt_dt;
    const float* weights;
    float* noise_state;                                 // [n][2], AH_SELECT_OU only
};

__device__ __forceinline__ void policy_stage(const PolicyDev& P, float* sw) {
    for (int t = threadIdx.x; t < P.n_floats; t += blockDim.x) sw[t] = P.weights[t];
}

// tanh(v) = 1 - 2 / (exp(2 v) + 1) on the fast exponential and reciprocal (|error| < 2e-7; saturates to +-1 through
// exp -> inf / 0): 6 instructions where libm's tanhf inlines ~40 at each of the unrolled call sites -- the closed loop's
// frame body must stay small enough for the instruction cache (14 warps per SM, each at its own PC)
__device__ __forceinline__ float policy_tanh(float v) {
    return 1.0f - __fdividef(2.0f, __expf(2.0f * v) + 1.0f);
}
__device__ __forceinline__ float policy_activate(float v, int act) {
    if (act == AH_ACTV_RELU) return fmaxf(v, 0.0f);
    if (act == AH_ACTV_TANH) return policy_tanh(v);
    return v;
}

// NARROW networks (every width <= 8: the reference's PPO actors 8-4-6-4-4 / 8-6-6-4-2, the A2C actor and the DDQN net
// 8-8-4): activations stay in REGISTERS.  A one-launch rollout of 65,536 envs has 3.5 warps per scheduler, so the closed
// loop is bound by the latency and the SIZE of this code, not by issue slots: the layer shapes those networks use are
// compiled as exact, fully unrolled bodies (the weight broadcasts -- LDS.128 -- of a layer are issued ahead of its
// multiply-adds, nothing is predicated off), any other shape <= 8 runs the guarded body; the loop over the layers is
// NOT unrolled (one copy of each body in the instruction stream).  Same multiply-add order per output as the general
// path: same bits.
template <int WIN, int OUTP>
__device__ __forceinline__ void narrow_layer(const float* __restrict__ W, const float* __restrict__ B, int act, float (&h)[8]) {
    float o[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < OUTP; j += 4) {
        float4 acc = *reinterpret_cast<const float4*>(B + j);
#pragma unroll
        for (int c = 0; c < WIN; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(W + c * OUTP + j);
            acc.x = fmaf(w.x, h[c], acc.x); acc.y = fmaf(w.y, h[c], acc.y);
            acc.z = fmaf(w.z, h[c], acc.z); acc.w = fmaf(w.w, h[c], acc.w);
        }
        o[j] = acc.x; o[j + 1] = acc.y; o[j + 2] = acc.z; o[j + 3] = acc.w;
    }
    if (act == AH_ACTV_RELU) {
#pragma unroll
        for (int q = 0; q < OUTP; ++q) o[q] = fmaxf(o[q], 0.0f);
    } else if (act == AH_ACTV_TANH) {
#pragma unroll
        for (int q = 0; q < OUTP; ++q) o[q] = policy_tanh(o[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) h[q] = o[q];
}

// any (win, outp) with win <= 8, outp in {4, 8}: the same body behind warp-uniform guards
__device__ __forceinline__ void narrow_layer_any(const float* __restrict__ W, const float* __restrict__ B, int win, int outp, int act,
                                                 float (&h)[8]) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
        float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (j < outp) {
            acc = *reinterpret_cast<const float4*>(B + j);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c < win) {
                    const float4 w = *reinterpret_cast<const float4*>(W + c * outp + j);
                    acc.x = fmaf(w.x, h[c], acc.x); acc.y = fmaf(w.y, h[c], acc.y);
                    acc.z = fmaf(w.z, h[c], acc.z); acc.w = fmaf(w.w, h[c], acc.w);
                }
            }
            acc.x = policy_activate(acc.x, act); acc.y = policy_activate(acc.y, act);
            acc.z = policy_activate(acc.z, act); acc.w = policy_activate(acc.w, act);
        }
        o[j] = acc.x; o[j + 1] = acc.y; o[j + 2] = acc.z; o[j + 3] = acc.w;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = o[c];
}

__device__ __forceinline__ void policy_forward_narrow(const PolicyDev& P, const float* __restrict__ sw, const float (&x)[8], float (&y)[8]) {
    float h[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = x[c];
#pragma unroll 1
    for (int l = 0; l < P.n_layers; ++l) {
        const int win = P.width[l], outp = (P.width[l + 1] + 3) & ~3, act = P.act[l];
        const float* W = sw + P.w_off[l];
        const float* B = sw + P.b_off[l];
        switch (win * 16 + outp) {
            case 8 * 16 + 4: narrow_layer<8, 4>(W, B, act, h); break;
            case 8 * 16 + 8: narrow_layer<8, 8>(W, B, act, h); break;
            case 6 * 16 + 4: narrow_layer<6, 4>(W, B, act, h); break;
            case 6 * 16 + 8: narrow_layer<6, 8>(W, B, act, h); break;
            case 4 * 16 + 4: narrow_layer<4, 4>(W, B, act, h); break;
            case 4 * 16 + 8: narrow_layer<4, 8>(W, B, act, h); break;
            default: narrow_layer_any(W, B, win, outp, act, h); break;
        }
    }
    const int wl = P.width[P.n_layers];
#pragma unroll
    for (int q = 0; q < 8; ++q) y[q] = q < wl ? h[q] : 0.0f;
}

// x[8] -> y[8] (the first width[n_layers] outputs of the last layer, zero-filled)
__device__ __forceinline__ void policy_forward(const PolicyDev& P, const float* __restrict__ sw, float* hs, const float (&x)[8],
                                               float (&y)[8]) {
    if (P.narrow) { policy_forward_narrow(P, sw, x, y); return; }
    const int T = blockDim.x, tid = threadIdx.x;
    float* in = hs;
    float* out = hs + kPolMaxW * T;
#pragma unroll
    for (int c = 0; c < 8; ++c) in[c * T + tid] = x[c];
    for (int l = 0; l < P.n_layers; ++l) {
        const int win = P.width[l], outp = (P.width[l + 1] + 3) & ~3, act = P.act[l];
        const float* W = sw + P.w_off[l];
        const float* B = sw + P.b_off[l];
        for (int j = 0; j < outp; j += 4) {
            float4 acc = *reinterpret_cast<const float4*>(B + j);
            for (int c = 0; c < win; ++c) {
                const float4 w = *reinterpret_cast<const float4*>(W + c * outp + j);
                const float h = in[c * T + tid];
                acc.x = fmaf(w.x, h, acc.x); acc.y = fmaf(w.y, h, acc.y);
                acc.z = fmaf(w.z, h, acc.z); acc.w = fmaf(w.w, h, acc.w);
            }
            out[(j + 0) * T + tid] = policy_activate(acc.x, act);
            out[(j + 1) * T + tid] = policy_activate(acc.y, act);
            out[(j + 2) * T + tid] = policy_activate(acc.z, act);
            out[(j + 3) * T + tid] = policy_activate(acc.w, act);
        }
        float* t = in; in = out; out = t;
    }
    const int wl = P.width[P.n_layers];
#pragma unroll
    for (int q = 0; q < 8; ++q) y[q] = q < wl ? in[q * T + tid] : 0.0f;
}

// head: y -> v[0..A) = probabilities (softmax head) or Q-values (values / dueling head); continuous: v[0..1] = mean
__device__ __forceinline__ void policy_head(const PolicyDev& P, const float (&y)[8], float (&v)[8]) {
    const int A = P.n_actions;
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.0f;
    if (P.head == AH_HEAD_SOFTMAX && A == 4) {          // the reference's four actions: the loop below without its guards
        const float m = fmaxf(fmaxf(fmaxf(y[0], y[1]), y[2]), y[3]);          // (same operation order: same bits)
        const float e0 = __expf(y[0] - m), e1 = __expf(y[1] - m), e2 = __expf(y[2] - m), e3 = __expf(y[3] - m);
        const float inv = 1.0f / (((e0 + e1) + e2) + e3);
        v[0] = e0 * inv; v[1] = e1 * inv; v[2] = e2 * inv; v[3] = e3 * inv;
    } else if (P.head == AH_HEAD_SOFTMAX) {
        float m = y[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) if (q < A) m = fmaxf(m, y[q]);
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) if (q < A) { v[q] = __expf(y[q] - m); s += v[q]; }
        const float inv = 1.0f / s;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] *= inv;
    } else if (P.head == AH_HEAD_DUELING) {             // q = v + a - mean(a)      dueling/model.py:58-66
        float mean = 0.0f;
#pragma unroll
        for (int q = 1; q < 8; ++q) if (q <= A) mean += y[q];
        mean /= (float)A;
#pragma unroll
        for (int q = 0; q < 7; ++q) if (q < A) v[q] = y[0] + (y[q + 1] - mean);
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = y[q];
    }
}

__device__ __forceinline__ int policy_argmax(const float (&v)[8], int A) {       // np.argmax: the first maximum
    int best = 0; float bv = v[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) if (q < A && v[q] > bv) { bv = v[q]; best = q; }
    return best;
}

// np.random.choice(A, p): the first index whose cumulative probability exceeds u
__device__ __forceinline__ int policy_choice(const float (&p)[8], int A, float u) {
    if (A == 4) {                                        // four actions: the loop below without its guards (same sums)
        const float c0 = p[0], c1 = c0 + p[1], c2 = c1 + p[2];
        return u < c0 ? 0 : (u < c1 ? 1 : (u < c2 ? 2 : 3));
    }
    float c = 0.0f; int act = A - 1;
    bool found = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (q < A) { c += p[q]; if (!found && u < c) { act = q; found = true; } }
    }
    return act;
}

// EpsilonGreedy.decrease (strategies.py:39-43): epsilon after the step() call number t (1-based)
__device__ __forceinline__ float policy_epsilon(const PolicyDev& P, unsigned long long t) {
    unsigned long long m = P.observe > 0 ? t / (unsigned long long)P.observe : 0ull;
    if (m > (unsigned long long)P.max_decrements) m = (unsigned long long)P.max_decrements;
    return P.eps0 - (float)m * P.eps_delta;
}

struct PolicyAction { int ia; float ax, ay; };

// The two uniform streams of an env, one 128-bit Philox block per 4 frames each: a caller that plays consecutive frames
// (the step kernel's closed loop) keeps the blocks across frames -- one draw per 4 frames instead of one per frame.
struct PolicyRng {
    Philox4 b0, b1;
    unsigned long long id0 = ~0ull, id1 = ~0ull;        // frame >> 2 of the cached block
};
__device__ __forceinline__ float policy_u0(PolicyRng& g, uint64_t seed, uint64_t env_id, unsigned long long f) {
    if ((f >> 2) != g.id0) { g.b0 = sample_block(seed, env_id, f); g.id0 = f >> 2; }
    return sample_uniform(g.b0, f);
}
__device__ __forceinline__ float policy_u1(PolicyRng& g, uint64_t seed, uint64_t env_id, unsigned long long f) {
    if ((f >> 2) != g.id1) { g.b1 = sample_block2(seed, env_id, f); g.id1 = f >> 2; }
    return sample_uniform(g.b1, f);
}

// One action for one env.  `f` = the global frame index (frames played before this one); v = policy_head's output.
__device__ __forceinline__ PolicyAction policy_select(const PolicyDev& P, const float (&v)[8], uint64_t seed, uint64_t env_id,
                                                      unsigned long long f, int64_t i, PolicyRng& g) {
    PolicyAction r; r.ia = -1; r.ax = 0.0f; r.ay = 0.0f;
    const int A = P.n_actions, sel = P.select;
    float u0 = 0.0f, u1 = 0.0f;
    if (sel != AH_SELECT_GREEDY) {
        u0 = policy_u0(g, seed, env_id, f);
        if (sel != AH_SELECT_SAMPLE && sel != AH_SELECT_BOLTZMANN) u1 = policy_u1(g, seed, env_id, f);
    }
    if (P.head == AH_HEAD_CONTINUOUS) {
        r.ax = v[0]; r.ay = v[1];
        if (sel == AH_SELECT_GAUSSIAN || sel == AH_SELECT_OU) {
            // two standard normals from (u0, u1): Box-Muller
            const float rad = sqrtf(-2.0f * __logf(1.0f - u0)), ang = 6.283185307179586f * u1;
            const float z0 = rad * __cosf(ang), z1 = rad * __sinf(ang);
            const float sigma = fmaxf(P.sigma_min, fmaf(P.sigma_m, (float)f, P.sigma0));      // random.py:26-28
            if (sel == AH_SELECT_GAUSSIAN) {                                                     // random.py:38-41
                r.ax += fmaf(sigma, z0, P.mu); r.ay += fmaf(sigma, z1, P.mu);
            } else {                                                                             // random.py:56-63
                float2 xp = reinterpret_cast<float2*>(P.noise_state)[i];
                xp.x = xp.x + P.theta * (P.mu - xp.x) * P.dt + sigma * P.sqrt_dt * z0;
                xp.y = xp.y + P.theta * (P.mu - xp.y) * P.dt + sigma * P.sqrt_dt * z1;
                reinterpret_cast<float2*>(P.noise_state)[i] = xp;
                r.ax += xp.x; r.ay += xp.y;
            }
        }
        return r;
    }
    if (sel == AH_SELECT_SAMPLE) { r.ia = policy_choice(v, A, u0); return r; }              // strategies.py:17-19
    if (sel == AH_SELECT_GREEDY) { r.ia = policy_argmax(v, A); return r; }
    float b[8];                                                                             // Boltzmann law (strategies.py:73-75)
    if (sel == AH_SELECT_BOLTZMANN || sel == AH_SELECT_MAX_BOLTZMANN) {
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) { b[q] = q < A ? __expf(fminf(fmaxf(v[q] * P.inv_tau, P.clip_lo), P.clip_hi)) : 0.0f; s += b[q]; }
        const float inv = 1.0f / s;
#pragma unroll
        for (int q = 0; q < 8; ++q) b[q] *= inv;
    }
    if (sel == AH_SELECT_BOLTZMANN) { r.ia = policy_choice(b, A, u0); return r; }
    const float eps = policy_epsilon(P, f + 1ull);                                          // self.t += 1; decrease()
    if (u0 < eps) {
        if (sel == AH_SELECT_EPS_GREEDY) { int a = (int)(u1 * (float)A); r.ia = a < A ? a : A - 1; }   // np.random.randint(0, A)
        else r.ia = policy_choice(b, A, u1);
    } else {
        r.ia = policy_argmax(v, A);
    }
    return r;
}

// ---- stand-alone kernel: obs [n][8] -> action (+ head values), one frame (ah_policy_act)
template <typename R>
__global__ void __launch_bounds__(kPolThreads) policy_act_kernel(int64_t n, PolicyDev P, const R* __restrict__ obs, const R* __restrict__ values_in,
                                                                 int8_t* __restrict__ actions, R* __restrict__ actions_real,
                                                                 R* __restrict__ values_out, uint64_t seed, uint64_t env_id0,
                                                                 const DevBlock* blk) {
    extern __shared__ __align__(16) float pol_smem[];
    float* sw = pol_smem;
    float* hs = pol_smem + policy_weight_floats(P.n_floats);
    if (!values_in) policy_stage(P, sw);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[8];
    if (values_in) {                                     // the caller's own network produced the head values
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = q < P.n_actions ? (float)values_in[i * P.n_actions + q] : 0.0f;
    } else {
        using V = typename Vec4T<R>::type;
        const V o0 = reinterpret_cast<const V*>(obs)[2 * i], o1 = reinterpret_cast<const V*>(obs)[2 * i + 1];
        const float x[8] = {(float)o0.x, (float)o0.y, (float)o0.z, (float)o0.w, (float)o1.x, (float)o1.y, (float)o1.z, (float)o1.w};
        float y[8];
        policy_forward(P, sw, hs, x, y);
        policy_head(P, y, v);
    }
    PolicyRng g;
    const PolicyAction a = policy_select(P, v, seed, env_id0 + (uint64_t)i, blk->frame, i, g);
    if (P.head == AH_HEAD_CONTINUOUS) {
        if (actions_real) { actions_real[2 * i] = (R)a.ax; actions_real[2 * i + 1] = (R)a.ay; }
    } else if (actions) {
        actions[i] = (int8_t)a.ia;
    }
    if (values_out) {
        const int nv = P.head == AH_HEAD_CONTINUOUS ? 2 : P.n_actions;
#pragma unroll
        for (int q = 0; q < 8; ++q) if (q < nv) values_out[i * nv + q] = (R)v[q];
    }
}

}  // namespace ah
