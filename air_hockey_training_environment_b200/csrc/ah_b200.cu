// B200-native batched air-hockey environment step: kernels + the C ABI of include/ah_b200.h.
//
// Mapping: ONE THREAD OWNS ONE ENVIRONMENT.  A thread loads its env's state from the SoA arrays with 128-bit
// coalesced accesses, keeps the 13 reals + counters in registers across K fused frames (3 physics sub-updates
// each), writes the per-frame outputs straight into caller-owned tensors, and stores the state back once.
// Persistent grid: SMs x resident-blocks CTAs, grid-stride over the envs.  Per-block statistics are reduced
// with warp shuffles / REDUX and leave the SM as one atomic set per block.  No tensor cores: there is no dense
// contraction anywhere in this path.
//
// Compile with -fmad=false (see build.py): the FP64 instantiation must not contract a*b+c.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <type_traits>

#include "../../include/ah_b200.h"
#include "ah_physics.cuh"
#include "ah_rng.cuh"

namespace ah {

// ------------------------------------------------------------------------------------------------ device block
struct DevBlock {
    unsigned long long stats[2][AH_NSTATS];   // AH_STAT_*, two banks (ah_step_io.stats_bank)
    unsigned long long frame;              // global frame counter (Philox action index)
    unsigned long long xchg_epoch;         // ah_stats_allreduce: exchanges done so far (all ranks call it equally often)
    unsigned int ticket[4];                // last-block-done detection, one per concurrent ordered launch (ring counter index)
    unsigned int pad[40];
};
static_assert(sizeof(DevBlock) == 512, "DevBlock is 512 bytes");

// One contribution to the fused statistics all-reduce: written by its source rank straight into every rank's receive
// buffer over NVLink (P2P stores), `epoch` last (release) -- the flag the receiver spins on.
struct PeerSlot {
    long long vals[AH_NSTATS];
    unsigned long long epoch;
    unsigned long long pad[11];
};
static_assert(sizeof(PeerSlot) == 256, "PeerSlot is 256 bytes");

template <typename R> struct Vec4T;
template <> struct Vec4T<float> { using type = float4; };
template <> struct Vec4T<double> { using type = double4; };

template <typename R>
struct StatePtrs {
    typename Vec4T<R>::type* puck;
    typename Vec4T<R>::type* robot;
    typename Vec4T<R>::type* opp;
    R* prev_dist;
    int4* meta;
};

template <typename R>
struct RawEnv {                       // one env's rows exactly as they sit in HBM
    typename Vec4T<R>::type p, r, o;
    int4 m;
    R prev;
};

template <typename R>
__device__ __forceinline__ RawEnv<R> load_raw(const StatePtrs<R>& s, int64_t i) {
    RawEnv<R> w;
    w.p = s.puck[i]; w.r = s.robot[i]; w.o = s.opp[i]; w.m = s.meta[i]; w.prev = s.prev_dist[i];
    return w;
}

template <typename R>
__device__ __forceinline__ Env<R> unpack_env(const RawEnv<R>& w) {
    Env<R> e;
    e.px = w.p.x; e.py = w.p.y; e.pdx = w.p.z; e.pdy = w.p.w;
    e.rx = w.r.x; e.ry = w.r.y; e.rdx = w.r.z; e.rdy = w.r.w;
    e.ox = w.o.x; e.oy = w.o.y; e.odx = w.o.z; e.ody = w.o.w;
    e.prev_dist = w.prev;
    e.robot_score = w.m.x & 0xffff; e.opp_score = (w.m.x >> 16) & 0xffff;
    e.resets = (uint32_t)w.m.y; e.game_frames = w.m.z; e.game_return = w.m.w;
    return e;
}

template <typename R>
__device__ __forceinline__ Env<R> load_env(const StatePtrs<R>& s, int64_t i) { return unpack_env<R>(load_raw<R>(s, i)); }

template <typename R>
__device__ __forceinline__ void store_env(const StatePtrs<R>& s, int64_t i, const Env<R>& e) {
    typename Vec4T<R>::type p, r, o;
    p.x = e.px; p.y = e.py; p.z = e.pdx; p.w = e.pdy;
    r.x = e.rx; r.y = e.ry; r.z = e.rdx; r.w = e.rdy;
    o.x = e.ox; o.y = e.oy; o.z = e.odx; o.w = e.ody;
    s.puck[i] = p; s.robot[i] = r; s.opp[i] = o;
    s.prev_dist[i] = e.prev_dist;
    s.meta[i] = make_int4((e.robot_score & 0xffff) | (e.opp_score << 16), (int)e.resets, e.game_frames, e.game_return);
}

template <typename R>
__device__ __forceinline__ void store_obs(R* base, int64_t row, const Obs<R>& o) {
    using V = typename Vec4T<R>::type;
    V a, b;
    a.x = o.v[0]; a.y = o.v[1]; a.z = o.v[2]; a.w = o.v[3];
    b.x = o.v[4]; b.y = o.v[5]; b.z = o.v[6]; b.w = o.v[7];
    V* dst = reinterpret_cast<V*>(base) + 2 * row;
    dst[0] = a; dst[1] = b;
}

}  // namespace ah
#include "ah_policy.cuh"
namespace ah {

// ------------------------------------------------------------------------------------------------ resets
struct ResetSrc {
    const double* table;   // [n][R][2] or nullptr => Philox
    int R;
    uint64_t seed, env_id0;
    ResetKey key;          // reset-velocity stream of `seed` (ah_rng.cuh)
};

// AirHockey.reset  airhockey.py:171-187 ; Puck.reset puck.py:111-118 ; Mallet.reset mallet.py:78-84
template <typename R>
__device__ __forceinline__ void env_reset(Env<R>& e, bool total, const ResetSrc& rs, int64_t i, unsigned& overflow) {
    if (total) { e.opp_score = 0; e.robot_score = 0; }
    R dx, dy;
    if (rs.table) {
        if (e.resets < (uint32_t)rs.R) {
            const double2 v = reinterpret_cast<const double2*>(rs.table)[(size_t)i * rs.R + e.resets];
            dx = (R)v.x; dy = (R)v.y;
        } else { dx = R(0); dy = R(0); overflow += 1; }
    } else {
        // (opaque copy of the env index: keeps ptxas from hoisting the first Philox round, which only depends on the
        // env id, out of the frame loop -- four registers and a dozen instructions per env for a rare path)
        unsigned int iu = (unsigned int)i;
        asm volatile("" : "+r"(iu));
        reset_velocity(rs.key, rs.env_id0 + (uint64_t)iu, e.resets, dx, dy);
    }
    e.resets += 1;
    e.px = K<R>::mid_x; e.py = K<R>::mid_y; e.pdx = dx; e.pdy = dy;   // puck to the centre, fresh velocity
    e.rdx = R(0); e.rdy = R(0); e.odx = R(0); e.ody = R(0);           // mallets: velocity only, positions kept
}

// ------------------------------------------------------------------------------------------------ step kernel
template <typename R>
struct StepArgs {
    StatePtrs<R> st;
    int64_t n;
    int K;
    const void* actions;
    int action_kind;
    ResetSrc rs;
    R* obs_state; R* obs_new_state; R* reward; uint8_t* done; int8_t* score; uint8_t* game_over;
    R* obs_next; int obs_next_per_frame;
    int collect_stats;
    // ring
    R* ring_state; R* ring_new_state; void* ring_action; R* ring_reward; uint8_t* ring_done; int64_t ring_capacity;
    unsigned long long* ring_appended;     // the ring's own device counter (ah_ring.appended)
    int dot_fma, friction;
    DevBlock* blk;
    int stats_bank;
    int8_t* actions_out; R* probs_out;     // GENERIC mode: the actions applied / the in-kernel actor's distribution
    PolicyDev pol; R* actions_real_out;    // AH_ACT_POLICY_MLP
};

// Block-level statistics: rare events (goals, wins, contacts ...) are counted with shared-memory atomics right
// where they happen (they cost nothing on the common path and hold no registers); the two per-frame sums
// (reward, reward^2) live in registers and are REDUX/shuffle-reduced per warp at the end of the launch.
struct BlockStats {
    int v[AH_NSTATS];                 // 32-bit: native shared-memory atomic add (a 64-bit one is a CAS loop)
    unsigned long long len_sq;        // the only sum that can outgrow 32 bits within one block and launch
    int warps_done;                   // the last warp of the block to finish flushes the block's statistics
};

// `red.shared.add` issued by exactly the lanes that saw the event.  (A plain atomicAdd makes the compiler wrap every
// call in a vote / REDUX / elect-one-lane sequence -- ~15 instructions executed by the whole warp -- to save an
// atomic that, for these rare events, is a single lane's anyway.)
__device__ __forceinline__ void stat_add(BlockStats& bs, int which, int x) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(&bs.v[which]);
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(x) : "memory");
}

// ---- epilogue WITHOUT block barriers: warps finish at different times (contact / goal branches), so each warp
// folds its per-frame sums into the block's shared counters (REDUX + one shared red per value) and leaves; the
// LAST warp of the block to arrive flushes the block's statistics (<= 16 global atomics) and takes the launch
// ticket; the last block of the launch advances the device-side frame / ring counters (graph-capturable: no host
// value is baked into the launch).
// ORDERED = the launch's own blocks read the frame / ring counters at their start (Philox actions, in-kernel policy,
// ring append): those may only advance once EVERY block is done (fence + ticket).  Kernels that never read them
// (the packed mode) let block 0 advance the frame counter: launches on a stream are serialised, so nobody races it.
template <bool ORDERED = true>
__device__ __forceinline__ void finish_stats(BlockStats& bs, DevBlock* blk, int bank, int collect_stats, int frames, int reward_sum,
                                             int reward_sq, int dones, int contacts_r, int contacts_o, int overflow, int bad,
                                             int ambiguous, int K, unsigned long long* ring_appended, int ring_counters = 1,
                                             int ticket = 0, int ring_frames = -1) {
    if (ring_frames < 0) ring_frames = K;
    const int fs = __reduce_add_sync(0xffffffffu, frames);
    const int rs = __reduce_add_sync(0xffffffffu, reward_sum);
    const int rq = __reduce_add_sync(0xffffffffu, reward_sq);
    const int dd = __reduce_add_sync(0xffffffffu, dones);
    const int cr = __reduce_add_sync(0xffffffffu, contacts_r);
    const int co = __reduce_add_sync(0xffffffffu, contacts_o);
    const int rare = __reduce_or_sync(0xffffffffu, overflow | bad | ambiguous);
    int ov = 0, nf = 0, am = 0;
    if (rare) {
        ov = __reduce_add_sync(0xffffffffu, overflow); nf = __reduce_add_sync(0xffffffffu, bad);
        am = __reduce_add_sync(0xffffffffu, ambiguous);
    }
    const int lane = threadIdx.x & 31;
    int arrived = 0;
    // every warp total is known to every lane (REDUX): lane q adds statistic q -- ONE shared atomic instruction with six
    // active lanes and six addresses instead of six single-lane atomics (each of which ptxas wraps in its vote / leader /
    // popc aggregation sequence: ~40 instructions per warp per launch)
    {
        int v = 0;
        v = lane == AH_STAT_FRAMES ? fs : v;
        v = lane == AH_STAT_DONES ? dd : v;
        v = lane == AH_STAT_REWARD_SUM ? rs : v;
        v = lane == AH_STAT_REWARD_SQ_SUM ? rq : v;
        v = lane == AH_STAT_CONTACTS_ROBOT ? cr : v;
        v = lane == AH_STAT_CONTACTS_OPPONENT ? co : v;
        if (rare) {
            v = lane == AH_STAT_RESET_TABLE_OVERFLOW ? ov : v;
            v = lane == AH_STAT_NONFINITE ? nf : v;
            v = lane == AH_STAT_DONE_AMBIGUOUS ? am : v;
        }
        if (v != 0) stat_add(bs, lane, v);
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        arrived = atomicAdd(&bs.warps_done, 1);
    }
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived != (int)(blockDim.x >> 5) - 1) return;
    __threadfence_block();
    if (collect_stats && lane < AH_NSTATS) {                    // lane q flushes statistic q
        const long long sv = lane == AH_STAT_GAME_LEN_SQ_SUM ? (long long)*(volatile unsigned long long*)&bs.len_sq
                                                             : (long long)*(volatile int*)&bs.v[lane];
        if (sv != 0) atomicAdd(&blk->stats[bank][lane], (unsigned long long)sv);
    }
    if (lane != 0) return;
    if (!ORDERED) {
        // (K == 0: a sub-range launch that does not own env 0 -- it must not touch the counter at all: a read-modify-write
        //  of "+ 0" would race with the concurrent slice that does advance it)
        if (blockIdx.x == 0 && K != 0) blk->frame = *(volatile unsigned long long*)&blk->frame + (unsigned long long)K;
        return;
    }
    __threadfence();
    if (atomicAdd(&blk->ticket[ticket], 1u) == gridDim.x - 1) {
        // (K == 0: a sub-range launch that does not own env 0 must not touch the frame counter, see above)
        if (K != 0) blk->frame = *(volatile unsigned long long*)&blk->frame + (unsigned long long)K;
        if (ring_appended) {                     // this launch's ring counters (ah_step_io.ring_counter_first / _count)
            const unsigned long long next = *(volatile unsigned long long*)ring_appended + (unsigned long long)ring_frames;
            for (int c = 0; c < ring_counters; ++c) ring_appended[c] = next;
        }
        blk->ticket[ticket] = 0;
        __threadfence();
    }
}

// ---- the reference's discrete PPO actor (lib/agents/ppo/model.py:56-74; BatchNorm folded into layer 2) and
// sampling from its softmax.  `sw` = the 114 weights in shared memory (warp-uniform reads -> LDS broadcasts).
__device__ __forceinline__ void actor_forward(const float* __restrict__ sw, const float (&x)[8], float (&p)[4]) {
    const float *W1 = sw, *B1 = sw + 32, *W2 = sw + 36, *B2 = sw + 60, *W3 = sw + 66, *B3 = sw + 90, *W4 = sw + 94, *B4 = sw + 110;
    float h1[4], h2[6], h3[4], z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { float s = B1[j];
#pragma unroll
        for (int c = 0; c < 8; ++c) s = fmaf(W1[j * 8 + c], x[c], s);
        h1[j] = fmaxf(s, 0.f); }
#pragma unroll
    for (int j = 0; j < 6; ++j) { float s = B2[j];
#pragma unroll
        for (int c = 0; c < 4; ++c) s = fmaf(W2[j * 4 + c], h1[c], s);
        h2[j] = fmaxf(s, 0.f); }
#pragma unroll
    for (int j = 0; j < 4; ++j) { float s = B3[j];
#pragma unroll
        for (int c = 0; c < 6; ++c) s = fmaf(W3[j * 6 + c], h2[c], s);
        h3[j] = fmaxf(s, 0.f); }
#pragma unroll
    for (int j = 0; j < 4; ++j) { float s = B4[j];
#pragma unroll
        for (int c = 0; c < 4; ++c) s = fmaf(W4[j * 4 + c], h3[c], s);
        z[j] = s; }
    const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { p[j] = __expf(z[j] - m); sum += p[j]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] *= inv;
}

__device__ __forceinline__ int actor_pick(const float (&p)[4], bool greedy, float u) {
    if (greedy) {                                   // np.argmax(q_values) when not training   ppo.py:114
        int act = 0; float best = p[0];
#pragma unroll
        for (int j = 1; j < 4; ++j) if (p[j] > best) { best = p[j]; act = j; }
        return act;
    }
    const float c0 = p[0], c1 = c0 + p[1], c2 = c1 + p[2];
    return (u >= c0) + (u >= c1) + (u >= c2);
}

#ifndef AH_THREADS
#define AH_THREADS 256
#endif
#ifndef AH_OCC_F32
#define AH_OCC_F32 4
#endif
constexpr int kThreads = AH_THREADS;
constexpr int kMaxBlocks = 1 << 20;

template <typename R> struct Occ { static constexpr int blocks = AH_OCC_F32; };
template <> struct Occ<double> { static constexpr int blocks = 2; };

// Output modes (compile-time, chosen by the host from which pointers were given):
//   MODE_LEAN     reward, done, score, game_over [K][n] and obs_next [n][8] are ALL present; no per-frame
//                 observations, no ring, no friction -> no pointer tests inside the frame loop (rollout mode)
//   MODE_SIM      no per-frame output at all (obs_next [n][8] optional): pure simulation
//   MODE_GENERIC  every output nullable, per-frame observations, fused ring append, friction flag
enum { MODE_LEAN = 0, MODE_SIM = 1, MODE_GENERIC = 2 };

// PF ("prefetch"): the variant for small K, where the launch is HBM-bound (176 B per env-step at K = 1 against
// ~300 instructions).  One env per thread keeps loads in flight only during each block's short load phase
// (measured 4.4 TB/s); the PF variant runs a persistent grid (SMs x 3 blocks) and requests the NEXT env's rows
// (4 x LDG.128 + LDG.32 + the action byte) before computing the current one, so every resident thread always has
// 69 bytes in flight.  Costs 18 registers (64 -> ~84, 3 blocks/SM), which K = 1 can afford.
template <typename R, int ACT, int MODE, bool PF = false>   // ACT: 1 discrete, 2 continuous, 3 philox, 4 in-kernel actor
// (ACT 5 is launched with blocks of kPolThreads, 4 of which fit an SM's shared memory: up to 128 registers, no spills)
__global__ void __launch_bounds__(ACT == 5 ? kPolThreads : kThreads,
                                  ACT == 5 ? (sizeof(R) == 4 ? 4 : 2) : ACT == 4 ? 2 : ((PF || MODE == MODE_GENERIC) ? (Occ<R>::blocks < 3 ? Occ<R>::blocks : 3) : Occ<R>::blocks))
step_kernel(const StepArgs<R> a) {
    using V2 = typename std::conditional<sizeof(R) == 4, float2, double2>::type;
    __shared__ BlockStats bs;
    // discrete action -> position nudge (airhockey.py:74-84) as a 5-entry shared table: one clamp + one LDS.64
    // instead of four compares and four selects per frame.  Entry 4 = "any other int": no nudge.
    __shared__ V2 s_nudge[8];
    __shared__ __align__(16) float s_actor[ACT == 4 ? 128 : 4];     // AH_ACT_POLICY: the actor's 114 weights
    extern __shared__ __align__(16) float pol_smem[];                                // AH_ACT_POLICY_MLP (ah_policy.cuh):
    float* s_pol = pol_smem;                                                         //   weights
    float* s_hs = pol_smem + (ACT == 5 ? policy_weight_floats(a.pol.n_floats) : 0);  //   activations (general path)
    if (ACT == 4)
        for (int t = threadIdx.x; t < AH_ACTOR_NWEIGHTS; t += blockDim.x) s_actor[t] = reinterpret_cast<const float*>(a.actions)[t];
    if (ACT == 5) policy_stage(a.pol, s_pol);
    const bool cont_act = (ACT == 2) || (ACT == 5 && a.pol.head == AH_HEAD_CONTINUOUS);
    if (threadIdx.x < AH_NSTATS) bs.v[threadIdx.x] = 0;
    if (threadIdx.x == 0) { bs.len_sq = 0ull; bs.warps_done = 0; }
    if (threadIdx.x < 8) {
        const int t = threadIdx.x;
        s_nudge[t] = V2{t == 3 ? K<R>::step : (t == 2 ? -K<R>::step : R(0)), t == 0 ? K<R>::step : (t == 1 ? -K<R>::step : R(0))};
    }
    __syncthreads();

    const uint32_t n = (uint32_t)a.n;                         // n < 2^31 (checked by ah_create)
    const uint32_t stride = gridDim.x * blockDim.x;
    // the device-side counters are read up front only by the variants that need them in the frame loop (Philox
    // actions, ring append): a dependent L2 round-trip at the head of every thread is what the others avoid
    unsigned long long frame0 = 0ull, ring0 = 0ull;
    if (ACT == 3 || ACT == 4 || ACT == 5) frame0 = a.blk->frame;
    if (MODE == MODE_GENERIC && a.ring_capacity > 0) ring0 = *a.ring_appended;
    const bool repeat = (a.action_kind == AH_ACT_DISCRETE_REPEAT) || (a.action_kind == AH_ACT_CONTINUOUS_REPEAT);
    const bool dot_fma = a.dot_fma != 0;
    const bool friction = (MODE == MODE_GENERIC) && (a.friction != 0);
    const bool want_obs = (MODE == MODE_GENERIC) && (a.obs_state || a.obs_new_state || a.ring_capacity > 0);
    const int K = a.K;
    const size_t act_step = repeat ? 0 : (size_t)n;           // elements between frame k and k+1 of the action tensor
    int reward_sum = 0, reward_sq = 0, frames = 0, dones = 0, contacts_r = 0, contacts_o = 0, bad = 0, amb = 0;
    unsigned overflow = 0;
    const int64_t ring_slot0 = (MODE == MODE_GENERIC && a.ring_capacity > 0) ? (int64_t)(ring0 % (unsigned long long)a.ring_capacity) : 0;

    const uint32_t i_first = blockIdx.x * blockDim.x + threadIdx.x;
    RawEnv<R> raw = {}, raw_next = {};
    int ia_pf = -1, ia_pf_next = -1;
    if (PF && i_first < n) { raw = load_raw<R>(a.st, i_first); ia_pf = reinterpret_cast<const int8_t*>(a.actions)[i_first]; }

    for (uint32_t i = i_first; i < n; i += stride) {
        Env<R> e;
        if (PF) {
            e = unpack_env<R>(raw);
            const uint32_t i_next = i + stride;
            if (i_next < n && i_next > i) {
                raw_next = load_raw<R>(a.st, i_next);
                ia_pf_next = reinterpret_cast<const int8_t*>(a.actions)[i_next];
            }
        } else {
            e = load_env<R>(a.st, i);
        }
        // 10-goal flag carried into the frame (only reachable through injected state; see game.py:169-177)
        int go = (e.robot_score == 10) ? 1 : ((e.opp_score == 10) ? 2 : 0);
        bool speed_dirty = speed_out_of_limits<R>(e.pdx, e.pdy);   // only injected states can start out of limits
        // game_frames / reward_sum are not incremented per frame: frames-in-game = gf_base + frames done in this
        // launch, and the launch's reward sum is recovered from game_return (+ the returns of finished games)
        int gf_base = e.game_frames;
        int ret_credit = -e.game_return;
        Philox4 ablock = {};
        uint32_t ablock_id = 0xffffffffu;
        unsigned long long sblock_id = ~0ull;                  // AH_ACT_POLICY: cached block of 4 sampling uniforms
        PolicyRng prng;                                        // AH_ACT_POLICY_MLP: the same, for both uniform streams
        int64_t ring_slot = ring_slot0;                        // slot of frame k: (appended so far + k) mod capacity
        // warp-uniform per-frame base pointers (advanced by n per frame); the thread adds its 32-bit env index
        const int8_t* act_i8 = reinterpret_cast<const int8_t*>(a.actions);
        const V2* act_v2 = reinterpret_cast<const V2*>(a.actions);
        R* reward_k = a.reward; uint8_t* done_k = a.done; int8_t* score_k = a.score; uint8_t* over_k = a.game_over;
        // software-pipelined action fetch: frame k+1's action is requested before frame k is computed
        int ia_next = -1; V2 av_next = {};
        if (ACT == 1) ia_next = PF ? ia_pf : (int)act_i8[i];
        if (ACT == 2) av_next = act_v2[i];
        for (int k = 0; k < K; ++k) {
            // ---- the robot's action for this frame
            int ia = -1; R ax = R(0), ay = R(0);
            if (ACT == 1) {
                ia = ia_next;
                act_i8 += act_step;
                if (k + 1 < K) ia_next = act_i8[i];
                const V2 nd = s_nudge[min((unsigned)ia, 4u)];
                ax = nd.x; ay = nd.y;
            } else if (ACT == 2) {
                ax = av_next.x; ay = av_next.y;
                act_v2 += act_step;
                if (k + 1 < K) av_next = act_v2[i];
            } else if (ACT == 4) {
                // closed loop: Agent.get_action (agent.py:42-48) -- the actor is evaluated on the env's current
                // observation straight from the registers, and an action is sampled from its softmax
                const Obs<R> ob = get_state<R>(e, true);
                const float x[8] = {(float)ob.v[0], (float)ob.v[1], (float)ob.v[2], (float)ob.v[3],
                                    (float)ob.v[4], (float)ob.v[5], (float)ob.v[6], (float)ob.v[7]};
                float p[4];
                actor_forward(s_actor, x, p);
                const unsigned long long f = frame0 + (unsigned long long)k;
                if ((f >> 2) != sblock_id) { ablock = sample_block(a.rs.seed, a.rs.env_id0 + (uint64_t)i, f); sblock_id = f >> 2; }
                ia = actor_pick(p, false, sample_uniform(ablock, f));
                const V2 nd = s_nudge[ia];
                ax = nd.x; ay = nd.y;
                if (a.probs_out) {
                    typename Vec4T<R>::type q; q.x = p[0]; q.y = p[1]; q.z = p[2]; q.w = p[3];
                    reinterpret_cast<typename Vec4T<R>::type*>(a.probs_out)[(int64_t)k * n + i] = q;
                }
            } else if (ACT == 5) {
                // closed loop with a generic dense policy + exploration strategy (ah_policy.cuh)
                const Obs<R> ob = get_state<R>(e, true);
                const float x[8] = {(float)ob.v[0], (float)ob.v[1], (float)ob.v[2], (float)ob.v[3],
                                    (float)ob.v[4], (float)ob.v[5], (float)ob.v[6], (float)ob.v[7]};
                float y[8], v[8];
                policy_forward(a.pol, s_pol, s_hs, x, y);
                policy_head(a.pol, y, v);
                const PolicyAction pa = policy_select(a.pol, v, a.rs.seed, a.rs.env_id0 + (uint64_t)i, frame0 + (unsigned long long)k, (int64_t)i, prng);
                if (cont_act) {
                    ax = (R)pa.ax; ay = (R)pa.ay;
                    if (a.actions_real_out) reinterpret_cast<V2*>(a.actions_real_out)[(int64_t)k * n + i] = V2{ax, ay};
                } else {
                    ia = pa.ia;
                    const V2 nd = s_nudge[min((unsigned)ia, 4u)];
                    ax = nd.x; ay = nd.y;
                }
                if (a.probs_out) {
                    const int nv = cont_act ? 2 : a.pol.n_actions;
                    R* dst = a.probs_out + ((int64_t)k * n + i) * nv;
                    if (sizeof(R) == 4 && nv == 4 && (reinterpret_cast<uintptr_t>(a.probs_out) & 15u) == 0) {       // one 128-bit store per env-frame
                        typename Vec4T<R>::type q4; q4.x = (R)v[0]; q4.y = (R)v[1]; q4.z = (R)v[2]; q4.w = (R)v[3];
                        *reinterpret_cast<typename Vec4T<R>::type*>(dst) = q4;
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) if (q < nv) dst[q] = (R)v[q];
                    }
                }
            } else {
                const unsigned long long f = frame0 + (unsigned long long)k;
                const uint32_t bid = (uint32_t)(f >> 6);
                if (bid != ablock_id) { ablock = action_block(a.rs.seed, a.rs.env_id0 + (uint64_t)i, bid); ablock_id = bid; }
                ia = action_from_block(ablock, (uint32_t)(f & 63u));
                const V2 nd = s_nudge[ia];
                ax = nd.x; ay = nd.y;
            }
            // ---- robot.move(a): sub-update #1                      game.py:136 -> agent.py:32-36
            move_robot<R>(e, cont_act, ax, ay);
            physics_tail<R, false>(e, dot_fma, friction, speed_dirty, contacts_r, contacts_o);
            // observations are stored the moment they exist (not held in registers until the end of the frame)
            const int64_t row = (int64_t)k * n + i;
            const int64_t rrow = ring_slot * (int64_t)n + i;
            if (MODE == MODE_GENERIC && want_obs) {                             // Observation.state      agent.py:61
                const Obs<R> s0 = get_state<R>(e, true);
                if (a.obs_state) store_obs<R>(a.obs_state, row, s0);
                if (a.ring_capacity > 0) store_obs<R>(a.ring_state, rrow, s0);
            }
            // ---- robot.step(a) -> self.move(a): sub-update #2      agent.py:64
            move_robot<R>(e, cont_act, ax, ay);
            physics_tail<R, false>(e, dot_fma, friction, speed_dirty, contacts_r, contacts_o);
            if (MODE == MODE_GENERIC && want_obs) {                             // Observation.new_state  agent.py:67
                const Obs<R> s1 = get_state<R>(e, true);
                if (a.obs_new_state) store_obs<R>(a.obs_new_state, row, s1);
                if (a.ring_capacity > 0) store_obs<R>(a.ring_new_state, rrow, s1);
            }
            // ---- Rewards                                           agent.py:70-71
            int rw, sc; bool dn;
            rewards_call<R>(e, dot_fma, rw, sc, dn, amb);
            e.game_return += rw;
            reward_sq += rw * rw;
            // ---- env.update_score(score)                           game.py:142 -> airhockey.py:149-169
            if (sc != 0) {
                if (sc > 0) { e.robot_score += 1; stat_add(bs, AH_STAT_ROBOT_GOALS, 1); }
                else { e.opp_score += 1; stat_add(bs, AH_STAT_OPPONENT_GOALS, 1); }
                env_reset<R>(e, false, a.rs, i, overflow);
                // stats(): win flags are read BEFORE the opponent moves and before the total reset (game.py:101-109)
                go = (e.robot_score == 10) ? 1 : go;
                go = (e.opp_score == 10) ? 2 : go;
            }
            dones += dn ? 1 : 0;
            // ---- env.update_state(agent_name="computer"): sub-update #3   game.py:166
            computer_ai<R>(e);
            physics_tail<R, true>(e, dot_fma, friction, speed_dirty, contacts_r, contacts_o);
            // ---- outputs, straight into the caller's tensors
            if (MODE == MODE_LEAN) {
                reward_k[i] = (R)rw; done_k[i] = dn ? 1 : 0; score_k[i] = (int8_t)sc; over_k[i] = (uint8_t)go;
                reward_k += n; done_k += n; score_k += n; over_k += n;
            } else if (MODE == MODE_GENERIC) {
                if (a.reward) a.reward[row] = (R)rw;
                if (a.done) a.done[row] = dn ? 1 : 0;
                if (a.score) a.score[row] = (int8_t)sc;
                if (a.game_over) a.game_over[row] = (uint8_t)go;
                if (!cont_act && a.actions_out) a.actions_out[row] = (int8_t)ia;
                if (a.ring_capacity > 0) {      // fused MemoryBuffer.append(Observation(...))   lib/buffer.py:24-32
                    ring_slot = (ring_slot + 1 == a.ring_capacity) ? 0 : ring_slot + 1;
                    if (cont_act) reinterpret_cast<V2*>(a.ring_action)[rrow] = V2{ax, ay};
                    else reinterpret_cast<int8_t*>(a.ring_action)[rrow] = (int8_t)ia;
                    a.ring_reward[rrow] = (R)rw;
                    a.ring_done[rrow] = dn ? 1 : 0;
                }
            }
            // ---- 10-goal game over                                 game.py:169-177
            if (go != 0) {
                stat_add(bs, AH_STAT_GAMES, 1);
                stat_add(bs, go == 1 ? AH_STAT_ROBOT_WINS : AH_STAT_OPPONENT_WINS, 1);
                const int len = gf_base + k + 1;
                stat_add(bs, AH_STAT_GAME_LEN_SUM, len);
                atomicAdd(&bs.len_sq, (unsigned long long)len * (unsigned long long)len);
                stat_add(bs, AH_STAT_GAME_RETURN_SUM, e.game_return);
                ret_credit += e.game_return;
                gf_base = -(k + 1); e.game_return = 0;
                env_reset<R>(e, true, a.rs, i, overflow);
                speed_dirty = a.rs.table != nullptr;     // Philox draws are within the limits; table values may not be
                go = 0;
            }
            if (MODE == MODE_GENERIC && a.obs_next && a.obs_next_per_frame)
                store_obs<R>(a.obs_next, (int64_t)k * n + i, get_state<R>(e, true));
        }
        // the next frame's policy input (agent.py:46)
        if (a.obs_next && !(MODE == MODE_GENERIC && a.obs_next_per_frame)) store_obs<R>(a.obs_next, i, get_state<R>(e, true));
        frames += K;
        e.game_frames = gf_base + K;
        reward_sum += e.game_return + ret_credit;
        bad += (isfinite(e.px) && isfinite(e.py) && isfinite(e.pdx) && isfinite(e.pdy)) ? 0 : 1;
        store_env<R>(a.st, i, e);
        if (PF) { raw = raw_next; ia_pf = ia_pf_next; }
    }

    finish_stats(bs, a.blk, a.stats_bank, a.collect_stats, frames, reward_sum, reward_sq, dones, contacts_r, contacts_o, (int)overflow, bad, amb, K,
                 a.ring_capacity > 0 ? a.ring_appended : nullptr, AH_RING_COUNTERS);
}

}  // namespace ah
#include "ah_step_packed.cuh"
namespace ah {

// ------------------------------------------------------------------------------------------------ small kernels
template <typename R>
__global__ void init_state_kernel(StatePtrs<R> st, int64_t n) {     // airhockey.py:25-63, puck.py:16, rewards.py:31
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Env<R> e;
    e.px = K<R>::mid_x; e.py = K<R>::mid_y; e.pdx = R(-3); e.pdy = R(0);
    e.rx = K<R>::mid_x - R(100); e.ry = K<R>::mid_y; e.rdx = R(0); e.rdy = R(0);
    e.ox = K<R>::mid_x + R(100); e.oy = K<R>::mid_y; e.odx = R(0); e.ody = R(0);
    e.prev_dist = (R)INFINITY;
    e.robot_score = 0; e.opp_score = 0; e.resets = 0; e.game_frames = 0; e.game_return = 0;
    store_env<R>(st, i, e);
}

template <typename R>
__global__ void reset_kernel(StatePtrs<R> st, int64_t n, int total, const uint8_t* mask, ResetSrc rs, DevBlock* blk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    Env<R> e = load_env<R>(st, i);
    unsigned overflow = 0;
    env_reset<R>(e, total != 0, rs, i, overflow);
    if (total) { e.game_frames = 0; e.game_return = 0; }
    store_env<R>(st, i, e);
    if (overflow) atomicAdd(&blk->stats[0][AH_STAT_RESET_TABLE_OVERFLOW], 1ull);
}

// AirHockey.update_state  airhockey.py:96-134 : one sub-update
template <typename R>
__global__ void update_state_kernel(StatePtrs<R> st, int64_t n, const void* actions, int action_kind, int agent,
                                    int dot_fma, int friction) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using V2 = typename std::conditional<sizeof(R) == 4, float2, double2>::type;
    Env<R> e = load_env<R>(st, i);
    if (agent == AH_AGENT_ROBOT) {                                   // :100-101
        if (action_kind == AH_ACT_CONTINUOUS) {
            const V2 v = reinterpret_cast<const V2*>(actions)[i];
            move_robot<R>(e, true, v.x, v.y);
        } else {
            const int ia = actions ? (int)reinterpret_cast<const int8_t*>(actions)[i] : -1;
            R nx, ny;
            action_nudge<R>(ia, nx, ny);
            move_robot<R>(e, false, nx, ny);
        }
    } else if (agent == AH_AGENT_HUMAN) {                            // :102-105 teleport, then update()
        const V2 v = reinterpret_cast<const V2*>(actions)[i];
        e.ox = v.x; e.oy = v.y;
        mallet_update<R>(e.ox, e.oy, e.odx, e.ody, K<R>::opp_lo, K<R>::opp_hi);
    } else {                                                         // :106-107
        computer_ai<R>(e);
    }
    int c0 = 0, c1 = 0;
    bool dirty = true;
    physics_tail<R, true>(e, dot_fma != 0, friction != 0, dirty, c0, c1);
    store_env<R>(st, i, e);
}

template <typename R>
__global__ void get_state_kernel(StatePtrs<R> st, int64_t n, int robot, R* obs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Env<R> e = load_env<R>(st, i);
    store_obs<R>(obs, i, get_state<R>(e, robot != 0));
}

// AirHockey.update_score  airhockey.py:149-169
template <typename R>
__global__ void update_score_kernel(StatePtrs<R> st, int64_t n, const int8_t* score, ResetSrc rs, DevBlock* blk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int sc = score[i];
    if (sc == 0) return;
    Env<R> e = load_env<R>(st, i);
    unsigned overflow = 0;
    if (sc > 0) e.robot_score += 1; else e.opp_score += 1;
    env_reset<R>(e, false, rs, i, overflow);
    store_env<R>(st, i, e);
    if (overflow) atomicAdd(&blk->stats[0][AH_STAT_RESET_TABLE_OVERFLOW], 1ull);
}

// Rewards.__call__  lib/rewards.py:118-142
template <typename R>
__global__ void rewards_kernel(StatePtrs<R> st, int64_t n, int dot_fma, R* reward, int8_t* score, uint8_t* done) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Env<R> e = load_env<R>(st, i);
    int rw, sc, amb = 0; bool dn;
    rewards_call<R>(e, dot_fma != 0, rw, sc, dn, amb);
    st.prev_dist[i] = e.prev_dist;
    if (reward) reward[i] = (R)rw;
    if (score) score[i] = (int8_t)sc;
    if (done) done[i] = dn ? 1 : 0;
}

// ---- on-device action selection: the reference's discrete PPO actor + softmax sampling, one thread per env.
// Weights (114 floats) are staged in shared memory once per block and read as warp-uniform LDS.128 broadcasts.
template <typename R>
__global__ void __launch_bounds__(256) actor_sample_kernel(int64_t n, const float* __restrict__ w, const R* __restrict__ obs,
                                                           R* __restrict__ probs, int8_t* __restrict__ actions, int greedy,
                                                           uint64_t seed, uint64_t env_id0, const DevBlock* blk) {
    __shared__ __align__(16) float sw[128];
    if (threadIdx.x < AH_ACTOR_NWEIGHTS) sw[threadIdx.x] = w[threadIdx.x];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using V = typename Vec4T<R>::type;
    const V o0 = reinterpret_cast<const V*>(obs)[2 * i], o1 = reinterpret_cast<const V*>(obs)[2 * i + 1];
    const float x[8] = {(float)o0.x, (float)o0.y, (float)o0.z, (float)o0.w, (float)o1.x, (float)o1.y, (float)o1.z, (float)o1.w};
    float p[4];
    actor_forward(sw, x, p);
    float u = 0.f;
    if (!greedy) {      // uniform in [0,1) keyed on (seed, global env id), counter = the current frame (stream 2)
        const unsigned long long f = blk->frame;
        u = sample_uniform(sample_block(seed, env_id0 + (uint64_t)i, f), f);
    }
    const int act = actor_pick(p, greedy != 0, u);
    actions[i] = (int8_t)act;
    if (probs) { V q; q.x = p[0]; q.y = p[1]; q.z = p[2]; q.w = p[3]; reinterpret_cast<V*>(probs)[i] = q; }
}

// host-transfer packing (see ah_pack_host in include/ah_b200.h): 4 envs per thread, 128-bit loads of the rewards
template <typename R>
__global__ void pack_host_kernel(int64_t n, int K, const R* __restrict__ reward, const uint8_t* __restrict__ done,
                                 const R* __restrict__ obs, uint8_t* __restrict__ events, short4* __restrict__ pos,
                                 float4* __restrict__ vel) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (events) {
        for (int k = 0; k < K; ++k) {
            const int64_t row = (int64_t)k * n + i;
            const int rw = (int)reward[row] + 4;
            events[row] = (uint8_t)((rw & 15) | ((done[row] ? 1 : 0) << 4));
        }
    }
    const R* o = obs + 8 * i;
    pos[i] = make_short4((short)o[0], (short)o[1], (short)o[2], (short)o[3]);
    vel[i] = make_float4((float)o[4], (float)o[5], (float)o[6], (float)o[7]);
}

__global__ void stats_copy_kernel(DevBlock* blk, int bank, int64_t* out, int clear) {
    const int q = threadIdx.x;
    if (q < AH_NSTATS) {
        out[q] = (int64_t)blk->stats[bank][q];
        if (clear) blk->stats[bank][q] = 0ull;
    }
}

// ---- statistics copy + clear + all-reduce over NVLink peer memory in ONE kernel (no NCCL launch, no host round trip):
// every rank pushes its vector into its slot of every rank's receive buffer (P2P stores, then a system-scope fence, then
// the epoch flag), spins until all `world` slots of its OWN buffer carry the current epoch, and sums them in rank order
// (so every rank computes the same sum).  Two parities: a rank can be at most one exchange ahead of a peer (it needs the
// peer's contribution to finish its own), so the slot it overwrites two exchanges later has been consumed.  A spin that
// exceeds ~2 s (a peer that never calls) reports -1 in out[AH_NSTATS - 1] instead of hanging the GPU.
struct PeerPtrs { PeerSlot* p[AH_MAX_PEERS]; };

__global__ void stats_allreduce_kernel(DevBlock* blk, int bank, int clear, int rank, int world, PeerPtrs peers, int64_t* out) {
    const int q = threadIdx.x;                       // one warp: lane q owns statistic q
    __shared__ unsigned long long s_epoch;
    if (q == 0) { s_epoch = blk->xchg_epoch + 1ull; blk->xchg_epoch = s_epoch; }
    __syncwarp();
    const unsigned long long epoch = s_epoch;
    const int par = (int)(epoch & 1ull);
    long long mine = 0;
    if (q < AH_NSTATS) {
        mine = (long long)blk->stats[bank][q];
        if (clear) blk->stats[bank][q] = 0ull;
        for (int r = 0; r < world; ++r) peers.p[r][par * AH_MAX_PEERS + rank].vals[q] = mine;
    }
    __threadfence_system();
    __syncwarp();
    if (q < world) {
        volatile unsigned long long* flag = &peers.p[q][par * AH_MAX_PEERS + rank].epoch;
        *flag = epoch;
    }
    __threadfence_system();
    // wait for every source's slot in MY buffer
    bool ok = true;
    if (q < world) {
        volatile unsigned long long* flag = &peers.p[rank][par * AH_MAX_PEERS + q].epoch;
        const long long t0 = clock64();
        while (*flag != epoch) {
            if (clock64() - t0 > 4000000000ll) { ok = false; break; }
            __nanosleep(200);
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    __threadfence_system();
    if (q < AH_NSTATS) {
        long long sum = 0;
        for (int r = 0; r < world; ++r) sum += ((volatile long long*)peers.p[rank][par * AH_MAX_PEERS + r].vals)[q];
        out[q] = (ok || q != AH_NSTATS - 1) ? sum : -1;
    }
}

__global__ void counters_set_kernel(DevBlock* blk, unsigned long long frame) { blk->frame = frame; }
__global__ void counters_get_kernel(const DevBlock* blk, unsigned long long* out1) { out1[0] = blk->frame; }

// ---- MemoryBuffer.sample (lib/buffer.py:47-50) on the device: a keyed bijection of [0, population) instead of a
// materialised permutation.  6-round balanced Feistel network over 2^bits >= population (bits even), cycle-walked.
struct RingKeys { uint32_t k[8]; };

__host__ __device__ __forceinline__ uint32_t ring_round(uint32_t v, uint32_t key) {
    v = (v + key) * 0x9E3779B1u;
    v ^= v >> 15;
    v *= 0x85EBCA77u;
    v ^= v >> 13;
    v *= 0xC2B2AE3Du;
    v ^= v >> 16;
    return v;
}

__device__ __forceinline__ uint64_t ring_permute(uint64_t x, int half, const RingKeys& rk) {
    const uint32_t mask = (half >= 32) ? 0xffffffffu : ((1u << half) - 1u);
    uint32_t L = (uint32_t)(x >> half) & mask, Rr = (uint32_t)x & mask;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const uint32_t t = L ^ (ring_round(Rr, rk.k[r]) & mask);
        L = Rr; Rr = t;
    }
    return ((uint64_t)L << half) | Rr;
}

template <typename R>
__global__ void ring_sample_kernel(int64_t n, int64_t capacity, const unsigned long long* appended_p, int continuous,
                                   int64_t batch, uint64_t seed, uint64_t draw,
                                   const R* __restrict__ r_state, const R* __restrict__ r_new_state, const void* __restrict__ r_action,
                                   const R* __restrict__ r_reward, const uint8_t* __restrict__ r_done,
                                   R* state, void* action, R* reward, uint8_t* done, R* new_state, int64_t* index,
                                   const short* __restrict__ r_state_pos, const short* __restrict__ r_new_state_pos) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= batch) return;
    using V = typename Vec4T<R>::type;
    using V2 = typename std::conditional<sizeof(R) == 4, float2, double2>::type;
    const unsigned long long appended = *appended_p;
    const uint64_t held = appended < (unsigned long long)capacity ? appended : (unsigned long long)capacity;
    const uint64_t population = held * (uint64_t)n;
    if (population == 0) return;
    int bits = 2;
    while (bits < 64 && (1ull << bits) < population) bits += 2;
    RingKeys rk;
    const Philox4 a = philox4x32_10((uint32_t)draw, (uint32_t)(draw >> 32), 0x52494E47u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const Philox4 b = philox4x32_10((uint32_t)draw, (uint32_t)(draw >> 32), 0x52494E47u, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
    rk.k[0] = a.x; rk.k[1] = a.y; rk.k[2] = a.z; rk.k[3] = a.w; rk.k[4] = b.x; rk.k[5] = b.y; rk.k[6] = b.z; rk.k[7] = b.w;
    uint64_t f = (uint64_t)s % population;                // batch <= population is the caller's contract
    do { f = ring_permute(f, bits >> 1, rk); } while (f >= population);
    const uint64_t age = f / (uint64_t)n, env = f % (uint64_t)n;
    const uint64_t slot = (appended - 1ull - age) % (unsigned long long)capacity;     // age 0 = the newest tuple
    const int64_t row = (int64_t)(slot * (uint64_t)n + env);
    V* dst0 = reinterpret_cast<V*>(state) + 2 * s;
    V* dst1 = reinterpret_cast<V*>(new_state) + 2 * s;
    if (r_state_pos) {                                    // AH_RING_COMPACT: int16 locations + velocities, int8 reward
        const short4 p0 = reinterpret_cast<const short4*>(r_state_pos)[row], p1 = reinterpret_cast<const short4*>(r_new_state_pos)[row];
        V a0, a1;
        a0.x = (R)p0.x; a0.y = (R)p0.y; a0.z = (R)p0.z; a0.w = (R)p0.w;
        a1.x = (R)p1.x; a1.y = (R)p1.y; a1.z = (R)p1.z; a1.w = (R)p1.w;
        dst0[0] = a0; dst0[1] = reinterpret_cast<const V*>(r_state)[row];
        dst1[0] = a1; dst1[1] = reinterpret_cast<const V*>(r_new_state)[row];
        reward[s] = (R)reinterpret_cast<const int8_t*>(r_reward)[row];
    } else {
        const V* src0 = reinterpret_cast<const V*>(r_state) + 2 * row;
        const V* src1 = reinterpret_cast<const V*>(r_new_state) + 2 * row;
        dst0[0] = src0[0]; dst0[1] = src0[1];
        dst1[0] = src1[0]; dst1[1] = src1[1];
        reward[s] = r_reward[row];
    }
    if (continuous) reinterpret_cast<V2*>(action)[s] = reinterpret_cast<const V2*>(r_action)[row];
    else reinterpret_cast<int8_t*>(action)[s] = reinterpret_cast<const int8_t*>(r_action)[row];
    done[s] = r_done[row];
    if (index) index[s] = (int64_t)f;
}

// ---- C51's categorical projection (lib/agents/c51/c51.py:175-190) as ONE kernel: a thread owns sample i and writes the
// A x atoms entries of m_prob[.][i][.] -- zeros except the row of the action taken, which receives the projected
// distribution of the optimal next action, accumulated in the reference's order (j ascending, lower atom first).
// As written there: m_l = floor(bj), m_u = ceil(bj); when bj is an integer both weights are 0 and the mass is dropped.
template <typename R>
__global__ void c51_project_kernel(int64_t B, int A, int atoms, R v_min, R v_max, R gamma, const R* __restrict__ z_opt,
                                   const int8_t* __restrict__ action, const R* __restrict__ reward, const uint8_t* __restrict__ done,
                                   R* __restrict__ m_prob) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const R dz = (v_max - v_min) / (R)(atoms - 1);
    for (int a = 0; a < A; ++a)
        for (int j = 0; j < atoms; ++j) m_prob[((int64_t)a * B + i) * atoms + j] = R(0);
    const int act = action[i];
    if (act < 0 || act >= A) return;
    R* row = m_prob + ((int64_t)act * B + i) * atoms;
    const R rw = reward[i];
    if (done[i]) {                                                   // :181-187 the distribution collapses to one point
        const R tz = fmin(v_max, fmax(v_min, rw));
        const R bj = (tz - v_min) / dz;
        const R ml = floor(bj), mu = ceil(bj);
        row[(int)ml] += mu - bj;
        row[(int)mu] += bj - ml;
    } else {
        for (int j = 0; j < atoms; ++j) {                            // :189-194
            const R zj = v_min + (R)j * dz;
            const R tz = fmin(v_max, fmax(v_min, rw + gamma * zj));
            const R bj = (tz - v_min) / dz;
            const R ml = floor(bj), mu = ceil(bj);
            const R p = z_opt[i * atoms + j];
            row[(int)ml] += p * (mu - bj);
            row[(int)mu] += p * (bj - ml);
        }
    }
}

// a*b+c must NOT be contracted in this translation unit (FP64 validation build): discriminating probe
__global__ void nofma_probe_kernel(double a, double b, double c, double* out) { out[0] = a * b + c; }

}  // namespace ah

// ================================================================================================ C ABI
struct ah_env {
    int64_t n;
    int dtype, device;
    uint64_t seed, env_id0;
    ah_config cfg;
    void *puck, *robot, *opp, *prev_dist;
    int32_t* meta;
    ah::DevBlock* blk;
    int sms, blocks_f32, blocks_f64;
    int last_cuda;
    // peer statistics exchange (ah_peer_export / ah_peer_attach / ah_stats_allreduce)
    ah::PeerSlot* xchg;                    // this rank's receive buffer [2 parities][AH_MAX_PEERS]
    ah::PeerSlot* peer[AH_MAX_PEERS];      // every rank's receive buffer, mapped into this process (peer[rank] == xchg)
    int rank, world;
};

namespace {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

inline int cuda_fail(ah_env* env, cudaError_t e) { if (env) env->last_cuda = (int)e; return AH_ECUDA; }

template <typename R> ah::StatePtrs<R> state_ptrs(const ah_env* env) {
    ah::StatePtrs<R> s;
    using V = typename ah::Vec4T<R>::type;
    s.puck = reinterpret_cast<V*>(env->puck); s.robot = reinterpret_cast<V*>(env->robot);
    s.opp = reinterpret_cast<V*>(env->opp); s.prev_dist = reinterpret_cast<R*>(env->prev_dist);
    s.meta = reinterpret_cast<int4*>(env->meta);
    return s;
}

inline ah::ResetSrc reset_src(const ah_env* env, const double* table, int R) {
    ah::ResetSrc rs; rs.table = table; rs.R = R; rs.seed = env->seed; rs.env_id0 = env->env_id0;
    rs.key = ah::reset_key(env->seed); return rs;
}

inline unsigned small_grid(int64_t n) { return (unsigned)((n + 255) / 256); }

inline size_t policy_smem_bytes(const ah::PolicyDev& d, int threads) {
    return sizeof(float) * ((size_t)ah::policy_weight_floats(d.n_floats) + (d.narrow ? 0 : 2 * (size_t)ah::kPolMaxW * threads));
}

// ah_policy (host struct, validated) -> the by-value kernel argument
int make_policy_dev(const ah_policy* p, ah::PolicyDev* d, bool need_weights) {
    if (!p || !d) return AH_EINVAL;
    memset(d, 0, sizeof *d);
    if (p->head < AH_HEAD_SOFTMAX || p->head > AH_HEAD_CONTINUOUS || p->select < AH_SELECT_SAMPLE || p->select > AH_SELECT_OU) return AH_EINVAL;
    const bool cont = p->head == AH_HEAD_CONTINUOUS;
    if (cont != (p->select == AH_SELECT_GAUSSIAN || p->select == AH_SELECT_OU || (cont && p->select == AH_SELECT_GREEDY))) return AH_EINVAL;
    if (!cont && (p->n_actions < 1 || p->n_actions > 7)) return AH_EINVAL;
    if (p->select == AH_SELECT_OU && (!p->noise_state || (reinterpret_cast<uintptr_t>(p->noise_state) & 7u))) return AH_EINVAL;
    d->head = p->head; d->select = p->select; d->n_actions = cont ? 2 : p->n_actions;
    if (need_weights) {
        if (p->n_layers < 1 || p->n_layers > AH_POLICY_MAX_LAYERS || p->width[0] != 8 || !p->weights || !aligned16(p->weights)) return AH_EINVAL;
        const int last = p->width[p->n_layers];
        if (last != (cont ? 2 : (p->head == AH_HEAD_DUELING ? p->n_actions + 1 : p->n_actions))) return AH_EINVAL;
        int off = 0;
        d->n_layers = p->n_layers;
        for (int l = 0; l <= p->n_layers; ++l) {
            if (p->width[l] < 1 || p->width[l] > AH_POLICY_MAX_WIDTH) return AH_EINVAL;
            d->width[l] = p->width[l];
        }
        for (int l = 0; l < p->n_layers; ++l) {
            if (p->activation[l] < AH_ACTV_LINEAR || p->activation[l] > AH_ACTV_TANH) return AH_EINVAL;
            d->act[l] = p->activation[l];
            const int outp = (p->width[l + 1] + 3) & ~3;
            d->w_off[l] = off; off += p->width[l] * outp;
            d->b_off[l] = off; off += outp;
        }
        d->n_floats = off;
        d->weights = p->weights;
        d->narrow = 1;
        for (int l = 0; l <= p->n_layers; ++l) if (p->width[l] > 8) d->narrow = 0;
        if (getenv("AH_POLICY_NO_NARROW")) d->narrow = 0;      // tuning / test aid: force the general path
    }
    // epsilon schedule: the number of decrements after which epsilon <= final_epsilon (strategies.py:39-43)
    d->eps0 = p->epsilon; d->eps_final = p->final_epsilon; d->observe = p->observe;
    d->eps_delta = p->explore > 0 ? (p->initial_epsilon - p->final_epsilon) / (float)p->explore : 0.0f;
    d->max_decrements = 0;
    if (d->eps_delta > 0.0f && p->epsilon > p->final_epsilon) {
        double m = ceil(((double)p->epsilon - (double)p->final_epsilon) / (double)d->eps_delta);
        d->max_decrements = m > 2e9 ? 2000000000 : (int)m;
    }
    d->inv_tau = p->tau != 0.0f ? 1.0f / p->tau : 1.0f; d->clip_lo = p->clip_lo; d->clip_hi = p->clip_hi;
    d->mu = p->mu; d->sigma0 = p->sigma; d->theta = p->theta; d->dt = p->dt; d->sqrt_dt = sqrtf(p->dt > 0.0f ? p->dt : 0.0f);
    if (p->sigma_min >= 0.0f && p->n_steps_annealing > 0) {      // random.py:17-24
        d->sigma_m = -(p->sigma - p->sigma_min) / (float)p->n_steps_annealing; d->sigma_min = p->sigma_min;
    } else { d->sigma_m = 0.0f; d->sigma_min = p->sigma; }
    d->noise_state = p->noise_state;
    return AH_OK;
}

template <typename R>
int launch_step(ah_env* env, const ah_step_io* io, int K, cudaStream_t s) {
    ah::StepArgs<R> a;
    memset(&a, 0, sizeof a);
    a.st = state_ptrs<R>(env); a.n = env->n; a.K = K;
    a.actions = io->actions; a.action_kind = io->action_kind;
    a.rs = reset_src(env, io->reset_table, io->reset_table_len);
    a.obs_state = (R*)io->obs_state; a.obs_new_state = (R*)io->obs_new_state; a.reward = (R*)io->reward;
    a.done = io->done; a.score = io->score; a.game_over = io->game_over;
    a.obs_next = (R*)io->obs_next; a.obs_next_per_frame = io->obs_next_per_frame;
    a.collect_stats = io->collect_stats;
    if (io->ring) {
        a.ring_state = (R*)io->ring->state; a.ring_new_state = (R*)io->ring->new_state; a.ring_action = io->ring->action;
        a.ring_reward = (R*)io->ring->reward; a.ring_done = io->ring->done; a.ring_capacity = io->ring->capacity;
        a.ring_appended = reinterpret_cast<unsigned long long*>(io->ring->appended);
    }
    a.dot_fma = env->cfg.dot_fma; a.friction = env->cfg.friction; a.blk = env->blk; a.stats_bank = io->stats_bank;
    a.actions_out = io->actions_out; a.probs_out = (R*)io->probs_out; a.actions_real_out = (R*)io->actions_real_out;
    if (io->action_kind == AH_ACT_POLICY_MLP) {
        const int prc = make_policy_dev(io->policy, &a.pol, true);
        if (prc != AH_OK) return prc;
    }
    const int max_blocks = sizeof(R) == 4 ? env->blocks_f32 : env->blocks_f64;
    // block size: 256 threads, halved (down to 64) while the grid would leave SMs unevenly filled (< 4 blocks per SM):
    // a 65,536-env rollout is 256 blocks of 256 (1.7 per SM: some SMs get twice the work of others) but 1,024 of 64
    int threads = ah::kThreads;
    while (threads > 64 && (env->n + threads - 1) / threads < 4 * (int64_t)env->sms) threads >>= 1;
    int64_t want = (env->n + threads - 1) / threads;
    const unsigned grid = (unsigned)(want < max_blocks ? want : max_blocks);
    // the packed-core kernel (ah_step_packed.cuh): one env per thread, blocks of kPackedThreads
    ah::PackedArgs<R> p;
    memset(&p, 0, sizeof p);
    p.st = a.st; p.n = env->n; p.K = K;
    p.obs_next = (R*)io->obs_next; p.rs = a.rs; p.collect_stats = io->collect_stats; p.dot_fma = env->cfg.dot_fma; p.blk = env->blk; p.stats_bank = io->stats_bank;
    const int64_t count = io->env_count > 0 ? io->env_count : env->n;
    p.i0 = (uint32_t)io->env_first; p.i_end = (uint32_t)(io->env_first + count);
    int pth = ah::kPackedThreads;
    while (pth > 64 && (count + pth - 1) / pth < 4 * (int64_t)env->sms) pth >>= 1;
    const int64_t pwant = (count + pth - 1) / pth;
    const unsigned pgrid = (unsigned)(pwant < max_blocks ? pwant : max_blocks);
    if (io->events) {                                        // PACKED rollout mode
        p.actions = reinterpret_cast<const uint32_t*>(io->actions); p.events = reinterpret_cast<uint32_t*>(io->events);
        if (io->ring) {
            p.ring_state = (R*)io->ring->state; p.ring_new_state = (R*)io->ring->new_state;
            p.ring_action = reinterpret_cast<int8_t*>(io->ring->action);
            p.ring_reward = (R*)io->ring->reward; p.ring_done = io->ring->done; p.ring_capacity = io->ring->capacity;
            p.ring_appended = reinterpret_cast<unsigned long long*>(io->ring->appended);
            if (io->ring_counter_count > 0) {           // a slice of a step: its own counter(s) and ticket
                p.ring_appended += io->ring_counter_first; p.ring_counters = io->ring_counter_count; p.ticket = io->ring_counter_first;
            } else { p.ring_counters = AH_RING_COUNTERS; p.ticket = 0; }
            if (io->ring->format == AH_RING_COMPACT) {
                p.ring_state_pos = reinterpret_cast<short*>(io->ring->state_pos);
                p.ring_new_state_pos = reinterpret_cast<short*>(io->ring->new_state_pos);
                p.ring_reward8 = reinterpret_cast<int8_t*>(io->ring->reward);
                ah::step_packed_kernel<R, ah::IO_RING_COMPACT><<<pgrid, pth, 0, s>>>(p);
            } else
            ah::step_packed_kernel<R, ah::IO_RING><<<pgrid, pth, 0, s>>>(p);
        } else {
            ah::step_packed_kernel<R, ah::IO_PACKED><<<pgrid, pth, 0, s>>>(p);
        }
        const cudaError_t pe = cudaGetLastError();
        return pe == cudaSuccess ? AH_OK : cuda_fail(env, pe);
    }
    // pick the specialised variant the given pointers allow
    const bool per_frame_any = io->reward || io->done || io->score || io->game_over;
    const bool per_frame_all = io->reward && io->done && io->score && io->game_over;
    const bool heavy = io->obs_state || io->obs_new_state || io->ring || env->cfg.friction || (io->obs_next && io->obs_next_per_frame) ||
                       io->actions_out || io->probs_out || io->action_kind == AH_ACT_POLICY || io->action_kind == AH_ACT_POLICY_MLP;
    int mode = ah::MODE_GENERIC;
    if (!heavy && per_frame_all && io->obs_next) mode = ah::MODE_LEAN;
    else if (!heavy && !per_frame_any) mode = ah::MODE_SIM;
    int act = 0;
    switch (io->action_kind) {
        case AH_ACT_DISCRETE: case AH_ACT_DISCRETE_REPEAT: act = 1; break;
        case AH_ACT_CONTINUOUS: case AH_ACT_CONTINUOUS_REPEAT: act = 2; break;
        case AH_ACT_PHILOX: act = 3; break;
        case AH_ACT_POLICY: act = 4; break;
        case AH_ACT_POLICY_MLP: act = 5; break;
        default: return AH_EINVAL;                            // AH_ACT_DISCRETE_PACKED4 is only valid with io->events
    }
    // the same arithmetic core behind the general API's tensors (K > 2: the small-K launches are HBM-bound and keep the
    // prefetching variant below) and for the pure simulation with in-kernel Philox actions
    if (mode == ah::MODE_LEAN && io->action_kind == AH_ACT_DISCRETE && K > 2 && !env->cfg.no_packed_core) {
        p.actions8 = reinterpret_cast<const int8_t*>(io->actions);
        p.reward = (R*)io->reward; p.done = io->done; p.score = io->score; p.game_over = io->game_over;
        ah::step_packed_kernel<R, ah::IO_UNPACKED><<<pgrid, pth, 0, s>>>(p);
        const cudaError_t pe = cudaGetLastError();
        return pe == cudaSuccess ? AH_OK : cuda_fail(env, pe);
    }
    if (mode == ah::MODE_SIM && act == 3 && !env->cfg.no_packed_core) {
        ah::step_packed_kernel<R, ah::IO_PHILOX><<<pgrid, pth, 0, s>>>(p);
        const cudaError_t pe = cudaGetLastError();
        return pe == cudaSuccess ? AH_OK : cuda_fail(env, pe);
    }
#define AH_LAUNCH(ACT_, MODE_) ah::step_kernel<R, ACT_, MODE_><<<grid, threads, 0, s>>>(a)
#define AH_LAUNCH_MODE(ACT_) do { if (mode == ah::MODE_LEAN) AH_LAUNCH(ACT_, ah::MODE_LEAN); \
        else if (mode == ah::MODE_SIM) AH_LAUNCH(ACT_, ah::MODE_SIM); else AH_LAUNCH(ACT_, ah::MODE_GENERIC); } while (0)
    bool launched = false;
    if constexpr (sizeof(R) == 4) {
        if (act == 1 && mode == ah::MODE_LEAN && K <= 2 && env->n >= 65536 && !env->cfg.no_prefetch) {
            // HBM-bound regime: persistent grid with next-env prefetch (see step_kernel<..., PF>)
            const int64_t pf_want = (env->n + ah::kThreads - 1) / ah::kThreads;
            const unsigned pf_grid = (unsigned)(env->sms * 3 < pf_want ? env->sms * 3 : pf_want);
            ah::step_kernel<R, 1, ah::MODE_LEAN, true><<<pf_grid, ah::kThreads, 0, s>>>(a);
            launched = true;
        }
    }
    if (launched) {
    } else if (act == 4) {
        ah::step_kernel<R, 4, ah::MODE_GENERIC><<<grid, threads, 0, s>>>(a);
    } else if (act == 5) {                                   // blocks of <= kPolThreads; dynamic shared memory sized by the network
        int pt = ah::kPolThreads;
        while (pt > 64 && (env->n + pt - 1) / pt < 4 * (int64_t)env->sms) pt >>= 1;
        const unsigned pg = (unsigned)((env->n + pt - 1) / pt);
        ah::step_kernel<R, 5, ah::MODE_GENERIC><<<pg < (unsigned)max_blocks ? pg : (unsigned)max_blocks, pt, policy_smem_bytes(a.pol, pt), s>>>(a);
    } else if (act == 1) AH_LAUNCH_MODE(1); else if (act == 2) AH_LAUNCH_MODE(2); else AH_LAUNCH_MODE(3);
#undef AH_LAUNCH_MODE
#undef AH_LAUNCH
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? AH_OK : cuda_fail(env, e);
}

}  // namespace

extern "C" {

const char* ah_strerror(int code) {
    switch (code) {
        case AH_OK: return "ok";
        case AH_EINVAL: return "invalid argument";
        case AH_ECUDA: return "CUDA runtime error";
        case AH_ENOMEM: return "out of memory";
        case AH_EAGENT: return "invalid agent";
        case AH_ESTATE: return "state not bound";
        default: return "unknown error";
    }
}

int ah_abi_version(void) { return AH_ABI_VERSION; }

int ah_last_cuda_error(const ah_env* env) { return env ? env->last_cuda : 0; }

int ah_create(ah_env** out, int64_t n, int dtype, int device, uint64_t seed, uint64_t env_id0, const ah_config* cfg) {
    if (!out || n <= 0 || n >= (int64_t(1) << 31) || (dtype != AH_F32 && dtype != AH_F64) || device < 0) return AH_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count) return AH_ECUDA;   // no CPU fallback
    ah_env* env = new (std::nothrow) ah_env();
    if (!env) return AH_ENOMEM;
    memset(env, 0, sizeof *env);
    env->n = n; env->dtype = dtype; env->device = device; env->seed = seed; env->env_id0 = env_id0;
    env->cfg.dot_fma = 1; env->cfg.friction = 0;
    if (cfg) env->cfg = *cfg;
    DeviceGuard g(device);
    cudaError_t e = g.ok ? cudaSuccess : cudaErrorInvalidDevice;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&env->sms, cudaDevAttrMultiProcessorCount, device);
    int occ32 = 0, occ64 = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, ah::step_kernel<float, 1, ah::MODE_LEAN>, ah::kThreads, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, ah::step_kernel<double, 1, ah::MODE_GENERIC>, ah::kThreads, 0);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&env->blk), sizeof(ah::DevBlock));
    if (e == cudaSuccess) e = cudaMemset(env->blk, 0, sizeof(ah::DevBlock));
    if (e != cudaSuccess) { if (env->blk) cudaFree(env->blk); delete env; return AH_ECUDA; }
    // Grid: one thread per environment, blocks of 256, left to the hardware block scheduler.  Measured on B200
    // at n = 2^20, K = 8 (FP32): 148 x 4 persistent blocks 123 us, 4096 blocks 115 us -- the frame loop is
    // ALU-pipe bound, so finer-grained blocks only help by evening out the tail.  The grid-stride loop in the
    // kernel stays for n > kMaxBlocks * 256 and for the AH_GRID_BLOCKS tuning override.
    (void)occ32; (void)occ64;
    env->blocks_f32 = ah::kMaxBlocks;
    env->blocks_f64 = ah::kMaxBlocks;
    if (const char* ov = getenv("AH_GRID_BLOCKS")) {       // tuning aid: override the grid size
        const int b = atoi(ov);
        if (b > 0) { env->blocks_f32 = b; env->blocks_f64 = b; }
    }
    *out = env;
    return AH_OK;
}

int ah_destroy(ah_env* env) {
    if (!env) return AH_EINVAL;
    DeviceGuard g(env->device);
    if (env->blk) cudaFree(env->blk);
    for (int r = 0; r < env->world; ++r)
        if (r != env->rank && env->peer[r]) cudaIpcCloseMemHandle(env->peer[r]);
    if (env->xchg) cudaFree(env->xchg);
    delete env;
    return AH_OK;
}

int ah_bind_state(ah_env* env, void* d_puck, void* d_robot, void* d_opponent, void* d_prev_dist, int32_t* d_meta) {
    if (!env || !d_puck || !d_robot || !d_opponent || !d_prev_dist || !d_meta) return AH_EINVAL;
    if (!aligned16(d_puck) || !aligned16(d_robot) || !aligned16(d_opponent) || !aligned16(d_meta) ||
        (reinterpret_cast<uintptr_t>(d_prev_dist) & 7u)) return AH_EINVAL;
    env->puck = d_puck; env->robot = d_robot; env->opp = d_opponent; env->prev_dist = d_prev_dist; env->meta = d_meta;
    return AH_OK;
}

#define AH_CHECK_BOUND(env) do { if (!(env)) return AH_EINVAL; if (!(env)->puck) return AH_ESTATE; } while (0)
#define AH_DISPATCH(env, call_f32, call_f64) do { if ((env)->dtype == AH_F32) { call_f32; } else { call_f64; } } while (0)
#define AH_LAUNCH_RC(env) do { const cudaError_t e_ = cudaGetLastError(); return e_ == cudaSuccess ? AH_OK : cuda_fail((env), e_); } while (0)

int ah_init_state(ah_env* env, void* stream) {
    AH_CHECK_BOUND(env);
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    AH_DISPATCH(env, (ah::init_state_kernel<float><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<float>(env), env->n)),
                (ah::init_state_kernel<double><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<double>(env), env->n)));
    AH_LAUNCH_RC(env);
}

int ah_reset(ah_env* env, int total, const uint8_t* d_mask, const double* d_reset_table, int reset_table_len, void* stream) {
    AH_CHECK_BOUND(env);
    if (d_reset_table && (reset_table_len <= 0 || !aligned16(d_reset_table))) return AH_EINVAL;
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const ah::ResetSrc rs = reset_src(env, d_reset_table, reset_table_len);
    AH_DISPATCH(env, (ah::reset_kernel<float><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<float>(env), env->n, total, d_mask, rs, env->blk)),
                (ah::reset_kernel<double><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<double>(env), env->n, total, d_mask, rs, env->blk)));
    AH_LAUNCH_RC(env);
}

int ah_update_state(ah_env* env, const void* d_actions, int action_kind, int agent, void* stream) {
    AH_CHECK_BOUND(env);
    if (agent != AH_AGENT_ROBOT && agent != AH_AGENT_HUMAN && agent != AH_AGENT_COMPUTER) return AH_EAGENT;
    if (agent == AH_AGENT_HUMAN && (!d_actions || action_kind != AH_ACT_CONTINUOUS)) return AH_EINVAL;
    if (agent == AH_AGENT_ROBOT) {
        if (action_kind == AH_ACT_CONTINUOUS) { if (!d_actions || (reinterpret_cast<uintptr_t>(d_actions) & (env->dtype == AH_F32 ? 7u : 15u))) return AH_EINVAL; }
        else if (action_kind == AH_ACT_DISCRETE) { if (!d_actions) return AH_EINVAL; }
        else if (action_kind != AH_ACT_NONE) return AH_EINVAL;   // AH_ACT_NONE: `update_state("", "robot")`, no nudge
    }
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    AH_DISPATCH(env, (ah::update_state_kernel<float><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<float>(env), env->n, d_actions, action_kind, agent, env->cfg.dot_fma, env->cfg.friction)),
                (ah::update_state_kernel<double><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<double>(env), env->n, d_actions, action_kind, agent, env->cfg.dot_fma, env->cfg.friction)));
    AH_LAUNCH_RC(env);
}

int ah_get_state(ah_env* env, int agent, void* d_obs, void* stream) {
    AH_CHECK_BOUND(env);
    if (!d_obs || !aligned16(d_obs)) return AH_EINVAL;
    // get_state: any name other than "robot" selects the opponent's mallet (airhockey.py:139)
    const int robot = agent == AH_AGENT_ROBOT;
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    AH_DISPATCH(env, (ah::get_state_kernel<float><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<float>(env), env->n, robot, (float*)d_obs)),
                (ah::get_state_kernel<double><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<double>(env), env->n, robot, (double*)d_obs)));
    AH_LAUNCH_RC(env);
}

int ah_update_score(ah_env* env, const int8_t* d_score, const double* d_reset_table, int reset_table_len, void* stream) {
    AH_CHECK_BOUND(env);
    if (!d_score) return AH_EINVAL;
    if (d_reset_table && (reset_table_len <= 0 || !aligned16(d_reset_table))) return AH_EINVAL;
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const ah::ResetSrc rs = reset_src(env, d_reset_table, reset_table_len);
    AH_DISPATCH(env, (ah::update_score_kernel<float><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<float>(env), env->n, d_score, rs, env->blk)),
                (ah::update_score_kernel<double><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<double>(env), env->n, d_score, rs, env->blk)));
    AH_LAUNCH_RC(env);
}

int ah_rewards(ah_env* env, void* d_reward, int8_t* d_score, uint8_t* d_done, void* stream) {
    AH_CHECK_BOUND(env);
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    AH_DISPATCH(env, (ah::rewards_kernel<float><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<float>(env), env->n, env->cfg.dot_fma, (float*)d_reward, d_score, d_done)),
                (ah::rewards_kernel<double><<<small_grid(env->n), 256, 0, s>>>(state_ptrs<double>(env), env->n, env->cfg.dot_fma, (double*)d_reward, d_score, d_done)));
    AH_LAUNCH_RC(env);
}

int ah_step(ah_env* env, const ah_step_io* io, int K, void* stream) {
    AH_CHECK_BOUND(env);
    if (!io || K < 1 || K > 4096) return AH_EINVAL;
    const int kind = io->action_kind;
    if (kind != AH_ACT_DISCRETE && kind != AH_ACT_CONTINUOUS && kind != AH_ACT_PHILOX && kind != AH_ACT_DISCRETE_REPEAT &&
        kind != AH_ACT_CONTINUOUS_REPEAT && kind != AH_ACT_POLICY && kind != AH_ACT_DISCRETE_PACKED4 &&
        kind != AH_ACT_POLICY_MLP) return AH_EINVAL;
    if ((kind == AH_ACT_POLICY_MLP) != (io->policy != nullptr)) return AH_EINVAL;
    if (io->actions_real_out && kind != AH_ACT_POLICY_MLP) return AH_EINVAL;
    if ((kind == AH_ACT_DISCRETE_PACKED4) != (io->events != nullptr)) return AH_EINVAL;   // the packed mode is all-or-nothing
    if (io->stats_bank != 0 && io->stats_bank != 1) return AH_EINVAL;
    if (io->env_first < 0 || io->env_count < 0 || io->env_first + io->env_count > env->n) return AH_EINVAL;
    if ((io->env_first != 0 || io->env_count != 0) && !io->events) return AH_EINVAL;     // sub-ranges: packed mode only
    if (io->events) {
        if ((reinterpret_cast<uintptr_t>(io->events) & 3u) || (reinterpret_cast<uintptr_t>(io->actions) & 3u)) return AH_EINVAL;
        // ring append from a sub-range: the slice must name its own ring counter(s) (ah_step_io.ring_counter_first / _count)
        if (io->ring && (io->env_first != 0 || (io->env_count != 0 && io->env_count != env->n)) && io->ring_counter_count <= 0) return AH_EINVAL;
        if (io->reward || io->done || io->score || io->game_over || io->obs_state || io->obs_new_state ||
            io->actions_out || io->probs_out || (io->obs_next && io->obs_next_per_frame) || env->cfg.friction) return AH_EINVAL;
    }
    if ((io->probs_out && kind != AH_ACT_POLICY && kind != AH_ACT_POLICY_MLP) || (kind == AH_ACT_POLICY && !aligned16(io->probs_out))) return AH_EINVAL;
    if (kind != AH_ACT_PHILOX && kind != AH_ACT_POLICY_MLP && !io->actions) return AH_EINVAL;
    if ((kind == AH_ACT_CONTINUOUS || kind == AH_ACT_CONTINUOUS_REPEAT) &&
        (reinterpret_cast<uintptr_t>(io->actions) & (env->dtype == AH_F32 ? 7u : 15u))) return AH_EINVAL;
    if (io->reset_table && (io->reset_table_len <= 0 || !aligned16(io->reset_table))) return AH_EINVAL;
    if (!aligned16(io->obs_state) || !aligned16(io->obs_new_state) || !aligned16(io->obs_next)) return AH_EINVAL;
    if (io->ring) {
        const ah_ring* r = io->ring;
        if (r->capacity <= 0 || !r->state || !r->new_state || !r->action || !r->reward || !r->done || !r->appended) return AH_EINVAL;
        if (reinterpret_cast<uintptr_t>(r->appended) & 7u) return AH_EINVAL;
        if (!aligned16(r->state) || !aligned16(r->new_state)) return AH_EINVAL;
        if (r->format != AH_RING_ROWS && r->format != AH_RING_COMPACT) return AH_EINVAL;
        if (io->ring_counter_count < 0 || io->ring_counter_first < 0 || io->ring_counter_first + io->ring_counter_count > AH_RING_COUNTERS) return AH_EINVAL;
        if (io->ring_counter_count > 0 && !io->events) return AH_EINVAL;                       // sliced appends: packed mode only
        if (r->format == AH_RING_COMPACT) {      // the compact layout is written by the packed FP32 kernel only
            if (env->dtype != AH_F32 || !io->events || !r->state_pos || !r->new_state_pos) return AH_EINVAL;
            if ((reinterpret_cast<uintptr_t>(r->state_pos) | reinterpret_cast<uintptr_t>(r->new_state_pos)) & 7u) return AH_EINVAL;
        }
    }
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (env->dtype == AH_F32) return launch_step<float>(env, io, K, s);
    return launch_step<double>(env, io, K, s);
}

int ah_actor_sample(ah_env* env, const float* d_weights, const void* d_obs, void* d_probs, int8_t* d_actions,
                    int greedy, void* stream) {
    if (!env || !d_weights || !d_obs || !d_actions || !aligned16(d_obs) || !aligned16(d_probs)) return AH_EINVAL;
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    AH_DISPATCH(env, (ah::actor_sample_kernel<float><<<small_grid(env->n), 256, 0, s>>>(env->n, d_weights, (const float*)d_obs, (float*)d_probs, d_actions, greedy, env->seed, env->env_id0, env->blk)),
                (ah::actor_sample_kernel<double><<<small_grid(env->n), 256, 0, s>>>(env->n, d_weights, (const double*)d_obs, (double*)d_probs, d_actions, greedy, env->seed, env->env_id0, env->blk)));
    AH_LAUNCH_RC(env);
}

int ah_policy_act(ah_env* env, const ah_policy* policy, const void* d_obs, const void* d_values_in, int8_t* d_actions,
                  void* d_actions_real, void* d_values_out, void* stream) {
    if (!env || !policy || (!d_obs && !d_values_in) || (d_obs && !aligned16(d_obs))) return AH_EINVAL;
    ah::PolicyDev pd;
    const int prc = make_policy_dev(policy, &pd, d_values_in == nullptr);
    if (prc != AH_OK) return prc;
    if (policy->head == AH_HEAD_CONTINUOUS ? !d_actions_real : !d_actions) return AH_EINVAL;
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((env->n + ah::kPolThreads - 1) / ah::kPolThreads);
    const size_t psm = d_values_in ? 0 : policy_smem_bytes(pd, ah::kPolThreads);
    AH_DISPATCH(env, (ah::policy_act_kernel<float><<<grid, ah::kPolThreads, psm, s>>>(env->n, pd, (const float*)d_obs, (const float*)d_values_in, d_actions, (float*)d_actions_real, (float*)d_values_out, env->seed, env->env_id0, env->blk)),
                (ah::policy_act_kernel<double><<<grid, ah::kPolThreads, psm, s>>>(env->n, pd, (const double*)d_obs, (const double*)d_values_in, d_actions, (double*)d_actions_real, (double*)d_values_out, env->seed, env->env_id0, env->blk)));
    AH_LAUNCH_RC(env);
}

int ah_c51_project(ah_env* env, int64_t batch, int n_actions, int n_atoms, double v_min, double v_max, double gamma,
                   const void* d_z_opt, const int8_t* d_action, const void* d_reward, const uint8_t* d_done, void* d_m_prob,
                   void* stream) {
    if (!env || batch <= 0 || n_actions < 1 || n_atoms < 2 || !(v_max > v_min) || !d_z_opt || !d_action || !d_reward || !d_done ||
        !d_m_prob) return AH_EINVAL;
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((batch + 127) / 128);
    AH_DISPATCH(env, (ah::c51_project_kernel<float><<<grid, 128, 0, s>>>(batch, n_actions, n_atoms, (float)v_min, (float)v_max, (float)gamma,
                          (const float*)d_z_opt, d_action, (const float*)d_reward, d_done, (float*)d_m_prob)),
                (ah::c51_project_kernel<double><<<grid, 128, 0, s>>>(batch, n_actions, n_atoms, v_min, v_max, gamma,
                          (const double*)d_z_opt, d_action, (const double*)d_reward, d_done, (double*)d_m_prob)));
    AH_LAUNCH_RC(env);
}

int ah_pack_host(ah_env* env, int K, const void* d_reward, const uint8_t* d_done, const void* d_obs_next,
                 uint8_t* d_events, int16_t* d_obs_pos, float* d_obs_vel, void* stream) {
    if (!env || K < 1 || !d_obs_next || !d_obs_pos || !d_obs_vel) return AH_EINVAL;
    if ((d_events != nullptr) != (d_reward != nullptr) || (d_events != nullptr) != (d_done != nullptr)) return AH_EINVAL;
    if (!aligned16(d_obs_vel) || (reinterpret_cast<uintptr_t>(d_obs_pos) & 7u)) return AH_EINVAL;
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    AH_DISPATCH(env, (ah::pack_host_kernel<float><<<small_grid(env->n), 256, 0, s>>>(env->n, K, (const float*)d_reward, d_done, (const float*)d_obs_next, d_events, (short4*)d_obs_pos, (float4*)d_obs_vel)),
                (ah::pack_host_kernel<double><<<small_grid(env->n), 256, 0, s>>>(env->n, K, (const double*)d_reward, d_done, (const double*)d_obs_next, d_events, (short4*)d_obs_pos, (float4*)d_obs_vel)));
    AH_LAUNCH_RC(env);
}

int ah_stats_bank(ah_env* env, int bank, int64_t* d_out, int clear, void* stream) {
    if (!env || !d_out || (bank != 0 && bank != 1)) return AH_EINVAL;
    DeviceGuard g(env->device);
    ah::stats_copy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(env->blk, bank, d_out, clear);
    AH_LAUNCH_RC(env);
}

int ah_stats(ah_env* env, int64_t* d_out, int clear, void* stream) { return ah_stats_bank(env, 0, d_out, clear, stream); }

int ah_peer_export(ah_env* env, uint8_t* handle64) {
    if (!env || !handle64) return AH_EINVAL;
    DeviceGuard g(env->device);
    if (!env->xchg) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&env->xchg), 2 * AH_MAX_PEERS * sizeof(ah::PeerSlot));
        if (e == cudaSuccess) e = cudaMemset(env->xchg, 0, 2 * AH_MAX_PEERS * sizeof(ah::PeerSlot));
        if (e != cudaSuccess) return cuda_fail(env, e);
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == AH_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, env->xchg);
    if (e != cudaSuccess) return cuda_fail(env, e);
    memcpy(handle64, &h, sizeof h);
    return AH_OK;
}

int ah_peer_attach(ah_env* env, int rank, int world, const uint8_t* handles) {
    if (!env || !handles || world < 1 || world > AH_MAX_PEERS || rank < 0 || rank >= world || !env->xchg) return AH_EINVAL;
    DeviceGuard g(env->device);
    for (int r = 0; r < world; ++r) {
        if (r == rank) { env->peer[r] = env->xchg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * AH_IPC_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail(env, e);
        env->peer[r] = reinterpret_cast<ah::PeerSlot*>(p);
    }
    env->rank = rank; env->world = world;
    return AH_OK;
}

int ah_stats_allreduce(ah_env* env, int bank, int64_t* d_out, int clear, void* stream) {
    if (!env || !d_out || (bank != 0 && bank != 1)) return AH_EINVAL;
    if (env->world < 1) return AH_ESTATE;                    // ah_peer_attach first
    DeviceGuard g(env->device);
    ah::PeerPtrs pp;
    for (int r = 0; r < AH_MAX_PEERS; ++r) pp.p[r] = r < env->world ? env->peer[r] : nullptr;
    ah::stats_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(env->blk, bank, clear, env->rank, env->world, pp, d_out);
    AH_LAUNCH_RC(env);
}

int ah_set_counters(ah_env* env, uint64_t frame, void* stream) {
    if (!env) return AH_EINVAL;
    DeviceGuard g(env->device);
    ah::counters_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(env->blk, frame);
    AH_LAUNCH_RC(env);
}

int ah_get_counters(ah_env* env, uint64_t* d_out1, void* stream) {
    if (!env || !d_out1) return AH_EINVAL;
    DeviceGuard g(env->device);
    ah::counters_get_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(env->blk, reinterpret_cast<unsigned long long*>(d_out1));
    AH_LAUNCH_RC(env);
}

int ah_ring_sample(ah_env* env, const ah_ring* ring, int continuous, int64_t batch, uint64_t draw, void* d_state,
                   void* d_action, void* d_reward, uint8_t* d_done, void* d_new_state, int64_t* d_index, void* stream) {
    if (!env || !ring || batch <= 0 || !d_state || !d_action || !d_reward || !d_done || !d_new_state) return AH_EINVAL;
    if (ring->capacity <= 0 || !ring->state || !ring->new_state || !ring->action || !ring->reward || !ring->done || !ring->appended) return AH_EINVAL;
    if (!aligned16(d_state) || !aligned16(d_new_state) || !aligned16(ring->state) || !aligned16(ring->new_state)) return AH_EINVAL;
    const bool compact = ring->format == AH_RING_COMPACT;
    if (ring->format != AH_RING_ROWS && !compact) return AH_EINVAL;
    if (compact && (continuous || !ring->state_pos || !ring->new_state_pos ||
                    ((reinterpret_cast<uintptr_t>(ring->state_pos) | reinterpret_cast<uintptr_t>(ring->new_state_pos)) & 7u))) return AH_EINVAL;
    const short* sp0 = compact ? reinterpret_cast<const short*>(ring->state_pos) : nullptr;
    const short* sp1 = compact ? reinterpret_cast<const short*>(ring->new_state_pos) : nullptr;
    DeviceGuard g(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((batch + 255) / 256);
    const unsigned long long* ap = reinterpret_cast<const unsigned long long*>(ring->appended);
    AH_DISPATCH(env, (ah::ring_sample_kernel<float><<<grid, 256, 0, s>>>(env->n, ring->capacity, ap, continuous, batch, env->seed, draw,
                          (const float*)ring->state, (const float*)ring->new_state, ring->action, (const float*)ring->reward, ring->done,
                          (float*)d_state, d_action, (float*)d_reward, d_done, (float*)d_new_state, d_index, sp0, sp1)),
                (ah::ring_sample_kernel<double><<<grid, 256, 0, s>>>(env->n, ring->capacity, ap, continuous, batch, env->seed, draw,
                          (const double*)ring->state, (const double*)ring->new_state, ring->action, (const double*)ring->reward, ring->done,
                          (double*)d_state, d_action, (double*)d_reward, d_done, (double*)d_new_state, d_index, sp0, sp1)));
    AH_LAUNCH_RC(env);
}

int ah_launch_info(const ah_env* env, int* blocks, int* threads, int* sms) {
    if (!env) return AH_EINVAL;
    int th = ah::kThreads;                     // same rule as launch_step
    while (th > 64 && (env->n + th - 1) / th < 4 * (int64_t)env->sms) th >>= 1;
    int64_t want = (env->n + th - 1) / th;
    const int max_blocks = env->dtype == AH_F32 ? env->blocks_f32 : env->blocks_f64;
    if (blocks) *blocks = (int)(want < max_blocks ? want : max_blocks);
    if (threads) *threads = th;
    if (sms) *sms = env->sms;
    return AH_OK;
}

/* Self-test: AH_OK when this library was compiled without multiply-add contraction (required by the FP64
 * validation build).  d_scratch: device double[1]. */
int ah_selftest_nofma(ah_env* env, double* d_scratch, void* stream) {
    if (!env || !d_scratch) return AH_EINVAL;
    DeviceGuard g(env->device);
    // a*b + c with a = 1 + 2^-30, b = 1 - 2^-30, c = -1: exact product 1 - 2^-60 rounds to 1.0, so the
    // unfused result is 0.0 while a fused multiply-add returns -2^-60.
    ah::nofma_probe_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(1.0 + 9.313225746154785e-10, 1.0 - 9.313225746154785e-10, -1.0, d_scratch);
    AH_LAUNCH_RC(env);
}

}  // extern "C"
