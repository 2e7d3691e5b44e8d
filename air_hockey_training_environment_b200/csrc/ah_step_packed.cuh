// The PACKED step kernel: the headline rollout mode (BASELINE config 3) with byte-packed per-frame I/O.
//
// Same frame as ah::step_kernel (game.py:133-177 + lib/agents/agent.py:57-78), same device functions, same
// results -- what changes is how a frame's inputs and outputs cross the register file:
//
//   actions  u8 [ceil(K/4)][n][4]   "quad-major": byte j of word (q, i) is env i's action of frame 4q + j.  One
//                                   coalesced LDG.32 per env per 4 frames (a warp reads 128 contiguous bytes)
//                                   instead of a byte load + 64-bit pointer bump + prefetch logic per frame.
//   events   u8 [ceil(K/4)][n][4]   one byte per env-frame carrying EVERYTHING the reference's frame returns:
//                                       bits 0-2  (reward + 4) / 2     reward in {-4,-2,0,2,4,6}  lib/rewards.py:118-142
//                                       bit  3    done                 "mallet overlaps puck"     rewards.py:86-89,124
//                                       bits 4-5  score + 1            -1 / 0 / +1                rewards.py:64-81
//                                       bits 6-7  game_over            0 / 1 robot / 2 opponent   game.py:101-109,169-177
//                                   assembled in a register (one PRMT per frame) and written as one STG.32 per 4
//                                   frames, instead of four stores + four pointer bumps + conversions per frame.
//   The launch's statistics (reward sum, reward^2 sum, dones) are recovered from the packed words with two IDP.4A
//   and a POPC per 4 frames instead of three accumulations per frame.
//
// Other common-path savings over step_kernel (all result-preserving; FP64 stays bit-exact):
//   * the action -> nudge map is a 256-entry table indexed by the raw action byte (no clamp), in global memory read
//     through L1 (no per-block table to fill);
//   * "the puck's velocity may be out of limits" is carried as the contact-test threshold (1225, or +inf when
//     dirty) instead of a flag tested beside it;
//   * FP32 only: the goal / wall predicates of Rewards are single compares (valid because the puck has just been
//     advanced by Puck.update, so 5 <= x <= 895), and the computer sub-update's speed clamp is min(v, 10) unless
//     the rare branch ran (the velocity was within limits before computerAI's +2 nudge).
#pragma once

namespace ah {

template <typename R>
struct PackedArgs {
    StatePtrs<R> st;
    int64_t n;
    int K;
    const uint32_t* actions;      // u8 [ceil(K/4)][n][4]
    uint32_t* events;             // u8 [ceil(K/4)][n][4]
    R* obs_next;                  // [n][8], nullable
    ResetSrc rs;
    int collect_stats, dot_fma, stats_bank;
    DevBlock* blk;
    uint32_t i0, i_end;           // envs [i0, i_end) are stepped by this launch (row strides stay n)
    // RING variant: the fused MemoryBuffer.append of every frame's Observation tuple (lib/buffer.py:24-32), slot of
    // frame k = (appended + k) mod capacity, the ring's own device counter advanced by K at the end of the launch
    R* ring_state; R* ring_new_state;     // [capacity][n][8]
    int8_t* ring_action;                  // [capacity][n]
    R* ring_reward; uint8_t* ring_done;   // [capacity][n]
    int64_t ring_capacity;
    unsigned long long* ring_appended;
    // IO_RING_COMPACT (AH_RING_COMPACT): ring_state / ring_new_state hold the velocities [capacity][n][4], the locations go
    // to int16 [capacity][n][4], the reward to int8 [capacity][n]
    short* ring_state_pos; short* ring_new_state_pos; int8_t* ring_reward8;
    int ring_counters, ticket;            // how many of the ring's counters this launch advances; its last-block ticket
    // IO_UNPACKED: the general API's tensors -- time-major int8 actions in, four per-frame outputs out
    const int8_t* actions8;               // i8 [K][n]
    R* reward; uint8_t* done; int8_t* score; uint8_t* game_over;     // [K][n] each
};

// What crosses the register file per frame (compile-time; the arithmetic is the same in all four):
//   IO_PACKED    quad-major uint8 actions in, one event byte per env-frame out              (the headline mode)
//   IO_RING      IO_PACKED + the fused replay-ring append                                    (BASELINE config 4)
//   IO_UNPACKED  int8 [K][n] actions in; reward real / done u8 / score i8 / game_over u8 [K][n] out   (BatchedAirHockey.step's
//                default tensors: the event byte is decoded into the caller's four tensors as it is produced)
//   IO_PHILOX    actions drawn in the kernel (Philox stream 1, frame-indexed), no per-frame output: pure simulation
//   IO_RING_COMPACT  IO_RING with the compact tuple layout (include/ah_b200.h: AH_RING_COMPACT): 51 instead of 70 bytes per
//                env-frame and only whole-sector stores (one STG.128 + one STG.64 per observation, each contiguous per warp)
enum { IO_PACKED = 0, IO_RING = 1, IO_UNPACKED = 2, IO_PHILOX = 3, IO_RING_COMPACT = 4 };

// Observation -> compact ring row: velocities as they are, locations int()-truncated to int16 (exact: 0 <= x < 2^15)
__device__ __forceinline__ void store_obs_compact(float* vel, short* pos, int64_t row, const Env<float>& e) {
    reinterpret_cast<float4*>(vel)[row] = make_float4(e.rdx, e.rdy, e.pdx, e.pdy);
    const uint32_t a = (uint32_t)(__float2int_rz(e.rx) & 0xffff) | ((uint32_t)__float2int_rz(e.ry) << 16);
    const uint32_t b = (uint32_t)(__float2int_rz(e.px) & 0xffff) | ((uint32_t)__float2int_rz(e.py) << 16);
    reinterpret_cast<uint2*>(pos)[row] = make_uint2(a, b);
}
__device__ __forceinline__ void store_obs_compact(double*, short*, int64_t, const Env<double>&) {}      // (AH_F32 only)

// the actions of up to 4 consecutive frames of env i, gathered from the time-major int8 tensor into one word
__device__ __forceinline__ uint32_t gather_quad_actions(const int8_t* p, uint32_t n, int count) {
    uint32_t w = (uint32_t)(uint8_t)p[0];
    if (count > 1) w |= (uint32_t)(uint8_t)p[n] << 8;
    if (count > 2) w |= (uint32_t)(uint8_t)p[2 * (size_t)n] << 16;
    if (count > 3) w |= (uint32_t)(uint8_t)p[3 * (size_t)n] << 24;
    return w;
}

// ---- Rewards.__call__ as predicates (lib/rewards.py:118-142); see rewards_call for the line-by-line citations.
// IN_FRAME: the puck was just advanced by Puck.update (5 <= px <= 895), which lets the FP32 build test the goal and
// wall zones with one compare each: |900 - px| < 30 <=> px > 870, |0 - px| < 30 <=> px < 30, |px - 13| < 15 <=> px < 28,
// |px - 882| < 15 <=> px > 867 (the subtractions are exact next to their thresholds, so the equivalences are exact).
template <typename R>
__device__ __forceinline__ void rewards_bits(Env<R>& e, bool dot_fma, int& score, bool& done, bool& wl, bool& wr, bool& closer,
                                             int& ambiguous) {
    using M = Math<R>;
    const bool in_y = M::abs_(K<R>::mid_y - e.py) < R(95);
    bool right, left;
    if (M::exact) {
        right = (M::abs_(K<R>::table_w - e.px) < R(30)) && in_y;
        left = (M::abs_(R(0) - e.px) < R(30)) && in_y;
        wl = M::abs_(e.px - K<R>::left_wall) < K<R>::puck_r;
        wr = M::abs_(e.px - K<R>::right_wall) < K<R>::puck_r;
    } else {
        right = (e.px > R(870)) && in_y;
        left = (e.px < R(30)) && in_y;
        wl = e.px < R(28);
        wr = e.px > R(867);
    }
    score = left ? -1 : (right ? 1 : 0);
    R ex, ey;
    M::sub2(ex, ey, e.px, e.py, e.rx, e.ry);
    const R dsq = M::dot2(ex, ey, ex, ey, dot_fma);
    if (M::exact) {                               // see rewards_call: pow(x, 2) vs x*x
        const R ssq = ex * ex + ey * ey;
        done = M::sqrt_(ssq) < K<R>::radius;
        ambiguous += done_is_ambiguous(ex, ey, ssq) ? 1 : 0;
    } else {
        done = dsq < K<R>::radius_sq;
    }
    const R nd = M::sqrt_(dsq);
    closer = nd < e.prev_dist;
    e.prev_dist = nd;
}

// ---- FP32 only: Rewards.__call__ + the event byte WITHOUT the ALU pipe.  The frame loop is bound by the ALU pipe
// (compares, selects, min/max: one warp-instruction per 2 cycles), not by issue slots, so the seven compares and the
// integer assembly of the event byte are re-expressed on the FMA pipes: a predicate `a > b` becomes the 0/1 value
// sat((a - b) * 2^40) -- one FFMA.SAT; IEEE subtraction gives the exact sign and an exact zero, and any non-zero
// difference of these operands times 2^40 exceeds 1 -- and the byte is a short FFMA chain on those 0/1 values, biased
// by 2^23 so that the integer sits in the low mantissa bits of the result, from where PRMT takes it.  Same results as
// rewards_bits (tests/test_gpu_packed.py compares the packed kernel with step_kernel bit for bit).
__device__ __forceinline__ float gt01(float a, float b) {            // (a > b) ? 1.0f : 0.0f
    return __saturatef(fmaf(a, 1099511627776.0f, -b * 1099511627776.0f));
}
__device__ __forceinline__ float lt01(float a, float b) {            // (a < b) ? 1.0f : 0.0f
    return __saturatef(fmaf(a, -1099511627776.0f, b * 1099511627776.0f));
}
__device__ __forceinline__ uint32_t rewards_event_fma(Env<float>& e, float& score_f) {
    const float in_y = lt01(fabsf(240.0f - e.py), 95.0f);
    score_f = in_y * (gt01(e.px, 870.0f) - lt01(e.px, 30.0f));      // +1 right goal, -1 left goal (the zones are disjoint)
    float ex, ey;
    Math<float>::sub2(ex, ey, e.px, e.py, e.rx, e.ry);
    const float dsq = fmaf(ey, ey, ex * ex);
    const float dn = lt01(dsq, 1225.0f);
    const float nd = Math<float>::sqrt_(dsq);
    const float closer = __saturatef((e.prev_dist - nd) * 1099511627776.0f);     // prev_dist may be +inf: inf -> 1
    e.prev_dist = nd;
    const float wall = gt01(e.px, 867.0f) - lt01(e.px, 28.0f);
    // 2^23 + 0x11 + closer + wall + 10 done + 16 score: exact small integers, the byte is the low 8 mantissa bits
    const float b = fmaf(score_f, 16.0f, fmaf(dn, 10.0f, (closer + wall) + 8388625.0f));
    return __float_as_uint(b);
}

// ---- FP32 only: AirHockey.computerAI (environment/airhockey.py:270-308) without compares and selects, in the same
// spirit: the branch conditions become 0/1 values on the FMA pipes and the chosen velocity is a short arithmetic blend of
// small exact integers.  Uses what Mallet.update guarantees, 470 <= opponent.x <= 880 and 20 <= opponent.y <= 460:
//   x:  puck.x < 470 implies puck.x < opponent.x, puck.x > 880 implies puck.x > opponent.x, so
//       dx = +1 (px < 470) | -1 (px > 880) | -2 / +2 / unchanged by the sign of px - ox in between          :272-283
//   y:  py < oy -> (py < 20 ? +1 : -6)  [py < 20 implies py < oy];
//       py > oy -> (py <= 360 ? +6 : (oy > 200 ? -2 : 0));  equal: unchanged                                    :285-300
//   the +2, +2 puck nudge only when py > oy and |dy| < 40 and |dx| < 40                                       :303-305
// All values are small exact integers and every blend coefficient is 0 or 1, so the arithmetic is exact.
__device__ __forceinline__ float pos01(float d) { return __saturatef(d * 1099511627776.0f); }      // (d > 0) ? 1 : 0
__device__ __forceinline__ void computer_ai_fma(Env<float>& e) {
    float gx, gy;
    Math<float>::sub2(gx, gy, e.px, e.py, e.ox, e.oy);               // puck - opponent: the AI's compares AND the nudge test
    // x: dx = sign(gx) * (2 - 3 * [px < 470 or px > 880]); gx == 0 keeps the old value
    const float sx = pos01(gx) - pos01(-gx);
    const float far = lt01(e.px, 470.0f) + gt01(e.px, 880.0f);
    e.odx = fmaf(sx, fmaf(far, -3.0f, 2.0f), (1.0f - fabsf(sx)) * e.odx);
    // y: dy = 6 sign(gy) + 7 [py < 20] - [py > oy] [py > 360] (6 + 2 [oy > 200]); gy == 0 keeps the old value
    const float y_gt = pos01(gy);
    const float sy = y_gt - pos01(-gy);
    const float up = fmaf(lt01(e.py, 20.0f), 7.0f, sy * 6.0f);
    const float cap = (y_gt * gt01(e.py, 360.0f)) * fmaf(gt01(e.oy, 200.0f), 2.0f, 6.0f);
    e.ody = fmaf(1.0f - fabsf(sy), e.ody, up - cap);
    const float inc = (y_gt + y_gt) * (lt01(fabsf(gy), 40.0f) * lt01(fabsf(gx), 40.0f));      // 2 or 0      :303-305
    Math<float>::add2(e.pdx, e.pdy, inc, inc);
    mallet_update<float>(e.ox, e.oy, e.odx, e.ody, K<float>::opp_lo, K<float>::opp_hi);      // :307
}

// the common tail of a sub-update with the speed-limit bookkeeping carried in `thr` (see file header)
template <typename R, bool COMPUTER>
__device__ __forceinline__ void physics_tail_thr(Env<R>& e, bool dot_fma, R& thr, int& contacts) {
    R rdx_, rdy_, odx_, ody_;
    const R dr = contact_dist_sq<R>(e.px, e.py, e.rx, e.ry, dot_fma, rdx_, rdy_);
    R dq = contact_dist_sq<R>(e.px, e.py, e.ox, e.oy, dot_fma, odx_, ody_);
    if (!(dr > thr) || !(dq > thr)) {                                            // :210, or velocity possibly out of limits
        R dr_in = dr;
        opaque(dr_in); opaque(dq);
        if (!(dr_in > K<R>::radius_sq)) {
            contacts += 1;
            contact_response<R>(e.px, e.py, e.pdx, e.pdy, e.rx, e.ry, e.rdx, e.rdy, rdx_, rdy_, dr_in, dot_fma);
            dq = contact_dist_sq<R>(e.px, e.py, e.ox, e.oy, dot_fma, odx_, ody_);
        }
        if (!(dq > K<R>::radius_sq)) {
            contacts += 0x10000;
            contact_response<R>(e.px, e.py, e.pdx, e.pdy, e.ox, e.oy, e.odx, e.ody, odx_, ody_, dq, dot_fma);
        }
        limit_puck_speed<R>(e.pdx, e.pdy);                                       // :117
        thr = K<R>::radius_sq;
    } else if (COMPUTER) {
        if (Math<R>::exact) limit_puck_speed<R>(e.pdx, e.pdy);
    }
    if (COMPUTER && !Math<R>::exact) {
        // within limits before computerAI's nudge (+2 or +0 on both axes) => only the upper limit can bind; after the
        // rare branch the velocity is already limited and this is the identity
        e.pdx = fminf((float)e.pdx, 10.0f);
        e.pdy = fminf((float)e.pdy, 10.0f);
    }
    puck_update<R>(e.px, e.py, e.pdx, e.pdy);                                    // :118
    mallet_update<R>(e.rx, e.ry, e.rdx, e.rdy, K<R>::robot_lo, K<R>::robot_hi);  // :121-122
    mallet_update<R>(e.ox, e.oy, e.odx, e.ody, K<R>::opp_lo, K<R>::opp_hi);
}

// raw action byte -> position nudge (airhockey.py:74-84): 0: y += 10, 1: y -= 10, 2: x -= 10, 3: x += 10, any other
// value: no nudge.  A 256-entry table in GLOBAL memory read through L1 (ld.global.nc): no per-block table to fill and no
// index clamp; 2 KB / 4 KB that stay L1-resident.
__device__ const float2 g_nudge_f32[256] = {{0.f, 10.f}, {0.f, -10.f}, {-10.f, 0.f}, {10.f, 0.f}};
__device__ const double2 g_nudge_f64[256] = {{0., 10.}, {0., -10.}, {-10., 0.}, {10., 0.}};
__device__ __forceinline__ float2 nudge_of(const float2* t, uint32_t byte) { return __ldg(t + byte); }
__device__ __forceinline__ double2 nudge_of(const double2* t, uint32_t byte) { return __ldg(t + byte); }

// Launch geometry (measured on B200, n = 2^20, K = 8, FP32; gpurun_out r2c-r2e): blocks of 128
// threads, 9 per SM (56 registers; the kernel needs 51) beat 256 x 4 by 1.5-2 us per launch -- a block's registers are
// held until its slowest warp is done, and the warps of a block finish at different times (contact / goal branches);
// 10 or 12 blocks per SM (48 / 40 registers, spills outside the loop) are slower: the loop is pipe-bound, not
// latency-bound.
#ifndef AH_PACKED_THREADS
#define AH_PACKED_THREADS 128
#endif
#ifndef AH_PACKED_OCC_F32
#define AH_PACKED_OCC_F32 9
#endif
#ifndef AH_FRAME_UNROLL
#define AH_FRAME_UNROLL 1
#endif
constexpr int kFrameUnroll = AH_FRAME_UNROLL;        // frames of a quad unrolled together (tuning: section 4.5 of DESIGN.md)
constexpr int kPackedThreads = AH_PACKED_THREADS;
#ifndef AH_PACKED_OCC_RING
#define AH_PACKED_OCC_RING 6
#endif
// RING: the append variant is bound by the write path (70 B per env-frame) and carries the ring cursor.  Measured (2^20 envs,
// K = 8): 6 (66 registers, 7 resident blocks) 140.7 us, 8 (60 registers) 154.5 us, 9 (56, spills) 191 us -- more resident
// warps put more half-row stores in flight than L1 can merge before they leave for L2.
// (IO_UNPACKED / IO_PHILOX need 58-60 registers: 8 blocks per SM instead of spilling at 9)
#ifndef AH_PACKED_OCC_RINGC
#define AH_PACKED_OCC_RINGC 8
#endif
template <typename R, int IO> struct PackedOcc { static constexpr int blocks = IO == 1 ? AH_PACKED_OCC_RING : (IO == 4 ? AH_PACKED_OCC_RINGC : (IO >= 2 ? 8 : AH_PACKED_OCC_F32)); };
template <int IO> struct PackedOcc<double, IO> { static constexpr int blocks = 512 / AH_PACKED_THREADS; };

// AirHockey.reset with the env's NEXT Philox draw precomputed: the draw for reset index `resets` depends only on
// (seed, env id, resets), so every thread computes it once at the top of the launch -- all 32 lanes together, in the
// shadow of the state loads -- instead of one lane of the warp running the 80-instruction generator inside the goal
// branch (23 % of warp-frames).  A second reset of the same env in the launch (a game-over frame, or two goals within
// K frames) falls back to the in-branch draw.  Same values either way: bitwise-identical results.
template <typename R>
__device__ __forceinline__ void env_reset_cached(Env<R>& e, bool total, const ResetSrc& rs, int64_t i, unsigned& overflow,
                                                 R& next_dx, R& next_dy, bool& next_valid) {
    if (rs.table != nullptr || !next_valid) { env_reset<R>(e, total, rs, i, overflow); return; }
    if (total) { e.opp_score = 0; e.robot_score = 0; }
    next_valid = false;
    e.resets += 1;
    e.px = K<R>::mid_x; e.py = K<R>::mid_y; e.pdx = next_dx; e.pdy = next_dy;
    e.rdx = R(0); e.rdy = R(0); e.odx = R(0); e.ody = R(0);
}

template <typename R, int IO>
__global__ void __launch_bounds__(kPackedThreads, PackedOcc<R, IO>::blocks)
step_packed_kernel(const PackedArgs<R> a) {
    constexpr bool RING = IO == IO_RING || IO == IO_RING_COMPACT;
    constexpr bool COMPACT = IO == IO_RING_COMPACT;
    constexpr bool ORDERED = RING || IO == IO_PHILOX;      // the launch reads a device counter at its start
    using V2 = typename std::conditional<sizeof(R) == 4, float2, double2>::type;
    __shared__ BlockStats bs;
    if (threadIdx.x < AH_NSTATS) bs.v[threadIdx.x] = 0;
    if (threadIdx.x == 0) { bs.len_sq = 0ull; bs.warps_done = 0; }
    __syncthreads();
    const V2* nudge_tab = sizeof(R) == 4 ? reinterpret_cast<const V2*>(g_nudge_f32) : reinterpret_cast<const V2*>(g_nudge_f64);

    const uint32_t n = (uint32_t)a.n;
    const uint32_t stride = gridDim.x * blockDim.x;
    const bool dot_fma = a.dot_fma != 0;
    const int nK = a.K;
    int r3sum = 0, r3sq = 0, dones = 0, contacts = 0, frames = 0, bad = 0, amb = 0;
    unsigned overflow = 0;
    // RING: every block reads the ring's counter at its start, so it may only advance once every block is done
    // (finish_stats<true>: fence + ticket); a slice of a step reads and advances its own counter (ring_counter_first)
    int64_t slot0 = 0;
    if constexpr (RING) slot0 = (int64_t)(*a.ring_appended % (unsigned long long)a.ring_capacity);
    unsigned long long frame0 = 0ull;
    if constexpr (IO == IO_PHILOX) frame0 = a.blk->frame;

    for (uint32_t i = a.i0 + blockIdx.x * blockDim.x + threadIdx.x; i < a.i_end; i += stride) {
        Env<R> e = load_env<R>(a.st, i);
        uint32_t aw = 0;
        const int8_t* ap = nullptr;                    // IO_UNPACKED: &actions8[first frame of the quad][i]
        if constexpr (IO == IO_PACKED || RING) aw = a.actions[i];
        if constexpr (IO == IO_UNPACKED) { ap = a.actions8 + i; aw = gather_quad_actions(ap, n, nK < 4 ? nK : 4); }
        Philox4 ablock = {};
        uint32_t ablock_id = 0xffffffffu;
        int64_t row = (int64_t)i;                      // IO_UNPACKED: element (frame, i) of the [K][n] output tensors
        int go = (e.robot_score == 10) ? 1 : ((e.opp_score == 10) ? 2 : 0);   // only reachable through injected state
        // contact-test threshold: radius^2, or +inf while the velocity may be outside the speed limits (injected
        // state, table resets): the rare branch then runs once and clamps (limit_puck_speed is idempotent)
        R thr = speed_out_of_limits<R>(e.pdx, e.pdy) ? (R)INFINITY : K<R>::radius_sq;
        R next_dx = R(0), next_dy = R(0);
#ifdef AH_RESET_CACHE                                      /* measured: no gain (section 4.5 of DESIGN.md); off by default */
        bool next_valid = a.rs.table == nullptr;
#else
        bool next_valid = false;
#endif
        if (next_valid) reset_velocity(a.rs.key, a.rs.env_id0 + (uint64_t)i, e.resets, next_dx, next_dy);
        int gf_base = e.game_frames;        // frames-in-game = gf_base + frames done in this launch
        // The current game's return is kept as "sum of r3 since r3_mark": reward = 2 r3 - 4, so
        // return = 2 (r3sum - r3_mark) - 4 frames_in_game.  One accumulator (r3sum) serves the launch statistics too.
        int r3_mark = r3sum - ((e.game_return + 4 * e.game_frames) >> 1);
        uint32_t off = i;                   // word (q, i) of the quad-major action / event arrays: q * n + i
        int left = nK;                      // frames left at the start of the quad
        int64_t slot = slot0, rrow = slot0 * (int64_t)n + i;      // RING: row (slot, i) of the [capacity][n] ring arrays
#pragma unroll 1
        do {
            const int kc = left < 4 ? left : 4;
            uint32_t aw_next = 0;
            if constexpr (IO == IO_PACKED || RING) { if (left > 4) aw_next = a.actions[off + n]; }
            if constexpr (IO == IO_UNPACKED) {
                if (left > 4) { ap += 4 * (size_t)n; aw_next = gather_quad_actions(ap, n, left - 4 < 4 ? left - 4 : 4); }
            }
            uint32_t ev = 0;
            int jr = kc;                    // frames left in the quad, counting the current one
#pragma unroll (kFrameUnroll)
            do {
                if constexpr (IO == IO_PHILOX) {       // the frame's action from the env's action stream (csrc/ah_rng.cuh)
                    const unsigned long long f = frame0 + (unsigned long long)(nK - left + kc - jr);
                    const uint32_t bid = (uint32_t)(f >> 6);
                    if (bid != ablock_id) { ablock = action_block(a.rs.seed, a.rs.env_id0 + (uint64_t)i, bid); ablock_id = bid; }
                    aw = (uint32_t)action_from_block(ablock, (uint32_t)(f & 63u));
                }
                V2 nd;
                if constexpr (RING) {
                    // no table load here: with 70 B of stores per frame in flight, a load that the frame's first instruction
                    // depends on queues behind them in the L1 pipeline (ncu: long_scoreboard 6.5 warps per issue slot)
                    const uint32_t ab = aw & 0xffu;
                    nd.y = ab == 0u ? K<R>::step : (ab == 1u ? -K<R>::step : R(0));
                    nd.x = ab == 3u ? K<R>::step : (ab == 2u ? -K<R>::step : R(0));
                    a.ring_action[rrow] = (int8_t)ab;
                } else {
                    nd = nudge_of(nudge_tab, aw & 0xffu);
                }
                aw >>= 8;
                // ---- robot.move(a): sub-update #1                      game.py:136 -> agent.py:32-36
                move_robot<R>(e, false, nd.x, nd.y);
                physics_tail_thr<R, false>(e, dot_fma, thr, contacts);
                if constexpr (COMPACT) store_obs_compact(a.ring_state, a.ring_state_pos, rrow, e);
                else if constexpr (RING) store_obs<R>(a.ring_state, rrow, get_state<R>(e, true));  // Observation.state      agent.py:61
                // ---- robot.step(a) -> self.move(a): sub-update #2      agent.py:64
                move_robot<R>(e, false, nd.x, nd.y);
                physics_tail_thr<R, false>(e, dot_fma, thr, contacts);
                if constexpr (COMPACT) store_obs_compact(a.ring_new_state, a.ring_new_state_pos, rrow, e);
                else if constexpr (RING) store_obs<R>(a.ring_new_state, rrow, get_state<R>(e, true));   // Observation.new_state  agent.py:67
                // ---- Rewards                                           agent.py:70-71
                // event byte: (reward + 4) / 2 = 1 + closer + wall + 2 done, done << 3, (score + 1) << 4, game_over << 6
                uint32_t byte; bool scored; int sc = 0;
                if constexpr (Math<R>::exact) {
                    bool dn, wl, wr, closer;
                    rewards_bits<R>(e, dot_fma, sc, dn, wl, wr, closer, amb);
                    byte = (closer ? 0x12u : 0x11u) + (wr ? 1u : (wl ? 0xffffffffu : 0u)) + (dn ? 10u : 0u) + (uint32_t)(sc * 16);
                    scored = sc != 0;
                } else {
                    float score_f;
                    byte = rewards_event_fma(e, score_f);     // (only byte 0 is meaningful: PRMT takes exactly that)
                    scored = score_f != 0.0f;
                    if (scored) sc = score_f > 0.0f ? 1 : -1;
                }
                if constexpr (RING) {        // the rest of the Observation tuple: reward = 2 (byte & 7) - 4, done = bit 3
                    if constexpr (COMPACT) a.ring_reward8[rrow] = (int8_t)(2 * (int)(byte & 7u) - 4);
                    else a.ring_reward[rrow] = (R)(2 * (int)(byte & 7u) - 4);
                    a.ring_done[rrow] = (uint8_t)((byte >> 3) & 1u);
                    slot += 1; rrow += (int64_t)n;
                    if (slot == a.ring_capacity) { slot = 0; rrow = i; }
                }
                // ---- env.update_score(score)                           game.py:142 -> airhockey.py:149-169
                if (scored) {
                    if (sc > 0) { e.robot_score += 1; stat_add(bs, AH_STAT_ROBOT_GOALS, 1); }
                    else { e.opp_score += 1; stat_add(bs, AH_STAT_OPPONENT_GOALS, 1); }
                    env_reset_cached<R>(e, false, a.rs, i, overflow, next_dx, next_dy, next_valid);
                    if (a.rs.table != nullptr) thr = (R)INFINITY;         // table values may exceed the limits
                    go = (e.robot_score == 10) ? 1 : go;                  // stats(): game.py:101-109
                    go = (e.opp_score == 10) ? 2 : go;
                }
                if constexpr (IO == IO_UNPACKED) {     // the event byte decoded into the caller's four tensors
                    a.reward[row] = (R)(2 * (int)(byte & 7u) - 4);
                    a.done[row] = (uint8_t)((byte >> 3) & 1u);
                    a.score[row] = (int8_t)((int)((byte >> 4) & 3u) - 1);
                    a.game_over[row] = (uint8_t)go;
                    row += (int64_t)n;
                }
                byte += (uint32_t)go << 6;
                ev = __byte_perm(ev, byte, 0x4321);                       // frames land in byte order
                // ---- env.update_state(agent_name="computer"): sub-update #3   game.py:166
                if constexpr (Math<R>::exact) computer_ai<R>(e);
                else computer_ai_fma(e);
                physics_tail_thr<R, true>(e, dot_fma, thr, contacts);
                // ---- 10-goal game over                                 game.py:169-177
                if (go != 0) {
                    int jr_o = jr, left_o = left;                          // (opaque: no induction variables for the rare path)
                    asm volatile("" : "+r"(jr_o), "+r"(left_o));
                    const int jdone = (left_o < 4 ? left_o : 4) - jr_o + 1;   // frames of this quad done, with this one
                    const int kk = nK - left_o + jdone;                    // frames done in this launch
                    stat_add(bs, AH_STAT_GAMES, 1);
                    stat_add(bs, go == 1 ? AH_STAT_ROBOT_WINS : AH_STAT_OPPONENT_WINS, 1);
                    const int len = gf_base + kk;
                    stat_add(bs, AH_STAT_GAME_LEN_SUM, len);
                    atomicAdd(&bs.len_sq, (unsigned long long)len * (unsigned long long)len);
                    // r3 of the finished game = whole quads so far + this quad's frames so far (unfilled bytes are 0)
                    const int part = (int)__dp4a(ev & 0x07070707u, 0x01010101u, 0u);
                    stat_add(bs, AH_STAT_GAME_RETURN_SUM, 2 * (r3sum + part - r3_mark) - 4 * len);
                    r3_mark = r3sum + part;
                    gf_base = -kk;
                    env_reset_cached<R>(e, true, a.rs, i, overflow, next_dx, next_dy, next_valid);
                    if (a.rs.table != nullptr) thr = (R)INFINITY;
                    go = 0;
                }
            } while (--jr > 0);
            if (left < 4) ev >>= 8 * (4 - left);                           // partial last quad: frame 0 in byte 0
            if constexpr (IO == IO_PACKED || RING) a.events[off] = ev;
            // statistics of the quad from the packed word
            const uint32_t r3 = ev & 0x07070707u;
            r3sum = (int)__dp4a(r3, 0x01010101u, (uint32_t)r3sum);
            r3sq = (int)__dp4a(r3, r3, (uint32_t)r3sq);
            dones += __popc(ev & 0x08080808u);
            off += n;
            left -= 4;
            aw = aw_next;
        } while (left > 0);
        if (a.obs_next) store_obs<R>(a.obs_next, i, get_state<R>(e, true));   // the next frame's policy input (agent.py:46)
        frames += nK;
        e.game_frames = gf_base + nK;
        e.game_return = 2 * (r3sum - r3_mark) - 4 * e.game_frames;
        bad += isfinite((e.px + e.py) + (e.pdx + e.pdy)) ? 0 : 1;           // any NaN / inf component poisons the sum
        store_env<R>(a.st, i, e);
    }

    // reward = 2 r3 - 4  =>  sum = 2 S1 - 4 F,  sum of squares = 4 S2 - 16 S1 + 16 F
    const int reward_sum = 2 * r3sum - 4 * frames;
    const int reward_sq = 4 * r3sq - 16 * r3sum + 16 * frames;
    finish_stats<ORDERED>(bs, a.blk, a.stats_bank, a.collect_stats, frames, reward_sum, reward_sq, dones, contacts & 0xffff, contacts >> 16,
                          (int)overflow, bad, amb, a.i0 == 0 ? nK : 0, RING ? a.ring_appended : nullptr, a.ring_counters, RING ? a.ticket : 0, nK);
}

}  // namespace ah
