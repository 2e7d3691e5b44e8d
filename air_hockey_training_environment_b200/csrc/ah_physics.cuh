// Device-side rules of one air-hockey game, one environment per thread, state in registers.
//
// Two arithmetic policies share this source:
//   Math<double>  VALIDATION build: every operation is the reference's IEEE binary64 operation in the
//                 reference's order.  The translation unit is compiled with -fmad=false, so the only fused
//                 multiply-adds are the explicit ones that reproduce numpy/OpenBLAS' 2-vector dot
//                 (np.dot([a,b],[c,d]) == fma(b, d, fl(a*c))), switchable at run time.
//   Math<float>   THROUGHPUT build: same rules and branch structure in binary32; min/max for clamps,
//                 MUFU rsqrt/sqrt on the (rare) contact path and for the reward distance.
//
// Each function cites the reference file:line it implements (paths relative to the reference root).
#pragma once
#include <stdint.h>
#include <math.h>

namespace ah {

// ---- constants: environment/config.py:4-13, table.py:8-15, mallet.py:33-44, airhockey.py:35-49,63 ----
template <typename R> struct K {
    static constexpr R table_w = R(900), table_h = R(480);
    static constexpr R puck_r = R(15), mallet_r = R(20);
    static constexpr R radius = R(35);            // mallet.radius + puck.radius   airhockey.py:206
    static constexpr R radius_sq = R(1225);       //                               airhockey.py:207
    static constexpr R puck_imass = R(10);        // 1.0 / 0.1 == 10.0 in binary64 puck.py:36
    static constexpr R mallet_imass = R(2);       // 1.0 / 0.5                     mallet.py:27
    static constexpr R imass_sum = R(12);         //                               airhockey.py:223
    static constexpr R speed = R(10);             // config.puck["default_speed"]
    static constexpr R step = R(10);              // airhockey.py:63
    static constexpr R left_wall = R(13), right_wall = R(882);           // table.py:14-15
    static constexpr R mid_x = R(450), mid_y = R(240);                   // table.py:11
    static constexpr R robot_lo = R(20), robot_hi = R(430);              // mallet.py:36-38
    static constexpr R opp_lo = R(470), opp_hi = R(880);                 // mallet.py:39-41
    static constexpr R y_lo = R(20), y_hi = R(460);                      // mallet.py:33-34
    static constexpr R puck_x_lo = R(15), puck_x_hi = R(885);            // puck.py:49-54
    static constexpr R puck_y_lo = R(15), puck_y_hi = R(465);            // puck.py:56-61
};

template <typename R> struct Math;

template <> struct Math<double> {
    static constexpr bool exact = true;
    // np.dot on two 2-vectors (airhockey.py:203,246; rewards.py:94 through np.linalg.norm)
    static __device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1, bool dot_fma) {
        const double p = __dmul_rn(a0, b0);
        return dot_fma ? __fma_rn(a1, b1, p) : __dadd_rn(p, __dmul_rn(a1, b1));
    }
    static __device__ __forceinline__ double sqrt_(double v) { return __dsqrt_rn(v); }
    static __device__ __forceinline__ double trunc_(double v) { return trunc(v); }
    // `if x < lo: x = lo elif x > hi: x = hi`   (mallet.py:65-74)
    static __device__ __forceinline__ double clamp(double x, double lo, double hi) {
        return (x < lo) ? lo : ((x > hi) ? hi : x);
    }
    static __device__ __forceinline__ double abs_(double v) { return fabs(v); }
    // (x, y) +/- (bx, by): plain IEEE adds
    static __device__ __forceinline__ void add2(double& x, double& y, double bx, double by) { x = x + bx; y = y + by; }
    static __device__ __forceinline__ void sub2(double& ox, double& oy, double ax, double ay, double bx, double by) {
        ox = ax - bx; oy = ay - by;
    }
};

template <> struct Math<float> {
    static constexpr bool exact = false;
    static __device__ __forceinline__ float dot2(float a0, float a1, float b0, float b1, bool) {
        return fmaf(a1, b1, a0 * b0);
    }
    static __device__ __forceinline__ float sqrt_(float v) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
        return r;
    }
    static __device__ __forceinline__ float rsqrt_(float v) {
        float r;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
        return r;
    }
    static __device__ __forceinline__ float trunc_(float v) { return truncf(v); }
    static __device__ __forceinline__ float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
    static __device__ __forceinline__ float abs_(float v) { return fabsf(v); }
    // Blackwell packed FP32: one FADD2 / FFMA2 (sm_100+ `add.f32x2` / `fma.rn.f32x2`) advances both coordinates of a
    // body -- every position/velocity here is an (x, y) pair, so the pairing is natural and halves the issue slots
    // of the additive part of the update.  a - b is fma(b, -1, a): one rounding, identical to a plain subtraction.
    static __device__ __forceinline__ void add2(float& x, float& y, float bx, float by) {
        const float2 r = __fadd2_rn(make_float2(x, y), make_float2(bx, by));
        x = r.x; y = r.y;
    }
    static __device__ __forceinline__ void sub2(float& ox, float& oy, float ax, float ay, float bx, float by) {
        const float2 r = __ffma2_rn(make_float2(bx, by), make_float2(-1.0f, -1.0f), make_float2(ax, ay));
        ox = r.x; oy = r.y;
    }
};

// All state of one game, in registers.
template <typename R> struct Env {
    R px, py, pdx, pdy;      // puck       environment/puck.py:16-43
    R rx, ry, rdx, rdy;      // robot      environment/mallet.py:11-48 (name "robot": left half)
    R ox, oy, odx, ody;      // opponent   (name "opponent": right half)
    R prev_dist;             // Rewards.prev_mallet_to_puck_distance   lib/rewards.py:31
    int robot_score, opp_score;                                     // airhockey.py:52
    uint32_t resets;         // reset draws consumed
    int game_frames, game_return;
};

// ---- Mallet.update  environment/mallet.py:54-76 ----
template <typename R>
__device__ __forceinline__ void mallet_update(R& x, R& y, R dx, R dy, R lo, R hi) {
    Math<R>::add2(x, y, dx, dy);
    x = Math<R>::clamp(x, lo, hi);
    y = Math<R>::clamp(y, K<R>::y_lo, K<R>::y_hi);
}

// ---- AirHockey._move  environment/airhockey.py:65-94 ----
// The nudge never enters the velocity: after the first update() the velocity is recomputed as the
// displacement actually made (== the old velocity unless a limit shortened it).
// Discrete action -> (nx, ny) nudge  (:74-84; any other int: no nudge)
template <typename R>
__device__ __forceinline__ void action_nudge(int ia, R& nx, R& ny) {
    ny = (ia == 0) ? K<R>::step : ((ia == 1) ? -K<R>::step : R(0));
    nx = (ia == 3) ? K<R>::step : ((ia == 2) ? -K<R>::step : R(0));
}

// continuous: (ax, ay) is added to the velocity (:69-71); discrete: (ax, ay) is the position nudge
template <typename R>
__device__ __forceinline__ void move_robot(Env<R>& e, bool continuous, R ax, R ay) {
    if (continuous) Math<R>::add2(e.rdx, e.rdy, ax, ay);
    else Math<R>::add2(e.rx, e.ry, ax, ay);
    const R lx = e.rx, ly = e.ry;                  // mallet.py:57-58 (last_x, last_y)
    mallet_update<R>(e.rx, e.ry, e.rdx, e.rdy, K<R>::robot_lo, K<R>::robot_hi);   // :87
    Math<R>::sub2(e.rdx, e.rdy, e.rx, e.ry, lx, ly);                              // :90-91
    mallet_update<R>(e.rx, e.ry, e.rdx, e.rdy, K<R>::robot_lo, K<R>::robot_hi);   // :92
}

// ---- AirHockey.collision  environment/airhockey.py:189-268 (correction=True) ----
// Split in two: the distance test that every sub-update pays (5 flops), and the contact response that runs
// for ~0.5 % of the tests.  `d = mallet - puck` and `dist_sq = np.dot(d, d)` are handed over from the test.
template <typename R>
__device__ __forceinline__ R contact_dist_sq(R px, R py, R mx, R my, bool dot_fma, R& d_x, R& d_y) {
    Math<R>::sub2(d_x, d_y, mx, my, px, py);                       // :198-199
    return Math<R>::dot2(d_x, d_y, d_x, d_y, dot_fma);             // :203
}

template <typename R>
__device__ __forceinline__ void contact_response(R& px, R& py, R& pdx, R& pdy, R& mx, R& my, R& mdx, R& mdy,
                                                 R d_x, R d_y, R dist_sq, bool dot_fma) {
    using M = Math<R>;
    if (M::exact) {
        R n_x, n_y;
        const R distance = M::sqrt_(dist_sq);                      // :214
        if (distance > R(0)) { n_x = d_x / distance; n_y = d_y / distance; }   // :217
        else { n_x = d_x; n_y = d_y; }
        // :220 `dcoll = radius - d` is a 2-VECTOR in the reference and :230 `np.max(dcoll - slop, 0)` is the
        // axis-0 maximum of its two components (not a clamp at zero).  Reproduced as written.
        const R a = (K<R>::radius - d_x) - R(0.01);
        const R b = (K<R>::radius - d_y) - R(0.01);
        const R m = (a >= b) ? a : b;
        const R s = (m / K<R>::imass_sum) * R(0.2);                // :230 ((max / imass_sum) * percent)
        const R sep_x = s * n_x, sep_y = s * n_y;
        px -= sep_x * K<R>::puck_imass;                            // :235
        py -= sep_y * K<R>::puck_imass;                            // :236
        mx += sep_x * K<R>::mallet_imass;                          // :237
        my += sep_y * K<R>::mallet_imass;                          // :238
        const R v_x = mdx - pdx;                                   // :241
        const R v_y = mdy - pdy;                                   // :242
        const R vn = M::dot2(v_x, v_y, n_x, n_y, dot_fma);         // :246
        if (vn > R(0)) return;                                     // :249 moving apart
        const R j = -(R(1.0) + R(0.9)) * (vn / K<R>::imass_sum);   // :256
        const R i_x = j * n_x, i_y = j * n_y;                      // :259
        pdx -= i_x * K<R>::puck_imass;                             // :262
        pdy -= i_y * K<R>::puck_imass;                             // :263
        mdx += i_x * K<R>::mallet_imass;                           // :265
        mdy += i_y * K<R>::mallet_imass;                           // :266
    } else {
        // FP32 build: the same response with MUFU.RSQ for the normal, the constant factors folded
        // ((m / 12) * 0.2 * 10 etc.), and both coordinates of every update in one packed FFMA2 / FMUL2
        const float inv = (float)dist_sq > 0.0f ? Math<float>::rsqrt_((float)dist_sq) : 1.0f;
        const float2 n = __fmul2_rn(make_float2((float)d_x, (float)d_y), make_float2(inv, inv));
        const float2 ab = __fadd2_rn(__ffma2_rn(make_float2((float)d_x, (float)d_y), make_float2(-1.0f, -1.0f), make_float2(35.0f, 35.0f)),
                                     make_float2(-0.01f, -0.01f));       // (35 - d) - 0.01, per component
        const float m = fmaxf(ab.x, ab.y);
        const float sp = m * (float)(-10.0 * 0.2 / 12.0), sm = m * (float)(2.0 * 0.2 / 12.0);
        const float2 p = __ffma2_rn(make_float2(sp, sp), n, make_float2((float)px, (float)py));
        const float2 q = __ffma2_rn(make_float2(sm, sm), n, make_float2((float)mx, (float)my));
        px = (R)p.x; py = (R)p.y; mx = (R)q.x; my = (R)q.y;
        float v_x, v_y;
        Math<float>::sub2(v_x, v_y, (float)mdx, (float)mdy, (float)pdx, (float)pdy);
        const float vn = fmaf(v_y, n.y, v_x * n.x);
        if (vn > 0.0f) return;
        const float jp = vn * (float)(10.0 * 1.9 / 12.0), jm = vn * (float)(-2.0 * 1.9 / 12.0);
        const float2 pv = __ffma2_rn(make_float2(jp, jp), n, make_float2((float)pdx, (float)pdy));
        const float2 mv = __ffma2_rn(make_float2(jm, jm), n, make_float2((float)mdx, (float)mdy));
        pdx = (R)pv.x; pdy = (R)pv.y; mdx = (R)mv.x; mdy = (R)mv.y;
    }
}

// ---- Puck.friction_on_puck  environment/puck.py:73-88 (dead code in the reference; opt-in flag) ----
template <typename R>
__device__ __forceinline__ void friction_on_puck(R& pdx, R& pdy) {
    if (pdx > R(1)) pdx -= R(0.5);
    if (pdx < R(-1)) pdx += R(0.5);
    if (pdy > R(1)) pdy -= R(0.5);
    else if (pdy < R(-1)) pdy += R(0.5);
}

// ---- Puck.limit_puck_speed  environment/puck.py:90-109 ----
template <typename R>
__device__ __forceinline__ void limit_puck_speed(R& pdx, R& pdy) {
    if (Math<R>::exact) {
        pdx = (pdx > K<R>::speed) ? K<R>::speed : pdx;             // :94-95
        pdx = (pdx < -K<R>::speed) ? -K<R>::speed : pdx;           // :96-97
        pdy = (pdy > K<R>::speed) ? K<R>::speed : pdy;             // :100-101
        pdy = (pdy < -K<R>::speed) ? K<R>::speed : pdy;            // :102-103: -> +speed, as in the reference
    } else {
        pdx = Math<R>::clamp(pdx, -K<R>::speed, K<R>::speed);
        pdy = (pdy < -K<R>::speed) ? K<R>::speed : fminf(pdy, K<R>::speed);
    }
}

// ---- Puck.update  environment/puck.py:45-71: clamp + reflect FIRST, then advance ----
template <typename R>
__device__ __forceinline__ void puck_update(R& px, R& py, R& pdx, R& pdy) {
    if (Math<R>::exact) {
        const bool in_x = (px > K<R>::puck_x_lo) && (px < K<R>::puck_x_hi);     // not (x <= 15 or x >= 885)
        px = Math<R>::clamp(px, K<R>::puck_x_lo, K<R>::puck_x_hi);
        pdx = in_x ? pdx : -pdx;                                       // `dx *= -1`
        const bool in_y = (py > K<R>::puck_y_lo) && (py < K<R>::puck_y_hi);
        py = Math<R>::clamp(py, K<R>::puck_y_lo, K<R>::puck_y_hi);
        pdy = in_y ? pdy : -pdy;
    } else {
        // FP32 build: the same predicate without compares.  t = (lo - x) * (x - hi) is positive strictly inside
        // the table and has its SIGN BIT set when x <= lo or x >= hi (at x == lo or x == hi one factor is +0 and the
        // other negative, so t = -0).  IEEE subtraction and multiplication give exact signs, so
        // `reflect == signbit(t)` is the reference's `x <= lo or x >= hi`; the velocity flip is one LOP3
        // (v ^ (t & 0x80000000)) on the ALU pipe instead of two FSETP and an FSEL, and the products pair into
        // FADD2/FMUL2 on the FMA pipes.
        float ax, ay, bx, by;
        Math<float>::sub2(ax, ay, (float)K<R>::puck_x_lo, (float)K<R>::puck_y_lo, (float)px, (float)py);   // lo - x
        Math<float>::sub2(bx, by, (float)px, (float)py, (float)K<R>::puck_x_hi, (float)K<R>::puck_y_hi);   // x - hi
        const float2 t = __fmul2_rn(make_float2(ax, ay), make_float2(bx, by));
        pdx = (R)__uint_as_float(__float_as_uint((float)pdx) ^ (__float_as_uint(t.x) & 0x80000000u));
        pdy = (R)__uint_as_float(__float_as_uint((float)pdy) ^ (__float_as_uint(t.y) & 0x80000000u));
        px = Math<R>::clamp(px, K<R>::puck_x_lo, K<R>::puck_x_hi);
        py = Math<R>::clamp(py, K<R>::puck_y_lo, K<R>::puck_y_hi);
    }
    Math<R>::add2(px, py, pdx, pdy);                               // :68-69
}

// ---- AirHockey.computerAI  environment/airhockey.py:270-308 (written with selects: no divergence) ----
template <typename R>
__device__ __forceinline__ void computer_ai(Env<R>& e) {
    using M = Math<R>;
    const bool x_lt = e.px < e.ox, x_gt = e.px > e.ox;
    const R dx_lt = (e.px < K<R>::opp_lo) ? R(1) : R(-2);                      // :273-277
    const R dx_gt = (e.px > K<R>::opp_hi) ? R(-1) : R(2);                      // :279-283
    e.odx = x_lt ? dx_lt : (x_gt ? dx_gt : e.odx);
    const bool y_lt = e.py < e.oy, y_gt = e.py > e.oy;
    const R dy_lt = (e.py < K<R>::y_lo) ? R(1) : R(-6);                        // :285-289
    const R dy_gt = (e.py <= R(360)) ? R(6) : ((e.oy > R(200)) ? R(-2) : R(0));   // :292-300
    e.ody = y_lt ? dy_lt : (y_gt ? dy_gt : e.ody);
    // :303-305, only inside the `puck.y > opponent.y` branch
    R gx, gy;
    M::sub2(gx, gy, e.px, e.py, e.ox, e.oy);
    const bool near = y_gt && (M::abs_(gy) < R(40)) && (M::abs_(gx) < R(40));
    if (M::exact) {
        e.pdx = near ? e.pdx + R(2) : e.pdx;
        e.pdy = near ? e.pdy + R(2) : e.pdy;
    } else {
        const R inc = near ? R(2) : R(0);          // one select, one packed add with a broadcast operand
        M::add2(e.pdx, e.pdy, inc, inc);
    }
    mallet_update<R>(e.ox, e.oy, e.odx, e.ody, K<R>::opp_lo, K<R>::opp_hi);    // :307
}

__device__ __forceinline__ void opaque(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void opaque(double& v) { asm volatile("" : "+d"(v)); }

// ---- the common tail of AirHockey.update_state  environment/airhockey.py:112-122 ----
// The two contact tests (robot first, then opponent, :113-114) share ONE rarely-taken branch; inside it the
// opponent's test is re-evaluated against the puck as the robot's response left it, as the reference does.
//
// Speed clamp (:117).  limit_puck_speed is idempotent and the puck's velocity changes only in a contact response,
// in computerAI's nudge, or in a reset.  `speed_dirty` says "the velocity may be outside the limits": when it is
// false and no contact happens, the clamp is a no-op and is skipped.  ALWAYS_LIMIT (the computer's sub-update,
// whose nudge precedes it) clamps unconditionally.  Either way the result equals the reference's.
template <typename R, bool ALWAYS_LIMIT>
__device__ __forceinline__ void physics_tail(Env<R>& e, bool dot_fma, bool friction, bool& speed_dirty,
                                             int& c_robot, int& c_opp) {
    R rdx_, rdy_, odx_, ody_;
    const R dr = contact_dist_sq<R>(e.px, e.py, e.rx, e.ry, dot_fma, rdx_, rdy_);
    R dq = contact_dist_sq<R>(e.px, e.py, e.ox, e.oy, dot_fma, odx_, ody_);
    const bool slow = !(dr > K<R>::radius_sq) || !(dq > K<R>::radius_sq) || (!ALWAYS_LIMIT && (speed_dirty || friction));
    if (slow) {                                                                  // :210 (touching counts as contact)
        // (opaque copies: keep the two inner compares INSIDE the rare branch instead of hoisted onto the common path)
        R dr_in = dr;
        opaque(dr_in); opaque(dq);
        if (!(dr_in > K<R>::radius_sq)) {
            c_robot += 1;
            contact_response<R>(e.px, e.py, e.pdx, e.pdy, e.rx, e.ry, e.rdx, e.rdy, rdx_, rdy_, dr_in, dot_fma);
            dq = contact_dist_sq<R>(e.px, e.py, e.ox, e.oy, dot_fma, odx_, ody_);
        }
        if (!(dq > K<R>::radius_sq)) {
            c_opp += 1;
            contact_response<R>(e.px, e.py, e.pdx, e.pdy, e.ox, e.oy, e.odx, e.ody, odx_, ody_, dq, dot_fma);
        }
        if (!ALWAYS_LIMIT) {
            if (friction) friction_on_puck<R>(e.pdx, e.pdy);
            limit_puck_speed<R>(e.pdx, e.pdy);                                   // :117
        }
    }
    if (ALWAYS_LIMIT) {
        if (friction) friction_on_puck<R>(e.pdx, e.pdy);
        limit_puck_speed<R>(e.pdx, e.pdy);                                       // :117
    }
    speed_dirty = false;
    puck_update<R>(e.px, e.py, e.pdx, e.pdy);                                    // :118
    mallet_update<R>(e.rx, e.ry, e.rdx, e.rdy, K<R>::robot_lo, K<R>::robot_hi);  // :121-122
    mallet_update<R>(e.ox, e.oy, e.odx, e.ody, K<R>::opp_lo, K<R>::opp_hi);
}

// true when limit_puck_speed would change the velocity
template <typename R>
__device__ __forceinline__ bool speed_out_of_limits(R pdx, R pdy) {
    return !(Math<R>::abs_(pdx) <= K<R>::speed) || !(pdy <= K<R>::speed) || !(pdy >= -K<R>::speed);
}

// ---- AirHockey.get_state + serialize_state  airhockey.py:136-147, puck.py:120-133, helpers.py:72-81 ----
// obs = [int(agent.x), int(agent.y), int(puck.x), int(puck.y), agent.dx, agent.dy, puck.dx, puck.dy]
template <typename R> struct Obs { R v[8]; };
template <typename R>
__device__ __forceinline__ Obs<R> get_state(const Env<R>& e, bool robot) {
    using M = Math<R>;
    Obs<R> o;
    o.v[0] = M::trunc_(robot ? e.rx : e.ox); o.v[1] = M::trunc_(robot ? e.ry : e.oy);
    o.v[2] = M::trunc_(e.px); o.v[3] = M::trunc_(e.py);
    o.v[4] = robot ? e.rdx : e.odx; o.v[5] = robot ? e.rdy : e.ody;
    o.v[6] = e.pdx; o.v[7] = e.pdy;
    return o;
}

// pow(x, 2) vs x*x (see rewards_call): inside the band AND at least one square inexact (its rounding error != 0)
__device__ __forceinline__ bool done_is_ambiguous(double ex, double ey, double ssq) {
    if (!(fabs(ssq - 1225.0) < 1e-9)) return false;
    return __fma_rn(ex, ex, -(ex * ex)) != 0.0 || __fma_rn(ey, ey, -(ey * ey)) != 0.0;
}
__device__ __forceinline__ bool done_is_ambiguous(float, float, float) { return false; }

// ---- Rewards.__call__ for agent "robot"  lib/rewards.py:118-142 ----
template <typename R>
__device__ __forceinline__ void rewards_call(Env<R>& e, bool dot_fma, int& reward, int& score, bool& done, int& ambiguous) {
    using M = Math<R>;
    // compute_score_reward  rewards.py:64-81 with Goal.__contains__  goal.py:13-18; its reward and its own
    // `done` are discarded by the caller (rewards.py:122).  The left-goal test comes second and wins ties.
    const bool in_y = M::abs_(K<R>::mid_y - e.py) < R(95);
    const bool right = (M::abs_(K<R>::table_w - e.px) < R(30)) && in_y;
    const bool left = (M::abs_(R(0) - e.px) < R(30)) && in_y;
    score = left ? -1 : (right ? 1 : 0);
    // compute_mallet_hit_puck  rewards.py:83-89 with Puck.__and__  puck.py:135-146
    R ex, ey;
    M::sub2(ex, ey, e.px, e.py, e.rx, e.ry);
    const R dsq = M::dot2(ex, ey, ex, ey, dot_fma);
    if (M::exact) {
        // sqrt((px-mx)**2 + (py-my)**2) < 35 : plain (non-fused) sum of squares; `** 2` is libm pow() in the
        // reference, evaluated here as x*x.  pow(x, 2) is within an ulp of x*x (and equal to it when the square is exactly
        // representable, e.g. for the integer-valued states before the first contact), so the two can only decide
        // differently when the sum is within a few ulps (5e-13) of 1225 and a square is inexact: such frames are COUNTED
        // (AH_STAT_DONE_AMBIGUOUS, band 1e-9); a run whose count is 0 has `done` provably equal to the reference's.
        const R ssq = ex * ex + ey * ey;
        done = M::sqrt_(ssq) < K<R>::radius;
        ambiguous += done_is_ambiguous(ex, ey, ssq) ? 1 : 0;
    } else {
        done = dsq < K<R>::radius_sq;
    }
    // compute_wall_reward  rewards.py:45-62 with Puck.__lshift__/__rshift__  puck.py:148-164 (second `if` wins)
    const bool wl = M::abs_(e.px - K<R>::left_wall) < K<R>::puck_r;
    const bool wr = M::abs_(e.px - K<R>::right_wall) < K<R>::puck_r;
    // compute_mallet_to_puck_distance  rewards.py:91-101 (np.linalg.norm == sqrt(np.dot))
    const R nd = M::sqrt_(dsq);
    const bool closer = nd < e.prev_dist;
    e.prev_dist = nd;
    reward = (done ? 3 : -1) + (wr ? 2 : (wl ? -2 : 0)) + (closer ? 1 : -1);
}

}  // namespace ah
