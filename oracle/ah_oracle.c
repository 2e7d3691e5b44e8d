/*
 * oracle/ah_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A CPU restatement, in plain C / IEEE binary64, of the environment step of
 * Positronic-IO/air-hockey-training-environment (pure Python).  It exists so that the CUDA path can be
 * checked at sizes and on machines where the Python reference cannot run (the reference does not
 * travel to the GPU box).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it; the product package never does.
 *
 * Parity status: PINNED.  oracle/gen_golden.py runs the live Python reference (through
 * oracle/refshim.py) and commits trajectories under tests/golden/; tests/test_oracle_golden.py checks
 * this file against them bit-for-bit, and tests/test_oracle_vs_reference.py re-runs the comparison
 * live whenever /root/reference is present.  The reference's own unit-test pins (tests/test_puck.py,
 * test_mallet.py, test_goal.py, test_environment.py) are re-expressed in tests/test_reference_pins.py.
 *
 * Arithmetic that lives in a third-party dependency of the reference: numpy (Pipfile.lock pins
 * numpy==1.17.0; the container that generated the golden files has numpy 2.3.5 on
 * scipy-openblas 0.3.30).  np.dot of two 2-vectors goes through OpenBLAS ddot whose scalar tail is
 * FMA-contracted: np.dot([a,b],[c,d]) == fma(b, d, fl(a*c)); np.linalg.norm is sqrt of that.  This is
 * a property of the numpy build, so it is a run-time switch here (`dot_fma`), probed by
 * refshim.numpy_dot_is_fma() and recorded in every golden file.  `x ** 2` on floats is libm pow().
 *
 * Every function cites the reference file:line it restates (paths relative to the reference root).
 * The in-kernel RNG (Philox4x32-10, Salmon et al., SC'11 -- the Random123 algorithm) is NOT part of
 * the reference (which draws from numpy's global MT19937, environment/puck.py:115-116); it is
 * restated here only so that GPU runs that draw resets/actions on device can be checked exactly.
 *
 * Build: see oracle/build.py  (gcc -O2 -std=c11 -ffp-contract=off -fno-builtin-pow -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- constants: environment/config.py:4-13, environment/table.py:8-15, environment/mallet.py:33-44,
 *      environment/airhockey.py:35-49,63 ---- */
#define TABLE_W 900.0
#define TABLE_H 480.0
#define MALLET_R 20.0
#define PUCK_R 15.0
#define PUCK_IMASS 10.0      /* 1.0 / 0.1 == 10.0 exactly in binary64 (puck.py:36) */
#define MALLET_IMASS 2.0     /* 1.0 / 0.5 (mallet.py:27) */
#define DEFAULT_SPEED 10.0   /* config.puck["default_speed"] */
#define RESTITUTION 0.9
#define SLOP 0.01
#define PERCENT 0.2
#define LEFT_WALL 13.0       /* int(27 / 2) (table.py:14) */
#define RIGHT_WALL 882.0     /* int(900 - 38 + 40 / 2) (table.py:15) */
#define MID_X 450.0
#define MID_Y 240.0
#define STEP_SIZE 10.0       /* airhockey.py:63 */
#define ROBOT_LEFT 20.0
#define ROBOT_RIGHT 430.0
#define OPP_LEFT 470.0
#define OPP_RIGHT 880.0
#define U_LIM 20.0
#define B_LIM 460.0

typedef struct { double x, y, dx, dy; } body_t;

typedef struct {
    body_t puck, robot, opp;
    double prev_dist;            /* Rewards.prev_mallet_to_puck_distance, rewards.py:31 */
    int32_t robot_score, opp_score;
} env_t;

/* ------------------------------------------------------------------ Philox4x32-10 (not reference) */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

/* counter = (env_id lo, env_id hi, index, stream); key = (seed lo, seed hi).
 * stream 0: reset velocities, index = per-env reset counter.
 * stream 1: actions, index = frame / 64; 64 two-bit actions per block. */
static void rng_reset_pair(uint64_t seed, uint64_t env_id, uint32_t reset_idx, double* dx, double* dy) {
    uint32_t ctr[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), reset_idx, 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    /* 53-bit uniform in [0,1), same construction as MT19937 genrand_res53 used by np.random.uniform */
    double u0 = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
    double u1 = ((double)(o[2] >> 5) * 67108864.0 + (double)(o[3] >> 6)) * (1.0 / 9007199254740992.0);
    *dx = -10.0 + 20.0 * u0;   /* np.random.uniform(-10, 10) == low + (high-low)*u  (puck.py:115) */
    *dy = -10.0 + 20.0 * u1;   /* (puck.py:116) */
}

/* The FP32 THROUGHPUT build's reset draw (not reference either): pcg3d (Jarzynski & Olano, JCGT 9(3), 2020) on
 * (env_id lo ^ mix lo, reset index ^ env_id hi, mix hi) with mix = splitmix64(seed); 24-bit uniforms, one binary32 fma each
 * (csrc/ah_rng.cuh).  The result is a binary32 value, returned exactly in a double. */
static uint64_t splitmix64(uint64_t seed) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_reset_pair_f32(uint64_t seed, uint64_t env_id, uint32_t reset_idx, double* dx, double* dy) {
    const uint64_t mix = splitmix64(seed);
    uint32_t x = (uint32_t)env_id ^ (uint32_t)mix, y = reset_idx ^ (uint32_t)(env_id >> 32), z = (uint32_t)(mix >> 32);
    x = x * 1664525u + 1013904223u; y = y * 1664525u + 1013904223u; z = z * 1664525u + 1013904223u;
    x += y * z; y += z * x; z += x * y;
    x ^= x >> 16; y ^= y >> 16; z ^= z >> 16;
    x += y * z; y += z * x;
    *dx = (double)fmaf((float)(x >> 8), 20.0f / 16777216.0f, -10.0f);
    *dy = (double)fmaf((float)(y >> 8), 20.0f / 16777216.0f, -10.0f);
}

static int rng_action(uint64_t seed, uint64_t env_id, uint64_t frame) {
    uint32_t ctr[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)(frame >> 6), 1u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    uint32_t f = (uint32_t)(frame & 63u);
    return (int)((o[f >> 4] >> (2u * (f & 15u))) & 3u);
}

void orc_rng_reset_pair(uint64_t seed, uint64_t env_id, uint32_t reset_idx, double* out2) {
    rng_reset_pair(seed, env_id, reset_idx, &out2[0], &out2[1]);
}
void orc_rng_reset_pair_f32(uint64_t seed, uint64_t env_id, uint32_t reset_idx, double* out2) {
    rng_reset_pair_f32(seed, env_id, reset_idx, &out2[0], &out2[1]);
}
int orc_rng_action(uint64_t seed, uint64_t env_id, uint64_t frame) { return rng_action(seed, env_id, frame); }

/* ------------------------------------------------------------------ numpy primitives */
/* np.dot on 2-vectors: OpenBLAS ddot scalar tail, fma-contracted when dot_fma (see header). */
static inline double dot2(double a0, double a1, double b0, double b1, int dot_fma) {
    double p = a0 * b0;
    return dot_fma ? fma(a1, b1, p) : p + a1 * b1;
}

/* ------------------------------------------------------------------ environment/mallet.py:54-76 */
static inline void mallet_update(body_t* m, double left_lim, double right_lim, double* last_x, double* last_y) {
    *last_x = m->x;
    *last_y = m->y;
    m->x += m->dx;
    m->y += m->dy;
    if (m->x < left_lim) m->x = left_lim;
    else if (m->x > right_lim) m->x = right_lim;
    if (m->y < U_LIM) m->y = U_LIM;
    else if (m->y > B_LIM) m->y = B_LIM;
}

/* ------------------------------------------------------------------ environment/airhockey.py:65-94 */
/* kind 0: discrete int action `ia`; kind 1: continuous (ax, ay) added to the velocity. */
static void move_robot(env_t* e, int kind, int ia, double ax, double ay) {
    body_t* m = &e->robot;
    double lx, ly;
    if (kind == 1) {            /* :69-71 */
        m->dx += ax;
        m->dy += ay;
    } else {                    /* :74-84, any other int: no nudge */
        if (ia == 0) m->y += STEP_SIZE;
        if (ia == 1) m->y += -STEP_SIZE;
        if (ia == 2) m->x += -STEP_SIZE;
        if (ia == 3) m->x += STEP_SIZE;
    }
    mallet_update(m, ROBOT_LEFT, ROBOT_RIGHT, &lx, &ly);   /* :87 */
    m->dx = m->x - lx;                                      /* :90 */
    m->dy = m->y - ly;                                      /* :91 */
    mallet_update(m, ROBOT_LEFT, ROBOT_RIGHT, &lx, &ly);   /* :92 */
}

/* ------------------------------------------------------------------ environment/airhockey.py:189-268 */
static int collision(body_t* puck, body_t* mallet, int dot_fma, double* dmin) {
    double d_x = mallet->x - puck->x;                       /* :198 */
    double d_y = mallet->y - puck->y;                       /* :199 */
    double distance_squared = dot2(d_x, d_y, d_x, d_y, dot_fma);   /* :203 */
    const double radius = MALLET_R + PUCK_R;                /* :206 */
    const double radius_squared = radius * radius;          /* :207 */
    if (distance_squared > radius_squared) return 0;        /* :210 */
    double distance = sqrt(distance_squared);               /* :214 */
    if (dmin && distance < *dmin) *dmin = distance;         /* test diagnostics only */
    double n_x, n_y;
    if (distance > 0) { n_x = d_x / distance; n_y = d_y / distance; }   /* :217 */
    else { n_x = d_x; n_y = d_y; }
    double dcoll_x = radius - d_x;                          /* :220 -- a VECTOR, as in the reference */
    double dcoll_y = radius - d_y;
    const double imass_sum = PUCK_IMASS + MALLET_IMASS;     /* :223 */
    double a = dcoll_x - SLOP, b = dcoll_y - SLOP;          /* :230 dcoll - slop */
    double m = (a >= b || a != a) ? a : b;                  /* np.max(arr, 0): axis-0 max of the two */
    double s = (m / imass_sum) * PERCENT;                   /* :230 */
    double sep_x = s * n_x, sep_y = s * n_y;
    puck->x -= sep_x * PUCK_IMASS;                          /* :235 */
    puck->y -= sep_y * PUCK_IMASS;                          /* :236 */
    mallet->x += sep_x * MALLET_IMASS;                      /* :237 */
    mallet->y += sep_y * MALLET_IMASS;                      /* :238 */
    double vcoll_x = mallet->dx - puck->dx;                 /* :241 */
    double vcoll_y = mallet->dy - puck->dy;                 /* :242 */
    double vn = dot2(vcoll_x, vcoll_y, n_x, n_y, dot_fma);  /* :246 */
    if (vn > 0) return 1;                                   /* :249 */
    double j = -(1.0 + RESTITUTION) * (vn / imass_sum);     /* :256 */
    double imp_x = j * n_x, imp_y = j * n_y;                /* :259 */
    puck->dx -= imp_x * PUCK_IMASS;                         /* :262 */
    puck->dy -= imp_y * PUCK_IMASS;                         /* :263 */
    mallet->dx += imp_x * MALLET_IMASS;                     /* :265 */
    mallet->dy += imp_y * MALLET_IMASS;                     /* :266 */
    return 1;
}

/* ------------------------------------------------------------------ environment/puck.py:90-109 */
static inline void limit_puck_speed(body_t* p) {
    if (p->dx > DEFAULT_SPEED) p->dx = DEFAULT_SPEED;
    if (p->dx < -DEFAULT_SPEED) p->dx = -DEFAULT_SPEED;
    if (p->dy > DEFAULT_SPEED) p->dy = DEFAULT_SPEED;
    if (p->dy < -DEFAULT_SPEED) p->dy = DEFAULT_SPEED;     /* :102-103, sign as in the reference */
}

/* ------------------------------------------------------------------ environment/puck.py:73-88 (dead code in the reference; opt-in) */
static inline void friction_on_puck(body_t* p) {
    if (p->dx > 1) p->dx -= 0.5;
    if (p->dx < -1) p->dx += 0.5;
    if (p->dy > 1) p->dy -= 0.5;
    else if (p->dy < -1) p->dy += 0.5;
}

/* ------------------------------------------------------------------ environment/puck.py:45-71 */
static inline void puck_update(body_t* p) {
    if (p->x <= PUCK_R) { p->x = PUCK_R; p->dx *= -1; }
    else if (p->x >= TABLE_W - PUCK_R) { p->x = TABLE_W - PUCK_R; p->dx *= -1; }
    if (p->y <= PUCK_R) { p->y = PUCK_R; p->dy *= -1; }
    else if (p->y >= TABLE_H - PUCK_R) { p->y = TABLE_H - PUCK_R; p->dy *= -1; }
    p->x += p->dx;
    p->y += p->dy;
}

/* ------------------------------------------------------------------ environment/airhockey.py:270-308 */
static void computer_ai(env_t* e) {
    body_t* p = &e->puck;
    body_t* o = &e->opp;
    double lx, ly;
    if (p->x < o->x) { if (p->x < OPP_LEFT) o->dx = 1; else o->dx = -2; }          /* :273-277 */
    if (p->x > o->x) { if (p->x > OPP_RIGHT) o->dx = -1; else o->dx = 2; }         /* :279-283 */
    if (p->y < o->y) { if (p->y < U_LIM) o->dy = 1; else o->dy = -6; }             /* :285-289 */
    if (p->y > o->y) {                                                             /* :291 */
        if (p->y <= 360) o->dy = 6;                                                /* :292-293 */
        else { if (o->y > 200) o->dy = -2; else o->dy = 0; }                       /* :296-300 */
        if (fabs(p->y - o->y) < 40 && fabs(p->x - o->x) < 40) {                    /* :303 */
            p->dx += 2;                                                            /* :304 */
            p->dy += 2;                                                            /* :305 */
        }
    }
    mallet_update(o, OPP_LEFT, OPP_RIGHT, &lx, &ly);                               /* :307 */
}

/* ------------------------------------------------------------------ environment/airhockey.py:96-134 */
/* agent: 0 robot, 1 human, 2 computer.  Returns 0, or -1 for an unknown agent (InvalidAgentError, :108-110). */
static int update_state_c(env_t* e, int agent, int kind, int ia, double ax, double ay, int dot_fma, int friction,
                          int* contacts, double* dmin) {
    double lx, ly;
    if (agent == 0) move_robot(e, kind, ia, ax, ay);                               /* :100-101 */
    else if (agent == 1) {                                                         /* :102-105 */
        e->opp.x = ax; e->opp.y = ay;
        mallet_update(&e->opp, OPP_LEFT, OPP_RIGHT, &lx, &ly);
    } else if (agent == 2) computer_ai(e);                                         /* :106-107 */
    else return -1;
    int c0 = collision(&e->puck, &e->robot, dot_fma, dmin);                        /* :113-114, robot first */
    int c1 = collision(&e->puck, &e->opp, dot_fma, dmin);
    if (contacts) { contacts[0] += c0; contacts[1] += c1; }
    if (friction) friction_on_puck(&e->puck);   /* not in the reference's live path; off by default */
    limit_puck_speed(&e->puck);                                                    /* :117 */
    puck_update(&e->puck);                                                         /* :118 */
    mallet_update(&e->robot, ROBOT_LEFT, ROBOT_RIGHT, &lx, &ly);                   /* :121-122 */
    mallet_update(&e->opp, OPP_LEFT, OPP_RIGHT, &lx, &ly);
    return 0;
}

static int update_state(env_t* e, int agent, int kind, int ia, double ax, double ay, int dot_fma, int friction) {
    return update_state_c(e, agent, kind, ia, ax, ay, dot_fma, friction, NULL, NULL);
}

/* ------------------------------------------------------------------ airhockey.py:136-147 + puck.py:120-133 + mallet.py:86-99 */
static inline double py_int(double v) { return trunc(v); }   /* int(): truncation toward zero */
static void get_state(const env_t* e, int agent_is_robot, double* obs8) {
    const body_t* a = agent_is_robot ? &e->robot : &e->opp;
    obs8[0] = py_int(a->x); obs8[1] = py_int(a->y);
    obs8[2] = py_int(e->puck.x); obs8[3] = py_int(e->puck.y);
    obs8[4] = a->dx; obs8[5] = a->dy;
    obs8[6] = e->puck.dx; obs8[7] = e->puck.dy;
}

/* ------------------------------------------------------------------ lib/rewards.py:118-142 (agent "robot") */
static void rewards_call(env_t* e, int dot_fma, int* reward, int* score, int* done, int* ambiguous) {
    const body_t* p = &e->puck;
    const body_t* m = &e->robot;
    int sc = 0;
    /* compute_score_reward, rewards.py:64-81 with goal.py:13-18; goals at (900,240) and (0,240) */
    if (fabs(TABLE_W - p->x) < 2 * PUCK_R && fabs(MID_Y - p->y) < 95) sc = 1;
    if (fabs(0.0 - p->x) < 2 * PUCK_R && fabs(MID_Y - p->y) < 95) sc = -1;
    /* compute_mallet_hit_puck, rewards.py:83-89 with Puck.__and__, puck.py:135-146 (`** 2` is libm pow) */
    double ex = p->x - m->x, ey = p->y - m->y;
    double distance = sqrt(pow(ex, 2.0) + pow(ey, 2.0));
    /* The device evaluates `** 2` as x*x.  pow(x, 2) is within an ulp of x*x, so the two can only decide `done`
     * differently when the sum of squares is within a few ulps (5e-13) of 1225.  ambiguous[0]: frames inside the
     * (far wider) band 1e-9 the device counts as AH_STAT_DONE_AMBIGUOUS; ambiguous[1]: frames where the two
     * evaluations actually disagree. */
    if (ambiguous) {
        const double ssq = ex * ex + ey * ey;
        /* (exactly representable squares are exact under pow too: only inexact ones can differ) */
        ambiguous[0] += fabs(ssq - 1225.0) < 1e-9 && (fma(ex, ex, -(ex * ex)) != 0.0 || fma(ey, ey, -(ey * ey)) != 0.0);
        ambiguous[1] += (sqrt(ssq) < PUCK_R + MALLET_R) != (distance < PUCK_R + MALLET_R);
    }
    int rw, dn;
    if (distance < PUCK_R + MALLET_R) { rw = 3; dn = 1; } else { rw = -1; dn = 0; }
    /* compute_wall_reward, rewards.py:45-62 with puck.py:148-164 */
    int wall = 0;
    if (fabs(p->x - LEFT_WALL) < PUCK_R) wall = -1 * 2;
    if (fabs(p->x - RIGHT_WALL) < PUCK_R) wall = 2;
    rw += wall;
    /* compute_mallet_to_puck_distance, rewards.py:91-101 (np.linalg.norm == sqrt(dot)) */
    double nd = sqrt(dot2(ex, ey, ex, ey, dot_fma));
    if (nd < e->prev_dist) { e->prev_dist = nd; rw += 1; }
    else { e->prev_dist = nd; rw += -1; }
    *reward = rw; *score = sc; *done = dn;
}

/* ------------------------------------------------------------------ airhockey.py:171-187, puck.py:111-118, mallet.py:78-84 */
typedef struct {
    const double* table;  /* [R][2] for this env, or NULL => Philox (rng 0) / the FP32 build's pcg3d (rng 1) */
    int R, rng;
    int32_t cursor;       /* resets consumed */
    uint64_t seed, env_id;
    int overflow;
} reset_src_t;

static void env_reset(env_t* e, int total, reset_src_t* rs) {
    if (total) { e->opp_score = 0; e->robot_score = 0; }
    double dx, dy;
    if (rs->table) {
        if (rs->cursor < rs->R) { dx = rs->table[2 * rs->cursor]; dy = rs->table[2 * rs->cursor + 1]; }
        else { dx = 0.0; dy = 0.0; rs->overflow = 1; }
    } else {
        if (rs->rng == 1) rng_reset_pair_f32(rs->seed, rs->env_id, (uint32_t)rs->cursor, &dx, &dy);
        else rng_reset_pair(rs->seed, rs->env_id, (uint32_t)rs->cursor, &dx, &dy);
    }
    rs->cursor += 1;
    e->puck.x = MID_X; e->puck.y = MID_Y; e->puck.dx = dx; e->puck.dy = dy;
    e->robot.dx = 0; e->robot.dy = 0;     /* velocity only; positions are kept (mallet.py:78-84) */
    e->opp.dx = 0; e->opp.dy = 0;
}

/* ------------------------------------------------------------------ initial state: airhockey.py:25-63, puck.py:16 */
void orc_init(int64_t n, double* state13, int32_t* scores2) {
    for (int64_t i = 0; i < n; ++i) {
        double* s = state13 + 13 * i;
        s[0] = MID_X; s[1] = MID_Y; s[2] = -3; s[3] = 0;
        s[4] = MID_X - 100; s[5] = MID_Y; s[6] = 0; s[7] = 0;
        s[8] = MID_X + 100; s[9] = MID_Y; s[10] = 0; s[11] = 0;
        s[12] = INFINITY;
        scores2[2 * i] = 0; scores2[2 * i + 1] = 0;
    }
}

static inline void load_env(env_t* e, const double* s, const int32_t* sc) {
    e->puck = (body_t){s[0], s[1], s[2], s[3]};
    e->robot = (body_t){s[4], s[5], s[6], s[7]};
    e->opp = (body_t){s[8], s[9], s[10], s[11]};
    e->prev_dist = s[12];
    e->robot_score = sc[0]; e->opp_score = sc[1];
}
static inline void store_env(const env_t* e, double* s, int32_t* sc) {
    s[0] = e->puck.x; s[1] = e->puck.y; s[2] = e->puck.dx; s[3] = e->puck.dy;
    s[4] = e->robot.x; s[5] = e->robot.y; s[6] = e->robot.dx; s[7] = e->robot.dy;
    s[8] = e->opp.x; s[9] = e->opp.y; s[10] = e->opp.dx; s[11] = e->opp.dy;
    s[12] = e->prev_dist;
    sc[0] = e->robot_score; sc[1] = e->opp_score;
}

/* ------------------------------------------------------------------ one sub-update for tests of the low-level API */
int orc_update_state(int64_t n, double* state13, int32_t* scores2, int agent, int kind,
                     const int32_t* ia /*[n] or NULL*/, const double* axy /*[n][2] or NULL*/, int dot_fma, int friction) {
    int rc = 0;
    for (int64_t i = 0; i < n; ++i) {
        env_t e;
        load_env(&e, state13 + 13 * i, scores2 + 2 * i);
        int a = ia ? ia[i] : -1;
        double ax = axy ? axy[2 * i] : 0.0, ay = axy ? axy[2 * i + 1] : 0.0;
        if (update_state(&e, agent, kind, a, ax, ay, dot_fma, friction) != 0) rc = -1;
        store_env(&e, state13 + 13 * i, scores2 + 2 * i);
    }
    return rc;
}

void orc_get_state(int64_t n, const double* state13, int agent_is_robot, double* obs8) {
    int32_t z[2] = {0, 0};
    for (int64_t i = 0; i < n; ++i) {
        env_t e;
        load_env(&e, state13 + 13 * i, z);
        get_state(&e, agent_is_robot, obs8 + 8 * i);
    }
}

/* ------------------------------------------------------------------ the frame loop: game.py:133-177 + agent.py:57-78 */
typedef struct {
    /* per-frame outputs, each nullable; layout [T][n](...) */
    double* obs_state;      /* [T][n][8]  Observation.state      agent.py:61 */
    double* obs_new_state;  /* [T][n][8]  Observation.new_state  agent.py:67 */
    double* obs_next;       /* [T][n][8]  get_state("robot") at the end of the frame */
    int32_t* reward;        /* [T][n] */
    int8_t* score;          /* [T][n] */
    uint8_t* done;          /* [T][n] */
    uint8_t* game_over;     /* [T][n]  0 none, 1 robot reached 10, 2 opponent reached 10 (game.py:101-109) */
    double* states_rec;     /* [T/record_every][n][13] */
    int32_t* scores_rec;    /* [T/record_every][n][2] */
    int64_t* stats;         /* [16] summed over all envs and frames, see AH_STAT_* in include/ah_b200.h */
    uint8_t* contacts;      /* [T][n]  test diagnostics: collision() == True count in the frame (0..6) */
    double* contact_dmin;   /* [T][n]  test diagnostics: smallest centre distance among them (inf if none) */
} orc_outputs_t;

/*
 * actions_i8   [T][n] discrete actions (kind 0), or NULL
 * actions_f64  [T][n][2] continuous actions (kind 1), or NULL
 * both NULL => Philox actions, frame index = frame0 + t
 * reset_table  [n][R][2] or NULL => Philox resets
 * reset_cursor [n] in/out
 * returns 0, or 1 if a reset table overflowed.
 */
/* reset_rng: which in-kernel generator the table-less resets restate: 0 Philox4x32-10 (the FP64 build), 1 the FP32
 * build's pcg3d */
int orc_frames_ex(int64_t n, int T, double* state13, int32_t* scores2,
                  const int8_t* actions_i8, const double* actions_f64,
                  const double* reset_table, int R, int32_t* reset_cursor,
                  uint64_t seed, uint64_t env_id0, uint64_t frame0,
                  int dot_fma, int friction, int record_every, int nthreads, int reset_rng,
                  const orc_outputs_t* out) {
    int overflow = 0;
    int64_t stats_acc[20];
    memset(stats_acc, 0, sizeof stats_acc);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(| : overflow) reduction(+ : stats_acc[:20])
#endif
    for (int64_t i = 0; i < n; ++i) {
        env_t e;
        load_env(&e, state13 + 13 * i, scores2 + 2 * i);
        reset_src_t rs;
        rs.table = reset_table ? reset_table + (size_t)i * R * 2 : NULL;
        rs.R = R; rs.rng = reset_rng; rs.cursor = reset_cursor[i]; rs.seed = seed; rs.env_id = env_id0 + (uint64_t)i; rs.overflow = 0;
        for (int t = 0; t < T; ++t) {
            size_t o = (size_t)t * n + i;
            int kind = 0, ia = -1; double ax = 0, ay = 0;
            if (actions_i8) ia = actions_i8[o];
            else if (actions_f64) { kind = 1; ax = actions_f64[2 * o]; ay = actions_f64[2 * o + 1]; }
            else ia = rng_action(seed, rs.env_id, frame0 + (uint64_t)t);
            int cts[2] = {0, 0}; double dmin = INFINITY;
            update_state_c(&e, 0, kind, ia, ax, ay, dot_fma, friction, cts, &dmin);   /* robot.move(a)  game.py:136 */
            if (out->obs_state) get_state(&e, 1, out->obs_state + 8 * o); /* state                agent.py:61 */
            update_state_c(&e, 0, kind, ia, ax, ay, dot_fma, friction, cts, &dmin);   /* self.move(a)   agent.py:64 */
            if (out->obs_new_state) get_state(&e, 1, out->obs_new_state + 8 * o);   /*            agent.py:67 */
            int rw, sc, dn, amb[2] = {0, 0};
            rewards_call(&e, dot_fma, &rw, &sc, &dn, amb);                /*                      agent.py:71 */
            if (sc > 0) { e.robot_score += 1; env_reset(&e, 0, &rs); }    /* update_score     airhockey.py:152-158 */
            if (sc < 0) { e.opp_score += 1; env_reset(&e, 0, &rs); }      /*                  airhockey.py:161-168 */
            int go = 0;                                                   /* stats()             game.py:101-109 */
            if (e.robot_score == 10) go = 1;
            if (e.opp_score == 10) go = 2;
            update_state_c(&e, 2, 0, -1, 0, 0, dot_fma, friction, cts, &dmin);        /* computer       game.py:166 */
            if (e.opp_score == 10) env_reset(&e, 1, &rs);                 /*                     game.py:169-172 */
            if (e.robot_score == 10) env_reset(&e, 1, &rs);               /*                     game.py:174-177 */
            if (out->obs_next) get_state(&e, 1, out->obs_next + 8 * o);
            if (out->reward) out->reward[o] = rw;
            if (out->score) out->score[o] = (int8_t)sc;
            if (out->done) out->done[o] = (uint8_t)dn;
            if (out->game_over) out->game_over[o] = (uint8_t)go;
            if (out->contacts) out->contacts[o] = (uint8_t)(cts[0] + cts[1]);
            if (out->contact_dmin) out->contact_dmin[o] = dmin;
            stats_acc[14] += cts[0];           /* collision() true, robot mallet */
            stats_acc[15] += cts[1];           /* collision() true, opponent mallet */
            stats_acc[0] += 1;                 /* frames */
            stats_acc[1] += (sc > 0);          /* robot goals */
            stats_acc[2] += (sc < 0);          /* opponent goals */
            stats_acc[3] += (go == 1);         /* robot wins */
            stats_acc[4] += (go == 2);         /* opponent wins */
            stats_acc[5] += dn;                /* dones */
            stats_acc[6] += rw;                /* sum reward */
            stats_acc[7] += (int64_t)rw * rw;  /* sum reward^2 */
            stats_acc[16] += amb[0];           /* `done` within the pow-vs-multiply ambiguity band */
            stats_acc[17] += amb[1];           /* `done` by libm pow != `done` by x*x */
            if (record_every > 0 && (t + 1) % record_every == 0) {
                size_t r = (size_t)((t + 1) / record_every - 1) * n + i;
                if (out->states_rec) { int32_t tmp[2]; store_env(&e, out->states_rec + 13 * r, out->scores_rec ? out->scores_rec + 2 * r : tmp); }
            }
        }
        store_env(&e, state13 + 13 * i, scores2 + 2 * i);
        reset_cursor[i] = rs.cursor;
        overflow |= rs.overflow;
    }
    if (out->stats) for (int k = 0; k < 20; ++k) out->stats[k] += stats_acc[k];
    return overflow;
}

int orc_frames(int64_t n, int T, double* state13, int32_t* scores2,
               const int8_t* actions_i8, const double* actions_f64,
               const double* reset_table, int R, int32_t* reset_cursor,
               uint64_t seed, uint64_t env_id0, uint64_t frame0,
               int dot_fma, int friction, int record_every, int nthreads,
               const orc_outputs_t* out) {
    return orc_frames_ex(n, T, state13, scores2, actions_i8, actions_f64, reset_table, R, reset_cursor, seed, env_id0, frame0,
                         dot_fma, friction, record_every, nthreads, 0, out);
}

/* predicates pinned by the reference's unit tests (tests/test_puck.py, tests/test_goal.py, tests/test_mallet.py) */
int orc_puck_overlaps_mallet(double px, double py, double mx, double my) {        /* puck.py:135-146 */
    return sqrt(pow(px - mx, 2.0) + pow(py - my, 2.0)) < PUCK_R + MALLET_R;
}
int orc_puck_hits_left_wall(double px) { return fabs(px - LEFT_WALL) < PUCK_R; }   /* puck.py:148-155 */
int orc_puck_hits_right_wall(double px) { return fabs(px - RIGHT_WALL) < PUCK_R; } /* puck.py:157-164 */
int orc_puck_in_goal(double gx, double gy, double px, double py) {                /* goal.py:13-18 */
    return fabs(gx - px) < 2 * PUCK_R && fabs(gy - py) < 95;
}
/* which: 0 generic, 1 robot, 2 opponent (mallet.py:33-44) */
void orc_friction_on_puck(double* dxdy) {                                         /* puck.py:73-88 */
    body_t p; p.x = p.y = 0; p.dx = dxdy[0]; p.dy = dxdy[1];
    friction_on_puck(&p);
    dxdy[0] = p.dx; dxdy[1] = p.dy;
}
void orc_mallet_update(int which, double* xydxdy) {
    body_t m = {xydxdy[0], xydxdy[1], xydxdy[2], xydxdy[3]};
    double lx, ly;
    double l = which == 2 ? OPP_LEFT : MALLET_R, r = which == 1 ? ROBOT_RIGHT : TABLE_W - MALLET_R;
    mallet_update(&m, l, r, &lx, &ly);
    xydxdy[0] = m.x; xydxdy[1] = m.y; xydxdy[2] = m.dx; xydxdy[3] = m.dy;
}
