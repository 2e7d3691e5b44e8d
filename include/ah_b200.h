/*
 * include/ah_b200.h -- C ABI of the B200-native batched air-hockey environment step.
 *
 * This is the drop-in boundary for the ONE hot path this project accelerates: the environment step of
 * Positronic-IO/air-hockey-training-environment (`AirHockey` reset / update_state / get_state, the
 * `Rewards` reward and the `computerAI` heuristic opponent, looped as game.py does), for N independent
 * games resident in HBM.  The reference has no FFI for this path (it is pure Python); the entry points
 * below are what a ctypes binding of the reference's environment API would bind -- each one cites the
 * reference interface it replaces (paths relative to the reference root).  INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - plain C: opaque handle, raw pointers and sizes; no torch / C++ types cross this boundary;
 *   - every pointer named `d_*` or documented "device" is a CUDA device address owned by the CALLER
 *     (e.g. torch tensors).  The library never allocates state or outputs: its only allocations are the
 *     host handle and one 256-byte device block (statistics + counters);
 *   - every call enqueues on the caller's stream (`void* stream` is a cudaStream_t; NULL = legacy default
 *     stream), never synchronises and never allocates, so all of them can be captured in a CUDA graph;
 *   - return value: AH_OK (0) or a negative AH_E* code; ah_strerror() names it.  No exception or signal
 *     crosses the boundary.  There is NO CPU fallback: without a CUDA device every compute call fails with
 *     AH_ECUDA.
 *   - a handle is not thread-safe; distinct handles (one per GPU) are independent.
 *
 * State layout in HBM (structure of arrays, one row per environment, all rows 16-byte aligned):
 *   puck      real[n][4]   x, y, dx, dy                          environment/puck.py:16-43
 *   robot     real[n][4]   x, y, dx, dy   (left mallet)           environment/mallet.py:11-48, airhockey.py:46
 *   opponent  real[n][4]   x, y, dx, dy   (right mallet)          airhockey.py:49
 *   prev_dist real[n]      Rewards.prev_mallet_to_puck_distance   lib/rewards.py:31,96-100
 *   meta      int32[n][4]  [0] robot_score | opponent_score<<16   airhockey.py:52
 *                          [1] resets consumed (RNG counter / reset-table cursor)
 *                          [2] frames played in the current game, [3] reward summed over the current game
 *   `real` is float (AH_F32, throughput build) or double (AH_F64, validation build that preserves the
 *   reference's operation order and numpy's FMA-contracted dot products).
 */
#ifndef AH_B200_H
#define AH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AH_ABI_VERSION 3
#define AH_ACTOR_NWEIGHTS 114   /* W1[4][8] b1[4] W2[6][4] b2[6] W3[4][6] b3[4] W4[4][4] b4[4], see ah_actor_sample */

typedef struct ah_env ah_env;

/* ---- error codes ---- */
enum {
    AH_OK = 0,
    AH_EINVAL = -1,   /* NULL / misaligned pointer, n <= 0, K out of range, bad enum */
    AH_ECUDA = -2,    /* a CUDA runtime call failed; ah_last_cuda_error() has the cudaError_t */
    AH_ENOMEM = -3,
    AH_EAGENT = -4,   /* unknown agent: the reference's InvalidAgentError (airhockey.py:108-110) */
    AH_ESTATE = -5    /* state not bound yet */
};

/* ---- enums ---- */
enum { AH_F32 = 0, AH_F64 = 1 };
/* agent_name of AirHockey.update_state / get_state (airhockey.py:100-110, :139) */
enum { AH_AGENT_ROBOT = 0, AH_AGENT_HUMAN = 1, AH_AGENT_COMPUTER = 2 };
/* where the robot's actions come from (Action = Union[int, Tuple], lib/types.py:5; airhockey.py:65-84) */
enum {
    AH_ACT_NONE = 0,         /* no action argument (agent "computer")                                   */
    AH_ACT_DISCRETE = 1,     /* int8 [K][n]: 0 y+=10, 1 y-=10, 2 x-=10, 3 x+=10, anything else no nudge  */
    AH_ACT_CONTINUOUS = 2,   /* real [K][n][2]: added to the mallet velocity (airhockey.py:69-71); for     */
                             /* agent "human": real [n][2] = the new mallet location (airhockey.py:104)   */
    AH_ACT_PHILOX = 3,       /* uniform {0,1,2,3} drawn in-kernel (synthetic random-action workload)     */
    AH_ACT_DISCRETE_REPEAT = 4,   /* int8 [n]: one action applied to all K frames (frame-skip)            */
    AH_ACT_CONTINUOUS_REPEAT = 5, /* real [n][2]: one action applied to all K frames                      */
    AH_ACT_POLICY = 6,       /* closed loop: every frame the reference's discrete PPO actor is evaluated IN the  */
                             /* kernel on the env's current observation and an action is sampled from its       */
                             /* softmax (see ah_actor_sample); `actions` = device float[AH_ACTOR_NWEIGHTS]       */
    AH_ACT_POLICY_MLP = 8,   /* closed loop with ANY small dense policy: ah_step_io.policy (an ah_policy) is evaluated IN   */
                             /* the kernel every frame and its exploration strategy picks the action; `actions` unused  */
    AH_ACT_DISCRETE_PACKED4 = 7 /* uint8 [ceil(K/4)][n][4], "quad-major": byte j of word (q, i) is env i's action of */
                             /* frame 4q + j (same action codes as AH_ACT_DISCRETE; 32-bit aligned).  The input     */
                             /* format of the packed rollout mode (ah_step_io.events)                               */
};

/* One byte per env-frame carrying everything a frame of Game.play returns (ah_step_io.events):
 *   bits 0-2  (reward + 4) / 2    reward in {-4,-2,0,2,4,6}      lib/rewards.py:118-142
 *   bit  3    done                "mallet overlaps puck"          rewards.py:86-89,124
 *   bits 4-5  score + 1           -1 / 0 / +1                     rewards.py:64-81
 *   bits 6-7  game_over           0 / 1 robot won / 2 opponent    game.py:101-109,169-177 */
#define AH_EVENT_REWARD(b)    ((int)(((b) & 7u) << 1) - 4)
#define AH_EVENT_DONE(b)      ((int)(((b) >> 3) & 1u))
#define AH_EVENT_SCORE(b)     ((int)(((b) >> 4) & 3u) - 1)
#define AH_EVENT_GAME_OVER(b) ((int)(((b) >> 6) & 3u))

/* indices into the statistics vector (int64[AH_NSTATS]); the device-side equivalent of the per-frame
 * scores.csv row (game.py:82-112) and of the rewards CSV (lib/rewards.py:129-140) */
enum {
    AH_STAT_FRAMES = 0,
    AH_STAT_ROBOT_GOALS = 1,
    AH_STAT_OPPONENT_GOALS = 2,
    AH_STAT_ROBOT_WINS = 3,
    AH_STAT_OPPONENT_WINS = 4,
    AH_STAT_DONES = 5,
    AH_STAT_REWARD_SUM = 6,
    AH_STAT_REWARD_SQ_SUM = 7,
    AH_STAT_GAMES = 8,            /* games finished (robot_wins + opponent_wins) */
    AH_STAT_GAME_LEN_SUM = 9,     /* frames of the finished games */
    AH_STAT_GAME_LEN_SQ_SUM = 10,
    AH_STAT_GAME_RETURN_SUM = 11, /* reward summed over the finished games */
    AH_STAT_NONFINITE = 12,       /* envs whose puck state was non-finite at the end of a launch */
    AH_STAT_RESET_TABLE_OVERFLOW = 13, /* resets that found their reset table exhausted (velocity set to 0) */
    AH_STAT_CONTACTS_ROBOT = 14,  /* AirHockey.collision() returned True, robot mallet (airhockey.py:189-268) */
    AH_STAT_CONTACTS_OPPONENT = 15,
    AH_STAT_DONE_AMBIGUOUS = 16,  /* AH_F64 only: frames whose `done` test sqrt(dx**2 + dy**2) < 35 (puck.py:138) had its   */
                                  /* sum of squares within 1e-9 of 1225 with an inexact square -- the only frames where libm */
                                  /* pow(x, 2) (the reference) and x*x (the device) could decide differently; 0 proves `done` */
                                  /* exact for the run                                                                     */
    AH_NSTATS = 20                /* 17-19 reserved */
};

/* ---- a small dense policy + exploration strategy, evaluated on the device (ah_policy_act, AH_ACT_POLICY_MLP).
 * Replaces the per-frame host hand-off Agent.get_action -> model.predict -> exploration strategy
 * (lib/agents/agent.py:42-48, lib/exploration/strategies.py, lib/exploration/random.py). ---- */
#define AH_POLICY_MAX_LAYERS 4
#define AH_POLICY_MAX_WIDTH 32
enum { AH_ACTV_LINEAR = 0, AH_ACTV_RELU = 1, AH_ACTV_TANH = 2 };
enum {
    AH_HEAD_SOFTMAX = 0,     /* last layer -> softmax probabilities (PPO / A2C actors, the reference's DDQN net)       */
    AH_HEAD_VALUES = 1,      /* last layer = Q-values (or probabilities) as they are                                   */
    AH_HEAD_DUELING = 2,     /* last layer = (v, a_0 .. a_{A-1}); q = v + a - mean(a)   lib/agents/dueling/model.py:50-66 */
    AH_HEAD_CONTINUOUS = 3   /* last layer = (mx, my): the continuous action before noise  lib/agents/ppo/model.py:84-110 */
};
enum {
    AH_SELECT_SAMPLE = 0,        /* SoftmaxPolicy: np.random.choice(A, p)                 strategies.py:7-19   */
    AH_SELECT_GREEDY = 1,        /* np.argmax (what every agent does when not training); continuous: the mean  */
    AH_SELECT_EPS_GREEDY = 2,    /* EpsilonGreedy, with its epsilon schedule              strategies.py:22-56  */
    AH_SELECT_BOLTZMANN = 3,     /* BoltzmannQPolicy                                      strategies.py:59-76  */
    AH_SELECT_MAX_BOLTZMANN = 4, /* MaxBoltzmannQPolicy                                   strategies.py:79-123 */
    AH_SELECT_GAUSSIAN = 5,      /* continuous: mean + GaussianWhiteNoiseProcess.sample() random.py:31-41      */
    AH_SELECT_OU = 6             /* continuous: mean + OrnsteinUhlenbeckProcess.sample()  random.py:44-67      */
};
typedef struct ah_policy {
    int32_t n_layers;            /* 1 .. AH_POLICY_MAX_LAYERS dense layers                                          */
    int32_t width[5];            /* width[0] = 8 (the observation), width[l] = outputs of layer l, each <= 32       */
    int32_t activation[4];       /* AH_ACTV_* per layer (fold BatchNorm into the neighbouring layer)               */
    int32_t head, select, n_actions;   /* n_actions <= 7; the last width is n_actions (softmax / values),          */
                                 /* n_actions + 1 (dueling) or 2 (continuous)                                       */
    /* epsilon schedule (strategies.py:31-43): epsilon starts at `epsilon`; on step() call t = 1, 2, ... it is lowered */
    /* by (initial_epsilon - final_epsilon) / explore whenever t % observe == 0 and epsilon > final_epsilon            */
    float epsilon, initial_epsilon, final_epsilon;
    int32_t observe, explore;
    float tau, clip_lo, clip_hi; /* Boltzmann: exp(clip(q / tau, clip_lo, clip_hi))        strategies.py:66-75      */
    /* noise (random.py:11-67): sigma(t) = max(sigma_min, sigma - (sigma - sigma_min) * t / n_steps_annealing);    */
    /* sigma_min < 0 means "no annealing" (the reference's sigma_min=None)                                          */
    float mu, sigma, sigma_min; int32_t n_steps_annealing;
    float theta, dt;             /* Ornstein-Uhlenbeck                                                               */
    const float* weights;        /* device float[]: per layer Wt[in][outp] (TRANSPOSED, outputs zero-padded to a     */
                                 /* multiple of 4 = outp) followed by b[outp]; 16-byte aligned                       */
    float* noise_state;          /* device float [n][2]: the OU process' x_prev per env (AH_SELECT_OU), else NULL    */
} ah_policy;

/* ---- configuration (flags only; the physical constants of environment/config.py are compile-time) ---- */
typedef struct ah_config {
    int32_t dot_fma;    /* 1: the 2-vector dot products are fma(b,d,fl(a*c)) like numpy/OpenBLAS (default) */
    int32_t friction;   /* 1: apply Puck.friction_on_puck (puck.py:73-88) before the speed clamp; the      */
                        /*    reference never calls it (dead code), so the default is 0                    */
    int32_t no_prefetch; /* 1: never use the small-K persistent/prefetching variant of the step kernel (tuning / tests) */
    int32_t no_packed_core; /* 1: the general API's tensors (int8 [K][n] actions, four per-frame outputs) and the Philox  */
                            /*    simulation run on round 1's step kernel instead of the packed-core kernel (tests compare */
                            /*    the two; same results either way)                                                        */
    int32_t reserved[4];
} ah_config;

/* ---- optional fused replay/rollout ring (replaces MemoryBuffer.append, lib/buffer.py:24-32, fed with the
 *      Observation tuple of lib/agents/agent.py:74).  Slot = (appended so far) mod capacity; the newest
 *      entry is read first, as `appendleft` + `retreive` do (lib/buffer.py:27,45). ---- */
typedef struct ah_ring {
    void* state;        /* real   [capacity][n][8]  Observation.state     */
    void* new_state;    /* real   [capacity][n][8]  Observation.new_state */
    void* action;       /* int8   [capacity][n]     (discrete) or real [capacity][n][2] (continuous) */
    void* reward;       /* real   [capacity][n] */
    uint8_t* done;      /* uint8  [capacity][n] */
    int64_t capacity;
    uint64_t* appended; /* device uint64[AH_RING_COUNTERS] (equal values; see ah_step_io.ring_counter_first), owned by the */
                        /* ring: Observation tuples appended so far.  The step kernel                                 */
                        /* reads the write slot from it and advances it (so captured graphs advance it too);          */
                        /* MemoryBuffer.purge (lib/buffer.py:52-61) = set it to 0                                     */
    /* AH_RING_COMPACT (AH_F32, discrete actions, packed rollout mode): the same tuples in 51 instead of 70 bytes, every    */
    /* array contiguous per warp -- the observation's four locations are int()-truncated in the reference                  */
    /* (airhockey.py:136-147), so they are kept as int16; `state` / `new_state` then hold only the four velocities,         */
    /* float [capacity][n][4]; `reward` is int8 [capacity][n] (lib/rewards.py:118-142 returns integers in [-4, 6]).          */
    /* ah_ring_sample decodes to the same outputs either way.                                                              */
    int32_t format;       /* AH_RING_ROWS (0) or AH_RING_COMPACT */
    int32_t reserved0;
    void* state_pos;      /* AH_RING_COMPACT: int16 [capacity][n][4]  int(agent.x), int(agent.y), int(puck.x), int(puck.y) */
    void* new_state_pos;  /* AH_RING_COMPACT: int16 [capacity][n][4] */
} ah_ring;
enum { AH_RING_ROWS = 0, AH_RING_COMPACT = 1 };

/* ---- inputs / outputs of the fused step ---- */
typedef struct ah_step_io {
    /* inputs */
    const void* actions;          /* device; see AH_ACT_* (NULL for AH_ACT_PHILOX) */
    int32_t action_kind;
    int32_t reset_table_len;      /* R */
    const double* reset_table;    /* device double [n][R][2] (dx, dy) per reset, or NULL => drawn in the kernel  */
                                  /* by a counter-based generator keyed (seed, global env id, reset index):      */
                                  /* Philox4x32-10 in the AH_F64 build, pcg3d in the AH_F32 build.               */
                                  /* Stands in for the global np.random.uniform draws of puck.py:115-116.       */
    /* per-frame outputs, each nullable; [K][n]... */
    void* obs_state;              /* real [K][n][8]  Observation.state     (agent.py:61)  */
    void* obs_new_state;          /* real [K][n][8]  Observation.new_state (agent.py:67)  */
    void* reward;                 /* real [K][n]     lib/rewards.py:118-142               */
    uint8_t* done;                /* u8   [K][n]     "mallet overlaps puck" (rewards.py:86-89,124) */
    int8_t* score;                /* i8   [K][n]     +1 robot scored, -1 opponent scored (rewards.py:64-81) */
    uint8_t* game_over;           /* u8   [K][n]     0, 1 robot reached 10, 2 opponent reached 10 (game.py:101-109,169-177) */
    /* observation at the end of the frame = the policy input of the next frame (agent.py:46) */
    void* obs_next;               /* real [n][8] (last frame only) or [K][n][8] when obs_next_per_frame */
    int32_t obs_next_per_frame;
    int32_t collect_stats;        /* accumulate the AH_STAT_* vector (warp-reduced, one atomic set per block) */
    const ah_ring* ring;          /* NULL => no append */
    /* the actions actually applied and, for AH_ACT_POLICY, the actor's distribution (PPO's old_prediction) */
    int8_t* actions_out;          /* i8   [K][n], nullable (discrete / Philox / policy actions) */
    void* probs_out;              /* real [K][n][4], nullable (AH_ACT_POLICY only) */
    /* PACKED rollout mode: all four per-frame outputs in one byte per env-frame (AH_EVENT_*), quad-major like the
     * actions: u8 [ceil(K/4)][n][4], 32-bit aligned.  Requires action_kind == AH_ACT_DISCRETE_PACKED4 and excludes the
     * other per-frame outputs (reward, done, score, game_over, obs_state, obs_new_state, per-frame obs_next,
     * actions_out, probs_out) and the friction flag; obs_next [n][8] and the fused ring append (`ring`, discrete action
     * layout; from sub-range launches too, see ring_counter_first) stay available. */
    uint8_t* events;
    /* AH_ACT_POLICY_MLP: the policy to evaluate (host struct; its weights are a device pointer), and where the continuous
     * actions it chose go (real [K][n][2], nullable; discrete ones go to actions_out, head values to probs_out as
     * real [K][n][n_actions], or [K][n][2] for the continuous head) */
    const struct ah_policy* policy;
    void* actions_real_out;
    /* Sub-range launch (packed mode only): step only the envs [env_first, env_first + env_count) of the bound arrays;
     * env_count == 0 means all n.  Array shapes and strides stay those of the full n, so a caller can split one step
     * into disjoint ranges on DIFFERENT streams: the tail of one range's launch then overlaps the head of the next
     * range's (host.InterleavedStepper).  The frame counter advances with the range that contains env 0. */
    int64_t env_first;
    int64_t env_count;
    /* Which of the handle's two statistics banks this launch accumulates into (0 or 1).  A caller that overlaps steps on
     * several streams alternates the bank per rollout, so that closing a rollout (ah_stats_bank: copy + clear of the
     * finished bank) never has to wait for -- or race with -- the launches of the next one. */
    int32_t stats_bank;
    /* Ring append from SUB-RANGE launches (packed mode): `ring->appended` is then an array of AH_RING_COUNTERS uint64 that
     * all hold the same count between steps; this launch reads its write slot from appended[ring_counter_first] and its
     * last block advances appended[ring_counter_first .. + ring_counter_count).  A step issued as P concurrent slices
     * gives slice j the counters {j} (the last slice: {P-1 .. AH_RING_COUNTERS-1}), so that no two concurrent launches
     * touch the same counter and all of them stay in lockstep.  count == 0 (the default): a whole-n launch, which reads
     * appended[0] and advances all of them.  ah_ring_sample reads appended[0]. */
    int32_t ring_counter_first;
    int32_t ring_counter_count;
    int32_t reserved0;
} ah_step_io;
#define AH_RING_COUNTERS 4

const char* ah_strerror(int code);
int ah_abi_version(void);
int ah_last_cuda_error(const ah_env* env);

/* AirHockey() for n games (airhockey.py:25-63).  `env_id0` is the global id of local env 0: the RNG is
 * keyed on (seed, env_id0 + i) so results do not depend on how the envs are sharded over GPUs. */
int ah_create(ah_env** out, int64_t n, int dtype, int device, uint64_t seed, uint64_t env_id0, const ah_config* cfg);
int ah_destroy(ah_env* env);

/* Attach caller-owned device state (layout above). */
int ah_bind_state(ah_env* env, void* d_puck, void* d_robot, void* d_opponent, void* d_prev_dist, int32_t* d_meta);

/* Write the constructor state (airhockey.py:39-52, puck.py:16, rewards.py:31) into the bound arrays. */
int ah_init_state(ah_env* env, void* stream);

/* AirHockey.reset(total) (airhockey.py:171-187) for the envs whose mask byte is non-zero (NULL = all). */
int ah_reset(ah_env* env, int total, const uint8_t* d_mask, const double* d_reset_table, int reset_table_len,
             void* stream);

/* AirHockey.update_state(action, agent_name) (airhockey.py:96-134): ONE physics sub-update. */
int ah_update_state(ah_env* env, const void* d_actions, int action_kind, int agent, void* stream);

/* AirHockey.get_state(agent_name) flattened by serialize_state (airhockey.py:136-147, helpers.py:72-81):
 * real [n][8] = int(agent.x), int(agent.y), int(puck.x), int(puck.y), agent.dx, agent.dy, puck.dx, puck.dy */
int ah_get_state(ah_env* env, int agent, void* d_obs, void* stream);

/* AirHockey.update_score(score) (airhockey.py:149-169): score int8 [n]. */
int ah_update_score(ah_env* env, const int8_t* d_score, const double* d_reset_table, int reset_table_len,
                    void* stream);

/* Rewards.__call__(puck, robot_mallet) (lib/rewards.py:118-142) on the current state: reward real [n],
 * score int8 [n] (+1 / -1 / 0), done uint8 [n]; updates prev_dist.  Outputs are nullable. */
int ah_rewards(ah_env* env, void* d_reward, int8_t* d_score, uint8_t* d_done, void* stream);

/* K fused frames of Game.play (game.py:133-177 + agent.py:57-78): per frame
 *   update_state(a,"robot"); state; update_state(a,"robot"); new_state; Rewards; update_score;
 *   update_state("computer"); 10-goal game over.            1 <= K <= 4096 */
int ah_step(ah_env* env, const ah_step_io* io, int K, void* stream);

/* On-device action selection for rollouts (the hand-off the reference does on the host: Agent.get_action,
 * lib/agents/agent.py:42-48 -> actor.predict -> exploration strategy, lib/agents/ppo/ppo.py:100-121).
 * Evaluates the reference's discrete PPO actor (lib/agents/ppo/model.py:56-74: Dense 4 relu, BatchNorm, Dense 6 relu,
 * Dense 4 relu, Dense 4 softmax) on obs real [n][8] and samples one action per env from the softmax with an in-kernel
 * Philox uniform (or takes the arg-max when `greedy`, as the reference does when not training, ppo.py:114).
 * d_weights: device float[AH_ACTOR_NWEIGHTS], row-major W1[4][8] b1[4] W2[6][4] b2[6] W3[4][6] b3[4] W4[4][4] b4[4]
 * with the BatchNorm already folded into W2/b2.  d_probs (real [n][4], PPO's `old_prediction`) may be NULL. */
int ah_actor_sample(ah_env* env, const float* d_weights, const void* d_obs, void* d_probs, int8_t* d_actions,
                    int greedy, void* stream);

/* The generic form of ah_actor_sample: evaluate `policy` on obs real [n][8] -- or, when d_values_in (real
 * [n][n_actions]) is given, take the head values from the caller's own network -- and select one action per env with
 * the policy's exploration strategy, at the env's current frame (random draws as in AH_ACT_POLICY_MLP).  Outputs:
 * d_actions int8 [n] (discrete heads) or d_actions_real real [n][2] (continuous head); d_values_out (nullable) receives
 * the head values (probabilities / Q-values / mean).  The caller advances time with ah_step. */
int ah_policy_act(ah_env* env, const ah_policy* policy, const void* d_obs, const void* d_values_in, int8_t* d_actions,
                  void* d_actions_real, void* d_values_out, void* stream);

/* C51's categorical projection of a replay minibatch (lib/agents/c51/c51.py:175-190) in one kernel: for sample i the row
 * m_prob[action[i]][i][:] receives the distribution d_z_opt[i][:] (the target network's distribution of the optimal next
 * action) shifted by reward + gamma * z and projected onto the fixed support of n_atoms atoms on [v_min, v_max] (terminal
 * samples collapse to one point); every other row of m_prob[.][i][:] is zeroed.  Arithmetic and accumulation order are the
 * reference's (including floor/ceil dropping the mass of integer bj).  d_z_opt real [batch][n_atoms], d_action int8
 * [batch], d_reward real [batch], d_done u8 [batch], d_m_prob real [n_actions][batch][n_atoms]. */
int ah_c51_project(ah_env* env, int64_t batch, int n_actions, int n_atoms, double v_min, double v_max, double gamma,
                   const void* d_z_opt, const int8_t* d_action, const void* d_reward, const uint8_t* d_done, void* d_m_prob,
                   void* stream);

/* Compact, lossless host-transfer encoding of one step's results (for callers that consume them on the HOST, where
 * PCIe, not the kernel, is the bound).  Inputs are the device outputs of ah_step:
 *   reward real [K][n] (integer-valued, -4..6: lib/rewards.py:118-142), done u8 [K][n], obs_next real [n][8].
 * Outputs (device, then copied to the host by the caller):
 *   events   u8    [K][n]   low nibble = reward + 4, bit 4 = done
 *   obs_pos  int16 [n][4]   the int()-truncated agent / puck locations of the observation (airhockey.py:136-147)
 *   obs_vel  f32   [n][4]   agent / puck velocities, rounded to binary32
 * 33 bytes instead of 72 per env for K = 8.  d_reward, d_done and d_events may all be NULL: only the observation is
 * packed then (the packed rollout mode already produces its per-frame results as one byte, ah_step_io.events). */
int ah_pack_host(ah_env* env, int K, const void* d_reward, const uint8_t* d_done, const void* d_obs_next,
                 uint8_t* d_events, int16_t* d_obs_pos, float* d_obs_vel, void* stream);

/* Copy the int64[AH_NSTATS] statistics block to a caller device buffer, optionally clearing it.  ah_stats reads bank 0
 * (the default of ah_step_io.stats_bank); ah_stats_bank reads the given bank (0 or 1). */
int ah_stats(ah_env* env, int64_t* d_out, int clear, void* stream);
int ah_stats_bank(ah_env* env, int bank, int64_t* d_out, int clear, void* stream);

/* The per-rollout exchange of the statistics vector, fused over NVLink peer memory for ranks on
 * ONE box (one process per GPU): ah_stats_allreduce = ah_stats_bank's copy + clear AND the sum over all ranks in a single
 * one-warp kernel -- every rank stores its vector straight into every rank's receive buffer (P2P), flags it, and spins
 * until all contributions of the current exchange have landed; no NCCL launch, every rank gets the same sum.  It is a
 * collective: every rank must call it equally often (a rank that never shows up makes the others report -1 in
 * d_out[AH_NSTATS - 1] after ~2 s instead of hanging).  Set-up, once: each rank exports a handle of its receive buffer
 * (ah_peer_export: cudaIpcGetMemHandle), the AH_IPC_HANDLE_BYTES x world handles are exchanged by the caller (e.g. an
 * all_gather) and attached (ah_peer_attach: cudaIpcOpenMemHandle). */
#define AH_MAX_PEERS 8
#define AH_IPC_HANDLE_BYTES 64
int ah_peer_export(ah_env* env, uint8_t* handle64);
int ah_peer_attach(ah_env* env, int rank, int world, const uint8_t* handles /* [world][AH_IPC_HANDLE_BYTES] */);
int ah_stats_allreduce(ah_env* env, int bank, int64_t* d_out, int clear, void* stream);

/* Global frame counter (Philox action / sampling index), kept on the device so that captured graphs advance it;
 * get copies it (uint64[1]) to a caller device buffer.  (The ring's append counter lives in the ring: ah_ring.appended.) */
int ah_set_counters(ah_env* env, uint64_t frame, void* stream);
int ah_get_counters(ah_env* env, uint64_t* d_out1, void* stream);

/* MemoryBuffer.sample(batch) (lib/buffer.py:47-50: random.sample, i.e. WITHOUT replacement) on the device, in one
 * kernel and without materialising a permutation: sample s is entry pi(s) of the population of
 * min(*ring->appended, capacity) x n Observation tuples, where pi is a keyed bijection of [0, population) (a 6-round
 * Feistel network over the next even power of two, cycle-walked; round keys from Philox4x32-10 keyed on the env's seed
 * with counter `draw`) -- so the `batch` samples are distinct, and a different `draw` gives a different sample.
 * Entry f = age * n + env, age 0 = the newest tuple (the order of appendleft + retreive, lib/buffer.py:27,45).
 * Outputs (device, caller-owned): state / new_state real [batch][8], action int8 [batch] (or real [batch][2] for a
 * continuous ring), reward real [batch], done u8 [batch], index int64 [batch] = f (nullable).  batch must not exceed
 * the population (the reference raises ValueError); entries beyond it wrap around. */
int ah_ring_sample(ah_env* env, const ah_ring* ring, int continuous, int64_t batch, uint64_t draw, void* d_state,
                   void* d_action, void* d_reward, uint8_t* d_done, void* d_new_state, int64_t* d_index, void* stream);

/* Launch geometry of the general step kernels (for reports): blocks, threads per block, SM count.  The packed rollout
 * kernel (AH_ACT_DISCRETE_PACKED4) always launches ceil(env_count / 128) blocks of 128 threads, one env per thread. */
int ah_launch_info(const ah_env* env, int* blocks, int* threads, int* sms);

/* Build self-test: writes a*b+c of a discriminating triple to d_scratch (device double[1]); the value is 0.0
 * iff the library was compiled WITHOUT multiply-add contraction, which the AH_F64 build relies on. */
int ah_selftest_nofma(ah_env* env, double* d_scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AH_B200_H */
