/* alphatank_b200.h - C ABI of the B200-native batched alphaTank environment step.
 *
 * The reference (georgong/alphaTank) is pure in-process Python: one env per
 * process, stepped by
 *     MultiAgentEnv.reset/step          env/gym_env.py:63-103
 *     MultiAgentTeamEnv.reset/step      env/gym_env_multi.py:192-223
 * over GamingENV / GamingTeamENV (env/gaming_env.py:48-386,
 * env/gaming_env_multi.py:48-97).  It has no FFI of its own; this header is
 * the boundary a maintainer binds instead of those two classes' bodies (the
 * ctypes stub is shown in INTEGRATION.md, the shipped one is
 * alphatank_b200/_capi.py).  N independent envs live as structure-of-arrays
 * state in HBM and every entry point below is one (or a few) kernel launches
 * on the caller's stream.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary.
 *   - d_* pointers are DEVICE pointers owned by the caller (e.g. torch tensors);
 *     h_* pointers are HOST pointers.  The env state is owned by the handle.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  All
 *     work is enqueued on it; no entry point synchronises unless it says so.
 *   - every function returns 0 on success or a negative AT_ERR_* code and
 *     records a thread-local message readable through at_last_error().
 *     Nothing throws; nothing is freed across the boundary.
 *   - a handle is single-writer (like a reference env instance).
 */
#ifndef ALPHATANK_B200_H
#define ALPHATANK_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AT_MAX_TANKS 8
#define AT_MAX_BULLETS 48 /* MAX_BULLETS(6) x tanks, configs/config_basic.py:24 */
#define AT_CELLS 121      /* MAZEWIDTH x MAZEHEIGHT, configs/config_basic.py:4 */

#define AT_OK 0
#define AT_ERR_ARG (-1)    /* invalid argument / unsupported config (e.g. mode "human") */
#define AT_ERR_CUDA (-2)   /* CUDA runtime error, see at_last_error() */
#define AT_ERR_NOGPU (-3)  /* no usable sm_100 device: the product has no CPU fallback */

/* which branch of GamingENV.step / GamingTeamENV.step runs */
enum at_mode {
    AT_MODE_AGENT = 0,     /* MultiAgentEnv(mode="agent"): "AI only" branch, gaming_env.py:309-374 */
    AT_MODE_BOT_AGENT = 1, /* MultiAgentEnv(mode="bot_agent", type="train"): gaming_env.py:139-199 */
    AT_MODE_ARENA = 2,     /* bot_arena.py:23-74: both tanks bot-driven, decisions on the pre-tick state */
    AT_MODE_TEAM = 3       /* MultiAgentTeamEnv(game_configs): gaming_env_multi.py:65-97 */
};
enum at_ctrl { AT_CTRL_AGENT = 0, AT_CTRL_BOT = 1 }; /* config_teams.py "mode": "agent" | "bot" */
enum at_bot {                                        /* env/bots/bot_factory.py:7-14 */
    AT_BOT_SMART = 0, AT_BOT_RANDOM = 1, AT_BOT_AGGRESSIVE = 2, AT_BOT_DEFENSIVE = 3, AT_BOT_DODGE = 4,
    AT_BOT_STATIONARY = 5
};
enum at_map {
    AT_MAP_OCTAGON = 0,     /* USE_OCTAGON=True, gaming_env.py:576-578 (40 walls) */
    AT_MAP_BATTLEFIELD = 1, /* USE_OCTAGON=False, gaming_env.py:579-606 (61 walls) */
    AT_MAP_MAZE = 2         /* env/maze.py:7-33 generate_maze per env at every reset (72 walls) */
};

/* The frozen form of configs/config_basic.py + one configs/config_teams.py dict. */
typedef struct at_config {
    int32_t mode;           /* enum at_mode */
    int32_t n_tanks;        /* 2 for the 1v1 modes, 2..8 for team mode */
    int32_t map;            /* enum at_map */
    int32_t terminate_time; /* TERMINATE_TIME, 0 = None (config_basic.py:127) */
    int32_t buff_on;        /* BUFF_ON   (config_basic.py:123) */
    int32_t debuff_on;      /* DEBUFF_ON (config_basic.py:124) */
    int32_t auto_reset;     /* 1: done envs are reset inside at_step and obs holds the post-reset observation */
    int32_t reserved;
    int32_t team[AT_MAX_TANKS];     /* team id per tank, config order */
    int32_t ctrl[AT_MAX_TANKS];     /* enum at_ctrl per tank */
    int32_t bot_type[AT_MAX_TANKS]; /* enum at_bot per bot tank */
    double weakness;                /* GamingENV(weakness=...), gaming_env.py:144 */
} at_config;

/* AoS snapshot of one env, for tests, checkpoints and state injection. */
typedef struct at_tank_state {
    double x, y, reward;
    int32_t angle, speed, alive, hitting_wall, in_buff, in_debuff, max_bullets, last_shot_ms, num_hit, num_be_hit;
} at_tank_state;
typedef struct at_bullet_state {
    double x, y, dx, dy, pending;
    int32_t owner, distance, bounces, cleared;
} at_bullet_state;
typedef struct at_bot_state {
    int32_t target, aiming, rot_dir, stuck_timer, has_next, next_r, next_c, pad0; /* SmartStrategyBot */
    double last_x, last_y;
    int32_t has_tpos, pad1; /* DefensiveBot.target_position */
    double tpos_x, tpos_y;
    int32_t dodge_timer, dodge_state, has_dodge_angle, pad2; /* DodgeBot */
    double dodge_angle;
    int32_t action_delay, cur_action[3]; /* RandomBot */
} at_bot_state;
typedef struct at_env_state {
    int32_t n_bullets, tick, training_step, episode_steps;
    uint32_t rng_ctr;
    int32_t n_walls;
    int32_t zones[8]; /* buff0 x,y buff1 x,y debuff0 x,y debuff1 x,y in pixels (multiples of 70) */
    uint8_t maze[AT_CELLS + 7];
    at_tank_state tanks[AT_MAX_TANKS];
    at_bot_state bots[AT_MAX_TANKS];
    at_bullet_state bullets[AT_MAX_BULLETS];
} at_env_state;

typedef struct at_env at_env; /* opaque handle: one batch of envs on one device */

/* Replaces MultiAgentEnv.__init__ (gym_env.py:11-23) / MultiAgentTeamEnv.__init__ (gym_env_multi.py:13-25).
 * Allocates SoA state for n_envs envs on `device` and runs the constructor's own reset (gaming_env.py:43).
 * Env i draws from the ATRNG stream keyed by (seed, env_id_base + i): a batch can be cut across GPUs
 * without changing any env's trajectory. */
int at_create(const at_config *cfg, int64_t n_envs, int64_t env_id_base, uint64_t seed, int device, at_env **out);
int at_destroy(at_env *env);

/* Geometry of the arrays: obs is float[n_envs][rows*dim], rewards float[n_envs][rows], done uint8[n_envs],
 * actions int32[n_envs][action_rows][3].  rows = n_tanks (1v1 wrapper, gym_env.py:119) or #agent tanks
 * (team wrapper, gym_env_multi.py:239); action_rows = 2 (1v1), #agents (team), 0 (arena). */
int at_obs_dim(const at_env *env);
int at_obs_rows(const at_env *env);
int at_action_rows(const at_env *env);
int at_num_walls(const at_env *env);
int64_t at_num_envs(const at_env *env);

/* Replaces env.reset() (gym_env.py:63-71, gym_env_multi.py:192-199) for the envs whose d_mask byte is
 * non-zero (d_mask NULL = all).  Writes the reset observation rows of those envs into d_obs (if non-NULL). */
int at_reset(at_env *env, const uint8_t *d_mask, float *d_obs, void *stream);

/* Replaces env.step(actions) (gym_env.py:73-103, gym_env_multi.py:202-223): the simulation kernel does bullets
 * pass 1, controllers (agents from d_actions, bots on device), bullets pass 2, rewards (cumulative, F1),
 * done / TERMINATE_TIME bookkeeping and the optional auto-reset; the rows kernel behind it the observation rows. */
int at_step(at_env *env, const int32_t *d_actions, float *d_obs, float *d_rewards, uint8_t *d_done, void *stream);

/* Same call with HOST action / reward / done buffers (pinned memory recommended): copies actions in,
 * steps, copies rewards+done out and synchronises the stream.  Observations stay in HBM (d_obs).
 * The call remembers which host buffers are pinned + mapped (the query is a driver call); a caller that frees
 * such a buffer and may get the same address back for pageable memory calls at_forget_host_buffers first. */
int at_step_host(at_env *env, const int32_t *h_actions, float *d_obs, float *h_rewards, uint8_t *h_done,
                 void *stream);
/* at_step_host with the actions packed one byte each (values 0 .. 2: gym_env.py:26 MultiDiscrete([3, 3, 2])): a quarter of
 * the bytes across PCIe, widened to int32 on the device right behind the copy. */
int at_step_host_u8(at_env *env, const uint8_t *h_actions, float *d_obs, float *h_rewards, uint8_t *h_done, void *stream);
int at_forget_host_buffers(at_env *env);

/* Replaces env._get_observation() (gym_env.py:105-183, gym_env_multi.py:225-304) on the CURRENT state: the
 * observation rows of every env into d_obs, nothing stepped.  (at_step / at_reset already return them.) */
int at_observe(at_env *env, float *d_obs, void *stream);

/* What at_step / at_step_norm / at_step_host write into the rewards array.  AT_REWARD_CUMULATIVE (default): the
 * reference's per-episode running sums (gym_env.py:186-187, gym_env_multi.py:307-308; tank.reward is only zeroed by
 * reset - quirk F1).  AT_REWARD_DELTA: the increment since the previous step's value (the first step of an episode
 * returns its value as it is) - what a learner that sums rewards itself needs; computed inside the step from the
 * previous value kept per env, same float32 arithmetic as subtracting consecutive cumulative outputs. */
enum at_reward_mode { AT_REWARD_CUMULATIVE = 0, AT_REWARD_DELTA = 1 };
int at_set_reward_mode(at_env *env, int mode);

/* info dict of step(): winner int32[n_envs] (gym_env.py:97, -1 = nobody alive), hits / be_hits
 * int32[n_envs][n_tanks] (gym_env_multi.py:220-222), change_time int32[n_envs][n_tanks] (gym_env.py:77-83).
 * Any pointer may be NULL. */
int at_get_info(at_env *env, int32_t *d_winner, int32_t *d_hits, int32_t *d_be_hits, int32_t *d_change_time,
                void *stream);

/* The info dict of the LATEST step as the reference's callers see it.  The reference builds info inside step()
 * (gym_env.py:93-98, gym_env_multi.py:218-222) and its callers reset afterwards (train_ppo_bot.py:121-129,
 * train_ppo_cycle.py:341-360, bot_arena.py:61-70), so at a terminal tick info describes the episode that just ended.
 * With auto_reset = 1 that state is replaced inside at_step; the step latches winner / hits / be_hits of the finished
 * episode first, and this call returns the latched values for exactly the envs whose episode ended in the latest step
 * (d_latched[e] = 1, uint8[n_envs], may be NULL) and the current values (= at_get_info) for all others. */
int at_get_terminal_info(at_env *env, int32_t *d_winner, int32_t *d_hits, int32_t *d_be_hits, int32_t *d_change_time,
                         uint8_t *d_latched, void *stream);

/* Per-tank state columns over the batch, for callers that read env.game_env.tanks[i].<attr> (inference.py:115-116
 * reads .alive every tick): d_out double[AT_TANK_COLS][n_envs] = alive, x, y, angle, reward, num_hit, num_be_hit of
 * tank `tank`.  One small kernel on the caller's stream; nothing is copied to the host. */
#define AT_TANK_COLS 7
int at_get_tank_columns(at_env *env, int tank, double *d_out, void *stream);

/* Snapshot / inject `count` envs starting at `first` (host AoS records).  Synchronous. */
int at_get_state(at_env *env, int64_t first, int64_t count, at_env_state *h_out);
int at_set_state(at_env *env, int64_t first, int64_t count, const at_env_state *h_in);

/* Synthetic uniform action stream used by bench / tests: ATRNG stream 1+a, counter = tick
 * (same function as oracle/atrng.py synthetic_actions).  d_actions int32[n_envs][action_rows][3]. */
int at_synthetic_actions(at_env *env, int64_t tick, int32_t *d_actions, void *stream);

/* Stand-alone primitives (parity tests of env/bfs.py:6-56 and env/maze.py:7-33).
 * at_bfs_batch: d_maze uint8[n_mazes][121]; d_query int32[n][5] = (maze index, start r, c, goal r, c);
 *               d_out int32[n][3] = (len(path) or 0 for None, path[1] row, col or -1).
 * at_generate_mazes: maze i is generated from ATRNG stream (seed, env_id_base + i) starting at counter 0;
 *               d_maze uint8[n][121], d_ctr uint32[n] = draws consumed. */
int at_bfs_batch(int device, const uint8_t *d_maze, const int32_t *d_query, int64_t n, int32_t *d_out, void *stream);
int at_generate_mazes(int device, uint64_t seed, int64_t env_id_base, int64_t n, uint8_t *d_maze, uint32_t *d_ctr,
                      void *stream);

/* at_step that also (d_obs != NULL) or only (d_obs == NULL) returns the observation rows as the reference's trainers
 * feed them to the policies: tick-normalised (see at_tick_normalize) - the normaliser runs on the rows while they are
 * still in shared memory, so the raw rows need not be written to, and read back from, HBM.  d_obs_norm has the layout
 * of d_obs.  Bit-identical to at_step followed by at_tick_normalize. */
int at_step_norm(at_env *env, const int32_t *d_actions, float *d_obs, float *d_obs_norm, float epsilon, float *d_rewards,
                 uint8_t *d_done, void *stream);

/* at_step_norm with the normalised rows written as IEEE fp16 (round to nearest of the same fp32 values): half the
 * observation traffic of a rollout, and the operand type of at_policy_forward16's tensor-core contraction.
 * d_obs_norm16: uint16[N][R * D] fp16 bits; d_obs (raw fp32 rows) may be NULL. */
int at_step_norm16(at_env *env, const int32_t *d_actions, float *d_obs, uint16_t *d_obs_norm16, float epsilon,
                   float *d_rewards, uint8_t *d_done, void *stream);

/* The trainers' per-tick observation normaliser (train_multi_ppo.py:89-91: a fresh RunningMeanStd((A, D)) from
 * models/ppo_utils.py:6-29 is updated with, and applied to, each env's own [rows][dim] block) in one pass over
 * d_obs float[n_envs][rows][dim] -> d_out (may alias d_obs).  epsilon = RunningMeanStd's 1e-4. */
int at_tick_normalize(int device, const float *d_obs, int64_t n_envs, int rows, int dim, float epsilon, float *d_out,
                      void *stream);

/* Categorical sampling of a policy's action heads (models/ppo_utils.py:53-61: Categorical(logits=...).sample() and
 * log_prob per head) in one pass: d_logits[k] is float[n_rows][dims[k]] (n_heads <= 4, dims <= 4, sum <= 8).  Writes
 * d_actions[row * act_stride + k] and d_logprobs[row * lp_stride + k] (strides in elements, so the outputs may be
 * columns of a [n_rows][agents][heads] tensor).  Gumbel-max draw from Philox4x32-10 keyed by seed with counter
 * (row, *d_counter, stream_id); d_counter is a device int64 the caller advances once per tick (nullable: 0). */
int at_sample_heads(int device, const float *const *d_logits, const int *dims, int n_heads, int64_t n_rows, uint64_t seed,
                    const int64_t *d_counter, uint32_t stream_id, int32_t *d_actions, int64_t act_stride,
                    float *d_logprobs, int64_t lp_stride, void *stream);

/* The tail of one policy's inference pass (models/ppo_utils.py:31-62, act_dim [3, 3, 2]) in one kernel: d_z2[k] =
 * float[n_rows][128] pre-activations of the second Linear of the critic (k = 0) and of the three actor heads; d_w3 =
 * float[9][128] / d_b3 = float[9] the output layers' rows (critic, 3 + 3 + 2 logits).  tanh, output layers, then as
 * at_sample_heads: actions -> d_actions[row * act_stride + head], log-probabilities -> d_logprobs[row * lp_stride +
 * head], and the critic's value -> d_values[row * val_stride]. */
int at_policy_tail(int device, const float *const *d_z2, const float *d_w3, const float *d_b3, int64_t n_rows, uint64_t seed,
                   const int64_t *d_counter, uint32_t stream_id, int32_t *d_actions, int64_t act_stride, float *d_logprobs,
                   int64_t lp_stride, float *d_values, int64_t val_stride, void *stream);

/* Everything of one policy's inference pass behind its first layers (models/ppo_utils.py:31-62, act_dim [3, 3, 2], 128
 * hidden units) in two kernels: d_z1 = float[n_rows][512] pre-activations of the four first layers (critic, three
 * heads; bias included), d_w2 = float[4][128][128] / d_b2 = float[4][128] the second layers, d_w3 / d_b3 as in
 * at_policy_tail, d_out9 = float[n_rows][9] scratch (value and logits on return).  tanh, the second layers on the
 * tensor cores (TF32, fp32 accumulation), tanh, the output layers, then sampling as at_sample_heads. */
int at_policy_mid_tail(int device, const float *d_z1, const float *d_w2, const float *d_b2, const float *d_w3, const float *d_b3,
                       float *d_out9, int64_t n_rows, uint64_t seed, const int64_t *d_counter, uint32_t stream_id,
                       int32_t *d_actions, int64_t act_stride, float *d_logprobs, int64_t lp_stride, float *d_values,
                       int64_t val_stride, void *stream);

/* The first layers of one policy's four MLPs (models/ppo_utils.py:31-45: critic + three heads share the input) as one
 * contraction on the tcgen05 tensor cores: d_z1 float[n_rows][512] = x . W1^T + b1, x = n_rows rows of `dim` floats
 * x_row_stride floats apart (so a column slice of the [N][A][D] observation tensor can be passed as it is), d_w1 =
 * float[512][dim], d_b1 = float[512].  TMA operand loads, TF32 multiply with fp32 accumulation in tensor memory - the
 * precision of the cuBLAS TF32 GEMM it replaces.  dim and x_row_stride must be multiples of 4. */
int at_policy_first(int device, const float *d_x, int64_t x_row_stride, int64_t n_rows, int dim, const float *d_w1,
                    const float *d_b1, float *d_z1, void *stream);

/* A policy's whole inference pass in one kernel on the tcgen05 tensor cores (models/ppo_utils.py:31-62 - critic + three
 * action heads, D -> 128 -> 128 -> {1, 3, 3, 2} - as train_multi_ppo.py:96-116 calls get_action_and_value per tank):
 * TMA-loaded observation tiles -> first layers (fp32 accumulators in tensor memory) -> tanh -> second layers -> tanh ->
 * output layers -> log-softmax -> Gumbel-max draw (the Philox stream of at_sample_heads).  Neither the [N, 512]
 * activations nor the logits go to HBM.  One launch serves 1 or 2 policies of n_rows rows each.
 *   x            n_rows rows of `dim` elements, x_row_stride elements apart (a column slice of [N][A][D] as it is):
 *                float (at_policy_forward: TF32 contraction) or fp16 bits (at_policy_forward16)
 *   w1 / b1      [512][dim] in x's type / float[512]: the four first layers side by side (critic, head 0, 1, 2)
 *   w2 / b2      fp16 bits [4][128][128] in BOTH entry points / float[4][128];   w3 / b3: float[9][128] / float[9]
 *   out9         optional float[n_rows][9]: value and the 8 logits
 * chunk_mask: bit c set = K chunk c takes part in the first layers' contraction (0 = all); a chunk is 32 columns for
 * float rows, 64 for fp16 rows.  Columns that are constant for the handle's lifetime (at_obs_static_range) may be left
 * out when their contribution W1[:, cols] . x is added to b1 by the caller - 2v2 rows: 9 of 17 (5 of 9) chunks remain.
 * Precision: 10-bit mantissas on both operands of both hidden layers (TF32 reads of the float rows, fp16 hidden
 * activations rounded to nearest), fp32 accumulation, tanh.approx (relative error 2^-11). */
typedef struct at_policy {
    const void *x; int64_t x_row_stride;      /* float (at_policy_forward) or fp16 bits (at_policy_forward16); stride in elements */
    const void *w1; const float *b1;          /* w1 in x's type, w2 fp16 bits, everything else float */
    const void *w2; const float *b2, *w3, *b3;
    int32_t *actions; int64_t act_stride;
    float *logprobs; int64_t lp_stride;
    float *values; int64_t val_stride;
    float *out9;
    uint32_t stream_id;
} at_policy;
/* d_counter: device int64[2] - [0] is the tick counter the draws are keyed by (seed, row, counter, stream_id), [1] must be
 * 0 (the kernel's finished-CTA count).  advance_counter != 0: the last CTA to finish adds 1 to d_counter[0], so a captured
 * rollout needs no separate kernel to move it on; 0: the caller advances it (or not: a pass whose draws are discarded). */
int at_policy_forward(int device, const at_policy *policies, int n_policies, int64_t n_rows, int dim, uint32_t chunk_mask,
                      uint64_t seed, int64_t *d_counter, int advance_counter, void *stream);

/* The same pass on fp16 rows - x = at_step_norm16's rows, w1 rounded to fp16 by the caller: tcgen05.mma.kind::f16 at
 * twice the TF32 rate and half the operand bytes.  chunk_mask counts chunks of 64 columns here. */
int at_policy_forward16(int device, const at_policy *policies, int n_policies, int64_t n_rows, int dim, uint32_t chunk_mask,
                        uint64_t seed, int64_t *d_counter, int advance_counter, void *stream);

/* Persistent output buffers.  With persistent_buffers != 0 the STEP calls (at_step, at_step_norm, at_step_norm16,
 * at_step_host) write only the columns of a row that can change - [0, off_walls) and [off_zones, D), widened to 16-byte
 * multiples - and leave the wall / maze block of a fixed map (at_obs_static_range: 281 of 528 columns of the 2v2 rows)
 * as it is.  The caller guarantees that every buffer it passes was written in full once before by this handle (at_reset,
 * at_observe, or a step call made before the switch): a ring of rollout slots, the env's own persistent buffer.
 * Resets and at_observe always write whole rows; per-env mazes and row lengths that are not 16-byte multiples ignore
 * the switch.  (The reference rebuilds the wall block of every observation, gym_env_multi.py:283-292.) */
int at_set_rows_mode(at_env *env, int persistent_buffers);

/* Diagnostics: while d_stamps (device, uint64[2][512]) is set, block 0 of at_policy_forward16 records (event id, SM
 * clock) pairs of its MMA-issuing thread ([0]) and of one epilogue thread ([1]); NULL switches them off (profiles/). */
int at_policy_debug_stamps(unsigned long long *d_stamps);

/* Observation columns [lo, hi) of every row that never change while the handle lives (the wall rectangles and the maze
 * block of a fixed map, gym_env_multi.py:283-292); lo == hi for per-env mazes. */
int at_obs_static_range(const at_env *env, int32_t *lo, int32_t *hi);

/* Diagnostics: while d_times (device, uint64[2 * ceil(n_envs / 64)]) is set, every block of the simulation kernel
 * stamps %globaltimer (ns) at its start and end into it; NULL switches the stamps off.  Used to study how the blocks'
 * finishing times spread (profiles/). */
int at_debug_block_times(at_env *env, unsigned long long *d_times);

/* Host copies of the device lookup tables (tests): cos/sin of math.radians(a), a = 0..360 -> double[361][2];
 * pygame Vector2.rotate corner offsets -> double[361][4]. */
int at_get_luts(const at_env *env, double *h_trig, double *h_corner);

/* Curriculum hook of the 1v1 bot_agent trainers (train_ppo_cycle.py:82-92, 296, 332-333; GamingENV(weakness=...),
 * gaming_env.py:144): probability that the scripted opponent's decision is applied on a tick.  Takes effect from the
 * next at_step; 0 <= weakness <= 1. */
int at_set_weakness(at_env *env, double weakness);

int64_t at_launch_count(const at_env *env); /* kernels launched through this handle so far */
const char *at_last_error(void);
const char *at_version(void);

#ifdef __cplusplus
}
#endif
#endif
