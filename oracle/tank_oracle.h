/* tank_oracle.h - CPU ORACLE for the alphaTank environment step.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, scalar, array-of-structs
 * restatement of the reference's Python algorithm (georgong/alphaTank, env/),
 * written to be compared against (a) traces recorded from the unmodified
 * reference (tests/golden/, tests/refharness.py) and (b) the CUDA kernels in
 * alphatank_b200/csrc/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product never does.
 *
 * Parity status: PINNED against the reference run in the build container
 * (tests/test_reference_pin.py live, tests/test_oracle_golden.py on committed
 * fixtures).  The reference itself ships no tests or golden vectors
 * (SURVEY.md 8c); pygame 2.6.1 Rect/Vector2 semantics are restated in
 * tests/refstubs/pygame and cannot be cross-checked against a real pygame here.
 */
#ifndef TANK_ORACLE_H
#define TANK_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATO_MAX_TANKS 8
#define ATO_MAX_BULLETS 48 /* MAX_BULLETS(6) per tank, configs/config_basic.py:24 */
#define ATO_CELLS 121      /* MAZEWIDTH*MAZEHEIGHT, configs/config_basic.py:4 */

enum { ATO_MODE_AGENT = 0, ATO_MODE_BOT_AGENT = 1, ATO_MODE_ARENA = 2, ATO_MODE_TEAM = 3 };
enum { ATO_CTRL_AGENT = 0, ATO_CTRL_BOT = 1 };
enum { ATO_BOT_SMART = 0, ATO_BOT_RANDOM = 1, ATO_BOT_AGGRESSIVE = 2, ATO_BOT_DEFENSIVE = 3, ATO_BOT_DODGE = 4,
       ATO_BOT_STATIONARY = 5 };
enum { ATO_MAP_OCTAGON = 0, ATO_MAP_BATTLEFIELD = 1, ATO_MAP_MAZE = 2 };

typedef struct ato_config {
    int32_t mode;           /* ATO_MODE_* */
    int32_t n_tanks;        /* 2 for the 1v1 modes */
    int32_t map;            /* ATO_MAP_* */
    int32_t terminate_time; /* 0 = None (configs/config_basic.py:127) */
    int32_t buff_on, debuff_on;
    int32_t auto_reset;     /* reset done envs inside step(); obs returned is the post-reset obs */
    int32_t aim_raymarch;   /* run the zero-weight Tank._aiming_reward ray-marches (F18) - timing realism only */
    int32_t team[ATO_MAX_TANKS];
    int32_t ctrl[ATO_MAX_TANKS];
    int32_t bot_type[ATO_MAX_TANKS];
    double weakness;        /* gaming_env.py:144 */
} ato_config;

typedef struct ato_tank_state {
    double x, y, reward;
    int32_t angle, speed, alive, hitting_wall, in_buff, in_debuff, max_bullets, last_shot_ms, num_hit, num_be_hit;
} ato_tank_state;

typedef struct ato_bullet_state {
    double x, y, dx, dy, pending;
    int32_t owner, distance, bounces, cleared;
} ato_bullet_state;

typedef struct ato_bot_state {
    /* smart */
    int32_t target, aiming, rot_dir, stuck_timer, has_next, next_r, next_c, pad0;
    double last_x, last_y;
    /* defensive */
    int32_t has_tpos, pad1;
    double tpos_x, tpos_y;
    /* dodge */
    int32_t dodge_timer, dodge_state, has_dodge_angle, pad2;
    double dodge_angle;
    /* random */
    int32_t action_delay, cur_action[3];
} ato_bot_state;

typedef struct ato_env_state {
    int32_t n_bullets, tick, training_step, episode_steps;
    uint32_t rng_ctr;
    int32_t n_walls;
    int32_t zones[8]; /* buff0 x,y buff1 x,y debuff0 x,y debuff1 x,y (pixels) */
    uint8_t maze[ATO_CELLS + 7];
    ato_tank_state tanks[ATO_MAX_TANKS];
    ato_bot_state bots[ATO_MAX_TANKS];
    ato_bullet_state bullets[ATO_MAX_BULLETS];
} ato_env_state;

void *ato_create(const ato_config *cfg, int64_t n_envs, int64_t env_id_base, uint64_t seed);
void ato_destroy(void *h);
int32_t ato_obs_dim(void *h);    /* D, per row */
int32_t ato_obs_rows(void *h);   /* R: n (1v1 wrapper) or A (team wrapper) */
int32_t ato_action_rows(void *h);/* rows of the actions array per env: 2 (1v1), A (team), 0 (arena) */
/* obs: float[N, R*D]; only rows of reset envs are written. mask NULL = all. */
void ato_reset(void *h, const uint8_t *mask, float *obs);
/* actions: int32[N, action_rows, 3]; obs float[N,R*D]; rewards float[N,R]; done uint8[N] */
void ato_step(void *h, const int32_t *actions, float *obs, float *rewards, uint8_t *done);
/* same, envs spread over n_threads host threads (OpenMP) - used by the CPU baseline */
void ato_step_mt(void *h, const int32_t *actions, float *obs, float *rewards, uint8_t *done, int32_t n_threads);
void ato_get_state(void *h, int64_t env, ato_env_state *out);
void ato_set_state(void *h, int64_t env, const ato_env_state *in);
/* info: winner int32[N] (1v1: first alive tank or -1), hits/be_hits int32[N,n] */
void ato_get_info(void *h, int32_t *winner, int32_t *hits, int32_t *be_hits, int32_t *change_time);
/* per-env BFS path length from tank i's cell to tank j's cell (len(path), 0 = None); F20 debug output */
int32_t ato_bfs_len_between(void *h, int64_t env, int32_t i, int32_t j);

/* primitives exposed for known-answer tests */
int32_t ato_bfs_path(const uint8_t *maze121, int32_t sr, int32_t sc, int32_t gr, int32_t gc, int32_t *path_rc);
void ato_generate_maze(uint64_t seed, int64_t env_id, uint32_t *ctr_inout, uint8_t *maze121);
int32_t ato_obb_vs_aabb(double cx, double cy, int32_t angle, int32_t rx, int32_t ry, int32_t rw, int32_t rh);
void ato_corners(double cx, double cy, int32_t angle, double *out8);
uint32_t ato_rng_word(uint64_t seed, int64_t env_id, uint32_t stream, uint32_t ctr);
void ato_trig(int32_t angle, double *cosv, double *sinv);

#ifdef __cplusplus
}
#endif
#endif
