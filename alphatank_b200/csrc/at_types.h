// at_types.h - shared POD types of the CUDA env step (kernel config, SoA state, packed fields).
//
// Data layout in HBM (DESIGN.md "State layout"): structure-of-arrays, env index
// fastest, so that the 32 lanes of a warp (= 32 consecutive envs) touch
// consecutive addresses.  Per-tank / per-bullet-slot arrays are [slot][N].
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AT_HD __host__ __device__ __forceinline__
#define AT_HDN __host__ __device__ __noinline__
#else
#define AT_HD inline
#define AT_HDN
#endif

namespace at {

constexpr int BS = 64;          // envs (threads) per block
constexpr int MAX_TANKS = 8;
constexpr int SLOTS_PER_TANK = 6;
constexpr int CELLS = 121;
constexpr int MAZE_W = 11;
constexpr int PEND_LUT = 300 * 7 + 1;

// tpack: angle[0:9) speed_code[9:11) alive 11 hitting_wall 12 in_buff 13 in_debuff 14 maxb1 15
constexpr uint32_t TP_ALIVE = 1u << 11, TP_HW = 1u << 12, TP_INBUFF = 1u << 13, TP_INDEBUFF = 1u << 14,
                   TP_MAXB1 = 1u << 15;
// bmeta: owner[0:4) angle[4:13) flipx 13 flipy 14 bounces[15:18) cleared 18 distance[19:28) rank[28:31)
//        pending_count[32:48)
constexpr int BM_ANGLE = 4, BM_FLIPX = 13, BM_FLIPY = 14, BM_BOUNCES = 15, BM_CLEARED = 18, BM_DIST = 19,
              BM_RANK = 28, BM_PEND = 32;

struct KCfg {
    int32_t mode, n_tanks, map, terminate_time, buff_on, debuff_on, auto_reset;
    int32_t n_agents, rows, obs_dim, act_rows, n_walls, team_layout, n_zone_vals;
    int32_t team[MAX_TANKS], ctrl[MAX_TANKS], bot_type[MAX_TANKS], agent_tank[MAX_TANKS];
    int32_t bot_tank[MAX_TANKS];       // bot-controlled tanks in config order
    int32_t bot_ord[MAX_TANKS];        // tank -> its ordinal among the bots (0 for agents)
    int32_t n_bots;
    int32_t first_bot_k;               // position in the acting order of the first deciding bot
    uint32_t los_hoist;                // smart bots whose line-of-sight query can run at that point (pooled)
    uint32_t team_members[MAX_TANKS];  // bit j set: tank j is on tank i's team (incl. i)
    // observation row layout (floats)
    int32_t self_len, off_ob, off_eb, eb_stride, off_et, et_len, off_walls, off_maze, off_zones, off_flags;
    int32_t row_bulk_ok;  // D*4 % 16 == 0: rows can leave shared memory as one bulk-async copy
    int32_t reward_delta; // at_set_reward_mode: rewards out = increment since the previous step (st.prev_rew)
    double weakness;
    uint64_t seed;
    int64_t env_id_base;
    uint64_t wall_lo, wall_hi;  // static map wall mask (bit r*11+c); per-env arrays are used when map == MAZE
    uint64_t near_lo, near_hi;  // static map: cells that are a wall or touch one (8-neighbourhood)
};

struct DevState {
    int64_t N;
    uint32_t *cnt, *rng_ctr, *tick, *tstep, *epstep, *zones, *prev_act;
    int32_t *change_time;          // [n][N]
    uint64_t *wall_lo, *wall_hi;   // [N], only when map == MAZE
    double *tx, *ty, *trew;        // [n][N]
    uint32_t *tpack, *thits;       // [n][N]
    // info latch of the episode an auto-reset replaced (at_get_terminal_info): valid while term_tick[e] == tick[e]
    uint32_t *term_tick, *term_winner;  // [N]  (term_winner: first alive tank + 1, 0 = nobody)
    uint32_t *term_hits;           // [n][N] thits at the terminal tick
    int32_t *tshot;                // [n][N]
    uint32_t *bpack;               // [n][N] bot FSM
    double *bf0, *bf1;             // [n][N] bot doubles (smart: last_x,last_y; dodge: dodge_angle)
    double *bx, *by;               // [6n][N]
    uint64_t *bmeta;               // [6n][N]
    const double *trig;            // [361][2] cos, sin of math.radians(a)
    const double *corn;            // [361][4] pygame-rotated corner offsets (c0x, c0y, c1x, c1y)
    const double *pend;            // [PEND_LUT] k-fold accumulation of BULLET_AWAY_PENALTY
    const uint8_t *udelta;         // [361] ulp offsets of the numpy-normalised bullet direction from (cos a, -sin a)
    const uint16_t *bfs_tab;       // [121*121] all-pairs bfs_first_step of a FIXED map: len | dir << 8 (null for mazes)
    float *prev_rew;               // [rows][N] rewards returned by the previous step (AT_REWARD_DELTA only)
    const double *atab;            // [361][5] K-degree angles with 106-bit sin / cos (at_sim.h atan2_refine)
    const double *dyn;             // [1] parameters a caller may change between steps (at_set_weakness): read from memory,
                                   //     so a captured CUDA graph sees the new value.  dyn[0] = bot weakness
};

// per-block working set of the tanks / bots (shared memory on the GPU)
struct BlockMem {
    double *trig;                  // [722] copy of st.trig
    double *tx, *ty, *trew;        // [n][BS]
    double *bf0, *bf1;             // [n][BS]
    uint32_t *tpack, *bpack;       // [n][BS]
    uint8_t *tlive;                // [n][BS] live bullets per owner (<= 6)
    uint32_t *nearlo, *nearhi;     // [n][BS] bullet slots that may be within 100 px of the tank when it acts (hi: slots >= 32)
    int32_t *tshot;                // [n][BS]
    const uint8_t *udelta;         // [361] copy of st.udelta
    uint8_t *los_q, *los_r;        // [warps][256] pooled line-of-sight job queue / answers
};

}  // namespace at
