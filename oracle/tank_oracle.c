/* tank_oracle.c - CPU ORACLE (test infrastructure, see tank_oracle.h).
 *
 * Scalar restatement of the reference's Python env step.  It deliberately
 * keeps the reference's data structures and control flow (lists of walls that
 * are scanned in order, a FIFO BFS queue, 4-axis SAT, per-object loops); the
 * CUDA kernels use different formulations (cell bitmasks, LUTs, bit-parallel
 * BFS) and are checked against this file.  Each function cites the reference
 * lines it follows.  Compile with -ffp-contract=off: the only fused
 * multiply-adds are the explicit fma() calls that restate numpy/OpenBLAS's
 * 2-element dot products (measured on the reference's numpy 2.3.5 /
 * OpenBLAS 0.3.30: v.dot(w) == fma(v1, w1, v0*w0), M(2x2).dot(v)[r] ==
 * fma(M[r][0], v0, M[r][1]*v1)).
 */
#include "tank_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- constants
 * configs/config_basic.py:3-37, 70-73, 98, 105-107 */
#define MAZEW 11
#define MAZEH 11
#define GRID 70.0
#define EPSILON 0.01
#define ROTATION_SPEED 2
#define BULLET_SPEED 1
#define BULLET_MAX_BOUNCES 6
#define BULLET_MAX_DISTANCE 300
#define MAX_BULLETS 6
#define BULLET_COOLDOWN 300
#define ROTATION_DEGREE 3
#define TANK_SPEED 10
#define TANK_WIDTH 30
#define TANK_HEIGHT 24
#define HIT_PENALTY (-50)
#define OPPONENT_HIT_REWARD 50
#define VICTORY_REWARD 200
#define DODGE_FACTOR 0.01
#define BULLET_AWAY_PENALTY 0.02
#define BULLET_CLOSE_REWARD 0.01
#define DISTANCE_CANCEL_THRESHOLD 100
#define PI 3.141592653589793

/* ------------------------------------------------------------------ ATRNG
 * oracle/atrng.py (Philox4x32-10, counter = (ctr, stream, env_lo, env_hi)) */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

uint32_t ato_rng_word(uint64_t seed, int64_t env_id, uint32_t stream, uint32_t ctr) {
    uint32_t c[4] = {ctr, stream, (uint32_t)((uint64_t)env_id & 0xFFFFFFFFu), (uint32_t)((uint64_t)env_id >> 32)};
    philox4x32_10(c, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32));
    return c[0];
}

typedef struct Rng {
    uint64_t seed;
    int64_t env_id;
    uint32_t ctr;
} Rng;

static uint32_t rng_u32(Rng *r) { return ato_rng_word(r->seed, r->env_id, 0, r->ctr++); }
static int rng_below(Rng *r, int n) { return (int)(((uint64_t)rng_u32(r) * (uint64_t)n) >> 32); }
static int rng_randint(Rng *r, int a, int b) { return a + rng_below(r, b - a + 1); }
static double rng_unit53(Rng *r) {
    uint32_t a = rng_u32(r) >> 5, b = rng_u32(r) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}
static double rng_uniform(Rng *r, double a, double b) { return a + (b - a) * rng_unit53(r); }

/* ------------------------------------------------------- Python arithmetic */
/* CPython float_rem (Objects/floatobject.c): result takes the divisor's sign */
static double py_fmod(double v, double w) {
    double m = fmod(v, w);
    if (m != 0.0) {
        if ((w < 0) != (m < 0)) m += w;
    } else {
        m = copysign(0.0, w);
    }
    return m;
}
/* CPython float_floor_div */
static double py_floordiv(double v, double w) {
    double mod = fmod(v, w);
    double div = (v - mod) / w;
    if (mod != 0.0 && ((w < 0) != (mod < 0))) div -= 1.0;
    double fl;
    if (div != 0.0) {
        fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
    } else {
        fl = copysign(0.0, v / w);
    }
    return fl;
}
static int py_imod(int a, int m) { /* Python int % for m > 0 */
    int r = a % m;
    return r < 0 ? r + m : r;
}
/* math.radians / math.degrees (Modules/mathmodule.c) */
static double py_radians(double x) { return x * (PI / 180.0); }
static double py_degrees(double x) { return x * (180.0 / PI); }
/* numpy: np.linalg.norm([a, b]) = sqrt(v.dot(v)), OpenBLAS ddot tail loop is FMA-contracted */
static double np_dot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }
static double np_norm2(double a0, double a1) { return sqrt(np_dot2(a0, a1, a0, a1)); }

void ato_trig(int32_t angle, double *cosv, double *sinv) {
    double rad = py_radians((double)angle);
    *cosv = cos(rad);
    *sinv = sin(rad);
}

/* ------------------------------------------------- pygame Rect / Vector2
 * tests/refstubs/pygame restates pygame 2.6.1 src_c/rect.c, src_c/math.c */
typedef struct Rect {
    int x, y, w, h;
} Rect;
static Rect mk_rect(double x, double y, double w, double h) {
    Rect r = {(int)x, (int)y, (int)w, (int)h}; /* C (int)double truncation, as pg_IntFromObj */
    return r;
}
static int colliderect(Rect a, Rect b) {
    if (a.w == 0 || a.h == 0 || b.w == 0 || b.h == 0) return 0;
    return a.x < b.x + b.w && a.y < b.y + b.h && a.x + a.w > b.x && a.y + a.h > b.y;
}
typedef struct V2 {
    double x, y;
} V2;
static double v2_dot(V2 a, V2 b) { return (0.0 + a.x * b.x) + a.y * b.y; }
static V2 v2_normalize(V2 a) {
    double ln = sqrt(v2_dot(a, a));
    V2 r = {a.x / ln, a.y / ln};
    return r;
}
static V2 v2_rotate(V2 v, double angle) {
    const double eps = 1e-6;
    double a = fmod(angle, 360.0);
    V2 r;
    if (a < 0) a += 360.0;
    if (fmod(a + eps, 90.0) < 2 * eps) {
        switch ((int)((a + eps) / 90.0)) {
        case 0: case 4: r.x = v.x; r.y = v.y; break;
        case 1: r.x = -v.y; r.y = v.x; break;
        case 2: r.x = -v.x; r.y = -v.y; break;
        default: r.x = v.y; r.y = -v.x; break;
        }
        return r;
    }
    double rad = a * PI / 180.0;
    double s = sin(rad), c = cos(rad);
    r.x = c * v.x - s * v.y;
    r.y = s * v.x + c * v.y;
    return r;
}

/* ------------------------------------------------------------------ types */
typedef struct Bullet {
    double x, y, dx, dy, pending;
    int owner, distance, bounces, cleared;
} Bullet;

typedef struct Tank {
    double x, y, reward;
    int angle, speed, alive, last_alive, hitting_wall, in_buff, in_debuff, max_bullets, last_shot, num_hit, num_be_hit,
        team;
    int aiming_counter;
} Tank;

typedef struct Bot {
    int type;
    int target, aiming, rot_dir, stuck_timer, has_next, next_r, next_c;
    double last_x, last_y;
    int has_tpos;
    double tpos_x, tpos_y;
    int dodge_timer, dodge_state, has_dodge_angle; /* dodge_state: 0 idle, 1 dodging, 2 dodging_wall */
    double dodge_angle;
    int action_delay, cur_action[3];
    int last_path_len;
} Bot;

typedef struct Env {
    uint8_t maze[ATO_CELLS];
    int n_walls, wall_r[ATO_CELLS], wall_c[ATO_CELLS];
    int n_empty, empty_x[ATO_CELLS], empty_y[ATO_CELLS];
    int n_buff, n_debuff;
    int zones[8];
    Tank tanks[ATO_MAX_TANKS];
    Bot bots[ATO_MAX_TANKS];
    int has_bot[ATO_MAX_TANKS];
    Bullet bullets[ATO_MAX_BULLETS + 8];
    int n_bullets;
    Rng rng;
    int tick, training_step, episode_steps;
    int prev_valid, prev_act[ATO_MAX_TANKS][3], change_time[ATO_MAX_TANKS];
    /* harness auto-reset only: info of the episode the latest step ended, as step() returned it before the caller's
     * reset (gym_env.py:93-98 winner, gym_env_multi.py:218-222 hits / be_hits) */
    int term_valid, term_winner, term_hits[ATO_MAX_TANKS], term_be_hits[ATO_MAX_TANKS];
} Env;

typedef struct Handle {
    ato_config cfg;
    int64_t n_envs;
    int n_agents, agent_tank[ATO_MAX_TANKS]; /* team: agent-mode tanks in config order */
    int n_walls_static;
    Env *envs;
} Handle;

/* ------------------------------------------------------------------- maze */
/* env/maze.py:7-33 (randomized Prim).  random.randrange(1, w-1, 2) and
 * random.randint(0, len-1) are served by ATRNG below(n). */
static void generate_maze(Rng *rng, uint8_t *maze) {
    int width = MAZEW, height = MAZEH;
    memset(maze, 1, ATO_CELLS);
    int start_x = 1 + 2 * rng_below(rng, (width - 1 - 1 + 1) / 2);
    int start_y = 1 + 2 * rng_below(rng, (height - 1 - 1 + 1) / 2);
    maze[start_y * MAZEW + start_x] = 0;
    static const int DX[4] = {0, 0, -2, 2}, DY[4] = {-2, 2, 0, 0};
    int wl[4 * ATO_CELLS][4], nw = 0;
#define ADD_WALLS(X, Y)                                                                     \
    for (int d = 0; d < 4; ++d) {                                                           \
        int nx = (X) + DX[d], ny = (Y) + DY[d];                                             \
        if (0 < nx && nx < width - 1 && 0 < ny && ny < height - 1 && maze[ny * MAZEW + nx] == 1) { \
            wl[nw][0] = nx; wl[nw][1] = ny; wl[nw][2] = (X); wl[nw][3] = (Y); ++nw;          \
        }                                                                                   \
    }
    ADD_WALLS(start_x, start_y)
    while (nw > 0) {
        int idx = rng_randint(rng, 0, nw - 1);
        int wx = wl[idx][0], wy = wl[idx][1], px = wl[idx][2], py = wl[idx][3];
        for (int i = idx; i < nw - 1; ++i) memcpy(wl[i], wl[i + 1], sizeof wl[0]); /* list.pop(idx) */
        --nw;
        if (maze[wy * MAZEW + wx] == 1) {
            maze[wy * MAZEW + wx] = 0;
            maze[((wy + py) / 2) * MAZEW + (wx + px) / 2] = 0;
            ADD_WALLS(wx, wy)
        }
    }
#undef ADD_WALLS
}

void ato_generate_maze(uint64_t seed, int64_t env_id, uint32_t *ctr_inout, uint8_t *maze121) {
    Rng r = {seed, env_id, *ctr_inout};
    generate_maze(&r, maze121);
    *ctr_inout = r.ctr;
}

/* env/gaming_env.py:566-617 (the live constructWall; octagon / battlefield),
 * :543-563 for ATO_MAP_MAZE (the shadowed generate_maze variant). */
static void construct_wall(Env *e, const ato_config *cfg) {
    uint8_t *m = e->maze;
    if (cfg->map == ATO_MAP_OCTAGON) {
        for (int r = 0; r < MAZEH; ++r)
            for (int c = 0; c < MAZEW; ++c) m[r * MAZEW + c] = !(r >= 1 && r < MAZEH - 1 && c >= 1 && c < MAZEW - 1);
    } else if (cfg->map == ATO_MAP_BATTLEFIELD) {
        memset(m, 0, ATO_CELLS);
        for (int i = 0; i < MAZEW; ++i) {
            m[0 * MAZEW + i] = 1; m[(MAZEH - 1) * MAZEW + i] = 1; m[i * MAZEW + 0] = 1; m[i * MAZEW + MAZEW - 1] = 1;
        }
        int cx = MAZEW / 2, cy = MAZEH / 2;
        m[cy * MAZEW + cx] = 1; m[(cy - 1) * MAZEW + cx] = 1; m[(cy + 1) * MAZEW + cx] = 1;
        m[cy * MAZEW + cx - 1] = 1; m[cy * MAZEW + cx + 1] = 1;
        static const int cover[16][2] = {{2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {MAZEW - 3, 3}, {MAZEW - 3, 4},
                                         {MAZEW - 3, 5}, {MAZEW - 3, 6}, {MAZEW - 3, 7}, {4, 8}, {5, 8}, {6, 8},
                                         {4, 2}, {5, 2}, {6, 2}};
        for (int i = 0; i < 16; ++i) m[cover[i][0] * MAZEW + cover[i][1]] = 1;
    } else {
        generate_maze(&e->rng, m);
    }
    e->n_walls = e->n_empty = 0;
    for (int r = 0; r < MAZEH; ++r)
        for (int c = 0; c < MAZEW; ++c) {
            if (m[r * MAZEW + c] == 1) {
                e->wall_r[e->n_walls] = r; e->wall_c[e->n_walls] = c; ++e->n_walls;
            } else {
                e->empty_x[e->n_empty] = (int)(c * GRID); e->empty_y[e->n_empty] = (int)(r * GRID); ++e->n_empty;
            }
        }
}

/* ------------------------------------------------------------------- BFS */
/* env/bfs.py:6-56.  Returns len(path) (0 = None) and the cells in path_rc. */
int32_t ato_bfs_path(const uint8_t *grid, int32_t sr, int32_t sc, int32_t gr, int32_t gc, int32_t *path_rc) {
    const int rows = MAZEH, cols = MAZEW;
    if (!(0 <= sr && sr < rows && 0 <= sc && sc < cols)) return 0;
    if (!(0 <= gr && gr < rows && 0 <= gc && gc < cols)) return 0;
    if (grid[sr * cols + sc] == 1 || grid[gr * cols + gc] == 1) return 0;
    uint8_t visited[ATO_CELLS];
    int parent[ATO_CELLS];
    int queue[ATO_CELLS], qh = 0, qt = 0;
    memset(visited, 0, sizeof visited);
    queue[qt++] = sr * cols + sc;
    visited[sr * cols + sc] = 1;
    int found = 0;
    static const int DR[4] = {-1, 1, 0, 0}, DC[4] = {0, 0, -1, 1};
    while (qh < qt && !found) {
        int cur = queue[qh++];
        int r = cur / cols, c = cur % cols;
        if (r == gr && c == gc) { found = 1; break; }
        for (int d = 0; d < 4; ++d) {
            int nr = r + DR[d], nc = c + DC[d];
            if (0 <= nr && nr < rows && 0 <= nc && nc < cols) {
                if (!visited[nr * cols + nc] && grid[nr * cols + nc] == 0) {
                    visited[nr * cols + nc] = 1;
                    parent[nr * cols + nc] = cur;
                    queue[qt++] = nr * cols + nc;
                }
            }
        }
    }
    if (!found) return 0;
    int tmp[ATO_CELLS], n = 0;
    int cur = gr * cols + gc, start = sr * cols + sc;
    while (cur != start) { tmp[n++] = cur; cur = parent[cur]; }
    tmp[n++] = start;
    if (path_rc)
        for (int i = 0; i < n; ++i) {
            path_rc[2 * i] = tmp[n - 1 - i] / cols;
            path_rc[2 * i + 1] = tmp[n - 1 - i] % cols;
        }
    return n;
}

/* ------------------------------------------------------------ tank helpers */
/* env/sprite.py:310-324 get_corners */
static void get_corners(double x, double y, int angle, V2 out[4]) {
    double hw = TANK_WIDTH / 2.0, hh = TANK_HEIGHT / 2.0;
    V2 c[4] = {{-hw, -hh}, {hw, -hh}, {hw, hh}, {-hw, hh}};
    for (int i = 0; i < 4; ++i) {
        V2 r = v2_rotate(c[i], (double)angle);
        out[i].x = x + r.x;
        out[i].y = y + r.y;
    }
}
void ato_corners(double cx, double cy, int32_t angle, double *out8) {
    V2 c[4];
    get_corners(cx, cy, angle, c);
    for (int i = 0; i < 4; ++i) { out8[2 * i] = c[i].x; out8[2 * i + 1] = c[i].y; }
}

/* env/util.py:15-48 */
static void project_polygon(const V2 *corners, V2 axis, double *mn, double *mx) {
    double lo = 0, hi = 0;
    for (int i = 0; i < 4; ++i) {
        double d = v2_dot(corners[i], axis);
        if (i == 0 || d < lo) lo = d;
        if (i == 0 || d > hi) hi = d;
    }
    *mn = lo; *mx = hi;
}
static int is_separating_axis(V2 axis, const V2 *c1, const V2 *c2) {
    double min1, max1, min2, max2;
    project_polygon(c1, axis, &min1, &max1);
    project_polygon(c2, axis, &min2, &max2);
    return max1 < min2 - EPSILON || max2 < min1 - EPSILON;
}
static int obb_vs_aabb(const V2 *obb, Rect r) {
    V2 aabb[4] = {{(double)r.x, (double)r.y}, {(double)(r.x + r.w), (double)r.y},
                  {(double)(r.x + r.w), (double)(r.y + r.h)}, {(double)r.x, (double)(r.y + r.h)}};
    V2 d0 = {obb[1].x - obb[0].x, obb[1].y - obb[0].y};
    V2 d1 = {obb[3].x - obb[0].x, obb[3].y - obb[0].y};
    V2 axes[4] = {v2_normalize(d0), v2_normalize(d1), {1.0, 0.0}, {0.0, 1.0}};
    for (int i = 0; i < 4; ++i)
        if (is_separating_axis(axes[i], obb, aabb)) return 0;
    return 1;
}
int32_t ato_obb_vs_aabb(double cx, double cy, int32_t angle, int32_t rx, int32_t ry, int32_t rw, int32_t rh) {
    V2 c[4];
    get_corners(cx, cy, angle, c);
    Rect r = {rx, ry, rw, rh};
    return obb_vs_aabb(c, r);
}

static Rect wall_rect(const Env *e, int i) { return mk_rect(e->wall_c[i] * GRID, e->wall_r[i] * GRID, GRID, GRID); }
static Rect tank_rect(const Tank *t) {
    return mk_rect(t->x - TANK_WIDTH / 2, t->y - TANK_HEIGHT / 2, TANK_WIDTH, TANK_HEIGHT);
}

/* env/sprite.py:131-190 BulletTrajectory.simulate_complete_trajectory -> will_hit_target */
static int trajectory_will_hit(const Env *e, const ato_config *cfg, double x, double y, double dx, double dy,
                               int owner) {
    int bounces = 0, distance = 0;
    const int owner_team = e->tanks[owner].team;
    for (;;) {
        double next_x = x + dx * BULLET_SPEED, next_y = y + dy * BULLET_SPEED;
        Rect br = mk_rect(next_x, next_y, 5, 5);
        for (int i = 0; i < cfg->n_tanks; ++i) {
            const Tank *t = &e->tanks[i];
            if (t->team != owner_team && t->alive)
                if (colliderect(br, tank_rect(t))) return 1;
        }
        int bounced = 0;
        for (int i = 0; i < e->n_walls; ++i) {
            Rect wr = wall_rect(e, i);
            if (colliderect(wr, br)) {
                Rect rx = mk_rect(x + dx * BULLET_SPEED, y, 5, 5), ry = mk_rect(x, y + dy * BULLET_SPEED, 5, 5);
                int bx = colliderect(wr, rx), by = colliderect(wr, ry);
                if (bx && by) { dx = -dx; dy = -dy; }
                else if (bx) dx = -dx;
                else if (by) dy = -dy;
                ++bounces;
                bounced = 1;
                break;
            }
        }
        if (!bounced) { x = next_x; y = next_y; distance += BULLET_SPEED; }
        if (bounces > BULLET_MAX_BOUNCES || distance > BULLET_MAX_DISTANCE) return 0;
    }
}

/* env/sprite.py:473-495 (zero weight; only run when cfg->aim_raymarch, F18) */
static void aiming_reward(Env *e, const ato_config *cfg, int ti) {
    Tank *t = &e->tanks[ti];
    if (!t->alive || !cfg->aim_raymarch) return;
    double rad = py_radians((double)t->angle);
    double bx = t->x + 10 * cos(rad), by = t->y - 10 * sin(rad);
    if (trajectory_will_hit(e, cfg, bx, by, cos(rad), -sin(rad), ti)) {
        if (++t->aiming_counter >= 17) t->aiming_counter = 0; /* += TRAJECTORY_AIM_REWARD (0) */
    } else {
        t->aiming_counter = 0;
    }
}

/* env/sprite.py:528-549 */
static void tank_rotate(Env *e, const ato_config *cfg, int ti, int direction) {
    Tank *t = &e->tanks[ti];
    if (!t->alive) return;
    t->angle = py_imod(t->angle + direction * ROTATION_SPEED, 360);
    aiming_reward(e, cfg, ti);
}

/* env/sprite.py:613-636 (F8) */
static void check_buff_debuff(Env *e, const ato_config *cfg, int ti) {
    Tank *t = &e->tanks[ti];
    Rect tr = tank_rect(t);
    t->in_buff = 0;
    int brk = 0;
    for (int i = 0; i < e->n_buff; ++i) {
        Rect z = mk_rect(e->zones[2 * i], e->zones[2 * i + 1], GRID * 3.5, GRID * 3.5);
        if (colliderect(tr, z) && cfg->buff_on) { t->max_bullets = 30; t->in_buff = 1; brk = 1; break; }
    }
    if (!brk) t->max_bullets = MAX_BULLETS;
    t->in_debuff = 0;
    brk = 0;
    for (int i = 0; i < e->n_debuff; ++i) {
        Rect z = mk_rect(e->zones[4 + 2 * i], e->zones[4 + 2 * i + 1], GRID * 3.5, GRID * 3.5);
        if (colliderect(tr, z) && cfg->debuff_on) { t->max_bullets = 1; t->in_debuff = 1; brk = 1; break; }
    }
    if (!brk) t->max_bullets = MAX_BULLETS;
}

/* env/sprite.py:579-609; the wall clock is the virtual clock tick*1000//60 (F9) */
static void tank_shoot(Env *e, const ato_config *cfg, int ti) {
    Tank *t = &e->tanks[ti];
    check_buff_debuff(e, cfg, ti);
    if (!t->alive) return;
    int now = e->tick * 1000 / 60;
    int active = 0;
    for (int i = 0; i < e->n_bullets; ++i) active += e->bullets[i].owner == ti;
    if (active >= t->max_bullets) return;
    if (now - t->last_shot < BULLET_COOLDOWN) return;
    double rad = py_radians((double)t->angle);
    Bullet *b = &e->bullets[e->n_bullets++];
    b->x = t->x + 10 * cos(rad);
    b->y = t->y - 10 * sin(rad);
    b->dx = cos(rad);
    b->dy = -sin(rad);
    b->owner = ti;
    b->distance = 0; b->bounces = 0; b->pending = 0.0; b->cleared = 0;
    /* F19: the render-only BulletTrajectory of shoot() is skipped unless timing realism is requested */
    if (cfg->aim_raymarch) (void)trajectory_will_hit(e, cfg, b->x, b->y, b->dx, b->dy, ti);
    t->last_shot = now;
}

/* env/sprite.py:497-525 (F17) */
static void dodge_reward(Env *e, int ti) {
    Tank *t = &e->tanks[ti];
    double reward = 0;
    double rad = py_radians((double)t->angle);
    double tvx = (double)(t->speed * 1) * cos(rad), tvy = (double)(t->speed * 1) * sin(rad); /* util.py:50-55 */
    for (int i = 0; i < e->n_bullets; ++i) {
        const Bullet *b = &e->bullets[i];
        double distance = np_norm2(t->x - b->x, t->y - b->y);
        if (distance >= 100) continue;
        double n = np_norm2(b->dx, b->dy);
        if (n == 0) continue;
        if (e->tanks[b->owner].team == t->team) continue;
        double d0 = b->dx / n, d1 = b->dy / n;
        double p0 = -d1, p1 = d0;
        double projection = np_dot2(tvx, tvy, p0, p1);
        reward += fabs(projection) * DODGE_FACTOR;
    }
    t->reward += reward;
}

/* env/sprite.py:326-357 */
static void tank_move(Env *e, const ato_config *cfg, int ti) {
    Tank *t = &e->tanks[ti];
    if (!t->alive) return;
    double rad = py_radians((double)t->angle);
    double new_x = t->x + t->speed * cos(rad);
    double new_y = t->y - t->speed * sin(rad);
    V2 corners[4];
    get_corners(new_x, new_y, t->angle, corners);
    int hit = 0;
    for (int i = 0; i < e->n_walls && !hit; ++i) hit = obb_vs_aabb(corners, wall_rect(e, i));
    if (!hit) { t->x = new_x; t->y = new_y; t->hitting_wall = 0; }
    else t->hitting_wall = 1;
    /* _stationary_penalty, _control_penalty: zero weight, unobservable state (App. A.1) */
    aiming_reward(e, cfg, ti);
    dodge_reward(e, ti);
}

/* ----------------------------------------------------------------- bullets */
/* env/gaming_env.py:522-541 (1v1, quirk F6) and env/gaming_env_multi.py:433-442 (team) */
static void update_reward_by_bullets(Env *e, const ato_config *cfg, int shooter, int victim) {
    e->tanks[shooter].reward += OPPONENT_HIT_REWARD; /* same-team branch unreachable (F7) */
    e->tanks[victim].reward += HIT_PENALTY;
    if (cfg->mode != ATO_MODE_TEAM) {
        int all_same = 1;
        for (int i = 1; i < cfg->n_tanks; ++i) all_same &= (e->tanks[i].alive == e->tanks[0].alive);
        if (all_same)
            for (int i = 0; i < cfg->n_tanks; ++i)
                if (e->tanks[i].alive) {
                    int max_steps = 1000;
                    int num = max_steps - e->episode_steps;
                    double q = (double)num / (double)max_steps;
                    double time_bonus = (q > 0) ? 5 * q : 0.0; /* 5 * max(0, q) */
                    e->tanks[i].reward += VICTORY_REWARD * time_bonus;
                }
    }
}

/* env/sprite.py:27-57 */
static void bullet_reward_logic(Env *e, const ato_config *cfg, Bullet *b) {
    double n = np_norm2(b->dx, b->dy);
    if (n == 0) return;
    double d0 = b->dx / n, d1 = b->dy / n;
    int owner_team = e->tanks[b->owner].team;
    for (int i = 0; i < cfg->n_tanks; ++i) {
        Tank *t = &e->tanks[i];
        if (!t->alive || t->team == owner_team) continue;
        double tx = t->x - b->x, ty = t->y - b->y;
        double dist = np_norm2(tx, ty);
        if (dist > BULLET_MAX_DISTANCE) continue;
        double u0 = tx / dist, u1 = ty / dist; /* dist == 0 -> nan, comparisons false */
        double dotp = np_dot2(d0, d1, u0, u1);
        if (dotp > 0.9) e->tanks[b->owner].reward += BULLET_CLOSE_REWARD;
        else if (dotp < 0.1) { if (!b->cleared) b->pending += BULLET_AWAY_PENALTY; }
        if (dist < DISTANCE_CANCEL_THRESHOLD) { b->pending = 0.0; b->cleared = 1; }
    }
}

/* env/sprite.py:59-105.  Returns 1 if the bullet must be removed. */
static int bullet_move(Env *e, const ato_config *cfg, Bullet *b) {
    double next_x = b->x + b->dx * BULLET_SPEED, next_y = b->y + b->dy * BULLET_SPEED;
    Rect br = mk_rect(next_x, next_y, 5, 5);
    bullet_reward_logic(e, cfg, b);
    for (int i = 0; i < e->n_walls; ++i) {
        Rect wr = wall_rect(e, i);
        if (colliderect(wr, br)) {
            Rect rx = mk_rect(b->x + b->dx * BULLET_SPEED, b->y, 5, 5);
            Rect ry = mk_rect(b->x, b->y + b->dy * BULLET_SPEED, 5, 5);
            int bx = colliderect(wr, rx), by = colliderect(wr, ry);
            if (bx && by) { b->dx = -b->dx; b->dy = -b->dy; }
            else if (bx) b->dx = -b->dx;
            else if (by) b->dy = -b->dy;
            b->bounces += 1;
            break;
        }
    }
    b->x = next_x; b->y = next_y;
    b->distance += BULLET_SPEED;
    int owner_team = e->tanks[b->owner].team;
    for (int i = 0; i < cfg->n_tanks; ++i) {
        Tank *t = &e->tanks[i];
        if (t->alive > 0 && t->team != owner_team) {
            if (colliderect(br, tank_rect(t))) {
                if (cfg->terminate_time == 0) t->alive = 0;
                e->tanks[b->owner].num_hit += 1;
                t->num_be_hit += 1;
                b->pending = 0.0; b->cleared = 1;
                update_reward_by_bullets(e, cfg, b->owner, i);
                return 1;
            }
        }
    }
    if (b->bounces >= BULLET_MAX_BOUNCES || b->distance >= BULLET_MAX_DISTANCE) {
        if (!b->cleared && b->pending > 0) { e->tanks[b->owner].reward -= b->pending; b->pending = 0.0; }
        return 1;
    }
    return 0;
}

/* `for bullet in self.bullets[:]: bullet.move()` with in-place removal (order preserved) */
static void bullets_pass(Env *e, const ato_config *cfg) {
    int n = e->n_bullets, w = 0;
    /* removal of bullet r must not disturb what later bullets see: later bullets only read tanks */
    for (int r = 0; r < n; ++r) {
        Bullet b = e->bullets[r];
        int rm = bullet_move(e, cfg, &b);
        if (!rm) e->bullets[w++] = b;
    }
    e->n_bullets = w;
}

/* -------------------------------------------------------------------- bots */
static int find_nearest_opponent(const Env *e, const ato_config *cfg, int ti, double *min_dist_out) {
    double min_dist = INFINITY;
    int nearest = -1;
    const Tank *me = &e->tanks[ti];
    for (int i = 0; i < cfg->n_tanks; ++i) {
        const Tank *t = &e->tanks[i];
        if (t->team != me->team && t->alive) {
            double ddx = t->x - me->x, ddy = t->y - me->y;
            double dist = sqrt(ddx * ddx + ddy * ddy);
            if (dist < min_dist) { min_dist = dist; nearest = i; }
        }
    }
    if (min_dist_out) *min_dist_out = min_dist;
    return nearest;
}

/* signed angle difference used by every bot: strategy_bot.py:40-56, simple_bots.py:72-84,202-214 */
static double angle_diff_to(const Tank *me, double tx, double ty) {
    double dx = tx - me->x, dy = ty - me->y;
    double target_angle = py_fmod(py_degrees(atan2(-dy, dx)), 360.0);
    int current_angle = py_imod(me->angle, 360);
    return py_fmod(target_angle - current_angle + 180, 360.0) - 180;
}
static int aim_at_target(const Tank *me, double tx, double ty, double thr) {
    double d = angle_diff_to(me, tx, ty);
    if (fabs(d) < thr) return 0;
    return d > 0 ? -1 : 1;
}
static void grid_position(const Tank *t, int *r, int *c) { /* sprite.py:692-697 */
    *r = (int)py_floordiv(t->y, GRID);
    *c = (int)py_floordiv(t->x, GRID);
}

/* env/bots/strategy_bot.py:75-203 */
static void smart_get_action(Env *e, const ato_config *cfg, int ti, int act[3]) {
    Bot *b = &e->bots[ti];
    Tank *me = &e->tanks[ti];
    /* update_state */
    if (b->target < 0 || !e->tanks[b->target].alive) {
        b->aiming = 0;
        b->target = find_nearest_opponent(e, cfg, ti, NULL);
    }
    if (b->target >= 0) {
        int mr, mc, tr, tc, path[2 * ATO_CELLS];
        grid_position(me, &mr, &mc);
        grid_position(&e->tanks[b->target], &tr, &tc);
        int len = ato_bfs_path(e->maze, mr, mc, tr, tc, path);
        b->last_path_len = len;
        if (len > 1) { b->has_next = 1; b->next_r = path[2]; b->next_c = path[3]; }
        else b->has_next = 0;
    }
    {
        double ddx = me->x - b->last_x, ddy = me->y - b->last_y;
        if (sqrt(ddx * ddx + ddy * ddy) < 1) b->stuck_timer += 1;
        else b->stuck_timer = 0;
        b->last_x = me->x; b->last_y = me->y;
    }
    act[0] = 1; act[1] = 1; act[2] = 0;
    if (!b->aiming) {
        b->target = find_nearest_opponent(e, cfg, ti, NULL);
        if (b->target >= 0) b->aiming = 1;
    }
    if (b->aiming && b->target >= 0) {
        double rad = py_radians((double)me->angle);
        double bx = me->x + 10 * cos(rad), by = me->y - 10 * sin(rad);
        int los = trajectory_will_hit(e, cfg, bx, by, cos(rad), -sin(rad), ti);
        const Tank *tg = &e->tanks[b->target];
        if (los) {
            int rotation = aim_at_target(me, tg->x, tg->y, 10);
            if (rotation != 0) act[0] = rotation > 0 ? 2 : 0;
            double ddx = tg->x - me->x, ddy = tg->y - me->y;
            double dist = sqrt(ddx * ddx + ddy * ddy);
            if (dist > 200) act[1] = 2;
            else if (dist < 20) act[1] = 0;
            if (rotation == 0) act[2] = 1;
        } else if (b->has_next) {
            double tx = b->next_c * GRID + (GRID / 2), ty = b->next_r * GRID + (GRID / 2);
            int rotation = aim_at_target(me, tx, ty, 10);
            if (rotation != 0) act[0] = rotation > 0 ? 2 : 0;
            act[1] = (aim_at_target(me, tx, ty, 10) == 0) ? 2 : 1; /* get_movement_to_cell */
        }
    }
    if (b->stuck_timer > 20) {
        act[1] = 2;
        act[0] = b->rot_dir > 0 ? 0 : 2;
        b->rot_dir *= -1;
        b->stuck_timer = 0;
    }
}

/* env/bots/simple_bots.py:32-46 */
static void random_get_action(Env *e, int ti, int act[3]) {
    Bot *b = &e->bots[ti];
    if (b->action_delay <= 0) {
        b->cur_action[0] = rng_randint(&e->rng, 0, 2);
        b->cur_action[1] = rng_randint(&e->rng, 0, 2);
        b->cur_action[2] = rng_randint(&e->rng, 0, 1);
        b->action_delay = rng_randint(&e->rng, 10, 30);
    }
    b->action_delay -= 1;
    act[0] = b->cur_action[0]; act[1] = b->cur_action[1]; act[2] = b->cur_action[2];
}

/* env/bots/simple_bots.py:86-125 */
static void aggressive_get_action(Env *e, const ato_config *cfg, int ti, int act[3]) {
    const Tank *me = &e->tanks[ti];
    double dist;
    int target = find_nearest_opponent(e, cfg, ti, &dist);
    act[0] = 1; act[1] = 1; act[2] = 0;
    if (target < 0) return;
    const Tank *tg = &e->tanks[target];
    double angle_diff = fabs(angle_diff_to(me, tg->x, tg->y));
    int rotation = aim_at_target(me, tg->x, tg->y, 15);
    act[0] = rotation > 0 ? 2 : (rotation < 0 ? 0 : 1);
    const double min_distance = GRID * 1;
    if (angle_diff < 15 && dist <= min_distance) act[2] = 1;
    if (dist > min_distance) { if (angle_diff < 45) act[1] = 2; }
    else act[1] = 1;
}

/* env/bots/simple_bots.py:151-200, 216-292 */
static void defensive_get_action(Env *e, const ato_config *cfg, int ti, int act[3]) {
    Bot *b = &e->bots[ti];
    const Tank *me = &e->tanks[ti];
    act[0] = 1; act[1] = 1; act[2] = 0;
    int target = find_nearest_opponent(e, cfg, ti, NULL);
    if (target < 0) return;
    const Tank *tg = &e->tanks[target];
    Rect tr = tank_rect(me);
    int in_buff = 0;
    for (int i = 0; i < e->n_buff && !in_buff; ++i)
        in_buff = colliderect(tr, mk_rect(e->zones[2 * i], e->zones[2 * i + 1], GRID * 3.5, GRID * 3.5));
    if (in_buff) {
        int rot = aim_at_target(me, tg->x, tg->y, 20);
        act[0] = rot > 0 ? 2 : (rot < 0 ? 0 : 1);
        if (fabs(angle_diff_to(me, tg->x, tg->y)) < 20) act[2] = 1;
    } else {
        if (!b->has_tpos) {
            double min_dist = INFINITY;
            for (int i = 0; i < e->n_buff; ++i) {
                double bw = GRID * 3.5, bh = GRID * 3.5;
                double mx = TANK_WIDTH * 0.75, my = TANK_HEIGHT * 0.75;
                double sx0 = e->zones[2 * i] + mx, sx1 = e->zones[2 * i] + bw - mx;
                double sy0 = e->zones[2 * i + 1] + my, sy1 = e->zones[2 * i + 1] + bh - my;
                double sx = (sx0 + sx1) / 2, sy = (sy0 + sy1) / 2;
                double ddx = sx - me->x, ddy = sy - me->y;
                double dist = sqrt(ddx * ddx + ddy * ddy);
                if (dist < min_dist) { min_dist = dist; b->has_tpos = 1; b->tpos_x = sx; b->tpos_y = sy; }
            }
        }
        if (b->has_tpos) {
            double diff = fabs(angle_diff_to(me, b->tpos_x, b->tpos_y));
            int rot = aim_at_target(me, b->tpos_x, b->tpos_y, 20);
            act[0] = rot > 0 ? 2 : (rot < 0 ? 0 : 1);
            if (diff < 45) act[1] = 2;
        }
    }
}

/* env/bots/simple_bots.py:309-329 */
static int dodge_will_hit_wall(const Env *e, const Tank *me, double angle) {
    double rad = py_radians(angle);
    double wcd = GRID * 1.5;
    double fx = me->x + cos(rad) * wcd, fy = me->y - sin(rad) * wcd;
    Rect fr = mk_rect(fx - TANK_WIDTH / 2, fy - TANK_HEIGHT / 2, TANK_WIDTH, TANK_HEIGHT);
    for (int i = 0; i < e->n_walls; ++i)
        if (colliderect(wall_rect(e, i), fr)) return 1;
    return 0;
}
/* env/bots/simple_bots.py:359-402 */
static double dodge_calc_angle(Env *e, const Tank *me, double thx, double thy) {
    double v0 = thx - me->x, v1 = thy - me->y;
    double n = np_norm2(v0, v1);
    if (n > 0) {
        n = np_norm2(v0, v1);
        v0 = v0 / n; v1 = v1 / n;
        int base[2] = {90, -90};
        { /* random.shuffle(base_angles) */
            int j = rng_below(&e->rng, 2);
            int t = base[1]; base[1] = base[j]; base[j] = t;
        }
        for (int bi = 0; bi < 2; ++bi)
            for (int k = 0; k < 3; ++k) {
                double var = rng_uniform(&e->rng, -(90 - 45), 90 - 120);
                double dodge_angle = base[bi] + var;
                double theta = py_radians(dodge_angle);
                double c = cos(theta), s = sin(theta);
                double r0 = fma(c, v0, (-s) * v1); /* rotation_matrix.dot(threat_vector), OpenBLAS dgemv */
                double r1 = fma(s, v0, c * v1);
                double final_angle = py_fmod(py_degrees(atan2(-r1, r0)), 360.0);
                if (!dodge_will_hit_wall(e, me, final_angle)) return final_angle;
            }
        return (double)me->angle;
    }
    return (double)me->angle;
}
/* env/bots/simple_bots.py:331-357, 404-455 */
static void dodge_get_action(Env *e, const ato_config *cfg, int ti, int act[3]) {
    Bot *b = &e->bots[ti];
    const Tank *me = &e->tanks[ti];
    const double radius = GRID * 4;
    act[0] = 1; act[1] = 1; act[2] = 1;
    double min_dist = INFINITY, px = 0, py = 0;
    int have = 0;
    for (int i = 0; i < cfg->n_tanks; ++i) {
        const Tank *t = &e->tanks[i];
        if (t->team != me->team && t->alive) {
            double ddx = t->x - me->x, ddy = t->y - me->y;
            double dist = sqrt(ddx * ddx + ddy * ddy);
            if (dist < min_dist && dist < radius) { min_dist = dist; px = t->x; py = t->y; have = 1; }
        }
    }
    for (int i = 0; i < e->n_bullets; ++i) {
        const Bullet *bl = &e->bullets[i];
        if (e->tanks[bl->owner].team != me->team) {
            double ddx = bl->x - me->x, ddy = bl->y - me->y;
            double dist = sqrt(ddx * ddx + ddy * ddy);
            if (dist < min_dist && dist < radius) { min_dist = dist; px = bl->x; py = bl->y; have = 1; }
        }
    }
    if (!have) { b->dodge_state = 0; b->dodge_timer = 0; b->has_dodge_angle = 0; return; }
    if (b->dodge_timer <= 0 || !b->has_dodge_angle) {
        b->dodge_angle = dodge_calc_angle(e, me, px, py);
        b->has_dodge_angle = 1;
        b->dodge_timer = 10;
        b->dodge_state = 1;
    }
    if (b->dodge_state == 1) {
        if (dodge_will_hit_wall(e, me, b->dodge_angle)) {
            b->dodge_angle = dodge_calc_angle(e, me, px, py);
            b->dodge_state = 2;
        }
        int current_angle = py_imod(me->angle, 360);
        double diff = py_fmod(b->dodge_angle - current_angle + 180, 360.0) - 180;
        int rotation = fabs(diff) < 5 ? 0 : (diff > 0 ? -1 : 1);
        act[0] = rotation > 0 ? 2 : (rotation < 0 ? 0 : 1);
        if (fabs(diff) < 45) act[1] = 2;
    }
    b->dodge_timer -= 1;
}

static void bot_get_action(Env *e, const ato_config *cfg, int ti, int act[3]) {
    switch (e->bots[ti].type) {
    case ATO_BOT_SMART: smart_get_action(e, cfg, ti, act); break;
    case ATO_BOT_RANDOM: random_get_action(e, ti, act); break;
    case ATO_BOT_AGGRESSIVE: aggressive_get_action(e, cfg, ti, act); break;
    case ATO_BOT_DEFENSIVE: defensive_get_action(e, cfg, ti, act); break;
    case ATO_BOT_DODGE: dodge_get_action(e, cfg, ti, act); break;
    default: act[0] = 1; act[1] = 1; act[2] = 0; break; /* stationary, simple_bots.py:465-467 */
    }
}

static void bot_init(Env *e, int ti, int type) {
    Bot *b = &e->bots[ti];
    memset(b, 0, sizeof *b);
    b->type = type;
    b->target = -1;
    b->rot_dir = 1;
    b->last_x = e->tanks[ti].x; /* strategy_bot.py:15 */
    b->last_y = e->tanks[ti].y;
    b->cur_action[0] = 1; b->cur_action[1] = 1; b->cur_action[2] = 0; /* simple_bots.py:14 */
    e->has_bot[ti] = 1;
}

/* ---------------------------------------------------------- reset and step */
/* env/gaming_env.py:48-66, 509-520; env/gaming_env_multi.py:48-63, 195-238; env/sprite.py:216-272;
 * wrapper reset env/gym_env.py:63-71, env/gym_env_multi.py:192-199 */
static void env_reset(Handle *h, Env *e) {
    const ato_config *cfg = &h->cfg;
    construct_wall(e, cfg);
    for (int i = 0; i < cfg->n_tanks; ++i) {
        int idx = rng_below(&e->rng, e->n_empty); /* np.random.choice(range(len(empty_space))) */
        Tank *t = &e->tanks[i];
        memset(t, 0, sizeof *t);
        t->x = e->empty_x[idx] + GRID / 2;
        t->y = e->empty_y[idx] + GRID / 2;
        t->angle = rng_randint(&e->rng, 0, 360);
        t->alive = 1; t->last_alive = 1;
        t->max_bullets = MAX_BULLETS;
        t->team = cfg->team[i];
        e->has_bot[i] = 0;
    }
    e->n_bullets = 0;
    if (cfg->mode == ATO_MODE_BOT_AGENT) bot_init(e, 0, cfg->bot_type[0]);
    if (cfg->mode == ATO_MODE_TEAM || cfg->mode == ATO_MODE_ARENA)
        for (int i = 0; i < cfg->n_tanks; ++i)
            if (cfg->ctrl[i] == ATO_CTRL_BOT) bot_init(e, i, cfg->bot_type[i]);
    e->n_buff = cfg->buff_on ? 2 : 0;
    e->n_debuff = cfg->debuff_on ? 2 : 0;
    memset(e->zones, 0, sizeof e->zones);
    for (int z = 0; z < 2; ++z) {
        if ((z == 0 && !cfg->buff_on) || (z == 1 && !cfg->debuff_on)) continue;
        int i = rng_below(&e->rng, e->n_empty); /* random.sample(empty_space, 2) */
        int j = rng_below(&e->rng, e->n_empty - 1);
        if (j >= i) ++j;
        e->zones[4 * z + 0] = e->empty_x[i]; e->zones[4 * z + 1] = e->empty_y[i];
        e->zones[4 * z + 2] = e->empty_x[j]; e->zones[4 * z + 3] = e->empty_y[j];
    }
    e->episode_steps = 0;
    e->prev_valid = 0;
    memset(e->change_time, 0, sizeof e->change_time);
}

/* rot/mov/shoot application shared by the branches of GamingENV.step and Tank.take_action */
static void apply_action(Env *e, const ato_config *cfg, int ti, const int a[3], int fwd_code, int back_code,
                         int stop_speed) {
    Tank *t = &e->tanks[ti];
    if (a[0] == 0) tank_rotate(e, cfg, ti, ROTATION_DEGREE);
    else if (a[0] == 2) tank_rotate(e, cfg, ti, -ROTATION_DEGREE);
    if (a[1] == fwd_code) t->speed = TANK_SPEED;
    else if (a[1] == back_code) t->speed = -TANK_SPEED;
    else t->speed = stop_speed;
    if (a[2] == 1) tank_shoot(e, cfg, ti);
    tank_move(e, cfg, ti);
}

static void game_step(Handle *h, Env *e, const int32_t *actions) {
    const ato_config *cfg = &h->cfg;
    int act[3];
    int arena_act[ATO_MAX_TANKS][3];
    e->tick += 1; /* virtual clock (F9) */
    if (cfg->mode == ATO_MODE_ARENA) /* bot_arena.py:51-52: both bots decide on the pre-tick state */
        for (int i = 0; i < cfg->n_tanks; ++i) bot_get_action(e, cfg, i, arena_act[i]);
    if (cfg->mode != ATO_MODE_TEAM) e->episode_steps += 1; /* gaming_env.py:69 */
    bullets_pass(e, cfg);                                  /* gaming_env.py:71-72 */
    switch (cfg->mode) {
    case ATO_MODE_BOT_AGENT: { /* gaming_env.py:139-199 (F14, F15) */
        bot_get_action(e, cfg, 0, act);
        int bot_acts = rng_unit53(&e->rng) < cfg->weakness;
        if (!bot_acts) { act[0] = 1; act[1] = 1; act[2] = 0; }
        apply_action(e, cfg, 0, act, 2, 0, 0);
        const int32_t *a = actions + 3; /* actions[1] when type == "train" */
        int aa[3] = {a[0], a[1], a[2]};
        apply_action(e, cfg, 1, aa, 0, 2, 1);
        break;
    }
    case ATO_MODE_TEAM: { /* gaming_env_multi.py:82-85, controller.py:81-113, sprite.py:638-657 */
        for (int k = 0; k < h->n_agents; ++k) {
            int aa[3] = {actions[3 * k], actions[3 * k + 1], actions[3 * k + 2]};
            apply_action(e, cfg, h->agent_tank[k], aa, 2, 0, 0);
        }
        for (int i = 0; i < cfg->n_tanks; ++i)
            if (cfg->ctrl[i] == ATO_CTRL_BOT) {
                bot_get_action(e, cfg, i, act);
                apply_action(e, cfg, i, act, 2, 0, 0);
            }
        break;
    }
    default: /* "AI only" branch gaming_env.py:309-374 (env-level BFS is zero-weight, F20) */
        for (int i = 0; i < cfg->n_tanks; ++i) {
            int aa[3];
            for (int k = 0; k < 3; ++k) aa[k] = cfg->mode == ATO_MODE_ARENA ? arena_act[i][k] : actions[3 * i + k];
            apply_action(e, cfg, i, aa, 2, 0, 0);
        }
        break;
    }
    bullets_pass(e, cfg); /* gaming_env.py:378-379 */
    for (int i = 0; i < cfg->n_tanks; ++i) e->tanks[i].last_alive = e->tanks[i].alive;
}

/* env/gym_env.py:105-183 and env/gym_env_multi.py:225-304 */
static void write_obs(Handle *h, Env *e, float *obs) {
    const ato_config *cfg = &h->cfg;
    const int team_layout = cfg->mode == ATO_MODE_TEAM;
    const int rows = team_layout ? h->n_agents : cfg->n_tanks;
    int p = 0;
    for (int row = 0; row < rows; ++row) {
        int ti = team_layout ? h->agent_tank[row] : row;
        const Tank *t = &e->tanks[ti];
        double c8[8], cv, sv;
        ato_corners(t->x, t->y, t->angle, c8);
        ato_trig(t->angle, &cv, &sv);
        obs[p++] = (float)t->x; obs[p++] = (float)t->y;
        for (int k = 0; k < 8; ++k) obs[p++] = (float)c8[k];
        obs[p++] = (float)((double)(t->speed * 1) * cv);
        obs[p++] = (float)((double)(t->speed * 1) * sv);
        obs[p++] = (float)t->hitting_wall;
        if (team_layout) obs[p++] = 1.0f;
        obs[p++] = t->alive ? 1.0f : 0.0f;
        int cnt = 0;
        for (int i = 0; i < e->n_bullets && cnt < MAX_BULLETS; ++i)
            if (e->bullets[i].owner == ti) {
                const Bullet *b = &e->bullets[i];
                obs[p++] = (float)b->x; obs[p++] = (float)b->y; obs[p++] = (float)b->dx; obs[p++] = (float)b->dy;
                ++cnt;
            }
        for (; cnt < MAX_BULLETS; ++cnt) { obs[p++] = 0; obs[p++] = 0; obs[p++] = 0; obs[p++] = 0; }
        for (int o = 0; o < cfg->n_tanks; ++o) {
            if (o == ti) continue;
            cnt = 0;
            for (int i = 0; i < e->n_bullets && cnt < MAX_BULLETS; ++i)
                if (e->bullets[i].owner == o) {
                    const Bullet *b = &e->bullets[i];
                    double rx = b->x - t->x, ry = b->y - t->y;
                    double dist = sqrt(rx * rx + ry * ry);
                    obs[p++] = (float)b->x; obs[p++] = (float)b->y; obs[p++] = (float)b->dx; obs[p++] = (float)b->dy;
                    obs[p++] = (float)rx; obs[p++] = (float)ry; obs[p++] = (float)dist;
                    if (team_layout) obs[p++] = (e->tanks[o].team == t->team) ? 1.0f : 0.0f;
                    ++cnt;
                }
            for (; cnt < MAX_BULLETS; ++cnt)
                for (int k = 0; k < (team_layout ? 8 : 7); ++k) obs[p++] = 0;
        }
        for (int o = 0; o < cfg->n_tanks; ++o) {
            if (o == ti) continue;
            const Tank *ot = &e->tanks[o];
            double rx = ot->x - t->x, ry = ot->y - t->y;
            double dist = sqrt(rx * rx + ry * ry);
            double oc[8], ocv, osv;
            ato_corners(ot->x, ot->y, ot->angle, oc);
            ato_trig(ot->angle, &ocv, &osv);
            obs[p++] = (float)ot->x; obs[p++] = (float)ot->y; obs[p++] = (float)rx; obs[p++] = (float)ry;
            obs[p++] = (float)dist;
            for (int k = 0; k < 8; ++k) obs[p++] = (float)oc[k];
            obs[p++] = (float)((double)(ot->speed * 1) * ocv);
            obs[p++] = (float)((double)(ot->speed * 1) * osv);
            obs[p++] = (float)ot->hitting_wall;
            if (team_layout) obs[p++] = (ot->team == t->team) ? 1.0f : 0.0f;
            obs[p++] = ot->alive ? 1.0f : 0.0f;
        }
        for (int i = 0; i < e->n_walls; ++i) {
            double wx = e->wall_c[i] * GRID, wy = e->wall_r[i] * GRID;
            obs[p++] = (float)wx; obs[p++] = (float)wy; obs[p++] = (float)(wx + GRID); obs[p++] = (float)(wy + GRID);
        }
        for (int i = 0; i < ATO_CELLS; ++i) obs[p++] = (float)e->maze[i];
        for (int i = 0; i < 2 * e->n_buff; ++i) obs[p++] = (float)e->zones[i];
        for (int i = 0; i < 2 * e->n_debuff; ++i) obs[p++] = (float)e->zones[4 + i];
        obs[p++] = t->in_buff ? 1.0f : 0.0f;
        obs[p++] = t->in_debuff ? 1.0f : 0.0f;
    }
}

static int count_alive_teams(const ato_config *cfg, const Env *e) {
    int seen[ATO_MAX_TANKS], n = 0;
    for (int i = 0; i < cfg->n_tanks; ++i) {
        if (!e->tanks[i].alive) continue;
        int dup = 0;
        for (int k = 0; k < n; ++k) dup |= seen[k] == e->tanks[i].team;
        if (!dup) seen[n++] = e->tanks[i].team;
    }
    return n;
}

/* env/gym_env.py:73-103,186-200; env/gym_env_multi.py:202-223,307-326 (F1, F4, F5) */
static void env_step(Handle *h, int64_t i, const int32_t *actions, float *obs, float *rewards, uint8_t *done) {
    const ato_config *cfg = &h->cfg;
    Env *e = &h->envs[i];
    const int team_layout = cfg->mode == ATO_MODE_TEAM;
    const int rows = team_layout ? h->n_agents : cfg->n_tanks;
    const int D = ato_obs_dim(h);
    const int arows = ato_action_rows(h);
    const int32_t *a = actions ? actions + i * arows * 3 : NULL;
    e->training_step += 1;
    e->term_valid = 0;
    if (!team_layout && a) { /* gym_env.py:77-85 change_time */
        if (e->prev_valid)
            for (int k = 0; k < arows; ++k)
                if (a[3 * k] == e->prev_act[k][0] && a[3 * k + 1] == e->prev_act[k][1]) e->change_time[k] += 1;
        for (int k = 0; k < arows; ++k) { e->prev_act[k][0] = a[3 * k]; e->prev_act[k][1] = a[3 * k + 1]; }
        e->prev_valid = 1;
    }
    game_step(h, e, a);
    write_obs(h, e, obs + i * rows * D);
    for (int r = 0; r < rows; ++r) rewards[i * rows + r] = (float)e->tanks[team_layout ? h->agent_tank[r] : r].reward;
    int d;
    if (cfg->terminate_time != 0) {
        if (e->training_step < cfg->terminate_time) d = 0;
        else { e->training_step = 0; d = 1; }
    } else {
        int alive_teams = count_alive_teams(cfg, e);
        d = alive_teams <= 1;
        if (team_layout && alive_teams == 1) /* gym_env_multi.py:319-326: lands AFTER the rewards were read (F5) */
            for (int k = 0; k < cfg->n_tanks; ++k) {
                int winner_team = -1;
                for (int q = 0; q < cfg->n_tanks; ++q)
                    if (e->tanks[q].alive) { winner_team = e->tanks[q].team; break; }
                if (e->tanks[k].team == winner_team) e->tanks[k].reward += VICTORY_REWARD;
            }
    }
    done[i] = (uint8_t)d;
    if (d && cfg->auto_reset) {
        e->term_valid = 1;
        e->term_winner = -1;
        for (int k = cfg->n_tanks - 1; k >= 0; --k) {
            if (e->tanks[k].alive) e->term_winner = k; /* first alive tank, gym_env.py:97 */
            e->term_hits[k] = e->tanks[k].num_hit;
            e->term_be_hits[k] = e->tanks[k].num_be_hit;
        }
        env_reset(h, e);
        write_obs(h, e, obs + i * rows * D);
    }
}

/* --------------------------------------------------------------- C entry */
int32_t ato_obs_rows(void *hh) {
    Handle *h = (Handle *)hh;
    return h->cfg.mode == ATO_MODE_TEAM ? h->n_agents : h->cfg.n_tanks;
}
int32_t ato_action_rows(void *hh) {
    Handle *h = (Handle *)hh;
    if (h->cfg.mode == ATO_MODE_TEAM) return h->n_agents;
    if (h->cfg.mode == ATO_MODE_ARENA) return 0;
    return h->cfg.n_tanks;
}
int32_t ato_obs_dim(void *hh) {
    Handle *h = (Handle *)hh;
    const ato_config *c = &h->cfg;
    int n = c->n_tanks, W = h->n_walls_static;
    int zones = (c->buff_on ? 4 : 0) + (c->debuff_on ? 4 : 0);
    if (c->mode == ATO_MODE_TEAM) return 15 + 24 + (n - 1) * 48 + (n - 1) * 18 + 4 * W + ATO_CELLS + zones + 2;
    return 14 + 24 + (n - 1) * 42 + (n - 1) * 17 + 4 * W + ATO_CELLS + zones + 2;
}

void *ato_create(const ato_config *cfg, int64_t n_envs, int64_t env_id_base, uint64_t seed) {
    Handle *h = (Handle *)calloc(1, sizeof(Handle));
    h->cfg = *cfg;
    h->n_envs = n_envs;
    for (int i = 0; i < cfg->n_tanks; ++i)
        if (cfg->ctrl[i] == ATO_CTRL_AGENT) h->agent_tank[h->n_agents++] = i;
    h->n_walls_static = cfg->map == ATO_MAP_OCTAGON ? 40 : (cfg->map == ATO_MAP_BATTLEFIELD ? 61 : 72);
    h->envs = (Env *)calloc((size_t)n_envs, sizeof(Env));
    for (int64_t i = 0; i < n_envs; ++i) {
        Env *e = &h->envs[i];
        e->rng.seed = seed;
        e->rng.env_id = env_id_base + i;
        e->rng.ctr = 0;
        env_reset(h, e); /* GamingENV.__init__ -> reset() (gaming_env.py:43) */
    }
    return h;
}
void ato_destroy(void *hh) {
    Handle *h = (Handle *)hh;
    if (!h) return;
    free(h->envs);
    free(h);
}
void ato_reset(void *hh, const uint8_t *mask, float *obs) {
    Handle *h = (Handle *)hh;
    const int rd = ato_obs_rows(h) * ato_obs_dim(h);
    for (int64_t i = 0; i < h->n_envs; ++i)
        if (!mask || mask[i]) {
            env_reset(h, &h->envs[i]);
            if (obs) write_obs(h, &h->envs[i], obs + i * rd);
        }
}
void ato_step(void *hh, const int32_t *actions, float *obs, float *rewards, uint8_t *done) {
    Handle *h = (Handle *)hh;
    for (int64_t i = 0; i < h->n_envs; ++i) env_step(h, i, actions, obs, rewards, done);
}
typedef struct MtJob {
    Handle *h;
    const int32_t *actions;
    float *obs, *rewards;
    uint8_t *done;
    int64_t lo, hi;
} MtJob;
static void *mt_worker(void *p) {
    MtJob *j = (MtJob *)p;
    for (int64_t i = j->lo; i < j->hi; ++i) env_step(j->h, i, j->actions, j->obs, j->rewards, j->done);
    return NULL;
}
/* envs are independent, so the batch is simply cut into n_threads contiguous ranges */
void ato_step_mt(void *hh, const int32_t *actions, float *obs, float *rewards, uint8_t *done, int32_t n_threads) {
    Handle *h = (Handle *)hh;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    MtJob jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        MtJob j = {h, actions, obs, rewards, done, h->n_envs * t / n_threads, h->n_envs * (t + 1) / n_threads};
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    mt_worker(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], NULL);
}

void ato_get_state(void *hh, int64_t env, ato_env_state *o) {
    Handle *h = (Handle *)hh;
    const Env *e = &h->envs[env];
    memset(o, 0, sizeof *o);
    o->n_bullets = e->n_bullets; o->tick = e->tick; o->training_step = e->training_step;
    o->episode_steps = e->episode_steps; o->rng_ctr = e->rng.ctr; o->n_walls = e->n_walls;
    memcpy(o->zones, e->zones, sizeof o->zones);
    memcpy(o->maze, e->maze, ATO_CELLS);
    for (int i = 0; i < h->cfg.n_tanks; ++i) {
        const Tank *t = &e->tanks[i];
        ato_tank_state *s = &o->tanks[i];
        s->x = t->x; s->y = t->y; s->reward = t->reward; s->angle = t->angle; s->speed = t->speed; s->alive = t->alive;
        s->hitting_wall = t->hitting_wall; s->in_buff = t->in_buff; s->in_debuff = t->in_debuff;
        s->max_bullets = t->max_bullets; s->last_shot_ms = t->last_shot; s->num_hit = t->num_hit;
        s->num_be_hit = t->num_be_hit;
        const Bot *b = &e->bots[i];
        ato_bot_state *bs = &o->bots[i];
        bs->target = b->target; bs->aiming = b->aiming; bs->rot_dir = b->rot_dir; bs->stuck_timer = b->stuck_timer;
        bs->has_next = b->has_next; bs->next_r = b->next_r; bs->next_c = b->next_c; bs->last_x = b->last_x;
        bs->last_y = b->last_y; bs->has_tpos = b->has_tpos; bs->tpos_x = b->tpos_x; bs->tpos_y = b->tpos_y;
        bs->dodge_timer = b->dodge_timer; bs->dodge_state = b->dodge_state; bs->has_dodge_angle = b->has_dodge_angle;
        bs->dodge_angle = b->dodge_angle; bs->action_delay = b->action_delay;
        memcpy(bs->cur_action, b->cur_action, sizeof bs->cur_action);
    }
    for (int i = 0; i < e->n_bullets; ++i) {
        const Bullet *b = &e->bullets[i];
        ato_bullet_state *s = &o->bullets[i];
        s->x = b->x; s->y = b->y; s->dx = b->dx; s->dy = b->dy; s->pending = b->pending; s->owner = b->owner;
        s->distance = b->distance; s->bounces = b->bounces; s->cleared = b->cleared;
    }
}
void ato_set_state(void *hh, int64_t env, const ato_env_state *o) {
    Handle *h = (Handle *)hh;
    Env *e = &h->envs[env];
    e->n_bullets = o->n_bullets; e->tick = o->tick; e->training_step = o->training_step;
    e->episode_steps = o->episode_steps; e->rng.ctr = o->rng_ctr;
    memcpy(e->zones, o->zones, sizeof o->zones);
    memcpy(e->maze, o->maze, ATO_CELLS);
    e->n_walls = e->n_empty = 0;
    for (int r = 0; r < MAZEH; ++r)
        for (int c = 0; c < MAZEW; ++c) {
            if (e->maze[r * MAZEW + c]) { e->wall_r[e->n_walls] = r; e->wall_c[e->n_walls] = c; ++e->n_walls; }
            else { e->empty_x[e->n_empty] = (int)(c * GRID); e->empty_y[e->n_empty] = (int)(r * GRID); ++e->n_empty; }
        }
    for (int i = 0; i < h->cfg.n_tanks; ++i) {
        Tank *t = &e->tanks[i];
        const ato_tank_state *s = &o->tanks[i];
        t->x = s->x; t->y = s->y; t->reward = s->reward; t->angle = s->angle; t->speed = s->speed; t->alive = s->alive;
        t->last_alive = s->alive; t->hitting_wall = s->hitting_wall; t->in_buff = s->in_buff;
        t->in_debuff = s->in_debuff; t->max_bullets = s->max_bullets; t->last_shot = s->last_shot_ms;
        t->num_hit = s->num_hit; t->num_be_hit = s->num_be_hit;
        Bot *b = &e->bots[i];
        const ato_bot_state *bs = &o->bots[i];
        b->target = bs->target; b->aiming = bs->aiming; b->rot_dir = bs->rot_dir; b->stuck_timer = bs->stuck_timer;
        b->has_next = bs->has_next; b->next_r = bs->next_r; b->next_c = bs->next_c; b->last_x = bs->last_x;
        b->last_y = bs->last_y; b->has_tpos = bs->has_tpos; b->tpos_x = bs->tpos_x; b->tpos_y = bs->tpos_y;
        b->dodge_timer = bs->dodge_timer; b->dodge_state = bs->dodge_state; b->has_dodge_angle = bs->has_dodge_angle;
        b->dodge_angle = bs->dodge_angle; b->action_delay = bs->action_delay;
        memcpy(b->cur_action, bs->cur_action, sizeof bs->cur_action);
    }
    for (int i = 0; i < o->n_bullets; ++i) {
        Bullet *b = &e->bullets[i];
        const ato_bullet_state *s = &o->bullets[i];
        b->x = s->x; b->y = s->y; b->dx = s->dx; b->dy = s->dy; b->pending = s->pending; b->owner = s->owner;
        b->distance = s->distance; b->bounces = s->bounces; b->cleared = s->cleared;
    }
}
void ato_get_info(void *hh, int32_t *winner, int32_t *hits, int32_t *be_hits, int32_t *change_time) {
    Handle *h = (Handle *)hh;
    const int n = h->cfg.n_tanks;
    for (int64_t i = 0; i < h->n_envs; ++i) {
        const Env *e = &h->envs[i];
        if (winner) { /* gym_env.py:97 (F16) */
            winner[i] = -1;
            for (int k = 0; k < n; ++k)
                if (e->tanks[k].alive) { winner[i] = k; break; }
        }
        for (int k = 0; k < n; ++k) {
            if (hits) hits[i * n + k] = e->tanks[k].num_hit;
            if (be_hits) be_hits[i * n + k] = e->tanks[k].num_be_hit;
            if (change_time) change_time[i * n + k] = e->change_time[k];
        }
    }
}
/* info of the latest step: the finished episode's for envs the harness auto-reset in it (latched[i] = 1) */
void ato_get_terminal_info(void *hh, int32_t *winner, int32_t *hits, int32_t *be_hits, int32_t *change_time,
                           uint8_t *latched) {
    Handle *h = (Handle *)hh;
    const int n = h->cfg.n_tanks;
    ato_get_info(hh, winner, hits, be_hits, change_time);
    for (int64_t i = 0; i < h->n_envs; ++i) {
        const Env *e = &h->envs[i];
        if (latched) latched[i] = (uint8_t)e->term_valid;
        if (!e->term_valid) continue;
        if (winner) winner[i] = e->term_winner;
        for (int k = 0; k < n; ++k) {
            if (hits) hits[i * n + k] = e->term_hits[k];
            if (be_hits) be_hits[i * n + k] = e->term_be_hits[k];
        }
    }
}
int32_t ato_bfs_len_between(void *hh, int64_t env, int32_t i, int32_t j) {
    Handle *h = (Handle *)hh;
    const Env *e = &h->envs[env];
    int r0, c0, r1, c1;
    grid_position(&e->tanks[i], &r0, &c0);
    grid_position(&e->tanks[j], &r1, &c1);
    return ato_bfs_path(e->maze, r0, c0, r1, c1, NULL);
}
