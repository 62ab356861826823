// emul.cpp - TEST-ONLY host emulation of the CUDA env kernel (alphatank_b200/csrc/at_kernels.cu).
//
// This GPU-less build container cannot run the kernels, so this file compiles the kernels' own per-thread
// code (at_sim.h, __host__ __device__) with g++ and drives it the way env_kernel does: blocks of 64 envs = two
// warps, every warp 32 fibers of a small SIMT emulator (simt.h) that run step_lane<> - the kernel body - in
// lockstep at the warp collectives, so the pooled line-of-sight stage executes here exactly as written; the
// observation rows come from the job functions of rows_kernel (at_rows.h).  It lets the CPU test-suite compare
// the CUDA-side FORMULATION (cell bitmasks, integer-angle tables, packed bullets, pooled line of sight,
// bit-parallel BFS, counter penalties, row jobs) with the oracle before any GPU time is spent.  It is not part of the product, is never loaded by alphatank_b200/, and is not a fallback:
// the product fails loudly without a GPU.  What it cannot cover (real memory ordering, bulk-async stores) is
// covered by the -m gpu tests.
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../alphatank_b200/csrc/at_host.h"
#include "../../alphatank_b200/csrc/at_sim.h"
#include "../../alphatank_b200/csrc/at_rows.h"
#include "simt.h"

using namespace at;

struct em_env {
    KCfg k;
    DevState st;
    HostLuts luts;
    int64_t n_envs;
    std::vector<char> arena;
    std::vector<char> block;  // one block's working set
    std::vector<float> tmpl;
    BlockMem bm;
    simt::Warp warp;
};
static thread_local std::string g_err;

template <int A, int B, int C>
struct ShapeTag { static constexpr int nt = A, nr = B, tm = C; };
// rows_kernel's formulation (at_rows.h): the rows of env e rebuilt from the STORED state by independent jobs
static int g_lane_order = 0;  // simt lane order between collectives: 0 forward, 1 reverse, 2 shuffled
static void rows_from_state(em_env *env, int64_t e, float *img) {
    const KCfg &c = env->k;
    const DevState &st = env->st;
    const int n = c.n_tanks, S = n * SLOTS_PER_TANK;
    double tx[MAX_TANKS], ty[MAX_TANKS];
    uint32_t tp[MAX_TANKS];
    for (int i = 0; i < n; ++i) { tx[i] = st.tx[i * st.N + e]; ty[i] = st.ty[i * st.N + e]; tp[i] = st.tpack[i * st.N + e]; }
    std::vector<uint64_t> rm(S, ROWS_EMPTY);
    std::vector<double> rx(S, 0.0), ry(S, 0.0);
    for (uint32_t s = 0; s < st.cnt[e]; ++s) {   // phase 1 of the kernel: scatter the live records by (owner, rank)
        const uint64_t m = st.bmeta[(int64_t)s * st.N + e];
        const int p = (int)(m & 15u) * SLOTS_PER_TANK + (int)((m >> BM_RANK) & 7u);
        rm[p] = m; rx[p] = st.bx[(int64_t)s * st.N + e]; ry[p] = st.by[(int64_t)s * st.N + e];
    }
    if (c.map != AT_MAP_MAZE) {
        for (int r = 0; r < c.rows; ++r)
            memcpy(img + (int64_t)r * c.obs_dim + c.off_walls, env->tmpl.data(), (4 * c.n_walls + CELLS) * sizeof(float));
    } else {
        for (int cell = 0; cell < CELLS; ++cell) rows_maze_job(c, st.wall_lo[e], st.wall_hi[e], cell, img);
    }
    // the kernel's compile-time shapes (2v2 team, 1v1) and the generic build
    auto jobs = [&](auto shape) {
        constexpr int NT = decltype(shape)::nt, NR = decltype(shape)::nr, TM = decltype(shape)::tm;
        for (int p = 0; p < S; ++p)
            rows_position_job<NT, NR, TM>(c, st.trig, p / SLOTS_PER_TANK, p % SLOTS_PER_TANK, rm[p], rx[p], ry[p], tx, ty, 1,
                                          img, (int)((e + p) & 7));
        for (int r = 0; r < c.rows; ++r)
            for (int i = 0; i < n; ++i) rows_tank_job<NT, NR, TM>(c, st.trig, st.corn, r, i, tx, ty, tp, 1, st.zones[e], img);
    };
    if (n == 4 && c.rows == 2 && c.team_layout) jobs(ShapeTag<4, 2, 1>{});
    else if (n == 2 && c.rows == 2 && !c.team_layout) jobs(ShapeTag<2, 2, 0>{});
    else jobs(ShapeTag<0, 0, -1>{});
}

template <int SPEC>
static void run_t(em_env *env, bool is_step, const int32_t *actions, const uint8_t *mask, float *obs, float *rewards,
                uint8_t *done) {
    const KCfg &c = env->k;
    const DevState &st = env->st;
    const int64_t row_len = (int64_t)c.rows * c.obs_dim;
    env->warp.order_mode = g_lane_order;
    for (int64_t b0 = 0; b0 < env->n_envs; b0 += BS) {
        for (int wp = 0; wp < BS / 32; ++wp) {
            const int64_t e_warp0 = b0 + wp * 32;
            uint32_t valid_mask = 0;
            for (int l = 0; l < 32; ++l)
                if (e_warp0 + l < st.N) valid_mask |= 1u << l;
            if (!valid_mask) continue;
            if (is_step) {
                simt::run_warp(env->warp, valid_mask, [&](int lane) {
                    const int64_t e = e_warp0 + lane;
                    step_lane<simt::FiberWarp, SPEC>(c, st, env->bm, e_warp0, wp * 32 + lane, valid_mask,
                                                     actions ? actions + e * c.act_rows * 3 : nullptr, rewards, done);
                });
            } else {
                for (int l = 0; l < 32; ++l) {
                    const int64_t e = e_warp0 + l;
                    if (e < st.N && (!mask || mask[e])) reset_lane(c, st, env->bm, e, wp * 32 + l);
                }
            }
        }
    }
    if (obs == nullptr || c.rows == 0) return;
    for (int64_t e = 0; e < st.N; ++e) {   // rows_kernel: every stepped env; the masked ones of a reset
        if (!is_step && mask && !mask[e]) continue;
        for (int64_t q = 0; q < row_len; ++q) obs[e * row_len + q] = -12345.0f;  // every column must be written
        rows_from_state(env, e, obs + e * row_len);
    }
}

// steps run on the same SimT<SPEC> specialisation the library would pick for the config (resets: the generic build)
static int g_spec_cap = 3;
static void run(em_env *env, bool is_step, const int32_t *actions, const uint8_t *mask, float *obs, float *rewards,
                uint8_t *done) {
    const int spec = is_step ? sim_spec_for(env->k, g_spec_cap) : 0;
    if (spec == 5) run_t<5>(env, is_step, actions, mask, obs, rewards, done);
    else if (spec == 4) run_t<4>(env, is_step, actions, mask, obs, rewards, done);
    else if (spec == 3) run_t<3>(env, is_step, actions, mask, obs, rewards, done);
    else if (spec == 2) run_t<2>(env, is_step, actions, mask, obs, rewards, done);
    else if (spec == 1) run_t<1>(env, is_step, actions, mask, obs, rewards, done);
    else run_t<0>(env, is_step, actions, mask, obs, rewards, done);
}

extern "C" {
const char *em_last_error(void) { return g_err.c_str(); }

int em_create(const at_config *cfg, int64_t n_envs, int64_t env_id_base, uint64_t seed, em_env **out) {
    em_env *env = new em_env();
    std::string msg = build_kcfg(cfg, seed, env_id_base, env->k);
    if (!msg.empty()) { g_err = msg; delete env; return AT_ERR_ARG; }
    env->n_envs = n_envs;
    {
        const M128 nm = near_wall_mask(env->k.wall_lo, env->k.wall_hi);
        env->k.near_lo = nm.lo;
        env->k.near_hi = nm.hi;
    }
    build_luts(env->luts);
    const ArenaLayout L = arena_layout(env->k, n_envs);
    env->arena.assign(L.bytes, 0);
    bind_state(env->st, env->arena.data(), L, n_envs);
    memset(env->arena.data() + L.o_ttick, 0xFF, 4 * (size_t)n_envs);   // no terminal info latched yet
    memcpy(env->arena.data() + L.o_trig, env->luts.trig.data(), 722 * 8);
    memcpy(env->arena.data() + L.o_corn, env->luts.corn.data(), 361 * 4 * 8);
    memcpy(env->arena.data() + L.o_pend, env->luts.pend.data(), PEND_LUT * 8);
    memcpy(env->arena.data() + L.o_udir, env->luts.udelta.data(), 364);
    if (env->k.map != AT_MAP_MAZE) {
        fill_bfs_table(env->k.wall_lo, env->k.wall_hi, (uint16_t *)(env->arena.data() + L.o_bfs));
        env->st.bfs_tab = (const uint16_t *)(env->arena.data() + L.o_bfs);
    }
    const int n = env->k.n_tanks;
    env->block.assign((size_t)n * BS * (5 * 8 + 6 * 4) + 722 * 8 + (size_t)(BS / 32) * 512 + 64, 0);
    char *p = env->block.data();
    BlockMem &bm = env->bm;
    bm.trig = (double *)p; p += 722 * 8;
    memcpy(bm.trig, env->luts.trig.data(), 722 * 8);
    bm.tx = (double *)p; p += n * BS * 8; bm.ty = (double *)p; p += n * BS * 8; bm.trew = (double *)p; p += n * BS * 8;
    bm.bf0 = (double *)p; p += n * BS * 8; bm.bf1 = (double *)p; p += n * BS * 8;
    bm.tpack = (uint32_t *)p; p += n * BS * 4; bm.nearlo = (uint32_t *)p; p += n * BS * 4;
    bm.bpack = (uint32_t *)p; p += n * BS * 4; bm.tlive = (uint8_t *)p; p += n * BS * 4;
    bm.tshot = (int32_t *)p; p += n * BS * 4;
    bm.nearhi = (uint32_t *)p; p += n * BS * 4;
    bm.los_q = (uint8_t *)p; p += (BS / 32) * 256;
    bm.los_r = (uint8_t *)p; p += (BS / 32) * 256;
    bm.udelta = env->luts.udelta.data();
    env->tmpl.assign(4 * 72 + CELLS, 0.f);
    if (env->k.map != AT_MAP_MAZE) fill_static_template(env->k.wall_lo, env->k.wall_hi, env->tmpl.data());
    *(double *)(env->arena.data() + L.o_dyn) = env->k.weakness;
    fill_atan_table((double *)(env->arena.data() + L.o_atab));
    run(env, false, nullptr, nullptr, nullptr, nullptr, nullptr);  // the constructor's own reset
    *out = env;
    return AT_OK;
}
void em_destroy(em_env *env) { delete env; }
void em_set_lane_order(int m) { g_lane_order = m; }

int em_set_weakness(em_env *env, double w) {
    *(double *)env->st.dyn = w;
    env->k.weakness = w;
    return AT_OK;
}
uint64_t em_collectives(em_env *env) { return env->warp.collectives; }
void em_set_spec_cap(int cap) { g_spec_cap = cap; }
int em_obs_dim(em_env *env) { return env->k.obs_dim; }
int em_obs_rows(em_env *env) { return env->k.rows; }
int em_action_rows(em_env *env) { return env->k.act_rows; }
void em_reset(em_env *env, const uint8_t *mask, float *obs) { run(env, false, nullptr, mask, obs, nullptr, nullptr); }
void em_step(em_env *env, const int32_t *actions, float *obs, float *rewards, uint8_t *done) {
    run(env, true, actions, nullptr, obs, rewards, done);
}
void em_get_state(em_env *env, int64_t e, at_env_state *out) {
    PackedEnv p;
    memset(&p, 0, sizeof p);
    gather_env(env->k, env->st, e, p);
    unpack_env(env->k, env->luts, p, *out);
}
int em_set_state(em_env *env, int64_t e, const at_env_state *in) {
    PackedEnv p;
    std::string msg = pack_env(env->k, env->luts, *in, p);
    if (!msg.empty()) { g_err = msg; return AT_ERR_ARG; }
    scatter_env(env->k, env->st, e, p);
    return AT_OK;
}
void em_get_info(em_env *env, int32_t *winner, int32_t *hits, int32_t *be_hits, int32_t *change_time) {
    const KCfg &c = env->k;
    const DevState &st = env->st;
    for (int64_t e = 0; e < st.N; ++e) {
        int w = -1;
        for (int i = 0; i < c.n_tanks; ++i) {
            int64_t g = (int64_t)i * st.N + e;
            if (w < 0 && (st.tpack[g] & TP_ALIVE)) w = i;
            if (hits) hits[e * c.n_tanks + i] = (int32_t)(st.thits[g] & 0xFFFFu);
            if (be_hits) be_hits[e * c.n_tanks + i] = (int32_t)(st.thits[g] >> 16);
            if (change_time) change_time[e * c.n_tanks + i] = st.change_time[g];
        }
        if (winner) winner[e] = w;
    }
}
void em_get_terminal_info(em_env *env, int32_t *winner, int32_t *hits, int32_t *be_hits, int32_t *change_time,
                          uint8_t *latched) {   // = info_kernel<true>
    const KCfg &c = env->k;
    const DevState &st = env->st;
    em_get_info(env, winner, hits, be_hits, change_time);
    for (int64_t e = 0; e < st.N; ++e) {
        const bool lt = st.term_tick[e] == st.tick[e];
        if (latched) latched[e] = lt ? 1 : 0;
        if (!lt) continue;
        if (winner) winner[e] = (int)st.term_winner[e] - 1;
        for (int i = 0; i < c.n_tanks; ++i) {
            int64_t g = (int64_t)i * st.N + e;
            if (hits) hits[e * c.n_tanks + i] = (int32_t)(st.term_hits[g] & 0xFFFFu);
            if (be_hits) be_hits[e * c.n_tanks + i] = (int32_t)(st.term_hits[g] >> 16);
        }
    }
}
int em_bfs(const uint8_t *maze121, int sr, int sc, int gr, int gc, int *next_r, int *next_c) {
    uint64_t lo = 0, hi = 0;
    for (int i = 0; i < CELLS; ++i)
        if (maze121[i]) { if (i < 64) lo |= 1ull << i; else hi |= 1ull << (i - 64); }
    return bfs_first_step(lo, hi, sr, sc, gr, gc, next_r, next_c);
}
double em_jump_axis(double x, double d, int m) { return jump_axis(x, d, m); }
// atan2_refine on a starting value that is `ulps` units off the host's atan2 (the device's is good to 2 ulp)
double em_atan2_refine(double y, double x, int ulps) {
    static std::vector<double> tab;
    if (tab.empty()) { tab.resize(361 * ATAB_STRIDE); fill_atan_table(tab.data()); }
    double th0 = atan2(y, x);
    for (int i = 0; i < (ulps < 0 ? -ulps : ulps); ++i) th0 = nextafter(th0, ulps < 0 ? -10.0 : 10.0);
    const double t = th0 * (180.0 / 3.141592653589793), k = rint(t);
    if (!(fabs(t - k) < 1e-9)) return NAN;      // not a case the kernel refines
    return atan2_refine(y, x, th0, (int)k, tab.data());
}
void em_generate_maze(uint64_t seed, int64_t env_id, uint32_t *ctr, uint8_t *maze121) {
    const MazeOut mz = generate_maze_mask(seed, env_id, *ctr);
    const uint64_t lo = mz.lo, hi = mz.hi;
    *ctr = mz.ctr;
    for (int i = 0; i < CELLS; ++i) maze121[i] = (uint8_t)((i < 64 ? lo >> i : hi >> (i - 64)) & 1ull);
}
}
