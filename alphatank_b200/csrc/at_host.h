// at_host.h - CUDA-free host helpers shared by at_kernels.cu and the test-only g++ emulator
// (tests/cpu_emul): config freezing, lookup tables, the SoA arena layout and the AoS<->packed conversions
// behind at_get_state / at_set_state.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/alphatank_b200.h"
#include "at_types.h"

namespace at {

// AoS <-> SoA transfer record for at_get_state / at_set_state (tests, checkpoints)
struct PackedEnv {
    uint32_t cnt, rng, tick, tstep, epstep, zones, prev_act, pad;
    uint64_t wlo, whi;
    double tx[MAX_TANKS], ty[MAX_TANKS], trew[MAX_TANKS], bf0[MAX_TANKS], bf1[MAX_TANKS];
    uint32_t tpack[MAX_TANKS], thits[MAX_TANKS], bpack[MAX_TANKS];
    int32_t tshot[MAX_TANKS], change_time[MAX_TANKS];
    double bx[AT_MAX_BULLETS], by[AT_MAX_BULLETS];
    uint64_t bmeta[AT_MAX_BULLETS];
};

// pygame.math.Vector2.rotate (src_c/math.c _vector2_rotate_helper), restated as in tests/refstubs/pygame
inline void pg_rotate(double vx, double vy, double angle, double *ox, double *oy) {
    const double eps = 1e-6;
    double a = fmod(angle, 360.0);
    if (a < 0) a += 360.0;
    if (fmod(a + eps, 90.0) < 2 * eps) {
        switch ((int)((a + eps) / 90.0)) {
        case 0: case 4: *ox = vx; *oy = vy; return;
        case 1: *ox = -vy; *oy = vx; return;
        case 2: *ox = -vx; *oy = -vy; return;
        default: *ox = vy; *oy = -vx; return;
        }
    }
    // (host TU is compiled with -ffp-contract=off: plain multiplies and adds, as in pygame's wheels)
    double rad = a * M_PI / 180.0;
    double s = sin(rad), cc = cos(rad);
    *ox = cc * vx - s * vy;
    *oy = s * vx + cc * vy;
}

inline void build_map(int map, uint64_t *lo, uint64_t *hi, int *n_walls) {
    uint8_t m[CELLS];
    memset(m, 0, sizeof m);
    if (map == AT_MAP_OCTAGON) {  // gaming_env.py:576-578
        for (int r = 0; r < 11; ++r)
            for (int c = 0; c < 11; ++c) m[r * 11 + c] = !(r >= 1 && r < 10 && c >= 1 && c < 10);
    } else if (map == AT_MAP_BATTLEFIELD) {  // gaming_env.py:579-606
        for (int i = 0; i < 11; ++i) { m[i] = 1; m[10 * 11 + i] = 1; m[i * 11] = 1; m[i * 11 + 10] = 1; }
        m[5 * 11 + 5] = m[4 * 11 + 5] = m[6 * 11 + 5] = m[5 * 11 + 4] = m[5 * 11 + 6] = 1;
        const int cover[16][2] = {{2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {8, 3}, {8, 4}, {8, 5}, {8, 6}, {8, 7},
                                  {4, 8}, {5, 8}, {6, 8}, {4, 2}, {5, 2}, {6, 2}};
        for (auto &rc : cover) m[rc[0] * 11 + rc[1]] = 1;
    } else {
        memset(m, 1, sizeof m);  // placeholder until the first reset generates the maze
    }
    *lo = *hi = 0;
    int w = 0;
    for (int i = 0; i < CELLS; ++i)
        if (m[i]) { ++w; if (i < 64) *lo |= 1ull << i; else *hi |= 1ull << (i - 64); }
    *n_walls = map == AT_MAP_MAZE ? 72 : w;
}


struct HostLuts {
    std::vector<double> trig, corn, pend, udir;
    std::vector<uint8_t> udelta;
    bool udelta_ok = true;
};
// built with the host libm exactly as CPython's math module / pygame would
inline void build_luts(HostLuts &l) {
    l.trig.resize(722); l.corn.resize(361 * 4); l.pend.resize(PEND_LUT); l.udir.resize(722); l.udelta.resize(364);
    for (int a = 0; a <= 360; ++a) {
        double rad = (double)a * (M_PI / 180.0);  // math.radians
        l.trig[2 * a] = cos(rad);
        l.trig[2 * a + 1] = sin(rad);
        pg_rotate(-15.0, -12.0, (double)a, &l.corn[4 * a], &l.corn[4 * a + 1]);
        pg_rotate(15.0, -12.0, (double)a, &l.corn[4 * a + 2], &l.corn[4 * a + 3]);
        // bullet_dir / np.linalg.norm(bullet_dir) (sprite.py:32, 516): norm = sqrt(fma(dy, dy, dx*dx)), see oracle
        const double dx = l.trig[2 * a], dy = -l.trig[2 * a + 1];
        const double nrm = sqrt(fma(dy, dy, dx * dx));
        l.udir[2 * a] = dx / nrm;
        l.udir[2 * a + 1] = dy / nrm;
        // the normalised direction is within one ulp of (dx, dy): keep only the two ulp offsets (-1, 0, +1)
        int64_t b0, b1, v0, v1;
        memcpy(&b0, &l.udir[2 * a], 8); memcpy(&b1, &l.udir[2 * a + 1], 8);
        memcpy(&v0, &dx, 8); memcpy(&v1, &dy, 8);
        const int64_t d0 = b0 - v0, d1 = b1 - v1;
        l.udelta_ok = l.udelta_ok && d0 >= -1 && d0 <= 1 && d1 >= -1 && d1 <= 1;
        l.udelta[a] = (uint8_t)((d0 + 1) | ((d1 + 1) << 2));
    }
    double acc = 0.0;
    for (int i = 0; i < PEND_LUT; ++i) { l.pend[i] = acc; acc = acc + 0.02; }  // pending += BULLET_AWAY_PENALTY
}

// Freeze an at_config into the kernel config.  Returns "" or an error message.
inline std::string build_kcfg(const at_config *cfg, uint64_t seed, int64_t env_id_base, KCfg &k) {
    if (cfg->mode < 0 || cfg->mode > 3) return "unknown mode";
    if (cfg->n_tanks < 2 || cfg->n_tanks > AT_MAX_TANKS) return "n_tanks must be 2..8";
    if (cfg->mode != AT_MODE_TEAM && cfg->n_tanks != 2) return "1v1 modes need n_tanks == 2";
    if (cfg->map < 0 || cfg->map > 2) return "unknown map";
    if (cfg->terminate_time < 0) return "terminate_time < 0";
    memset(&k, 0, sizeof k);
    k.mode = cfg->mode; k.n_tanks = cfg->n_tanks; k.map = cfg->map; k.terminate_time = cfg->terminate_time;
    k.buff_on = cfg->buff_on != 0; k.debuff_on = cfg->debuff_on != 0; k.auto_reset = cfg->auto_reset != 0;
    k.weakness = cfg->weakness; k.seed = seed; k.env_id_base = env_id_base;
    const int n = cfg->n_tanks;
    for (int i = 0; i < n; ++i) {
        k.team[i] = cfg->team[i];
        k.ctrl[i] = cfg->ctrl[i];
        k.bot_type[i] = cfg->bot_type[i];
        if (k.team[i] < 0 || k.team[i] > 15) return "team id must be 0..15";
        if (k.ctrl[i] != AT_CTRL_AGENT && k.ctrl[i] != AT_CTRL_BOT) return "ctrl must be agent or bot (mode 'human' is out of scope)";
        if (k.bot_type[i] < 0 || k.bot_type[i] > 5) return "unknown bot type";
    }
    if (cfg->mode == AT_MODE_AGENT) { k.ctrl[0] = k.ctrl[1] = AT_CTRL_AGENT; }
    if (cfg->mode == AT_MODE_BOT_AGENT) { k.ctrl[0] = AT_CTRL_BOT; k.ctrl[1] = AT_CTRL_AGENT; }
    if (cfg->mode == AT_MODE_ARENA) { k.ctrl[0] = k.ctrl[1] = AT_CTRL_BOT; }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            if (k.team[i] == k.team[j]) k.team_members[i] |= 1u << j;
    k.team_layout = cfg->mode == AT_MODE_TEAM;
    int nb = 0;
    for (int i = 0; i < n; ++i) {
        if (k.ctrl[i] == AT_CTRL_AGENT) k.agent_tank[k.n_agents++] = i;
        else { k.bot_ord[i] = nb; k.bot_tank[nb++] = i; }
    }
    k.n_bots = nb;
    // pooled line-of-sight: a smart bot's query may run when the first bot decides if no enemy BOT acts before it
    // (agents have all acted by then in team mode; in the arena both bots decide before anybody acts)
    k.first_bot_k = k.team_layout ? k.n_agents : 0;
    k.los_hoist = 0;
    if (cfg->map != AT_MAP_MAZE)
        for (int b = 0; b < nb; ++b) {
            const int j = k.bot_tank[b];
            if (k.bot_type[j] != AT_BOT_SMART) continue;
            bool ok = true;
            if (cfg->mode != AT_MODE_ARENA)
                for (int b2 = 0; b2 < b; ++b2) ok = ok && k.team[k.bot_tank[b2]] == k.team[j];
            if (ok) k.los_hoist |= 1u << j;
        }
    k.rows = k.team_layout ? k.n_agents : n;
    k.act_rows = k.team_layout ? k.n_agents : (cfg->mode == AT_MODE_ARENA ? 0 : n);
    build_map(cfg->map, &k.wall_lo, &k.wall_hi, &k.n_walls);
    k.n_zone_vals = (k.buff_on ? 4 : 0) + (k.debuff_on ? 4 : 0);
    // row layout: gym_env.py:119-181 / gym_env_multi.py:239-301
    k.self_len = k.team_layout ? 15 : 14;
    k.off_ob = k.self_len;
    k.off_eb = k.off_ob + 24;
    k.eb_stride = 6 * (k.team_layout ? 8 : 7);
    k.off_et = k.off_eb + (n - 1) * k.eb_stride;
    k.et_len = k.team_layout ? 18 : 17;
    k.off_walls = k.off_et + (n - 1) * k.et_len;
    k.off_maze = k.off_walls + 4 * k.n_walls;
    k.off_zones = k.off_maze + CELLS;
    k.off_flags = k.off_zones + k.n_zone_vals;
    k.obs_dim = k.off_flags + 2;
    k.row_bulk_ok = ((k.obs_dim * 4) % 16) == 0;
    return "";
}

// Which SimT<SPEC> specialisation (at_sim.h) a frozen config can run on.  1: team mode, every scripted tank a smart bot,
// fixed map, no TERMINATE_TIME; 2: 1 with exactly four tanks; 3: 2 with the 2a_vs_2b layout (tanks 0, 1 = one team's
// agents, tanks 2, 3 = the other team's bots).  `cap` limits the level (diagnostics: AT_NO_SPEC).
inline int sim_spec_for(const KCfg &k, int cap = 3) {
    bool smart_only = true;
    for (int i = 0; i < k.n_tanks; ++i) smart_only = smart_only && (k.ctrl[i] != AT_CTRL_BOT || k.bot_type[i] == AT_BOT_SMART);
    int spec = (k.mode == AT_MODE_TEAM && smart_only && k.map != AT_MAP_MAZE && k.terminate_time == 0) ? 1 : 0;
    if (spec == 1 && k.n_tanks == 4) spec = 2;
    if (spec == 2 && k.ctrl[0] == AT_CTRL_AGENT && k.ctrl[1] == AT_CTRL_AGENT && k.ctrl[2] == AT_CTRL_BOT &&
        k.ctrl[3] == AT_CTRL_BOT && k.team[0] == k.team[1] && k.team[2] == k.team[3] && k.team[0] != k.team[2])
        spec = 3;
    // the 1v1 wrappers on a fixed map: agent mode (4), bot_agent against the smart bot (5); any cap below 3 = generic
    if (k.n_tanks == 2 && k.map != AT_MAP_MAZE && !k.team_layout) {
        if (k.mode == AT_MODE_AGENT && k.ctrl[0] == AT_CTRL_AGENT && k.ctrl[1] == AT_CTRL_AGENT) return cap >= 3 ? 4 : 0;
        if (k.mode == AT_MODE_BOT_AGENT && k.ctrl[0] == AT_CTRL_BOT && k.bot_type[0] == AT_BOT_SMART && k.ctrl[1] == AT_CTRL_AGENT)
            return cap >= 3 ? 5 : 0;
    }
    return spec < cap ? spec : cap;
}

// One arena for every SoA column (HBM on the GPU, malloc in the emulator).
struct ArenaLayout {
    size_t o_cnt, o_rng, o_tick, o_tstep, o_ep, o_zones, o_prev, o_change, o_wlo, o_whi, o_tx, o_ty, o_trew, o_tpack,
        o_thits, o_tshot, o_bpack, o_bf0, o_bf1, o_bx, o_by, o_bm, o_trig, o_corn, o_pend, o_udir, o_bfs, o_tmpl, o_act, o_rw, o_dn, o_ttick, o_twin, o_thit, o_dyn, o_prevrw, o_atab, bytes;
};
inline ArenaLayout arena_layout(const KCfg &k, int64_t N) {
    ArenaLayout L;
    size_t bytes = 0;
    auto take = [&](size_t b) { size_t o = bytes; bytes += (b + 255) & ~(size_t)255; return o; };
    const int n = k.n_tanks, slots = SLOTS_PER_TANK * n;
    L.o_cnt = take(4 * N); L.o_rng = take(4 * N); L.o_tick = take(4 * N); L.o_tstep = take(4 * N); L.o_ep = take(4 * N);
    L.o_zones = take(4 * N); L.o_prev = take(4 * N); L.o_change = take(4 * N * n); L.o_wlo = take(8 * N);
    L.o_whi = take(8 * N); L.o_tx = take(8 * N * n); L.o_ty = take(8 * N * n); L.o_trew = take(8 * N * n);
    L.o_tpack = take(4 * N * n); L.o_thits = take(4 * N * n); L.o_tshot = take(4 * N * n); L.o_bpack = take(4 * N * n);
    L.o_bf0 = take(8 * N * n); L.o_bf1 = take(8 * N * n); L.o_bx = take(8 * N * slots); L.o_by = take(8 * N * slots);
    L.o_bm = take(8 * N * slots); L.o_trig = take(722 * 8); L.o_corn = take(361 * 4 * 8); L.o_pend = take(PEND_LUT * 8);
    L.o_udir = take(364);
    L.o_bfs = take(CELLS * CELLS * 2);
    L.o_tmpl = take((4 * 72 + CELLS) * 4);
    L.o_act = take(4 * N * 3 * (k.act_rows ? k.act_rows : 1)); L.o_rw = take(4 * N * (k.rows ? k.rows : 1));
    L.o_dn = take(N);
    L.o_ttick = take(4 * N); L.o_twin = take(4 * N); L.o_thit = take(4 * N * n);
    L.o_dyn = take(64);
    L.o_atab = take(361 * 5 * 8);
    L.o_prevrw = take(4 * N * (k.rows ? k.rows : 1));
    L.bytes = bytes;
    return L;
}
inline void bind_state(DevState &st, char *b, const ArenaLayout &L, int64_t N) {
    st.N = N;
    st.cnt = (uint32_t *)(b + L.o_cnt); st.rng_ctr = (uint32_t *)(b + L.o_rng); st.tick = (uint32_t *)(b + L.o_tick);
    st.tstep = (uint32_t *)(b + L.o_tstep); st.epstep = (uint32_t *)(b + L.o_ep); st.zones = (uint32_t *)(b + L.o_zones);
    st.prev_act = (uint32_t *)(b + L.o_prev); st.change_time = (int32_t *)(b + L.o_change);
    st.wall_lo = (uint64_t *)(b + L.o_wlo); st.wall_hi = (uint64_t *)(b + L.o_whi);
    st.tx = (double *)(b + L.o_tx); st.ty = (double *)(b + L.o_ty); st.trew = (double *)(b + L.o_trew);
    st.tpack = (uint32_t *)(b + L.o_tpack); st.thits = (uint32_t *)(b + L.o_thits); st.tshot = (int32_t *)(b + L.o_tshot);
    st.bpack = (uint32_t *)(b + L.o_bpack); st.bf0 = (double *)(b + L.o_bf0); st.bf1 = (double *)(b + L.o_bf1);
    st.bx = (double *)(b + L.o_bx); st.by = (double *)(b + L.o_by); st.bmeta = (uint64_t *)(b + L.o_bm);
    st.trig = (const double *)(b + L.o_trig); st.corn = (const double *)(b + L.o_corn);
    st.pend = (const double *)(b + L.o_pend);
    st.term_tick = (uint32_t *)(b + L.o_ttick); st.term_winner = (uint32_t *)(b + L.o_twin);
    st.term_hits = (uint32_t *)(b + L.o_thit);
    st.udelta = (const uint8_t *)(b + L.o_udir);
    st.bfs_tab = nullptr;  // set by the owner for fixed maps (at_sim.h fill_bfs_table)
    st.prev_rew = (float *)(b + L.o_prevrw);
    st.atab = (const double *)(b + L.o_atab);   // the owner fills it (at_sim.h fill_atan_table)
    st.dyn = (const double *)(b + L.o_dyn);   // the owner writes dyn[0] = weakness
}

// gather / scatter one env between the SoA columns and a PackedEnv (kernel bodies and emulator loops)
AT_HD void gather_env(const KCfg &c, const DevState &st, int64_t e, PackedEnv &p) {
    p.cnt = st.cnt[e]; p.rng = st.rng_ctr[e]; p.tick = st.tick[e]; p.tstep = st.tstep[e]; p.epstep = st.epstep[e];
    p.zones = st.zones[e]; p.prev_act = st.prev_act[e];
    p.wlo = c.map == 2 ? st.wall_lo[e] : c.wall_lo;
    p.whi = c.map == 2 ? st.wall_hi[e] : c.wall_hi;
    for (int i = 0; i < c.n_tanks; ++i) {
        int64_t g = (int64_t)i * st.N + e;
        p.tx[i] = st.tx[g]; p.ty[i] = st.ty[g]; p.trew[i] = st.trew[g]; p.bf0[i] = st.bf0[g]; p.bf1[i] = st.bf1[g];
        p.tpack[i] = st.tpack[g]; p.thits[i] = st.thits[g]; p.bpack[i] = st.bpack[g]; p.tshot[i] = st.tshot[g];
        p.change_time[i] = st.change_time[g];
    }
    for (uint32_t s = 0; s < p.cnt; ++s) {
        int64_t g = (int64_t)s * st.N + e;
        p.bx[s] = st.bx[g]; p.by[s] = st.by[g]; p.bmeta[s] = st.bmeta[g];
    }
}
AT_HD void scatter_env(const KCfg &c, const DevState &st, int64_t e, const PackedEnv &p) {
    st.cnt[e] = p.cnt; st.rng_ctr[e] = p.rng; st.tick[e] = p.tick; st.tstep[e] = p.tstep; st.epstep[e] = p.epstep;
    st.zones[e] = p.zones; st.prev_act[e] = p.prev_act;
    if (c.map == 2) { st.wall_lo[e] = p.wlo; st.wall_hi[e] = p.whi; }
    for (int i = 0; i < c.n_tanks; ++i) {
        int64_t g = (int64_t)i * st.N + e;
        st.tx[g] = p.tx[i]; st.ty[g] = p.ty[i]; st.trew[g] = p.trew[i]; st.bf0[g] = p.bf0[i]; st.bf1[g] = p.bf1[i];
        st.tpack[g] = p.tpack[i]; st.thits[g] = p.thits[i]; st.bpack[g] = p.bpack[i]; st.tshot[g] = p.tshot[i];
        st.change_time[g] = p.change_time[i];
    }
    for (uint32_t s = 0; s < p.cnt; ++s) {
        int64_t g = (int64_t)s * st.N + e;
        st.bx[g] = p.bx[s]; st.by[g] = p.by[s]; st.bmeta[g] = p.bmeta[s];
    }
}

inline void unpack_env(const KCfg &k, const HostLuts &l, const PackedEnv &p, at_env_state &o) {
        memset(&o, 0, sizeof o);
        o.n_bullets = (int32_t)p.cnt; o.tick = (int32_t)p.tick; o.training_step = (int32_t)p.tstep;
        o.episode_steps = (int32_t)p.epstep; o.rng_ctr = p.rng; o.n_walls = k.n_walls;
        for (int z = 0; z < 4; ++z) {
            int cell = (int)((p.zones >> (7 * z)) & 127u);
            bool on = z < 2 ? k.buff_on : k.debuff_on;
            o.zones[2 * z] = on ? (cell % 11) * 70 : 0;
            o.zones[2 * z + 1] = on ? (cell / 11) * 70 : 0;
        }
        for (int i = 0; i < CELLS; ++i) o.maze[i] = (uint8_t)((i < 64 ? p.wlo >> i : p.whi >> (i - 64)) & 1ull);
        for (int i = 0; i < k.n_tanks; ++i) {
            at_tank_state &t = o.tanks[i];
            const uint32_t pk = p.tpack[i];
            t.x = p.tx[i]; t.y = p.ty[i]; t.reward = p.trew[i];
            t.angle = (int32_t)(pk & 511u);
            const uint32_t sc = (pk >> 9) & 3u;
            t.speed = sc == 0 ? 0 : (sc == 1 ? 10 : (sc == 2 ? -10 : 1));
            t.alive = (pk & TP_ALIVE) != 0; t.hitting_wall = (pk & TP_HW) != 0; t.in_buff = (pk & TP_INBUFF) != 0;
            t.in_debuff = (pk & TP_INDEBUFF) != 0; t.max_bullets = (pk & TP_MAXB1) ? 1 : 6;
            t.last_shot_ms = p.tshot[i]; t.num_hit = (int32_t)(p.thits[i] & 0xFFFFu);
            t.num_be_hit = (int32_t)(p.thits[i] >> 16);
            at_bot_state &b = o.bots[i];
            const uint32_t bp = p.bpack[i];
            b.target = -1; b.rot_dir = 1; b.cur_action[0] = 1; b.cur_action[1] = 1;
            if (k.ctrl[i] != AT_CTRL_BOT) continue;
            switch (k.bot_type[i]) {
            case AT_BOT_SMART:
                b.target = (int32_t)(bp & 15u) - 1; b.aiming = (bp >> 4) & 1u; b.rot_dir = ((bp >> 5) & 1u) ? 1 : -1;
                b.stuck_timer = (bp >> 6) & 63u; b.has_next = (bp >> 12) & 1u; b.next_r = (bp >> 13) & 15u;
                b.next_c = (bp >> 17) & 15u; b.last_x = p.bf0[i]; b.last_y = p.bf1[i];
                break;
            case AT_BOT_RANDOM:
                b.action_delay = bp & 63u; b.cur_action[0] = (bp >> 6) & 3u; b.cur_action[1] = (bp >> 8) & 3u;
                b.cur_action[2] = (bp >> 10) & 1u;
                break;
            case AT_BOT_DEFENSIVE:
                b.has_tpos = bp & 1u;
                if (b.has_tpos) {
                    int zi = (bp >> 1) & 1u;
                    int cell = (int)((p.zones >> (7 * zi)) & 127u);
                    double zx = (cell % 11) * 70, zy = (cell / 11) * 70;
                    double x0 = zx + 22.5, x1 = zx + 245.0 - 22.5, y0 = zy + 18.0, y1 = zy + 245.0 - 18.0;
                    b.tpos_x = (x0 + x1) / 2; b.tpos_y = (y0 + y1) / 2;
                }
                break;
            case AT_BOT_DODGE:
                b.dodge_timer = bp & 31u; b.dodge_state = (bp >> 5) & 3u; b.has_dodge_angle = (bp >> 7) & 1u;
                b.dodge_angle = p.bf0[i];
                break;
            default: break;
            }
        }
        for (uint32_t s = 0; s < p.cnt && s < AT_MAX_BULLETS; ++s) {
            at_bullet_state &bl = o.bullets[s];
            const uint64_t m = p.bmeta[s];
            const int ang = (int)((m >> BM_ANGLE) & 511u);
            const double cv = l.trig[2 * ang], sv = l.trig[2 * ang + 1];
            bl.x = p.bx[s]; bl.y = p.by[s];
            bl.dx = ((m >> BM_FLIPX) & 1u) ? -cv : cv;
            bl.dy = ((m >> BM_FLIPY) & 1u) ? sv : -sv;
            bl.owner = (int32_t)(m & 15u); bl.distance = (int32_t)((m >> BM_DIST) & 511u);
            bl.bounces = (int32_t)((m >> BM_BOUNCES) & 7u); bl.cleared = (int32_t)((m >> BM_CLEARED) & 1u);
            bl.pending = l.pend[(m >> BM_PEND) & 0xFFFFu];
        }
}

inline int pend_count_of(const HostLuts &l, double v) {
    for (int i = 0; i < PEND_LUT; ++i)
        if (l.pend[i] == v) return i;
    return -1;
}
// returns "" or an error message
inline std::string pack_env(const KCfg &k, const HostLuts &l, const at_env_state &o, PackedEnv &p) {
        memset(&p, 0, sizeof p);
        if (o.n_bullets < 0 || o.n_bullets > SLOTS_PER_TANK * k.n_tanks)
            return "at_set_state: n_bullets out of range";
        p.cnt = (uint32_t)o.n_bullets; p.rng = o.rng_ctr; p.tick = (uint32_t)o.tick; p.tstep = (uint32_t)o.training_step;
        p.epstep = (uint32_t)o.episode_steps;
        for (int z = 0; z < 4; ++z) {
            int zx = o.zones[2 * z], zy = o.zones[2 * z + 1];
            if (zx % 70 || zy % 70 || zx < 0 || zx > 700 || zy < 0 || zy > 700)
                return "at_set_state: zone corners must be cell corners (multiples of 70)";
            p.zones |= (uint32_t)((zy / 70) * 11 + zx / 70) << (7 * z);
        }
        for (int i = 0; i < CELLS; ++i)
            if (o.maze[i]) { if (i < 64) p.wlo |= 1ull << i; else p.whi |= 1ull << (i - 64); }
        if (k.map != AT_MAP_MAZE && (p.wlo != k.wall_lo || p.whi != k.wall_hi))
            return "at_set_state: maze differs from the handle's fixed map";
        uint32_t live[MAX_TANKS] = {0};
        for (int i = 0; i < k.n_tanks; ++i) {
            const at_tank_state &t = o.tanks[i];
            if (t.angle < 0 || t.angle > 360) return "at_set_state: angle must be 0..360";
            uint32_t sc;
            switch (t.speed) { case 0: sc = 0; break; case 10: sc = 1; break; case -10: sc = 2; break; case 1: sc = 3; break;
            default: return "at_set_state: speed must be one of 0, 10, -10, 1"; }
            p.tx[i] = t.x; p.ty[i] = t.y; p.trew[i] = t.reward;
            p.tpack[i] = (uint32_t)t.angle | (sc << 9) | (t.alive ? TP_ALIVE : 0u) | (t.hitting_wall ? TP_HW : 0u) |
                         (t.in_buff ? TP_INBUFF : 0u) | (t.in_debuff ? TP_INDEBUFF : 0u) |
                         (t.max_bullets == 1 ? TP_MAXB1 : 0u);
            p.tshot[i] = t.last_shot_ms;
            p.thits[i] = ((uint32_t)t.num_hit & 0xFFFFu) | ((uint32_t)t.num_be_hit << 16);
            const at_bot_state &b = o.bots[i];
            if (k.ctrl[i] != AT_CTRL_BOT) continue;
            switch (k.bot_type[i]) {
            case AT_BOT_SMART:
                p.bpack[i] = (uint32_t)(b.target + 1) | ((uint32_t)(b.aiming != 0) << 4) | ((uint32_t)(b.rot_dir > 0) << 5) |
                             ((uint32_t)b.stuck_timer << 6) | ((uint32_t)(b.has_next != 0) << 12) |
                             ((uint32_t)b.next_r << 13) | ((uint32_t)b.next_c << 17);
                p.bf0[i] = b.last_x; p.bf1[i] = b.last_y;
                break;
            case AT_BOT_RANDOM:
                p.bpack[i] = (uint32_t)b.action_delay | ((uint32_t)b.cur_action[0] << 6) | ((uint32_t)b.cur_action[1] << 8) |
                             ((uint32_t)b.cur_action[2] << 10);
                break;
            case AT_BOT_DEFENSIVE: {
                uint32_t zi = 0;
                if (b.has_tpos) {  // which buff zone's safe centre was cached (simple_bots.py:151-200)
                    int cell0 = (int)(p.zones & 127u);
                    double zx = (cell0 % 11) * 70, zy = (cell0 / 11) * 70;
                    double x0 = zx + 22.5, x1 = zx + 245.0 - 22.5, y0 = zy + 18.0, y1 = zy + 245.0 - 18.0;
                    zi = ((x0 + x1) / 2 == b.tpos_x && (y0 + y1) / 2 == b.tpos_y) ? 0u : 1u;
                }
                p.bpack[i] = (uint32_t)(b.has_tpos != 0) | (zi << 1);
                break;
            }
            case AT_BOT_DODGE:
                p.bpack[i] = (uint32_t)b.dodge_timer | ((uint32_t)b.dodge_state << 5) | ((uint32_t)(b.has_dodge_angle != 0) << 7);
                p.bf0[i] = b.dodge_angle;
                break;
            default: break;
            }
        }
        for (int s = 0; s < o.n_bullets; ++s) {
            const at_bullet_state &bl = o.bullets[s];
            if (bl.owner < 0 || bl.owner >= k.n_tanks) return "at_set_state: bullet owner out of range";
            int ang = -1, fx = 0, fy = 0;
            for (int a = 0; a <= 360 && ang < 0; ++a) {  // bullets only ever carry (+-cos a, -+sin a)
                const double cv = l.trig[2 * a], sv = l.trig[2 * a + 1];
                for (int f = 0; f < 4; ++f) {
                    double dx = (f & 1) ? -cv : cv, dy = (f & 2) ? sv : -sv;
                    if (dx == bl.dx && dy == bl.dy) { ang = a; fx = f & 1; fy = (f >> 1) & 1; break; }
                }
            }
            if (ang < 0) return "at_set_state: bullet direction is not (+-cos a, -+sin a) of an integer angle";
            int pc = pend_count_of(l, bl.pending);
            if (pc < 0) return "at_set_state: pending penalty is not a k-fold sum of 0.02";
            uint32_t rank = live[bl.owner]++;
            p.bx[s] = bl.x; p.by[s] = bl.y;
            p.bmeta[s] = (uint64_t)bl.owner | ((uint64_t)ang << BM_ANGLE) | ((uint64_t)fx << BM_FLIPX) |
                         ((uint64_t)fy << BM_FLIPY) | ((uint64_t)(bl.bounces & 7) << BM_BOUNCES) |
                         ((uint64_t)(bl.cleared != 0) << BM_CLEARED) | ((uint64_t)(bl.distance & 511) << BM_DIST) |
                         ((uint64_t)rank << BM_RANK) | ((uint64_t)pc << BM_PEND);
        }
    return "";
}


}  // namespace at
