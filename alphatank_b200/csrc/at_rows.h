// at_rows.h - observation rows rebuilt from the STORED state (the job functions of rows_kernel, at_kernels.cu).
//
// gym_env.py:105-183 (1v1 layout) / gym_env_multi.py:225-304 (team layout).  Sim::emit_rows (at_sim.h) walks a row in
// layout order, one thread per env; here the same values are produced by independent JOBS that write straight into
// an image of the env's R x D floats, so that a block can build the images of a group of envs with all its threads
// and ship them with one bulk-async copy:
//   tank job (row r, tank i)        the row's own block (+ zone corners + flags) when i is the row's tank, else
//                                   the "other tank" block of tank i
//   position job (owner o, rank k)  bullet k (firing order) of tank o, written into EVERY row: 4 values into the
//                                   row's own-bullet block if o is the row's tank, else 7 / 8 values into the block
//                                   of the o-th other tank; zeros when the owner has fewer live bullets
//   maze job (cell)                 per-env maps only: the maze column of the cell and, for a wall, its rectangle
// The arithmetic is the emitter's (same f64 expressions, same casts), so both produce identical bits; the host
// emulator (tests/cpu_emul) runs these jobs against the oracle, the GPU tests run the kernel.
#pragma once
#include "at_types.h"

namespace at {

constexpr uint64_t ROWS_EMPTY = ~0ull;  // record-table marker: (owner, rank) has no live bullet (bmeta bits 48.. are 0)

AT_HD int rows_speed(uint32_t pack) {
    const uint32_t s = (pack >> 9) & 3u;
    return s == 0 ? 0 : (s == 1 ? 10 : (s == 2 ? -10 : 1));
}

// Compile-time shapes: NT tanks, NR rows, TEAM layout (1 / 0); 0 / 0 / -1 = read them from the config (generic build).
template <int NT, int NR, int TEAM>
struct RowsShape {
    AT_HD static int n(const KCfg &c) { return NT ? NT : c.n_tanks; }
    AT_HD static int rows(const KCfg &c) { return NR ? NR : c.rows; }
    AT_HD static bool team(const KCfg &c) { return TEAM >= 0 ? TEAM != 0 : c.team_layout != 0; }
};

// tank job (row r, tank i): x, y, the 8 corner coordinates, (vx, vy) and hittingWall of tank i (gym_env.py:121-131 /
// gym_env_multi.py:241-252) - as the row's own block (+ zone corners + flags) when i is the row's tank, else as the
// "other tank" block of tank i.  The kernel orders the jobs so that a warp holds one kind.
// tx / ty / tp: the env's n tanks (stride `ts` between tanks); img: the env's R x D image
template <int NT, int NR, int TEAM>
AT_HD void rows_tank_job(const KCfg &c, const double *trig, const double *corn, int r, int i, const double *tx,
                         const double *ty, const uint32_t *tp, int ts, uint32_t zones, float *img) {
    typedef RowsShape<NT, NR, TEAM> SH;
    const bool team = SH::team(c);
    const double x = tx[i * ts], y = ty[i * ts];
    const uint32_t pack = tp[i * ts];
    const float alive = (pack & TP_ALIVE) ? 1.0f : 0.0f;
    const int t = team ? c.agent_tank[r] : r;
    float *row = img + (int64_t)r * c.obs_dim;
    float *o = row;          // x, y, [rx, ry, dist,] 11 common values, [1.0 | same_team,] alive
    float *cm = row + 2;
    if (i != t) {
        o = row + c.off_et + (i < t ? i : i - 1) * c.et_len;
        cm = o + 5;
        const double rx = x - tx[t * ts], ry = y - ty[t * ts];
        o[2] = (float)rx; o[3] = (float)ry;
        o[4] = (float)sqrt(rx * rx + ry * ry);
        if (team) o[16] = c.team[i] == c.team[t] ? 1.0f : 0.0f;
        o[c.et_len - 1] = alive;
    } else {
        if (team) row[13] = 1.0f;
        row[c.self_len - 1] = alive;
#pragma unroll
        for (int z = 0; z < 2; ++z) {   // zone corners, buff first (gym_env.py:168-173)
            if (!(z == 0 ? c.buff_on : c.debuff_on)) continue;
            float *d = row + c.off_zones + ((z == 1 && c.buff_on) ? 4 : 0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cell = (int)((zones >> (14 * z + 7 * h)) & 127u);
                const int cr = cell / 11;
                d[2 * h] = (float)((cell - 11 * cr) * 70);
                d[2 * h + 1] = (float)(cr * 70);
            }
        }
        row[c.off_flags] = (pack & TP_INBUFF) ? 1.0f : 0.0f;
        row[c.off_flags + 1] = (pack & TP_INDEBUFF) ? 1.0f : 0.0f;
    }
    o[0] = (float)x; o[1] = (float)y;
    const int a = (int)(pack & 511u);
    const double sp = (double)rows_speed(pack);
    const double *co = corn + 4 * a;
    const double c0 = co[0], c1 = co[1], c2 = co[2], c3 = co[3];
    cm[0] = (float)(x + c0); cm[1] = (float)(y + c1); cm[2] = (float)(x + c2); cm[3] = (float)(y + c3);
    cm[4] = (float)(x - c0); cm[5] = (float)(y - c1); cm[6] = (float)(x - c2); cm[7] = (float)(y - c3);
    cm[8] = (float)(sp * trig[2 * a]);
    cm[9] = (float)(sp * trig[2 * a + 1]);
    cm[10] = (pack & TP_HW) ? 1.0f : 0.0f;
}

// rotate-left of a small register array by a lane-dependent amount (selects only, constant indices): lanes that write
// the same block of different envs then start at different elements, so their shared-memory stores hit different banks
template <int LEN>
AT_HD void rows_rotl(float (&v)[LEN], int r) {
#pragma unroll
    for (int s = 1; s < LEN; s <<= 1) {
        const bool b = (r & s) != 0;
        float w[LEN];
#pragma unroll
        for (int q = 0; q < LEN; ++q) w[q] = b ? v[(q + s) & (LEN - 1)] : v[q];
#pragma unroll
        for (int q = 0; q < LEN; ++q) v[q] = w[q];
    }
}

// position job: (m, bx, by) = record of tank o's k-th live bullet, m == ROWS_EMPTY when it has none (zeros are
// written).  `rot`: store-order rotation (any value gives the same image; the kernel passes the env's index in the group).
template <int NT, int NR, int TEAM>
AT_HD void rows_position_job(const KCfg &c, const double *trig, int o, int k, uint64_t m, double bx, double by,
                             const double *tx, const double *ty, int ts, float *img, int rot = 0) {
    typedef RowsShape<NT, NR, TEAM> SH;
    const bool team = SH::team(c);
    const bool has = m != ROWS_EMPTY;
    double x = 0.0, y = 0.0, dx = 0.0, dy = 0.0;
    if (has) {
        x = bx;
        y = by;
        const int ang = (int)((m >> BM_ANGLE) & 511u);
        const double cv = trig[2 * ang], sv = trig[2 * ang + 1];
        dx = ((m >> BM_FLIPX) & 1u) ? -cv : cv;
        dy = ((m >> BM_FLIPY) & 1u) ? sv : -sv;
    }
    const float fx = (float)x, fy = (float)y, fdx = (float)dx, fdy = (float)dy;
    const int per = team ? 8 : 7;
    const int nrows = SH::rows(c);
#pragma unroll
    for (int r = 0; r < nrows; ++r) {
        const int t = team ? c.agent_tank[r] : r;
        float *row = img + (int64_t)r * c.obs_dim;
        if (o == t) {
            float *d = row + c.off_ob + 4 * k;
            float v[4] = {fx, fy, fdx, fdy};
            rows_rotl<4>(v, rot);
#pragma unroll
            for (int q = 0; q < 4; ++q) d[(q + rot) & 3] = v[q];
        } else {
            float *d = row + c.off_eb + (o < t ? o : o - 1) * c.eb_stride + k * per;
            const double rx = x - tx[t * ts], ry = y - ty[t * ts];
            float v[8] = {fx, fy, fdx, fdy, has ? (float)rx : 0.0f, has ? (float)ry : 0.0f,
                          has ? (float)sqrt(rx * rx + ry * ry) : 0.0f, has && c.team[o] == c.team[t] ? 1.0f : 0.0f};
            if (team) {   // 8 values per bullet
                rows_rotl<8>(v, rot);
#pragma unroll
                for (int q = 0; q < 8; ++q) d[(q + rot) & 7] = v[q];
            } else {      // 7 values (1v1 layout has no same-team flag)
#pragma unroll
                for (int q = 0; q < 7; ++q) d[q] = v[q];
            }
        }
    }
}

// per-env maps (map == MAZE): maze column of `cell` and the rectangle of the wall in it, into every row
AT_HD void rows_maze_job(const KCfg &c, uint64_t wlo, uint64_t whi, int cell, float *img) {
    const bool wall = ((cell < 64 ? wlo >> cell : whi >> (cell - 64)) & 1ull) != 0;
    int k = 0;  // walls before this cell in row-major order
    if (wall) {
#ifdef __CUDA_ARCH__
        k = cell < 64 ? __popcll(wlo & ((1ull << cell) - 1ull)) : __popcll(wlo) + __popcll(whi & ((1ull << (cell - 64)) - 1ull));
#else
        k = cell < 64 ? __builtin_popcountll(wlo & ((1ull << cell) - 1ull))
                      : __builtin_popcountll(wlo) + __builtin_popcountll(whi & ((1ull << (cell - 64)) - 1ull));
#endif
    }
    const float X = (float)((cell % 11) * 70), Y = (float)((cell / 11) * 70);
    for (int r = 0; r < c.rows; ++r) {
        float *row = img + (int64_t)r * c.obs_dim;
        row[c.off_maze + cell] = wall ? 1.0f : 0.0f;
        if (wall) {
            float *d = row + c.off_walls + 4 * k;
            d[0] = X; d[1] = Y; d[2] = X + 70.0f; d[3] = Y + 70.0f;
        }
    }
}

}  // namespace at
