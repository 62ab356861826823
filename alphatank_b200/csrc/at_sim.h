// at_sim.h - simulation of one alphaTank tick, one thread (lane) per env; the smart bots' line-of-sight queries of a
// warp's 32 envs are pooled into a job queue that all lanes work through together (pool_los).
// Written as __host__ __device__ code against a warp policy (at_warp.h): the CUDA kernels in at_kernels.cu run it
// on the warp intrinsics; tests/cpu_emul compiles the same header with g++ and runs it on a 32-fiber SIMT emulator
// to check the formulation against the oracle in a GPU-less container (test infrastructure only - the product has
// no CPU path).
//
// This is NOT a transliteration of the reference's Python: walls are a 121-bit mask probed by cell, angles
// are integers indexing host-built libm tables, bullets carry a packed 64-bit record, BFS is a bit-parallel
// 4-label frontier sweep, the pending "away" penalty is a counter.  Each routine cites the reference lines
// whose results it must reproduce bit-for-bit.
//
// Numerics: compile with -fmad=false (nvcc) / -ffp-contract=off (g++).  The only fused multiply-adds are
// the explicit fma_rn() calls that restate numpy/OpenBLAS 2-element dot products (see oracle/tank_oracle.c).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "at_types.h"
#include "at_warp.h"

namespace at {

constexpr double PI = 3.141592653589793;
constexpr double GRID = 70.0;

AT_HD double fma_rn(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
AT_HD int d2i(double x) {  // C (int)double: truncation toward zero, as pygame.Rect does with float args
#ifdef __CUDA_ARCH__
    return __double2int_rz(x);
#else
    return (int)x;
#endif
}

// --- exact multi-step accumulation --------------------------------------------------------------------------
// The reference advances a (hypothetical) bullet by x = x + d once per sub-step.  While x stays inside one
// binade [2^e, 2^(e+1)) every representable x is a multiple of u = 2^(e-52); writing d = Q + r with Q the
// multiple of u that fl(x + d) - x yields, the sum x + d rounds to x + Q for EVERY x of the binade as long
// as |r| < u/2 (no tie), so m sequential additions give exactly x + m*Q, which one fma computes without
// rounding (the result is a multiple of u inside the binade).  jump_axis() applies that and falls back to
// the literal loop whenever a precondition fails, so it is bit-identical to m sequential additions.
AT_HD uint64_t d_bits(double v) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(v);
#else
    uint64_t b;
    memcpy(&b, &v, 8);
    return b;
#endif
}
AT_HD double bits_d(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}
AT_HD int d_expo(double v) { return (int)((d_bits(v) >> 52) & 0x7FFu); }
AT_HD double jump_axis(double x, double d, int m) {
    const double x1 = x + d;
    const double Q = x1 - x;
    const double r = d - Q;
    const int e = d_expo(x);
    const double xm = fma_rn((double)m, Q, x);
    const bool same_sign = (((d_bits(x) ^ d_bits(x1)) | (d_bits(x) ^ d_bits(xm))) >> 63) == 0;
    const bool ok = same_sign && e > 53 && e < 0x7FF && d_expo(x1) == e && d_expo(xm) == e &&
                    fabs(r) < bits_d((uint64_t)(e - 53) << 52);
    if (ok) return xm;
    for (int j = 0; j < m; ++j) x = x + d;
    return x;
}
template <class T>
AT_HD T ldg(const T *p) {  // read-only tables: lets the compiler hoist / batch the loads
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
AT_HD int fdiv70(int v) { return v >= 0 ? v / 70 : -((69 - v) / 70); }
AT_HD int imin(int a, int b) { return a < b ? a : b; }
AT_HD int imax(int a, int b) { return a > b ? a : b; }

// ------------------------------------------------------------------ ATRNG (oracle/atrng.py)
AT_HD void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
AT_HD uint32_t rng_word(uint64_t seed, int64_t env_id, uint32_t stream, uint32_t ctr) {
    uint32_t c0 = ctr, c1 = stream, c2 = (uint32_t)((uint64_t)env_id & 0xFFFFFFFFu),
             c3 = (uint32_t)((uint64_t)env_id >> 32);
    philox4x32_10(c0, c1, c2, c3, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32));
    return c0;
}

// ------------------------------------------------------------------ Python float helpers
AT_HD double py_fmod360(double v) {  // CPython float_rem with divisor 360.0
    double m = fmod(v, 360.0);
    if (m != 0.0) {
        if (m < 0) m += 360.0;
    } else {
        m = 0.0;
    }
    return m;
}
AT_HD double py_radians(double x) { return x * (PI / 180.0); }
AT_HD double py_degrees(double x) { return x * (180.0 / PI); }
AT_HD double np_dot2(double a0, double a1, double b0, double b1) { return fma_rn(a1, b1, a0 * b0); }
AT_HD double np_norm2(double a0, double a1) { return sqrt(np_dot2(a0, a1, a0, a1)); }
// Exact squared-distance forms of the reference's distance compares (sqrt is correctly rounded and monotonic):
//   sqrt(s) > 300  <=>  s > S300   (S300 = largest double whose root rounds to <= 300)
//   sqrt(s) < 100  <=>  s < 10000  (the double just below 10000 already roots to < 100)
constexpr double S300 = 0x1.5f90000000001p+16;
constexpr double S100 = 10000.0;
AT_HD double fast_rsqrt(double s) {
#ifdef __CUDA_ARCH__
    return rsqrt(s);
#else
    return 1.0 / sqrt(s);
#endif
}

// ---- correctly rounded atan2 near integer-degree bearings (the slow path of Sim::angle_diff_to).
// A bearing feeds integer-degree thresholds.  Bearings that are integers up to rounding are NOT measure-zero here:
// two tanks spawn in the same cell and one drives off along its (integer) heading, so the other sees it at exactly
// heading + 180 and |bearing - own angle| sits on a threshold up to the last bits of atan2.  The device atan2 is
// only good to 2 ulp, so such decisions are re-taken on the correctly rounded value: from theta0 = the device's
// atan2, one Newton step delta = (y cos theta0 - x sin theta0) / (x cos theta0 + y sin theta0), with sin / cos of
// theta0 = phi_K + d built by angle addition from a host table of 106-bit sin / cos of the K-degree angles phi_K
// (d < 1e-10, so first / second order terms in plain doubles suffice), and theta0 + delta rounded once.
// (Python's math.atan2 is glibc's: correctly rounded for 99.6 % of such inputs, measured against mpmath in
// tests/test_kernel_logic_cpu.py; the remaining inputs are where a libm-dependent reference has no single right
// answer - DESIGN.md "Transcendentals".)
struct DD {
    double hi, lo;
};
AT_HD DD dd_two_sum(double a, double b) {
    const double s = a + b, bb = s - a;
    return {s, (a - (s - bb)) + (b - bb)};
}
AT_HD DD dd_fast_two_sum(double a, double b) {  // |a| >= |b|
    const double s = a + b;
    return {s, b - (s - a)};
}
AT_HD DD dd_add(DD a, DD b) {
    DD s = dd_two_sum(a.hi, b.hi);
    const DD t = dd_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = dd_fast_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return dd_fast_two_sum(s.hi, s.lo);
}
AT_HD DD dd_neg(DD a) { return {-a.hi, -a.lo}; }
AT_HD DD dd_mul(DD a, DD b) {
    const double p = a.hi * b.hi;
    double e = fma_rn(a.hi, b.hi, -p);
    e = fma_rn(a.hi, b.lo, e);
    e = fma_rn(a.lo, b.hi, e);
    return dd_fast_two_sum(p, e);
}
AT_HD DD dd_mul_d(DD a, double b) {
    const double p = a.hi * b;
    double e = fma_rn(a.hi, b, -p);
    e = fma_rn(a.lo, b, e);
    return dd_fast_two_sum(p, e);
}
AT_HD DD dd_div_d(DD a, double d) {  // d: a small exact integer
    const double q1 = a.hi / d;
    const double p = q1 * d;
    const double r = ((a.hi - p) - fma_rn(q1, d, -p)) + a.lo;   // a - q1 d
    return dd_fast_two_sum(q1, r / d);
}
// host side: 106-bit sin / cos of a double by Taylor sums (|t| <= pi; the table of the K-degree angles below)
inline void dd_sincos_host(double t, DD &sn, DD &cs) {
    const DD x = {t, 0.0};
    const DD x2 = dd_mul(x, x);
    DD ts = x, tc = {1.0, 0.0};
    sn = ts; cs = tc;
    for (int k = 1; k <= 34; ++k) {
        tc = dd_neg(dd_div_d(dd_mul(tc, x2), (double)((2 * k - 1) * (2 * k))));
        ts = dd_neg(dd_div_d(dd_mul(ts, x2), (double)((2 * k) * (2 * k + 1))));
        cs = dd_add(cs, tc);
        sn = dd_add(sn, ts);
    }
}
constexpr int ATAB_STRIDE = 5;      // per K = -180 .. 180: phi_K, sin hi / lo, cos hi / lo
inline void fill_atan_table(double *tab) {
    for (int K = -180; K <= 180; ++K) {
        const double phi = (double)K * (3.141592653589793 / 180.0);
        DD sn, cs;
        dd_sincos_host(phi, sn, cs);
        double *r = tab + (K + 180) * ATAB_STRIDE;
        r[0] = phi; r[1] = sn.hi; r[2] = sn.lo; r[3] = cs.hi; r[4] = cs.lo;
    }
}
// th0 = an atan2(y, x) good to a few ulp whose value in degrees is within 1e-9 of the integer K: the correctly
// rounded atan2(y, x)
AT_HD double atan2_refine(double y, double x, double th0, int K, const double *tab) {
    const double *r = tab + (K + 180) * ATAB_STRIDE;
    const double d = th0 - ldg(r);                       // exact: both within 1e-10 of each other
    const DD S = {ldg(r + 1), ldg(r + 2)}, Cc = {ldg(r + 3), ldg(r + 4)};
    const double h = 0.5 * d * d;
    const DD sn = dd_add(S, dd_two_sum(Cc.hi * d, fma_rn(Cc.hi, d, -(Cc.hi * d)) + Cc.lo * d - S.hi * h));   // S + C d - S d^2 / 2
    const DD cs = dd_add(Cc, dd_two_sum(-(S.hi * d), -fma_rn(S.hi, d, -(S.hi * d)) - S.lo * d - Cc.hi * h)); // C - S d - C d^2 / 2
    const DD num = dd_add(dd_mul_d(cs, y), dd_neg(dd_mul_d(sn, x)));   // y cos - x sin  (= |v| sin(theta - th0))
    const double den = x * cs.hi + y * sn.hi;                            // x cos + y sin  (= |v| cos(theta - th0))
    return dd_two_sum(th0, num.hi / den).hi;
}

// atan2 with the lattice cases pinned to the correctly rounded doubles glibc returns (axis-aligned and
// diagonal bearings: tanks spawn on cell centres).  Elsewhere the device atan2 (<= 2 ulp); Sim::angle_diff_to
// re-takes the decisions that could depend on those last bits with atan2_cr.
AT_HD double at_atan2(double y, double x) {
    double ax = fabs(x), ay = fabs(y);
    if (ay == 0.0) {
        if (x > 0.0 || (x == 0.0 && !signbit(x))) return y;
        return copysign(3.141592653589793, y);
    }
    if (ax == 0.0) return copysign(1.5707963267948966, y);
    if (ax == ay) return copysign(x > 0 ? 0.7853981633974483 : 2.356194490192345, y);
    return atan2(y, x);
}
// Out-of-line copies for the generic build (SimT<0>): it serves every bot type, and with these three helpers inlined at
// every call site its SASS is 800 KB (at_atan2 60 KB, Philox 54 KB, py_fmod360 42 KB) - instruction fetch, not arithmetic.
AT_HDN double at_atan2_ni(double y, double x) { return at_atan2(y, x); }
AT_HDN double cos_ni(double v) { return cos(v); }   // (the dodge bot's non-integer angles: fp64 libm bodies, ~4 KB each)
AT_HDN double sin_ni(double v) { return sin(v); }
AT_HDN double py_fmod360_ni(double v) { return py_fmod360(v); }
AT_HDN uint32_t rng_word_ni(uint64_t seed, int64_t env_id, uint32_t stream, uint32_t ctr) { return rng_word(seed, env_id, stream, ctr); }

// ------------------------------------------------------------------ bit-parallel BFS (env/bfs.py:6-56)
// FIFO BFS with neighbour order up, down, left, right.  Queue order within a level is the lexicographic
// order of the direction strings from the start, so a cell's parent is the neighbour with the smallest
// string and its first letter (the child of `start` the path goes through) is the minimum first letter
// among its previous-level neighbours.  Four 121-bit frontier sets, one per first letter, are expanded
// level by level; returns len(path) (0 = None) and path[1].
struct M128 {
    uint64_t lo, hi;
};
AT_HD M128 m_and(M128 a, M128 b) { return {a.lo & b.lo, a.hi & b.hi}; }
AT_HD M128 m_or(M128 a, M128 b) { return {a.lo | b.lo, a.hi | b.hi}; }
AT_HD M128 m_andn(M128 a, M128 b) { return {a.lo & ~b.lo, a.hi & ~b.hi}; }
AT_HD bool m_any(M128 a) { return (a.lo | a.hi) != 0; }
AT_HD bool m_test(M128 a, int i) { return ((i < 64 ? a.lo >> i : a.hi >> (i - 64)) & 1ull) != 0; }
AT_HD M128 m_bit(int i) { return i < 64 ? M128{1ull << i, 0} : M128{0, 1ull << (i - 64)}; }
AT_HD M128 m_shl(M128 a, int k) { return {a.lo << k, (a.hi << k) | (a.lo >> (64 - k))}; }
AT_HD M128 m_shr(M128 a, int k) { return {(a.lo >> k) | (a.hi << (64 - k)), a.hi >> k}; }
AT_HD M128 col_mask(int col) {  // bits r*11+col, r = 0..10
    M128 m = {0, 0};
#pragma unroll
    for (int r = 0; r < 11; ++r) m = m_or(m, m_bit(r * 11 + col));
    return m;
}
AT_HD M128 m_neighbours(M128 f, M128 notc0, M128 notc10) {
    M128 up = m_shr(f, 11), down = m_shl(f, 11);
    M128 left = m_shr(m_and(f, notc0), 1), right = m_shl(m_and(f, notc10), 1);
    M128 r = m_or(m_or(up, down), m_or(left, right));
    r.hi &= (1ull << (CELLS - 64)) - 1;
    return r;
}
AT_HDN int bfs_first_step(uint64_t wall_lo, uint64_t wall_hi, int sr, int sc, int gr, int gc, int *next_r,
                          int *next_c) {
    *next_r = -1;
    *next_c = -1;
    if (sr < 0 || sr > 10 || sc < 0 || sc > 10 || gr < 0 || gr > 10 || gc < 0 || gc > 10) return 0;
    const M128 wall = {wall_lo, wall_hi};
    const int s = sr * 11 + sc, g = gr * 11 + gc;
    if (m_test(wall, s) || m_test(wall, g)) return 0;
    if (s == g) return 1;
    const M128 all = {~0ull, (1ull << (CELLS - 64)) - 1};
    const M128 freec = m_andn(all, wall);
    const M128 notc0 = m_andn(all, col_mask(0)), notc10 = m_andn(all, col_mask(10));
    M128 visited = m_bit(s);
    M128 F[4];
    const int dr[4] = {-1, 1, 0, 0}, dc[4] = {0, 0, -1, 1};
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        int r = sr + dr[d], cc = sc + dc[d];
        F[d] = {0, 0};
        if (r >= 0 && r <= 10 && cc >= 0 && cc <= 10) {
            M128 b = m_bit(r * 11 + cc);
            if (m_any(m_and(b, freec))) F[d] = b;
        }
        visited = m_or(visited, F[d]);
    }
    const M128 goal = m_bit(g);
    for (int level = 1; level < CELLS; ++level) {
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (m_any(m_and(F[d], goal))) {
                *next_r = sr + dr[d];
                *next_c = sc + dc[d];
                return level + 1;
            }
        if (!m_any(m_or(m_or(F[0], F[1]), m_or(F[2], F[3])))) return 0;
        M128 avail = m_andn(freec, visited);
        M128 n0 = m_and(m_neighbours(F[0], notc0, notc10), avail);
        M128 n1 = m_andn(m_and(m_neighbours(F[1], notc0, notc10), avail), n0);
        M128 n01 = m_or(n0, n1);
        M128 n2 = m_andn(m_and(m_neighbours(F[2], notc0, notc10), avail), n01);
        M128 n012 = m_or(n01, n2);
        M128 n3 = m_andn(m_and(m_neighbours(F[3], notc0, notc10), avail), n012);
        F[0] = n0; F[1] = n1; F[2] = n2; F[3] = n3;
        visited = m_or(visited, m_or(n012, n3));
    }
    return 0;
}

// ------------------------------------------------------------------ maze generation (env/maze.py:7-33)
// Randomized Prim with the reference's list.pop(idx) semantics; draws come from the env's ATRNG stream.
struct MazeOut {
    uint64_t lo, hi;
    uint32_t ctr;
};
AT_HDN MazeOut generate_maze_mask(uint64_t seed, int64_t env_id, uint32_t ctr) {
    M128 wall = {~0ull, (1ull << (CELLS - 64)) - 1};
    uint16_t wl[128];  // packed (nx, ny, px, py), 4 bits each
    int nw = 0;
    auto below = [&](int n) { return (int)(((uint64_t)rng_word(seed, env_id, 0, ctr++) * (uint64_t)n) >> 32); };
    int sx = 1 + 2 * below(5), sy = 1 + 2 * below(5);
    wall = m_andn(wall, m_bit(sy * 11 + sx));
    const int DX[4] = {0, 0, -2, 2}, DY[4] = {-2, 2, 0, 0};
    int cx = sx, cy = sy;
    for (;;) {
        for (int d = 0; d < 4; ++d) {  // add_walls(cx, cy)
            int nx = cx + DX[d], ny = cy + DY[d];
            if (0 < nx && nx < 10 && 0 < ny && ny < 10 && m_test(wall, ny * 11 + nx) && nw < 128)
                wl[nw++] = (uint16_t)(nx | (ny << 4) | (cx << 8) | (cy << 12));
        }
        bool carved = false;
        while (nw > 0 && !carved) {
            int idx = below(nw);  // random.randint(0, len(walls) - 1)
            uint16_t v = wl[idx];
            for (int i = idx; i < nw - 1; ++i) wl[i] = wl[i + 1];
            --nw;
            int wx = v & 15, wy = (v >> 4) & 15, px = (v >> 8) & 15, py = (v >> 12) & 15;
            if (m_test(wall, wy * 11 + wx)) {
                wall = m_andn(wall, m_bit(wy * 11 + wx));
                wall = m_andn(wall, m_bit(((wy + py) / 2) * 11 + (wx + px) / 2));
                cx = wx;
                cy = wy;
                carved = true;
            }
        }
        if (!carved) break;
    }
    return MazeOut{wall.lo, wall.hi, ctr};
}

// cells that are a wall or have one in their 8-neighbourhood: a tank whose centre lies in any other cell cannot
// touch a wall (its OBB stays within 19.3 px of the centre, the nearest wall is >= 70 px from the cell)
AT_HD M128 near_wall_mask(uint64_t wall_lo, uint64_t wall_hi) {
    const M128 all = {~0ull, (1ull << (CELLS - 64)) - 1};
    const M128 w = {wall_lo, wall_hi};
    const M128 notc0 = m_andn(all, col_mask(0)), notc10 = m_andn(all, col_mask(10));
    M128 h = m_or(w, m_or(m_shr(m_and(w, notc0), 1), m_shl(m_and(w, notc10), 1)));
    M128 v = m_or(h, m_or(m_shr(h, 11), m_shl(h, 11)));
    v.hi &= (1ull << (CELLS - 64)) - 1;
    return v;
}
// all-pairs table of bfs_first_step for a fixed map (built once on the host): len | dir << 8
AT_HD void fill_bfs_table(uint64_t wall_lo, uint64_t wall_hi, uint16_t *tab) {
    for (int s = 0; s < CELLS; ++s)
        for (int g = 0; g < CELLS; ++g) {
            int nr, nc;
            const int sr = s / 11, sc = s % 11;
            const int len = bfs_first_step(wall_lo, wall_hi, sr, sc, g / 11, g % 11, &nr, &nc);
            int dir = 0;
            if (len > 1) dir = nr < sr ? 0 : (nr > sr ? 1 : (nc < sc ? 2 : 3));
            tab[s * CELLS + g] = (uint16_t)(len | (dir << 8));
        }
}

AT_HD int popc64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
AT_HD int popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
AT_HD int ctz64(uint64_t v) {  // index of the lowest set bit (v != 0)
#ifdef __CUDA_ARCH__
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
AT_HD int ffs32(uint32_t v) {  // index of the lowest set bit (v != 0)
#ifdef __CUDA_ARCH__
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
// index of the k-th (0-based) set bit of a 128-bit mask
AT_HDN int select_bit(M128 m, int k) {   // (out of line: only resets call it, eight times)
    uint64_t w = m.lo;
    int base = 0;
    int pl = popc64(m.lo);
    if (k >= pl) { k -= pl; w = m.hi; base = 64; }
    for (int i = 0; i < k; ++i) w &= w - 1;  // drop k lowest set bits
#ifdef __CUDA_ARCH__
    return base + __ffsll((long long)w) - 1;
#else
    return base + __builtin_ctzll(w);
#endif
}

// ------------------------------------------------------------------ the per-env simulator
// SPEC selects a compile-time specialisation of the config-dependent branches (0: everything is read from the
// config; 2 = 1 + exactly four tanks; 3 = 2 + the 2a_vs_2b tank layout).  SPEC 1 = team mode, every scripted tank a smart bot, fixed map, no TERMINATE_TIME - the 2v2 / NvM
// team_vs_bot workloads: the other bots, the 1v1 modes, the maze generator and the arena ordering drop out of the
// kernel (the generic build is 630 KB of SASS and stalls on instruction fetch).
template <int SPEC>
struct SimT {
    static constexpr int CS = BS;   // stride of the block's working-set columns
    // SPEC 4 / 5: the 1v1 wrappers on a fixed map - 4 = agent mode (two policy-driven tanks: every bot drops out),
    // 5 = bot_agent against the smart bot (train_ppo_bot.py --bot-type smart, BASELINE config 2); TERMINATE_TIME stays a
    // run-time value there
    static constexpr bool kTeam = SPEC >= 1 && SPEC <= 3, kOne = SPEC >= 4;
    AT_HD int cfg_mode() const { return kTeam ? 3 : (SPEC == 4 ? 0 : (SPEC == 5 ? 1 : c.mode)); }
    AT_HD int cfg_bot_type(int i) const { return (kTeam || SPEC == 5) ? 0 : c.bot_type[i]; }
    AT_HD bool cfg_maze() const { return SPEC >= 1 ? false : c.map == 2; }
    AT_HD bool cfg_team_layout() const { return kTeam ? true : (kOne ? false : c.team_layout != 0); }
    AT_HD int cfg_terminate() const { return kTeam ? 0 : c.terminate_time; }
    AT_HD int cfg_n() const { return (SPEC == 2 || SPEC == 3) ? 4 : (kOne ? 2 : c.n_tanks); }   // SPEC 2 = SPEC 1 with exactly four tanks
    // SPEC 3 = SPEC 2 with the 2a_vs_2b layout: tanks 0, 1 = one team's agents, tanks 2, 3 = the other team's smart bots
    AT_HD int cfg_team(int i) const { return SPEC == 3 ? (i >> 1) : c.team[i]; }
    AT_HD uint32_t cfg_team_members(int i) const { return SPEC == 3 ? (i < 2 ? 3u : 12u) : c.team_members[i]; }
    AT_HD int cfg_ctrl(int i) const { return SPEC == 3 ? (i >> 1) : (SPEC == 4 ? 0 : (SPEC == 5 ? (i == 0 ? 1 : 0) : c.ctrl[i])); }
    AT_HD int cfg_agent_tank(int k) const { return SPEC == 3 ? k : c.agent_tank[k]; }
    AT_HD int cfg_bot_tank(int k) const { return SPEC == 3 ? 2 + k : c.bot_tank[k]; }
    AT_HD int cfg_bot_ord(int i) const { return SPEC == 3 ? (i >= 2 ? i - 2 : 0) : c.bot_ord[i]; }
    AT_HD uint32_t cfg_los_hoist() const { return SPEC == 3 ? 12u : c.los_hoist; }
    AT_HD int cfg_first_bot_k() const { return SPEC == 3 ? 2 : c.first_bot_k; }
    AT_HD int cfg_n_agents() const { return SPEC == 3 ? 2 : c.n_agents; }
    AT_HD int cfg_rows() const { return SPEC == 3 ? 2 : c.rows; }
    AT_HD int cfg_act_rows() const { return SPEC == 3 ? 2 : c.act_rows; }
    const KCfg &c;
    const DevState &st;
    const BlockMem &bm;
    int e;      // env index into the SoA arrays (32-bit column indices: at_create caps the handle at 2^31 / 48 envs)
    int n32;    // st.N
    int tid;    // slot in the block working set
    uint32_t cnt, rng, tick, tstep, epstep, zones;
    uint64_t wlo, whi, nlo, nhi;
    uint32_t alive_bits;   // bit i: tank i alive (mirror of TPACK)
    uint32_t warp_mask;    // lanes of this warp that own an env (the pooled stages of the controllers run on these)

    AT_HD SimT(const KCfg &c_, const DevState &st_, const BlockMem &bm_, int64_t e_, int tid_)
        : c(c_), st(st_), bm(bm_), e((int)e_), n32((int)st_.N), tid(tid_), cnt(0), rng(0), tick(0), tstep(0), epstep(0), zones(0), wlo(0),
          whi(0), nlo(0), nhi(0), alive_bits(0), warp_mask(0xffffffffu) {}

    AT_HD double &TX(int i) const { return bm.tx[i * CS + tid]; }
    AT_HD double &TY(int i) const { return bm.ty[i * CS + tid]; }
    AT_HD double &TREW(int i) const { return bm.trew[i * CS + tid]; }
    AT_HD uint32_t &TPACK(int i) const { return bm.tpack[i * CS + tid]; }
    AT_HD uint32_t &THITS(int i) const { return st.thits[G(i)]; }  // touched on hits / resets only: stays in HBM
    // bullets (slot indices) that may lie within 100 px of tank i when it acts this tick (dodge_reward candidates)
    AT_HD void near_clear(int i) const {
        bm.nearlo[i * CS + tid] = 0u;
        if (cfg_n() * SLOTS_PER_TANK > 32) bm.nearhi[i * CS + tid] = 0u;
    }
    AT_HD void near_set(int i, int slot) const {
        if (slot < 32) bm.nearlo[i * CS + tid] |= 1u << slot;
        else bm.nearhi[i * CS + tid] |= 1u << (slot - 32);
    }
    AT_HD uint64_t near_get(int i) const {
        uint64_t v = bm.nearlo[i * CS + tid];
        if (cfg_n() * SLOTS_PER_TANK > 32) v |= (uint64_t)bm.nearhi[i * CS + tid] << 32;
        return v;
    }
    AT_HD uint8_t &TLIVE(int i) const { return bm.tlive[i * CS + tid]; }
    AT_HD int32_t &TSHOT(int i) const { return bm.tshot[i * CS + tid]; }
    // bot FSM columns exist only for bot tanks (bot_ord: ordinal among the bots)
    AT_HD uint32_t &BPACK(int i) const { return bm.bpack[cfg_bot_ord(i) * CS + tid]; }
    AT_HD double &BF0(int i) const { return bm.bf0[cfg_bot_ord(i) * CS + tid]; }
    AT_HD double &BF1(int i) const { return bm.bf1[cfg_bot_ord(i) * CS + tid]; }
    AT_HD int G(int slot) const { return slot * n32 + e; }

    AT_HD int t_angle(int i) const { return (int)(TPACK(i) & 511u); }
    AT_HD int t_speed(int i) const {
        uint32_t s = (TPACK(i) >> 9) & 3u;
        return s == 0 ? 0 : (s == 1 ? 10 : (s == 2 ? -10 : 1));
    }
    AT_HD bool t_alive(int i) const { return (TPACK(i) & TP_ALIVE) != 0; }
    AT_HD void set_speed(int i, int sp) {
        uint32_t code = sp == 0 ? 0u : (sp == 10 ? 1u : (sp == -10 ? 2u : 3u));
        TPACK(i) = (TPACK(i) & ~(3u << 9)) | (code << 9);
    }
    // bullet_dir / np.linalg.norm(bullet_dir) of the firing direction (cos a, -sin a): one ulp offset per component
    AT_HD void unit_dir(int a, double &u0, double &u1) const {
        const uint32_t dl = bm.udelta[a];
        u0 = bits_d(d_bits(bm.trig[2 * a]) + (uint64_t)(int64_t)((int)(dl & 3u) - 1));
        u1 = bits_d(d_bits(-bm.trig[2 * a + 1]) + (uint64_t)(int64_t)((int)((dl >> 2) & 3u) - 1));
    }
    AT_HD double cosr(int a) const { return bm.trig[2 * a]; }
    AT_HD double sinr(int a) const { return bm.trig[2 * a + 1]; }

    // the kernel stages this env's action rows with asynchronous copies at launch (each thread copies and reads only
    // its own rows); they must have landed before the first read
    AT_HD void acts_ready() const {
#ifdef __CUDA_ARCH__
        asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
    }
    AT_HD uint32_t rng_u32() { return SPEC == 0 ? rng_word_ni(c.seed, c.env_id_base + e, 0, rng++) : rng_word(c.seed, c.env_id_base + e, 0, rng++); }
    AT_HD static double atan2_(double y, double x) { return SPEC == 0 ? at_atan2_ni(y, x) : at_atan2(y, x); }
    AT_HD static double fmod360_(double v) { return SPEC == 0 ? py_fmod360_ni(v) : py_fmod360(v); }
    AT_HD int rng_below(int n) { return (int)(((uint64_t)rng_u32() * (uint64_t)n) >> 32); }
    AT_HD double rng_unit53() {
        uint32_t a = rng_u32() >> 5, b = rng_u32() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }

    AT_HD bool wall_at(int idx) const { return ((idx < 64 ? wlo >> idx : whi >> (idx - 64)) & 1ull) != 0; }
    AT_HD bool near_at(int idx) const { return ((idx < 64 ? nlo >> idx : nhi >> (idx - 64)) & 1ull) != 0; }

    // The <= 2 x 2 cells under a w x h rect (w, h <= 70) at integer (ix, iy): index of the top-left cell and a
    // 13-bit selector with bit 0 / 1 / 11 / 12 = that cell / right / below / below-right (ascending bit order is
    // row-major order = the reference's wall list order, gaming_env.py:610-615).  sel == 0: outside the grid.
    AT_HD static void rect_cells(int ix, int iy, int w, int h, int &idx00, uint32_t &sel) {
        int c0, c1, r0, r1;
        if ((ix | iy) >= 0) {  // one division per axis; the far edge is at most one cell further
            c0 = ix / 70; c1 = imin(c0 + ((ix - 70 * c0 + w - 1) >= 70 ? 1 : 0), 10);
            r0 = iy / 70; r1 = imin(r0 + ((iy - 70 * r0 + h - 1) >= 70 ? 1 : 0), 10);
        } else {               // look-ahead rects of the dodge bot may leave the arena (App. H note)
            c0 = imax(fdiv70(ix), 0); c1 = imin(fdiv70(ix + w - 1), 10);
            r0 = imax(fdiv70(iy), 0); r1 = imin(fdiv70(iy + h - 1), 10);
        }
        const bool ok = c0 <= c1 && r0 <= r1 && c0 <= 10 && r0 <= 10;
        const uint32_t dc = c1 > c0 ? 1u : 0u, dr = r1 > r0 ? 1u : 0u;
        idx00 = ok ? r0 * 11 + c0 : 0;
        sel = ok ? (1u | (dc << 1) | (dr << 11) | ((dc & dr) << 12)) : 0u;
    }
    AT_HD static uint32_t window13(uint64_t lo, uint64_t hi, int idx) {  // bits [idx, idx + 13) of a 128-bit mask
        const uint64_t a = idx < 64 ? lo : hi, b = idx < 64 ? hi : 0ull;
        const int sh = idx & 63;
        const uint64_t v = (a >> sh) | ((b << 1) << (63 - sh));
        return (uint32_t)v & 0x1FFFu;
    }
    // first wall in list order that a w x h pygame.Rect at integer (ix, iy) collides with; -1 if none.
    // Replaces `for wall in walls: colliderect` scans.
    AT_HD int first_wall(int ix, int iy, int w, int h) const {
        int idx00;
        uint32_t sel;
        rect_cells(ix, iy, w, h, idx00, sel);
        const uint32_t ww = window13(wlo, whi, idx00) & sel;
        return ww ? idx00 + ffs32(ww) : -1;
    }
    AT_HD static bool rect5_hits_cell(int ix, int iy, int cell) {  // Rect(ix,iy,5,5) vs the 70x70 wall rect
        int X = (cell % 11) * 70, Y = (cell / 11) * 70;
        return X < ix + 5 && Y < iy + 5 && X + 70 > ix && Y + 70 > iy;
    }
    AT_HD bool rect5_hits_tank(int ix, int iy, int i) const {  // vs Rect(x - 15, y - 12, 30, 24) (F11)
        int Tx = d2i(TX(i) - 15), Ty = d2i(TY(i) - 12);
        return ix < Tx + 30 && iy < Ty + 24 && ix + 5 > Tx && iy + 5 > Ty;
    }

    // ---------------------------------------------------------------- bullets (env/sprite.py:27-111)
    // gaming_env.py:522-541 (1v1 incl. quirk F6) / gaming_env_multi.py:433-442 (team)
    AT_HD void reward_by_bullets(int shooter, int victim) {
        TREW(shooter) += 50;
        TREW(victim) += -50;
        if (cfg_mode() != AT_MODE_TEAM_) {
            bool all_same = true;
            for (int i = 1; i < cfg_n(); ++i) all_same = all_same && (t_alive(i) == t_alive(0));
            if (all_same)
                for (int i = 0; i < cfg_n(); ++i)
                    if (t_alive(i)) {
                        double q = (double)(1000 - (int)epstep) / 1000.0;
                        double time_bonus = q > 0 ? 5 * q : 0.0;
                        TREW(i) += 200 * time_bonus;
                    }
        }
    }

    static constexpr int AT_MODE_AGENT_ = 0, AT_MODE_BOT_AGENT_ = 1, AT_MODE_ARENA_ = 2, AT_MODE_TEAM_ = 3;

    // One Bullet.move() (sprite.py:59-105) of a bullet of the env in block slot q, as a PURE function of the record,
    // the env's tank columns and the alive mask: everything the move decides, nothing committed.  The commits - f64
    // reward additions, the kill, the compaction - are applied by the caller in firing order (bullet_move for one env
    // at a time).
    struct BOut {
        double nx, ny;
        uint64_t m;        // the record after the move (flips, bounces, distance, cleared, pending count)
        uint32_t near;     // alive enemies whose 113-px box holds the bullet (dodge_reward candidates)
        int n_hi;          // BULLET_TOWARD_REWARD additions to the owner (one per enemy the bullet flies at)
        int victim;        // first alive enemy (list order) whose rect the moved bullet overlaps, or -1
        bool expired;      // no victim, and bounces >= 6 or distance >= 300: leaves the list (settle_penalty)
    };
    AT_HD bool rect5_hits_tank_q(int ix, int iy, int i, int q) const {  // vs Rect(x - 15, y - 12, 30, 24) (F11)
        int Tx = d2i(bm.tx[i * CS + q] - 15), Ty = d2i(bm.ty[i * CS + q] - 12);
        return ix < Tx + 30 && iy < Ty + 24 && ix + 5 > Tx && iy + 5 > Ty;
    }
    AT_HD void bullet_eval(int q, double x, double y, uint64_t m, uint32_t alive, BOut &o) const {
        const int owner = (int)(m & 15u), ang = (int)((m >> BM_ANGLE) & 511u);
        const bool fx = (m >> BM_FLIPX) & 1u, fy = (m >> BM_FLIPY) & 1u;
        int bounces = (int)((m >> BM_BOUNCES) & 7u), dist = (int)((m >> BM_DIST) & 511u);
        bool cleared = (m >> BM_CLEARED) & 1u;
        uint32_t pcnt = (uint32_t)((m >> BM_PEND) & 0xFFFFu);
        const double cv = cosr(ang), sv = sinr(ang);
        double dx = fx ? -cv : cv, dy = fy ? sv : -sv;
        const double nx = x + dx, ny = y + dy;
        const int ix = d2i(nx), iy = d2i(ny);
        uint32_t near_tanks = 0, gate = 0;
        int n_hi = 0;
        // update_reward_logic on the pre-move position (sprite.py:27-57).  Same decisions as the reference's
        // norm / divide / dot sequence, reached cheaply: the unit direction comes from a table, distances are
        // compared through their squares (exact, see S300 / S100), and the cosine is first evaluated with fast
        // arithmetic - only when it lands within 1e-9 of a threshold is the reference's exact sequence run.
        {
            double u0, u1;
            unit_dir(ang, u0, u1);
            const double d0 = fx ? -u0 : u0, d1 = fy ? -u1 : u1;
            // one enemy: s2 = |t|^2 and q = d . t are passed in so that the caller decides when they are computed
            auto enemy = [&](double tx, double ty, double s2, double qd) {
                // cos > 0.9 <=> q > 0 and q^2 > 0.81 s2 ; cos < 0.1 <=> q <= 0 or q^2 < 0.01 s2   (q = d . t)
                const double q2 = qd * qd, a81 = 0.81 * s2, b01 = 0.01 * s2, eps = 1e-9 * s2;
                bool hi, lo;
                if (s2 > 0.0 && !(qd > 0.0 && (fabs(q2 - a81) <= eps || fabs(q2 - b01) <= eps))) {
                    hi = qd > 0.0 && q2 > a81;
                    lo = !(qd > 0.0 && q2 >= b01);
                } else {  // within 1e-9 of a threshold, or dist == 0 (NaN: both comparisons false): reference sequence
                    const double dt = sqrt(s2);
                    const double dp = np_dot2(d0, d1, tx / dt, ty / dt);
                    hi = dp > 0.9;
                    lo = dp < 0.1;
                }
                if (hi) ++n_hi;
                else if (lo) { if (!cleared) ++pcnt; }
                if (s2 < S100) { pcnt = 0; cleared = true; }          // dist < DISTANCE_CANCEL_THRESHOLD
            };
            // per enemy also: the dodge_reward candidate box (the tank moves <= 10 px and this bullet <= 1 px per axis
            // before that test) and the gate of the rect test below (the moved 5x5 rect can only overlap
            // Rect(x-15, y-12, 30, 24) when the pre-move centre distance is < 22 / 19 px per axis)
            if (SPEC == 3) {   // exactly two enemies (tanks 2, 3 or 0, 1): side by side and predicated, no divergence
                const int e0 = (owner & 2) ^ 2;
                bool use[2], hi[2], lo[2], in100[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int i = e0 + j;
                    const double tx = bm.tx[i * CS + q] - x, ty = bm.ty[i * CS + q] - y;
                    const double s2 = np_dot2(tx, ty, tx, ty);
                    const double qd = d0 * tx + d1 * ty;
                    const double atx = fabs(tx), aty = fabs(ty);
                    const bool al = (alive >> i) & 1u;
                    if (al && atx < 113.0 && aty < 113.0) near_tanks |= 1u << i;
                    if (al && atx < 24.0 && aty < 21.0) gate |= 1u << i;
                    use[j] = al && !(atx > 301.0 || aty > 301.0 || s2 > S300);
                    const double q2 = qd * qd, a81 = 0.81 * s2, b01 = 0.01 * s2, eps = 1e-9 * s2;
                    hi[j] = qd > 0.0 && q2 > a81;
                    lo[j] = !(qd > 0.0 && q2 >= b01);
                    in100[j] = s2 < S100;
                    if (use[j] && (!(s2 > 0.0) || (qd > 0.0 && (fabs(q2 - a81) <= eps || fabs(q2 - b01) <= eps)))) {
                        const double dt = sqrt(s2);   // within 1e-9 of a threshold, or dist == 0: the reference's sequence
                        const double dp = np_dot2(d0, d1, tx / dt, ty / dt);
                        hi[j] = dp > 0.9;
                        lo[j] = dp < 0.1;
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {   // the bullet's own state, enemy after enemy
                    n_hi += (use[j] && hi[j]) ? 1 : 0;
                    pcnt += (use[j] && !hi[j] && lo[j] && !cleared) ? 1u : 0u;
                    if (use[j] && in100[j]) { pcnt = 0; cleared = true; }
                }
            } else {
                for (uint32_t bits = alive & ~cfg_team_members(owner); bits; bits &= bits - 1) {  // alive enemies
                    const int i = ffs32(bits);
                    const double tx = bm.tx[i * CS + q] - x, ty = bm.ty[i * CS + q] - y;
                    const double atx = fabs(tx), aty = fabs(ty);
                    if (atx < 113.0 && aty < 113.0) near_tanks |= 1u << i;
                    if (atx < 24.0 && aty < 21.0) gate |= 1u << i;
                    if (atx > 301.0 || aty > 301.0) continue;  // norm >= max(|tx|, |ty|) > 300
                    const double s2 = np_dot2(tx, ty, tx, ty);
                    if (s2 > S300) continue;                              // dist > BULLET_MAX_DISTANCE
                    enemy(tx, ty, s2, d0 * tx + d1 * ty);
                }
            }
        }
        bool nfx = fx, nfy = fy;
        // the <= 4 cells under the 5x5 rect, branch-free: first wall in list order
        int cell;
        {
            int idx00;
            uint32_t sel;
            rect_cells(ix, iy, 5, 5, idx00, sel);
            const uint32_t ww = window13(wlo, whi, idx00) & sel;
            cell = ww ? idx00 + ffs32(ww) : -1;
        }
        if (cell >= 0) {  // sprite.py:66-82 (F12)
            bool bxh = rect5_hits_cell(ix, d2i(y), cell), byh = rect5_hits_cell(d2i(x), iy, cell);
            if (bxh) nfx = !nfx;
            if (byh) nfy = !nfy;
            ++bounces;
        }
        ++dist;
        int victim = -1;
        for (; gate; gate &= gate - 1) {  // sprite.py:88-100
            const int i = ffs32(gate);
            if (rect5_hits_tank_q(ix, iy, i, q)) { victim = i; break; }
        }
        o.nx = nx; o.ny = ny; o.near = near_tanks; o.n_hi = n_hi; o.victim = victim;
        o.expired = victim < 0 && (bounces >= 6 || dist >= 300);  // sprite.py:102-111
        o.m = (uint64_t)owner | ((uint64_t)ang << BM_ANGLE) | ((uint64_t)nfx << BM_FLIPX) | ((uint64_t)nfy << BM_FLIPY) |
              ((uint64_t)bounces << BM_BOUNCES) | ((uint64_t)cleared << BM_CLEARED) | ((uint64_t)dist << BM_DIST) |
              ((uint64_t)pcnt << BM_PEND);
    }
    // what a hit commits (sprite.py:88-100): the kill, the hit counters, the rewards
    AT_HD void commit_hit(int owner, int victim) {
        if (cfg_terminate() == 0) { TPACK(victim) &= ~TP_ALIVE; alive_bits &= ~(1u << victim); }
        THITS(owner) += 1u;
        THITS(victim) += 1u << 16;
        reward_by_bullets(owner, victim);
    }
    // settle_penalty (sprite.py:107-111) of an expired bullet with record m
    AT_HD void commit_expiry(uint64_t m) {
        const uint32_t pcnt = (uint32_t)((m >> BM_PEND) & 0xFFFFu);
        if (!((m >> BM_CLEARED) & 1u) && pcnt > 0) TREW((int)(m & 15u)) -= ldg(st.pend + pcnt);
    }
    // One Bullet.move() of this lane's env, committed.  Returns true when the bullet leaves the list.
    AT_HD bool bullet_move(double &x, double &y, uint64_t &m, uint32_t &near_tanks) {
        BOut o;
        bullet_eval(tid, x, y, m, alive_bits, o);
        const int owner = (int)(m & 15u);
        if (o.n_hi) {   // the same additions in the same order, through a register instead of a shared-memory round trip each
            double t = TREW(owner);
            for (int k = 0; k < o.n_hi; ++k) t += 0.01;
            TREW(owner) = t;
        }
        near_tanks |= o.near;
        x = o.nx; y = o.ny; m = o.m;
        if (o.victim >= 0) { commit_hit(owner, o.victim); return true; }
        if (o.expired) { commit_expiry(m); return true; }
        return false;
    }

    // `for bullet in self.bullets[:]: bullet.move()` (gaming_env.py:71-72, 378-379) of this lane's env from list
    // position r0 on: firing order, stable compaction in place (w = next free position); keeps the per-owner live
    // counts and ranks (the observation rows use them).
    AT_HD void bullets_serial(uint32_t r0, uint32_t w) {
        const uint32_t n = cnt;
        // software pipelining: the next bullet's record is requested before the current one is processed
        double nx_ = 0.0, ny_ = 0.0;
        uint64_t nm_ = 0;
        if (r0 < n) { nx_ = st.bx[G(r0)]; ny_ = st.by[G(r0)]; nm_ = st.bmeta[G(r0)]; }
        for (uint32_t r = r0; r < n; ++r) {
            double x = nx_, y = ny_;
            uint64_t m = nm_;
            if (r + 1 < n) { nx_ = st.bx[G(r + 1)]; ny_ = st.by[G(r + 1)]; nm_ = st.bmeta[G(r + 1)]; }
            uint32_t near_tanks = 0;
            if (!bullet_move(x, y, m, near_tanks)) {
                const int ow = (int)(m & 15u);
                for (; near_tanks; near_tanks &= near_tanks - 1) near_set(ffs32(near_tanks), (int)w);
                uint32_t rank = TLIVE(ow)++;
                m = (m & ~(7ull << BM_RANK)) | ((uint64_t)rank << BM_RANK);
                st.bx[G(w)] = x;
                st.by[G(w)] = y;
                st.bmeta[G(w)] = m;
                ++w;
            }
        }
        cnt = w;
    }
    AT_HD void bullets_pass() {
        for (int i = 0; i < cfg_n(); ++i) { TLIVE(i) = 0; near_clear(i); }
        bullets_serial(0, 0);
    }

    // ---------------------------------------------------------------- tanks (env/sprite.py:326-357, 528-657)
    AT_HD void zone_xy(int k, int &zx, int &zy) const {
        int cell = (int)((zones >> (7 * k)) & 127u);
        zx = (cell % 11) * 70;
        zy = (cell / 11) * 70;
    }
    AT_HD bool tank_in_zone(int i, int k) const {  // Rect(x-15,y-12,30,24) vs Rect(zx, zy, 245, 245)
        int Tx = d2i(TX(i) - 15), Ty = d2i(TY(i) - 12), zx, zy;
        zone_xy(k, zx, zy);
        return Tx < zx + 245 && Ty < zy + 245 && Tx + 30 > zx && Ty + 24 > zy;
    }
    AT_HD void check_buff_debuff(int i) {  // sprite.py:613-636 (F8: effective cap is 1 or 6)
        bool inb = c.buff_on && (tank_in_zone(i, 0) || tank_in_zone(i, 1));
        bool ind = c.debuff_on && (tank_in_zone(i, 2) || tank_in_zone(i, 3));
        uint32_t p = TPACK(i) & ~(TP_INBUFF | TP_INDEBUFF | TP_MAXB1);
        if (inb) p |= TP_INBUFF;
        if (ind) p |= TP_INDEBUFF | TP_MAXB1;
        TPACK(i) = p;
    }
    AT_HD void rotate(int i, int direction) {  // sprite.py:528-549 (aiming ray-march is zero-weight, F18)
        if (!t_alive(i)) return;
        int a = (t_angle(i) + direction * 2) % 360;
        if (a < 0) a += 360;
        TPACK(i) = (TPACK(i) & ~511u) | (uint32_t)a;
    }
    AT_HD void shoot(int i) {  // sprite.py:579-609 on the virtual clock (F9)
        check_buff_debuff(i);
        if (!t_alive(i)) return;
        const int now = (int)(((uint64_t)tick * 1000u) / 60u);
        const uint32_t maxb = (TPACK(i) & TP_MAXB1) ? 1u : 6u;
        if (TLIVE(i) >= maxb) return;
        if (now - TSHOT(i) < 300) return;
        const int a = t_angle(i);
        const double cv = cosr(a), sv = sinr(a);
        st.bx[G(cnt)] = TX(i) + 10 * cv;
        st.by[G(cnt)] = TY(i) - 10 * sv;
        uint32_t rank = TLIVE(i)++;
        st.bmeta[G(cnt)] = (uint64_t)i | ((uint64_t)a << BM_ANGLE) | ((uint64_t)rank << BM_RANK);
        for (uint32_t bits = ((1u << cfg_n()) - 1u) & ~cfg_team_members(i); bits; bits &= bits - 1)
            near_set(ffs32(bits), (int)cnt);  // a fresh bullet is a candidate for every enemy that acts later
        ++cnt;
        TSHOT(i) = now;
    }
    // `cand` = near_get(i); (pm, px, py) = the record of its lowest slot, requested by the caller before the wall
    // test so that the L2 round trip overlaps it
    AT_HD void dodge_reward(int i, uint64_t cand, uint64_t pm, double px, double py) {  // sprite.py:497-525 (F17)
        const int sp = t_speed(i);
        if (sp == 0) return;  // projection is exactly 0: reward unchanged
        const int a = t_angle(i);
        const double tvx = (double)sp * cosr(a), tvy = (double)sp * sinr(a);
        const int myteam = cfg_team(i);
        double reward = 0;
        const double mx = TX(i), my = TY(i);
        // only the candidates flagged by the bullet pass / shoot() are visited, in slot order (= list order);
        // every other enemy bullet is provably >= 100 px away and would hit `continue` (sprite.py:508)
        // (the next candidate's record is requested before the current one is processed: one L2 round trip in flight
        // behind the arithmetic instead of one exposed per candidate)
        uint64_t m = pm, nm = 0;
        double bxv = px, byv = py, nbx = 0.0, nby = 0.0;
        for (; cand; cand &= cand - 1, m = nm, bxv = nbx, byv = nby) {
            const uint64_t rest = cand & (cand - 1);
            if (rest) { const int r2 = ctz64(rest); nm = st.bmeta[G(r2)]; nbx = st.bx[G(r2)]; nby = st.by[G(r2)]; }
            if (cfg_team((int)(m & 15u)) == myteam) continue;
            const double ex_ = mx - bxv;
            if (fabs(ex_) >= 101.0) continue;                          // distance >= 100 for sure
            const double ey_ = my - byv;
            if (!(np_dot2(ex_, ey_, ex_, ey_) < S100)) continue;       // `distance >= 100: continue` (sprite.py:508)
            const int ba = (int)((m >> BM_ANGLE) & 511u);
            double u0, u1;
            unit_dir(ba, u0, u1);
            const double d0 = ((m >> BM_FLIPX) & 1u) ? -u0 : u0, d1 = ((m >> BM_FLIPY) & 1u) ? -u1 : u1;
            double proj = np_dot2(tvx, tvy, -d1, d0);
            reward += fabs(proj) * 0.01;
        }
        TREW(i) += reward;
    }
    // obb_vs_aabb over the wall list (util.py:15-48) -> exact 4-axis SAT against the few wall cells whose
    // AABB can reach the OBB (others are separated on the x or y axis by construction).
    AT_HD bool obb_hits_walls(double nx, double ny, int a) const {
        {   // quick reject: centre in a cell with no wall in its 3x3 neighbourhood
            const int cx = d2i(nx), cy = d2i(ny);
            if (cx >= 0 && cy >= 0 && cx < 770 && cy < 770 && !near_at((cy / 70) * 11 + cx / 70)) return false;
        }
        const double *co = st.corn + 4 * a;
        const double o0x = ldg(co), o0y = ldg(co + 1), o1x = ldg(co + 2), o1y = ldg(co + 3);
        double kx[4] = {nx + o0x, nx + o1x, nx - o0x, nx - o1x};
        double ky[4] = {ny + o0y, ny + o1y, ny - o0y, ny - o1y};
        double mnx = kx[0], mxx = kx[0], mny = ky[0], mxy = ky[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            mnx = kx[k] < mnx ? kx[k] : mnx; mxx = kx[k] > mxx ? kx[k] : mxx;
            mny = ky[k] < mny ? ky[k] : mny; mxy = ky[k] > mxy ? ky[k] : mxy;
        }
        // candidate cells: a conservative superset is enough (every candidate gets the exact test below), so the
        // bounds use a reciprocal multiply; the 0.02 vs 0.01 slack absorbs its last-bit difference from a division
        constexpr double INV70 = 1.0 / 70.0;
        const int c0 = imax((int)floor((mnx - 0.02) * INV70) - 1, 0), c1 = imin((int)floor((mxx + 0.02) * INV70), 10);
        const int r0 = imax((int)floor((mny - 0.02) * INV70) - 1, 0), r1 = imin((int)floor((mxy + 0.02) * INV70), 10);
        if (c0 > c1 || r0 > r1) return false;
        {
            const uint32_t colmask = ((2u << (c1 - c0)) - 1u) << c0;
            uint32_t any = 0u;
#pragma unroll 1
            for (int r = r0; r <= r1; ++r) any |= window13(wlo, whi, r * 11) & colmask;
            if (!any) return false;
        }
        // the reference tests 4 axes per wall (2 OBB axes, x, y) and needs all of them to overlap: test the cheap
        // x / y axes first and build the OBB axes (2 sqrt, 4 divisions) only for a wall that survives them
        bool have_axes = false;
        double ax[2] = {0, 0}, ay[2] = {0, 0}, omin[2] = {0, 0}, omax[2] = {0, 0};
#pragma unroll 1
        for (int r = r0; r <= r1; ++r)
#pragma unroll 1
            for (int cc = c0; cc <= c1; ++cc) {
                if (!wall_at(r * 11 + cc)) continue;
                const double X0 = (double)(cc * 70), X1 = (double)(cc * 70 + 70), Y0 = (double)(r * 70),
                             Y1 = (double)(r * 70 + 70);
                if ((mxx < X0 - 0.01) || (X1 < mnx - 0.01) || (mxy < Y0 - 0.01) || (Y1 < mny - 0.01)) continue;
                if (!have_axes) {  // computed from the absolute corners, as the reference does
                    have_axes = true;
                    double e0x = kx[1] - kx[0], e0y = ky[1] - ky[0], e1x = kx[3] - kx[0], e1y = ky[3] - ky[0];
                    double l0 = sqrt((0.0 + e0x * e0x) + e0y * e0y), l1 = sqrt((0.0 + e1x * e1x) + e1y * e1y);
                    ax[0] = e0x / l0; ax[1] = e1x / l1; ay[0] = e0y / l0; ay[1] = e1y / l1;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        double lo = 0, hi = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            double d = (0.0 + kx[k] * ax[q]) + ky[k] * ay[q];
                            lo = (k == 0 || d < lo) ? d : lo;
                            hi = (k == 0 || d > hi) ? d : hi;
                        }
                        omin[q] = lo; omax[q] = hi;
                    }
                }
                bool sep = false;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    double d0 = (0.0 + X0 * ax[q]) + Y0 * ay[q], d1 = (0.0 + X1 * ax[q]) + Y0 * ay[q];
                    double d2 = (0.0 + X1 * ax[q]) + Y1 * ay[q], d3 = (0.0 + X0 * ax[q]) + Y1 * ay[q];
                    double lo = fmin(fmin(d0, d1), fmin(d2, d3)), hi = fmax(fmax(d0, d1), fmax(d2, d3));
                    sep = sep || (omax[q] < lo - 0.01) || (hi < omin[q] - 0.01);
                }
                if (!sep) return true;
            }
        return false;
    }
    AT_HD void move(int i) {  // sprite.py:326-357
        if (!t_alive(i)) return;
        const int a = t_angle(i), sp = t_speed(i);
        const double nx = TX(i) + (double)sp * cosr(a), ny = TY(i) - (double)sp * sinr(a);
        const uint64_t cand = sp != 0 ? near_get(i) : 0ull;
        uint64_t pm = 0;
        double px = 0.0, py = 0.0;
        if (cand) { const int r = ctz64(cand); pm = st.bmeta[G(r)]; px = st.bx[G(r)]; py = st.by[G(r)]; }
        if (!obb_hits_walls(nx, ny, a)) {
            TX(i) = nx;
            TY(i) = ny;
            TPACK(i) &= ~TP_HW;
        } else {
            TPACK(i) |= TP_HW;
        }
        dodge_reward(i, cand, pm, px, py);
    }
    AT_HD void apply_action(int i, int a0, int a1, int a2, int fwd, int back, int stop_speed) {
        if (a0 == 0) rotate(i, 3);
        else if (a0 == 2) rotate(i, -3);
        set_speed(i, a1 == fwd ? 10 : (a1 == back ? -10 : stop_speed));
        if (a2 == 1) shoot(i);
        move(i);
    }

    // ---------------------------------------------------------------- bots (env/bots/*.py)
    AT_HD int nearest_opponent(int me, double *dist_out) const {
        double best = INFINITY;
        int idx = -1;
        auto visit = [&](int i) {
            if (cfg_team(i) != cfg_team(me) && t_alive(i)) {
                double dx = TX(i) - TX(me), dy = TY(i) - TY(me);
                double d = sqrt(dx * dx + dy * dy);
                if (d < best) { best = d; idx = i; }
            }
        };
        if (SPEC == 0) {   // the generic build keeps the loop rolled (100 KB of SASS otherwise: four call sites x eight tanks)
#pragma unroll 1
            for (int i = 0; i < cfg_n(); ++i) visit(i);
        } else {
            for (int i = 0; i < cfg_n(); ++i) visit(i);
        }
        if (dist_out) *dist_out = best;
        return idx;
    }
    AT_HD double angle_diff_to(int me, double tx, double ty) const {
        double dx = tx - TX(me), dy = ty - TY(me);
        double th = atan2_(-dy, dx);
#if defined(__CUDA_ARCH__) && !defined(AT_NO_ATAN2_CR)
        // every caller compares the difference with integer thresholds (and the tank's angle is an integer): when the
        // bearing is within 1e-9 degrees of an integer the last bits of atan2 decide, and the decision is taken on the
        // correctly rounded value (atan2_refine; pinned lattice cases are exact already)
        {
            const double t = py_degrees(th), k = rint(t);
            if (fabs(t - k) < 1e-9 && dx != 0.0 && dy != 0.0 && fabs(dx) != fabs(dy)) th = atan2_refine(-dy, dx, th, (int)k, st.atab);
        }
#endif
        double ta = fmod360_(py_degrees(th));
        int cur = t_angle(me) % 360;
        return fmod360_(ta - cur + 180) - 180;
    }
    AT_HD static int aim_from_diff(double d, double thr) { return fabs(d) < thr ? 0 : (d > 0 ? -1 : 1); }

    // BulletTrajectory.simulate_complete_trajectory -> will_hit_target (sprite.py:131-190), used by
    // SmartStrategyBot.check_line_of_sight (strategy_bot.py:58-73).  Written as a resumable state machine so that
    // the same iteration runs inline (one env per thread) or pooled across the lanes of a warp (`q` = block slot
    // of the env the job belongs to; wall masks are this thread's, which is the same map for fixed-map handles).
    struct Los {
        double x, y, dx, dy, inv_dx, inv_dy;
        uint32_t enemies;  // alive tanks of other teams
        int q, bx0, bx1, by0, by1, bounces, dist, plan_in, result;  // result: -1 running, 0 / 1 finished
        bool stat_x, stat_y, accepted;
    };
    AT_HD int los_ex(const Los &L, int i) const { return d2i(bm.tx[i * CS + L.q] - 15); }
    AT_HD int los_ey(const Los &L, int i) const { return d2i(bm.ty[i * CS + L.q] - 12); }
    AT_HD void los_init(Los &L, int q, int me) const {
        const uint32_t pk = bm.tpack[me * CS + q];
        const int a = (int)(pk & 511u);
        const double cv = cosr(a), sv = sinr(a);
        L.q = q;
        L.x = bm.tx[me * CS + q] + 10 * cv;
        L.y = bm.ty[me * CS + q] - 10 * sv;
        L.dx = cv;
        L.dy = -sv;
        L.enemies = 0;
        L.bx0 = 1 << 20; L.bx1 = -(1 << 20); L.by0 = 1 << 20; L.by1 = -(1 << 20);  // union of the enemy rect origins
        bool reachable = false;
        for (int i = 0; i < cfg_n(); ++i)
            if (cfg_team(i) != cfg_team(me) && (bm.tpack[i * CS + q] & TP_ALIVE)) {
                L.enemies |= 1u << i;
                const int exi = los_ex(L, i), eyi = los_ey(L, i);
                L.bx0 = imin(L.bx0, exi); L.bx1 = imax(L.bx1, exi); L.by0 = imin(L.by0, eyi); L.by1 = imax(L.by1, eyi);
                // Exact cull.  Every tested position lies within 301 unit sub-steps (a polyline, so <= 301 px) of the
                // muzzle, and its 5x5 rect can only overlap the enemy's Rect(x-15, y-12, 30, 24) when it is less than
                // 20 px / 17 px from the enemy's centre: the muzzle must be within 301 px of that box (1 px of slack).
                const double gx = fabs(bm.tx[i * CS + q] - L.x), gy = fabs(bm.ty[i * CS + q] - L.y);
                const double ux = fmax(gx - 21.0, 0.0), uy = fmax(gy - 18.0, 0.0);
                reachable = reachable || (ux * ux + uy * uy < 303.0 * 303.0);
            }
        L.result = reachable ? -1 : 0;
        // |dx|, |dy| never change (reflections only negate).  cos/sin of an integer angle are either
        // >= 0.017 or < 2e-16 in magnitude; the latter is absorbed by every addition (x >= 2): static axis.
        L.stat_x = fabs(L.dx) < 1e-15;
        L.stat_y = fabs(L.dy) < 1e-15;
        L.inv_dx = L.stat_x ? 0.0 : 1.0 / fabs(L.dx);
        L.inv_dy = L.stat_y ? 0.0 : 1.0 / fabs(L.dy);
        L.bounces = 0; L.dist = 0; L.plan_in = 0;
        // the muzzle position itself is never tested by the reference; the skip-ahead needs a wall-free rect
        L.accepted = first_wall(d2i(L.x), d2i(L.y), 5, 5) < 0;
    }
    // los_plan: the exact skip-ahead (an accelerator: it may run at any time, or never).
    AT_HD void los_plan(Los &L) const {
        if (L.result >= 0) return;
        double x = L.x, y = L.y;
        const double dx = L.dx, dy = L.dy;
        const bool stat_x = L.stat_x, stat_y = L.stat_y;
        const double inv_dx = L.inv_dx, inv_dy = L.inv_dy;
        // Exact skip-ahead.  The reference advances 1 px per sub-step and tests tanks and walls every
        // time.  From an accepted (wall-free) position the 5x5 rect can only START overlapping a wall
        // when one of its LEADING edges crosses a cell boundary (trailing edges only uncover cells), and
        // it cannot overlap an enemy rect before both axes have entered that rect's span.  Both kinds of
        // event are linear in the step count, so a DDA over the leading-edge crossings finds the first
        // crossing that uncovers a wall cell; all sub-steps more than 2 before it (and before the tank
        // bound, the next binade boundary and `dist > 300`) can only be "advance": one exact jump.
        {
            int m = 0;
            const int room = 300 - L.dist;
            if (room >= 6 && L.accepted) {
                double lim = (double)room;
                for (uint32_t bits = L.enemies; bits; bits &= bits - 1) {  // overlap needs x in [ex-4, ex+30), y in [ey-4, ey+24)
                    const int i = ffs32(bits);
                    const int exi = los_ex(L, i), eyi = los_ey(L, i);
                    const double lx = (double)(exi - 4), hx = (double)(exi + 30);
                    const double ly = (double)(eyi - 4), hy = (double)(eyi + 24);
                    const double sx = x < lx ? ((dx > 0 && !stat_x) ? (lx - x) * inv_dx : 1.0e9)
                                             : (x >= hx ? ((dx < 0 && !stat_x) ? (x - hx) * inv_dx : 1.0e9) : 0.0);
                    const double sy = y < ly ? ((dy > 0 && !stat_y) ? (ly - y) * inv_dy : 1.0e9)
                                             : (y >= hy ? ((dy < 0 && !stat_y) ? (y - hy) * inv_dy : 1.0e9) : 0.0);
                    lim = fmin(lim, fmax(sx, sy));
                }
                const int jx = d2i(x), jy = d2i(y);
                double tx = 1.0e9, ty = 1.0e9;  // step count of the next leading-edge crossing per axis
                int lcx = 0, lcy = 0;           // leading column / row
                const int sgx = dx > 0 ? 1 : -1, sgy = dy > 0 ? 1 : -1;
                if (!stat_x) {
                    lcx = dx > 0 ? (jx + 4) / 70 : jx / 70;
                    tx = (dx > 0 ? (double)(70 * (lcx + 1) - 4) - x : x - (double)(70 * lcx)) * inv_dx;
                    const int e = d_expo(x);  // jumps stay inside one binade (jump_axis)
                    lim = fmin(lim, (dx > 0 ? bits_d((uint64_t)(e + 1) << 52) - x : x - bits_d((uint64_t)e << 52)) * inv_dx);
                }
                if (!stat_y) {
                    lcy = dy > 0 ? (jy + 4) / 70 : jy / 70;
                    ty = (dy > 0 ? (double)(70 * (lcy + 1) - 4) - y : y - (double)(70 * lcy)) * inv_dy;
                    const int e = d_expo(y);
                    lim = fmin(lim, (dy > 0 ? bits_d((uint64_t)(e + 1) << 52) - y : y - bits_d((uint64_t)e << 52)) * inv_dy);
                }
                const double step_x = 70.0 * inv_dx, step_y = 70.0 * inv_dy;
                for (;;) {
                    const bool xev = tx <= ty;
                    const double t = xev ? tx : ty;
                    if (t >= lim) break;
                    bool hit = false;
                    if (xev) {  // column lcx + sgx becomes covered; rows around the predicted y (+-3 px slack)
                        lcx += sgx;
                        const double yt = y + t * dy;
                        const int ra = imax(d2i(yt - 3.0) / 70, 0), rb = imin(d2i(yt + 8.0) / 70, 10);
                        for (int r = ra; r <= rb; ++r) hit = hit || lcx < 0 || lcx > 10 || wall_at(r * 11 + lcx);
                        tx += step_x;
                    } else {
                        lcy += sgy;
                        const double xt = x + t * dx;
                        const int ca = imax(d2i(xt - 3.0) / 70, 0), cb = imin(d2i(xt + 8.0) / 70, 10);
                        for (int cc = ca; cc <= cb; ++cc) hit = hit || lcy < 0 || lcy > 10 || wall_at(lcy * 11 + cc);
                        ty += step_y;
                    }
                    if (hit) { lim = t; break; }
                }
                m = imax(d2i(lim) - 2, 0);
            }
            if (m >= 4) {
                if (!stat_x) x = jump_axis(x, dx, m);
                if (!stat_y) y = jump_axis(y, dy, m);
                L.dist += m;
                L.plan_in = 3;  // the event is 2..3 sub-steps ahead: walk through it before planning again
            } else {
                L.plan_in = m + 3;
            }
        }
        L.x = x;
        L.y = y;
    }
    // los_step: one reference sub-step (tank test, wall test / bounce, advance).  Returns true when done.
    AT_HD bool los_step(Los &L) const {
        if (L.result >= 0) return true;
        double x = L.x, y = L.y, dx = L.dx, dy = L.dy;
        --L.plan_in;
        const double nx = x + dx, ny = y + dy;
        const int ix = d2i(nx), iy = d2i(ny);
        if (ix < L.bx1 + 30 && iy < L.by1 + 24 && ix + 5 > L.bx0 && iy + 5 > L.by0)  // inside the union box: test each
            for (uint32_t bits = L.enemies; bits; bits &= bits - 1) {
                const int i = ffs32(bits);
                const int exi = los_ex(L, i), eyi = los_ey(L, i);
                if (ix < exi + 30 && iy < eyi + 24 && ix + 5 > exi && iy + 5 > eyi) { L.result = 1; return true; }
            }
        const int cell = first_wall(ix, iy, 5, 5);
        if (cell >= 0) {
            const bool bxh = rect5_hits_cell(ix, d2i(y), cell), byh = rect5_hits_cell(d2i(x), iy, cell);
            if (bxh) dx = -dx;
            if (byh) dy = -dy;
            ++L.bounces;
            L.plan_in = 0;
        } else {
            x = nx;
            y = ny;
            ++L.dist;
            L.accepted = true;
        }
        L.x = x; L.y = y; L.dx = dx; L.dy = dy;
        if (L.bounces > 6 || L.dist > 300) { L.result = 0; return true; }
        return false;
    }
    AT_HD bool line_of_sight(int me) const {
        Los L;
        los_init(L, tid, me);
        for (;;) {
            if (L.plan_in <= 0) los_plan(L);
            if (los_step(L)) break;
        }
        return L.result != 0;
    }

    // bpack layouts -----------------------------------------------------
    // smart:     target+1 [0:4) aiming 4 rot_dir>0 5 stuck [6:12) has_next 12 next_r [13:17) next_c [17:21)
    // defensive: has_tpos 0, zone index 1
    // dodge:     timer [0:5) state [5:7) has_angle 7
    // random:    delay [0:6) a0 [6:8) a1 [8:10) a2 10
    AT_HD static uint32_t bot_init_pack(int type) {
        if (type == 0) return 1u << 5;                 // rot_dir = +1
        if (type == 1) return (1u << 6) | (1u << 8);   // current_action = [1, 1, 0]
        return 0u;
    }
    AT_HD void bot_init(int i) {
        BPACK(i) = bot_init_pack(cfg_bot_type(i));
        BF0(i) = TX(i);  // strategy_bot.py:15 last_position
        BF1(i) = TY(i);
        if (cfg_bot_type(i) != 0) { BF0(i) = 0.0; BF1(i) = 0.0; }
    }

    // SmartStrategyBot.get_action (strategy_bot.py:75-203) split around its line-of-sight query so that the
    // ray-marches of a warp's envs can be pooled: smart_pre = update_state + target bookkeeping (returns whether
    // a line-of-sight query is needed), smart_post = the decision given the answer.
    AT_HD bool smart_pre(int me) {
        uint32_t p = BPACK(me);
        int target = (int)(p & 15u) - 1;
        bool aiming = (p >> 4) & 1u, has_next = (p >> 12) & 1u;
        int stuck = (int)((p >> 6) & 63u), next_r = (int)((p >> 13) & 15u), next_c = (int)((p >> 17) & 15u);
        if (target < 0 || !t_alive(target)) {
            aiming = false;
            target = nearest_opponent(me, nullptr);
        }
        if (target >= 0) {
            int nr, nc, len;
            const int sr = d2i(TY(me)) / 70, sc = d2i(TX(me)) / 70, gr = d2i(TY(target)) / 70, gc = d2i(TX(target)) / 70;
            if (st.bfs_tab != nullptr && (unsigned)sr < 11u && (unsigned)sc < 11u && (unsigned)gr < 11u && (unsigned)gc < 11u) {
                const uint32_t ent = ldg(st.bfs_tab + (sr * 11 + sc) * CELLS + gr * 11 + gc);  // fixed map: all-pairs table
                const int dir = (int)(ent >> 8);
                len = (int)(ent & 255u);
                nr = sr + (dir == 0 ? -1 : (dir == 1 ? 1 : 0));
                nc = sc + (dir == 2 ? -1 : (dir == 3 ? 1 : 0));
            } else {
                len = bfs_first_step(wlo, whi, sr, sc, gr, gc, &nr, &nc);
            }
            has_next = len > 1;
            if (has_next) { next_r = nr; next_c = nc; }
        }
        {
            double dx = TX(me) - BF0(me), dy = TY(me) - BF1(me);
            stuck = (sqrt(dx * dx + dy * dy) < 1) ? stuck + 1 : 0;
            BF0(me) = TX(me);
            BF1(me) = TY(me);
        }
        if (!aiming) {
            target = nearest_opponent(me, nullptr);
            if (target >= 0) aiming = true;
        }
        BPACK(me) = (uint32_t)(target + 1) | ((uint32_t)aiming << 4) | (p & (1u << 5)) | ((uint32_t)stuck << 6) |
                    ((uint32_t)has_next << 12) | ((uint32_t)next_r << 13) | ((uint32_t)next_c << 17);
        return aiming && target >= 0;
    }
    AT_HD void smart_post(int me, bool los, int act[3]) {
        const uint32_t p = BPACK(me);
        const int target = (int)(p & 15u) - 1;
        const bool aiming = (p >> 4) & 1u, has_next = (p >> 12) & 1u;
        bool rot_pos = (p >> 5) & 1u;
        int stuck = (int)((p >> 6) & 63u);
        const int next_r = (int)((p >> 13) & 15u), next_c = (int)((p >> 17) & 15u);
        act[0] = 1; act[1] = 1; act[2] = 0;
        if (aiming && target >= 0) {
            // one aim computation for both branches (strategy_bot.py:141-175): at the target when there is a line
            // of sight, else at the centre of the next BFS cell
            if (los || has_next) {
                const double px = los ? TX(target) : next_c * GRID + (GRID / 2);
                const double py = los ? TY(target) : next_r * GRID + (GRID / 2);
                const int rotation = aim_from_diff(angle_diff_to(me, px, py), 10);
                if (rotation != 0) act[0] = rotation > 0 ? 2 : 0;
                if (los) {
                    const double dx = px - TX(me), dy = py - TY(me);
                    const double d = sqrt(dx * dx + dy * dy);
                    if (d > 200) act[1] = 2;
                    else if (d < 20) act[1] = 0;
                    if (rotation == 0) act[2] = 1;
                } else {
                    act[1] = rotation == 0 ? 2 : 1;
                }
            }
        }
        if (stuck > 20) {
            act[1] = 2;
            act[0] = rot_pos ? 0 : 2;
            rot_pos = !rot_pos;
            stuck = 0;
        }
        BPACK(me) = (p & ~((1u << 5) | (63u << 6))) | ((uint32_t)rot_pos << 5) | ((uint32_t)stuck << 6);
    }
    // Line-of-sight queries of the warp's envs, pooled: the ray-march lengths differ wildly between envs (0 to ~40
    // iterations), so instead of every thread marching its own bots while its neighbours idle, the (env, bot) jobs
    // of the 32 envs go into a shared-memory queue and idle lanes keep dequeuing - one los_iter per lockstep round.
    // `need` has bit j set when bot tank j of this env needs an answer; returns the answers in the same bits.
    template <class WP>
    AT_HD uint32_t pool_los(uint32_t need) {
        const int lane = tid & 31, wbase = tid & ~31;
        const unsigned wm = warp_mask;
        uint8_t *queue = bm.los_q + (tid >> 5) * 256;
        uint8_t *res = bm.los_r + (tid >> 5) * 256;
        int total = 0;
        for (uint32_t hb = cfg_los_hoist(); hb; hb &= hb - 1) {
            const int j = ffs32(hb);
            const bool mine = (need >> j) & 1u;
            const unsigned m = WP::ballot(wm, mine);
            if (mine) queue[total + popc32(m & ((1u << lane) - 1u))] = (uint8_t)(lane | (j << 5));
            total += popc32(m);
        }
        WP::sync(wm);  // queue and the other envs' tank columns (shared memory) are visible
        int next = 0, myjob = 0;
        bool active = false;
        Los L;
        L.result = 0;
        for (;;) {
            const unsigned idle = WP::ballot(wm, !active);
            const int pos = next + popc32(idle & ((1u << lane) - 1u));
            if (!active && pos < total) {
                myjob = queue[pos];
                los_init(L, wbase + (myjob & 31), myjob >> 5);
                active = true;
            }
            next = imin(next + popc32(idle), total);
            if (WP::ballot(wm, active) == 0u) break;
            if (active) {  // phase-aligned: everybody plans now, then everybody takes plain sub-steps
                if (L.plan_in <= 0) los_plan(L);
                bool fin = false;
#pragma unroll 1
                for (int it = 0; it < 4 && !fin; ++it) fin = los_step(L);
                if (fin) {
                    res[myjob] = (uint8_t)(L.result != 0);
                    active = false;
                }
            }
        }
        WP::sync(wm);
        uint32_t out = 0;
        for (uint32_t hb = need; hb; hb &= hb - 1) {
            const int j = ffs32(hb);
            out |= (uint32_t)res[lane | (j << 5)] << j;
        }
        return out;
    }
    AT_HD void random_action(int me, int act[3]) {  // simple_bots.py:32-46
        uint32_t p = BPACK(me);
        int delay = (int)(p & 63u), a0 = (int)((p >> 6) & 3u), a1 = (int)((p >> 8) & 3u), a2 = (int)((p >> 10) & 1u);
        if (delay <= 0) {
            a0 = rng_below(3);
            a1 = rng_below(3);
            a2 = rng_below(2);
            delay = 10 + rng_below(21);
        }
        delay -= 1;
        BPACK(me) = (uint32_t)delay | ((uint32_t)a0 << 6) | ((uint32_t)a1 << 8) | ((uint32_t)a2 << 10);
        act[0] = a0; act[1] = a1; act[2] = a2;
    }
    AT_HD void aggressive_action(int me, int act[3]) const {  // simple_bots.py:86-125
        double dist;
        int target = nearest_opponent(me, &dist);
        act[0] = 1; act[1] = 1; act[2] = 0;
        if (target < 0) return;
        double d = angle_diff_to(me, TX(target), TY(target));
        double ad = fabs(d);
        int rotation = aim_from_diff(d, 15);
        act[0] = rotation > 0 ? 2 : (rotation < 0 ? 0 : 1);
        if (ad < 15 && dist <= 70.0) act[2] = 1;
        if (dist > 70.0) { if (ad < 45) act[1] = 2; }
    }
    AT_HD void defensive_action(int me, int act[3]) {  // simple_bots.py:151-292
        act[0] = 1; act[1] = 1; act[2] = 0;
        int target = nearest_opponent(me, nullptr);
        if (target < 0) return;
        const int nb = c.buff_on ? 2 : 0;
        bool in_buff = false;
        for (int k = 0; k < nb; ++k) in_buff = in_buff || tank_in_zone(me, k);
        if (in_buff) {
            double d = angle_diff_to(me, TX(target), TY(target));
            int rot = aim_from_diff(d, 20);
            act[0] = rot > 0 ? 2 : (rot < 0 ? 0 : 1);
            if (fabs(d) < 20) act[2] = 1;
            return;
        }
        uint32_t p = BPACK(me);
        bool has = p & 1u;
        int zi = (int)((p >> 1) & 1u);
        if (!has) {
            double best = INFINITY;
            for (int k = 0; k < nb; ++k) {
                double sx, sy;
                safe_centre(k, sx, sy);
                double dx = sx - TX(me), dy = sy - TY(me);
                double d = sqrt(dx * dx + dy * dy);
                if (d < best) { best = d; has = true; zi = k; }
            }
            BPACK(me) = (uint32_t)has | ((uint32_t)zi << 1);
        }
        if (has) {
            double sx, sy;
            safe_centre(zi, sx, sy);
            double d = angle_diff_to(me, sx, sy);
            int rot = aim_from_diff(d, 20);
            act[0] = rot > 0 ? 2 : (rot < 0 ? 0 : 1);
            if (fabs(d) < 45) act[1] = 2;
        }
    }
    AT_HD void safe_centre(int k, double &sx, double &sy) const {  // simple_bots.py:151-175
        int zx, zy;
        zone_xy(k, zx, zy);
        const double bw = GRID * 3.5, mx = 30 * 0.75, my = 24 * 0.75;
        double x0 = zx + mx, x1 = zx + bw - mx, y0 = zy + my, y1 = zy + bw - my;
        sx = (x0 + x1) / 2;
        sy = (y0 + y1) / 2;
    }
    AT_HD bool dodge_hits_wall(int me, double angle) const {  // simple_bots.py:309-329
        double rad = py_radians(angle);
        double wcd = GRID * 1.5;
        double fx = TX(me) + cos_ni(rad) * wcd, fy = TY(me) - sin_ni(rad) * wcd;
        return first_wall(d2i(fx - 15), d2i(fy - 12), 30, 24) >= 0;
    }
    AT_HD double dodge_calc_angle(int me, double thx, double thy) {  // simple_bots.py:359-402
        double v0 = thx - TX(me), v1 = thy - TY(me);
        double n = np_norm2(v0, v1);
        if (n > 0) {
            v0 = v0 / n;
            v1 = v1 / n;
            int base[2] = {90, -90};
            if (rng_below(2) == 0) { base[0] = -90; base[1] = 90; }  // random.shuffle of a 2-list
            for (int bi = 0; bi < 2; ++bi)
                for (int k = 0; k < 3; ++k) {
                    double var = -45.0 + (-30.0 - -45.0) * rng_unit53();  // random.uniform(-45, -30)
                    double theta = py_radians(base[bi] + var);
                    double cs = cos_ni(theta), sn = sin_ni(theta);
                    double r0 = fma_rn(cs, v0, (-sn) * v1), r1 = fma_rn(sn, v0, cs * v1);
                    double fa = fmod360_(py_degrees(atan2_(-r1, r0)));
                    if (!dodge_hits_wall(me, fa)) return fa;
                }
        }
        return (double)t_angle(me);
    }
    AT_HD void dodge_action(int me, int act[3]) {  // simple_bots.py:331-357, 404-455
        act[0] = 1; act[1] = 1; act[2] = 1;
        uint32_t p = BPACK(me);
        int timer = (int)(p & 31u), state = (int)((p >> 5) & 3u);
        bool has_angle = (p >> 7) & 1u;
        const double radius = GRID * 4;
        double best = INFINITY, px = 0, py = 0;
        bool have = false;
        for (int i = 0; i < cfg_n(); ++i)
            if (cfg_team(i) != cfg_team(me) && t_alive(i)) {
                double dx = TX(i) - TX(me), dy = TY(i) - TY(me);
                double d = sqrt(dx * dx + dy * dy);
                if (d < best && d < radius) { best = d; px = TX(i); py = TY(i); have = true; }
            }
        for (uint32_t r = 0; r < cnt; ++r) {
            const uint64_t m = st.bmeta[G(r)];
            if (cfg_team((int)(m & 15u)) == cfg_team(me)) continue;
            double bxv = st.bx[G(r)], byv = st.by[G(r)];
            double dx = bxv - TX(me), dy = byv - TY(me);
            double d = sqrt(dx * dx + dy * dy);
            if (d < best && d < radius) { best = d; px = bxv; py = byv; have = true; }
        }
        if (!have) {
            BPACK(me) = 0u;  // idle, timer 0, no angle
            return;
        }
        if (timer <= 0 || !has_angle) {
            BF0(me) = dodge_calc_angle(me, px, py);
            has_angle = true;
            timer = 10;
            state = 1;
        }
        if (state == 1) {
            if (dodge_hits_wall(me, BF0(me))) {
                BF0(me) = dodge_calc_angle(me, px, py);
                state = 2;
            }
            int cur = t_angle(me) % 360;
            double d = fmod360_(BF0(me) - cur + 180) - 180;
            int rotation = fabs(d) < 5 ? 0 : (d > 0 ? -1 : 1);
            act[0] = rotation > 0 ? 2 : (rotation < 0 ? 0 : 1);
            if (fabs(d) < 45) act[1] = 2;
        }
        timer -= 1;
        BPACK(me) = (uint32_t)timer | ((uint32_t)state << 5) | ((uint32_t)has_angle << 7);
    }
    AT_HD void bot_action(int me, int act[3], bool pre_done, bool los_hoisted) {
        switch (cfg_bot_type(me)) {
        case 0: {
            // runs for dead tanks too: take_action still sets their speed (sprite.py:646-651), which the
            // other tanks' observation rows expose as velocity
            bool los = los_hoisted;
            if (!pre_done) los = smart_pre(me) && line_of_sight(me);
            smart_post(me, los, act);
            break;
        }
        case 1: random_action(me, act); break;
        case 2: aggressive_action(me, act); break;
        case 3: defensive_action(me, act); break;
        case 4: dodge_action(me, act); break;
        default: act[0] = 1; act[1] = 1; act[2] = 0; break;
        }
    }

    // ---------------------------------------------------------------- reset (gaming_env.py:48-66, 509-520)
    AT_HD void env_reset() {
        if (cfg_maze()) {
            const MazeOut mz = generate_maze_mask(c.seed, c.env_id_base + e, rng);
            wlo = mz.lo; whi = mz.hi; rng = mz.ctr;
            set_near();
        }
        const M128 all = {~0ull, (1ull << (CELLS - 64)) - 1};
        const M128 freec = m_andn(all, M128{wlo, whi});
        const int n_empty = CELLS - c.n_walls;
        for (int i = 0; i < cfg_n(); ++i) {
            int cell = select_bit(freec, rng_below(n_empty));
            TX(i) = (double)((cell % 11) * 70) + GRID / 2;
            TY(i) = (double)((cell / 11) * 70) + GRID / 2;
            int a = rng_below(361);
            TPACK(i) = (uint32_t)a | TP_ALIVE;
            TREW(i) = 0.0;
            THITS(i) = 0u;
            TLIVE(i) = 0u;
            TSHOT(i) = 0;
            if (cfg_ctrl(i) == 1) { BPACK(i) = 0u; BF0(i) = 0.0; BF1(i) = 0.0; }
        }
        cnt = 0;
        alive_bits = (1u << cfg_n()) - 1u;
        for (int i = 0; i < cfg_n(); ++i) {
            bool has_bot = (cfg_mode() == AT_MODE_BOT_AGENT_ && i == 0) ||
                           ((cfg_mode() == AT_MODE_TEAM_ || cfg_mode() == AT_MODE_ARENA_) && cfg_ctrl(i) == 1);
            if (has_bot) bot_init(i);
        }
        zones = 0;
        for (int z = 0; z < 2; ++z) {
            if ((z == 0 && !c.buff_on) || (z == 1 && !c.debuff_on)) continue;
            int i = rng_below(n_empty), j = rng_below(n_empty - 1);
            if (j >= i) ++j;
            zones |= (uint32_t)select_bit(freec, i) << (14 * z);
            zones |= (uint32_t)select_bit(freec, j) << (14 * z + 7);
        }
        epstep = 0;
        if (c.reward_delta)
            for (int r = 0; r < cfg_rows(); ++r) st.prev_rew[r * n32 + e] = 0.0f;
        st.prev_act[e] = 0u;
        for (int k = 0; k < cfg_n(); ++k) st.change_time[G(k)] = 0;
    }

    // ---------------------------------------------------------------- one tick
    // GamingENV.step (gaming_env.py:68-386) / GamingTeamENV.step (gaming_env_multi.py:65-97).
    // Written as a 4-stage loop so that bullets_pass, bot_action and apply_action each have exactly ONE call
    // site (they are force-inlined; duplicating them made the kernel 700 KB of SASS and instruction-fetch bound):
    //   stage 0  arena only: both bots decide on the pre-tick state (bot_arena.py:51-52)
    //   stage 1  bullets pass 1      stage 2  controllers in the reference's order      stage 3  bullets pass 2
    // stage 0 (arena: bots decide on the pre-tick state) or stage 2 (controllers in the reference's order)
    template <class WP>
    AT_HD void controllers(int stage, const int32_t *a, int (&acts)[MAX_TANKS][3], uint32_t &pre_done, uint32_t &los_bits) {
        const bool arena = cfg_mode() == AT_MODE_ARENA_, team = cfg_mode() == AT_MODE_TEAM_, botagent = cfg_mode() == AT_MODE_BOT_AGENT_;
        // acting order: team mode = agent tanks in config order, then bot tanks in config order
        // (controller.py:81-113); every 1v1 mode = tank 0, tank 1
        for (int k = 0; k < cfg_n(); ++k) {
            int i = k;
            if (team) i = k < cfg_n_agents() ? cfg_agent_tank(k) : cfg_bot_tank(k - cfg_n_agents());
            const bool is_bot = cfg_ctrl(i) == 1;
            if (is_bot && (stage == 0 || !arena)) {
                if (k == cfg_first_bot_k() && cfg_los_hoist() != 0u) {
                    // every smart bot whose line of sight cannot change before it acts: query them together
                    uint32_t need = 0;
                    for (uint32_t hb = cfg_los_hoist(); hb; hb &= hb - 1) {
                        const int j = ffs32(hb);
                        if (smart_pre(j)) need |= 1u << j;
                    }
                    pre_done = cfg_los_hoist();
                    los_bits = pool_los<WP>(need);
                }
                bot_action(i, acts[i], (pre_done >> i) & 1u, (los_bits >> i) & 1u);
                // (the weakness lives in device memory: at_set_weakness changes it under a captured CUDA graph too)
                if (botagent && !(rng_unit53() < ldg(st.dyn))) { acts[i][0] = 1; acts[i][1] = 1; acts[i][2] = 0; }
            }
            if (stage == 0) continue;
            if (!is_bot) {
                const int row = team ? k : i;
                acts_ready();
                acts[i][0] = a[3 * row]; acts[i][1] = a[3 * row + 1]; acts[i][2] = a[3 * row + 2];
            }
            // bot_agent's agent tank uses the swapped movement table and creeps on "stop" (F14)
            const bool swapped = botagent && i == 1;
            apply_action(i, acts[i][0], acts[i][1], acts[i][2], swapped ? 0 : 2, swapped ? 2 : 0, swapped ? 1 : 0);
        }
    }
    template <class WP>
    AT_HD void game_step(const int32_t *a /* [act_rows][3] of this env, or null */) {
        int acts[MAX_TANKS][3];
        tick += 1;
        if (cfg_mode() != AT_MODE_TEAM_) epstep += 1;
        uint32_t pre_done = 0, los_bits = 0;
        if (SPEC >= 1) {   // straight-line: pass, controllers, pass (measured: the rolled stage loop costs 10 %)
            bullets_pass();
            controllers<WP>(2, a, acts, pre_done, los_bits);
            bullets_pass();
            return;
        }
        const bool arena = cfg_mode() == AT_MODE_ARENA_;
        for (int stage = arena ? 0 : 1; stage < 4; ++stage) {
            if (stage & 1) {
                bullets_pass();
                continue;
            }
            controllers<WP>(stage, a, acts, pre_done, los_bits);
        }
    }

    // The step's info dict (gym_env.py:93-98 winner; gym_env_multi.py:218-222 hits / be_hits) is read by the
    // reference's callers BEFORE they reset.  With auto-reset the terminal state is replaced inside the step, so
    // what the dict needs is latched here; `tick` (never reset, F9) stamps the latch: it is valid until the next step.
    AT_HD void latch_terminal_info() {
        int w = -1;
        for (int i = 0; i < cfg_n(); ++i) {
            if (w < 0 && t_alive(i)) w = i;
            st.term_hits[G(i)] = THITS(i);
        }
        st.term_winner[e] = (uint32_t)(w + 1);
        st.term_tick[e] = tick;
    }

    // wrapper step (gym_env.py:73-103, 186-200; gym_env_multi.py:202-223, 307-326).  Returns done.
    // the part of the wrapper step in front of the engine's tick
    AT_HD void wrapper_pre(const int32_t *a) {
        tstep += 1;
        if (!cfg_team_layout() && a) {  // change_time, gym_env.py:77-85
            acts_ready();
            uint32_t pv = st.prev_act[e];
            uint32_t nv = 1u;
            for (int k = 0; k < cfg_act_rows(); ++k) {
                uint32_t code = (uint32_t)(a[3 * k] * 3 + a[3 * k + 1]) & 15u;
                if ((pv & 1u) && ((pv >> (1 + 4 * k)) & 15u) == code) st.change_time[G(k)] += 1;
                nv |= code << (1 + 4 * k);
            }
            st.prev_act[e] = nv;
        }
    }
    // ... and the part behind it: the step's rewards and done flag
    AT_HD bool wrapper_post(float *rewards_row) {
        for (int r = 0; r < cfg_rows(); ++r)
            rewards_row[r] = (float)TREW(cfg_team_layout() ? cfg_agent_tank(r) : r);
        bool done;
        if (cfg_terminate() != 0) {
            done = !((int)tstep < cfg_terminate());
            if (done) tstep = 0;
        } else {
            uint32_t teams = 0;
            int winner_team = -1;
            for (int i = 0; i < cfg_n(); ++i)
                if (t_alive(i)) {
                    teams |= 1u << cfg_team(i);
                    if (winner_team < 0) winner_team = cfg_team(i);
                }
            int nalive = popc64(teams);
            done = nalive <= 1;
            if (cfg_team_layout() && nalive == 1)  // F5: lands after the rewards were read
                for (int i = 0; i < cfg_n(); ++i)
                    if (cfg_team(i) == winner_team) TREW(i) += 200;
        }
        return done;
    }
    template <class WP>
    AT_HD bool wrapper_step(const int32_t *a, float *rewards_row) {
        wrapper_pre(a);
        game_step<WP>(a);
        return wrapper_post(rewards_row);
    }
    // the step's results of this lane's env: rewards / done out, terminal info latched, auto-reset
    AT_HD void finish_step(float *rewards, uint8_t *done) {
        float rw[MAX_TANKS];
        const bool d = wrapper_post(rw);
        if (c.reward_delta)   // AT_REWARD_DELTA: the increment since the value the previous step returned
            for (int r = 0; r < cfg_rows(); ++r) {
                const float cum = rw[r];
                rw[r] = cum - st.prev_rew[r * n32 + e];
                st.prev_rew[r * n32 + e] = (d && c.auto_reset) ? 0.0f : cum;
            }
        if (cfg_rows() == 2) {
#ifdef __CUDA_ARCH__
            *reinterpret_cast<float2 *>(rewards + (int64_t)e * 2) = make_float2(rw[0], rw[1]);  // one 8-byte store
#else
            rewards[(int64_t)e * 2] = rw[0]; rewards[(int64_t)e * 2 + 1] = rw[1];
#endif
        } else {
            for (int r = 0; r < cfg_rows(); ++r) rewards[(int64_t)e * c.rows + r] = rw[r];
        }
        done[e] = d ? 1 : 0;
        if (d && c.auto_reset) { latch_terminal_info(); env_reset(); }
    }

    // ---------------------------------------------------------------- HBM <-> working set
    AT_HD void load() {
        cnt = st.cnt[e]; rng = st.rng_ctr[e]; tick = st.tick[e]; tstep = st.tstep[e]; epstep = st.epstep[e];
        zones = st.zones[e];
        if (cfg_maze()) { wlo = st.wall_lo[e]; whi = st.wall_hi[e]; set_near(); }
        else { wlo = c.wall_lo; whi = c.wall_hi; nlo = c.near_lo; nhi = c.near_hi; }
        // Four tanks at a time through registers: every load of a batch is issued before the first shared-memory
        // store (the compiler cannot tell that the working-set columns and the SoA columns never alias, and a
        // load / store / load / store sequence would be one L2 round trip per column - 40 of them in a row).
        const int n = cfg_n();
        for (int i0 = 0; i0 < n; i0 += 4) {
            double x[4], y[4], rw[4], f0[4], f1[4];
            uint32_t tp[4], bp[4];
            int32_t ts[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = i0 + k;
                x[k] = y[k] = rw[k] = f0[k] = f1[k] = 0.0; tp[k] = bp[k] = 0u; ts[k] = 0;
                if (i < n) {
                    x[k] = st.tx[G(i)]; y[k] = st.ty[G(i)]; rw[k] = st.trew[G(i)];
                    tp[k] = st.tpack[G(i)]; ts[k] = st.tshot[G(i)];
                    if (cfg_ctrl(i) == 1) { bp[k] = st.bpack[G(i)]; f0[k] = st.bf0[G(i)]; f1[k] = st.bf1[G(i)]; }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = i0 + k;
                if (i < n) {
                    TX(i) = x[k]; TY(i) = y[k]; TREW(i) = rw[k]; TPACK(i) = tp[k]; TSHOT(i) = ts[k];
                    if (cfg_ctrl(i) == 1) { BPACK(i) = bp[k]; BF0(i) = f0[k]; BF1(i) = f1[k]; }
                    if (tp[k] & TP_ALIVE) alive_bits |= 1u << i;
                }
            }
        }
    }
    AT_HD void set_near() { M128 nm = near_wall_mask(wlo, whi); nlo = nm.lo; nhi = nm.hi; }
    AT_HD void load_scalars_only() {
        cnt = st.cnt[e]; rng = st.rng_ctr[e]; tick = st.tick[e]; tstep = st.tstep[e]; epstep = st.epstep[e];
        zones = st.zones[e];
        if (cfg_maze()) { wlo = st.wall_lo[e]; whi = st.wall_hi[e]; }
        else { wlo = c.wall_lo; whi = c.wall_hi; }
    }
    AT_HD void store() {
        st.cnt[e] = cnt; st.rng_ctr[e] = rng; st.tick[e] = tick; st.tstep[e] = tstep; st.epstep[e] = epstep;
        st.zones[e] = zones;
        if (cfg_maze()) { st.wall_lo[e] = wlo; st.wall_hi[e] = whi; }
        for (int i = 0; i < cfg_n(); ++i) {
            st.tx[G(i)] = TX(i); st.ty[G(i)] = TY(i); st.trew[G(i)] = TREW(i);
            st.tpack[G(i)] = TPACK(i); st.tshot[G(i)] = TSHOT(i);
            if (cfg_ctrl(i) == 1) { st.bpack[G(i)] = BPACK(i); st.bf0[G(i)] = BF0(i); st.bf1[G(i)] = BF1(i); }
        }
    }
};
typedef SimT<0> Sim;   // the generic build (emulator, reset kernel, every config without a specialisation)

// One lane's part of a step: the body of env_kernel<true, SPEC> (and of the emulator's fibers).  e_warp0 = SoA index of
// the env of lane 0; valid_mask = lanes of the warp that own an env (the pooled line-of-sight stage runs on these).
template <class WP, int SPEC>
AT_HD void step_lane(const KCfg &c, const DevState &st, const BlockMem &bm, int64_t e_warp0, int tid, uint32_t valid_mask,
                     const int32_t *my_actions, float *rewards, uint8_t *done) {
    SimT<SPEC> s(c, st, bm, e_warp0 + (tid & 31), tid);
    s.warp_mask = valid_mask;
    s.load();
    s.wrapper_pre(my_actions);
    s.template game_step<WP>(my_actions);
    s.finish_step(rewards, done);
    s.store();
}
// One lane's part of a (masked) reset: the generic build.
AT_HD void reset_lane(const KCfg &c, const DevState &st, const BlockMem &bm, int64_t e, int tid) {
    Sim s(c, st, bm, e, tid);
    s.load();
    s.env_reset();
    s.store();
}

// wall rectangles (row-major = reference list order) + maze grid columns of a fixed map
AT_HD void fill_static_template(uint64_t wall_lo, uint64_t wall_hi, float *tmpl) {
    int w = 0;
    for (int cell = 0; cell < CELLS; ++cell)
        if ((cell < 64 ? wall_lo >> cell : wall_hi >> (cell - 64)) & 1ull) {
            const float X = (float)((cell % 11) * 70), Y = (float)((cell / 11) * 70);
            tmpl[4 * w] = X; tmpl[4 * w + 1] = Y; tmpl[4 * w + 2] = X + 70.0f; tmpl[4 * w + 3] = Y + 70.0f;
            ++w;
        }
    for (int cell = 0; cell < CELLS; ++cell)
        tmpl[4 * w + cell] = ((cell < 64 ? wall_lo >> cell : wall_hi >> (cell - 64)) & 1ull) ? 1.0f : 0.0f;
}

}  // namespace at
