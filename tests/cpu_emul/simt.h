// simt.h - TEST-ONLY SIMT emulator: one warp = 32 fibers (ucontext), warp collectives resolved by a scheduler.
//
// The simulation's cooperative stages (alphatank_b200/csrc/at_sim.h) are written against a warp policy
// (at_warp.h).  On the GPU the policy is the warp intrinsics; here it is FiberWarp: every lane of a warp runs the very
// same per-thread function on its own fiber until it reaches a collective, yields, and the scheduler completes the
// collective once every lane named in its mask has arrived.  Between collectives the lanes run one after the other in
// an order that can be forward, reverse or shuffled per scheduling round (set_order): code that reads another lane's
// shared-memory write without a barrier in between gives order-dependent results, which the tests look for by
// running the same step under several orders.
// Not part of the product; never loaded by alphatank_b200/.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <functional>
#include <vector>

namespace simt {

enum Op { OP_NONE = 0, OP_SYNC, OP_BALLOT, OP_SHFL, OP_SHFL_UP };

struct Lane {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = true, waiting = false;
    int op = OP_NONE;
    uint32_t mask = 0, arg = 0, result = 0;
    int src = 0;
};

struct Warp {
    Lane lanes[32];
    ucontext_t main;
    int cur = -1;
    int order_mode = 0;       // 0 forward, 1 reverse, 2 shuffled per round
    uint64_t order_state = 0x9E3779B97F4A7C15ull;
    std::function<void(int)> body;
    uint64_t collectives = 0;
};

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}

inline void lane_entry(int lane) {
    Warp *w = current();
    w->body(lane);
    w->lanes[lane].done = true;
    swapcontext(&w->lanes[lane].ctx, &w->main);
}

inline uint32_t collective(int op, uint32_t mask, uint32_t arg, int src) {
    Warp *w = current();
    Lane &L = w->lanes[w->cur];
    if (!((mask >> w->cur) & 1u)) {
        fprintf(stderr, "simt: lane %d calls a collective whose mask %08x does not name it\n", w->cur, mask);
        abort();
    }
    L.waiting = true; L.op = op; L.mask = mask; L.arg = arg; L.src = src;
    swapcontext(&L.ctx, &w->main);
    return L.result;
}

// complete every collective whose lanes have all arrived; returns whether anything was released
inline bool resolve(Warp *w) {
    bool progress = false;
    for (int l = 0; l < 32; ++l) {
        Lane &L = w->lanes[l];
        if (L.done || !L.waiting) continue;
        const uint32_t mask = L.mask;
        bool ready = true;
        for (int k = 0; k < 32 && ready; ++k)
            if ((mask >> k) & 1u) {
                const Lane &K = w->lanes[k];
                if (K.done || !K.waiting || K.mask != mask || K.op != L.op) ready = false;
            }
        if (!ready) continue;
        uint32_t bal = 0;
        if (L.op == OP_BALLOT)
            for (int k = 0; k < 32; ++k)
                if (((mask >> k) & 1u) && w->lanes[k].arg) bal |= 1u << k;
        uint32_t vals[32];
        for (int k = 0; k < 32; ++k) vals[k] = w->lanes[k].arg;
        for (int k = 0; k < 32; ++k)
            if ((mask >> k) & 1u) {
                Lane &K = w->lanes[k];
                switch (K.op) {
                case OP_BALLOT: K.result = bal; break;
                case OP_SHFL: { const int s = K.src & 31; K.result = ((mask >> s) & 1u) ? vals[s] : vals[k]; break; }
                case OP_SHFL_UP: { const int s = k - K.src; K.result = (s >= 0 && ((mask >> s) & 1u)) ? vals[s] : vals[k]; break; }
                default: K.result = 0; break;
                }
                K.waiting = false;
            }
        ++w->collectives;
        progress = true;
    }
    return progress;
}

// run body(lane) for the lanes of `active` as one warp
inline void run_warp(Warp &w, uint32_t active, const std::function<void(int)> &body) {
    w.body = body;
    Warp *prev = current();
    current() = &w;
    for (int l = 0; l < 32; ++l) {
        Lane &L = w.lanes[l];
        L.done = !((active >> l) & 1u);
        L.waiting = false;
        if (L.done) continue;
        if (L.stack.empty()) L.stack.resize(256 * 1024);
        getcontext(&L.ctx);
        L.ctx.uc_stack.ss_sp = L.stack.data();
        L.ctx.uc_stack.ss_size = L.stack.size();
        L.ctx.uc_link = nullptr;
        makecontext(&L.ctx, (void (*)())lane_entry, 1, l);
    }
    for (;;) {
        int order[32];
        for (int k = 0; k < 32; ++k) order[k] = w.order_mode == 1 ? 31 - k : k;
        if (w.order_mode == 2)
            for (int k = 31; k > 0; --k) {
                w.order_state = w.order_state * 6364136223846793005ull + 1442695040888963407ull;
                const int j = (int)((w.order_state >> 33) % (uint64_t)(k + 1));
                const int t = order[k]; order[k] = order[j]; order[j] = t;
            }
        bool ran = false, all_done = true;
        for (int k = 0; k < 32; ++k) {
            Lane &L = w.lanes[order[k]];
            if (L.done) continue;
            all_done = false;
            if (L.waiting) continue;
            w.cur = order[k];
            swapcontext(&w.main, &L.ctx);
            ran = true;
        }
        if (all_done) break;
        const bool rel = resolve(&w);
        if (!ran && !rel) {
            fprintf(stderr, "simt: deadlock - lanes wait on collectives that can never complete:\n");
            for (int l = 0; l < 32; ++l)
                if (!w.lanes[l].done)
                    fprintf(stderr, "  lane %2d op %d mask %08x\n", l, w.lanes[l].op, w.lanes[l].mask);
            abort();
        }
    }
    w.cur = -1;
    current() = prev;
}

// the warp policy (interface of at::DevWarp, alphatank_b200/csrc/at_warp.h)
struct FiberWarp {
    static void sync(uint32_t mask) { collective(OP_SYNC, mask, 0, 0); }
    static uint32_t ballot(uint32_t mask, bool pred) { return collective(OP_BALLOT, mask, pred ? 1u : 0u, 0); }
    static uint32_t shfl(uint32_t mask, uint32_t v, int src) { return collective(OP_SHFL, mask, v, src); }
};

}  // namespace simt
