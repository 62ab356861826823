// at_warp.h - the warp-collective operations the simulation uses, behind a policy class.
//
// The pooled stages of the simulation (at_sim.h: bullets_pass_pooled, pool_los, ...) are written once against a
// policy `WP`:  DevWarp (below) maps every operation onto the sm_100a warp intrinsics; the test-only SIMT emulator
// (tests/cpu_emul/simt.h) provides the same interface on top of 32 fibers per warp, so that the very same
// cooperative source runs - and is compared with the oracle - in the GPU-less build container.
//
//   WP::lane()                       lane index 0..31
//   WP::sync(mask)                   __syncwarp: execution + memory barrier among the lanes of `mask`
//   WP::ballot(mask, pred)           __ballot_sync
//   WP::shfl(mask, v, src)           __shfl_sync (32-bit)
//   WP::shfl_up(mask, v, d)          __shfl_up_sync (32-bit)
//   WP::atomic_or / atomic_min / atomic_add (32-bit) and atomic_add64 on shared memory
#pragma once
#include <stdint.h>

#include "at_types.h"

namespace at {

#if defined(__CUDACC__)
struct DevWarp {
    static constexpr bool kDevice = true;
    __device__ __forceinline__ static int lane() { return (int)(threadIdx.x & 31u); }
    __device__ __forceinline__ static void sync(uint32_t mask) { __syncwarp(mask); }
    __device__ __forceinline__ static uint32_t ballot(uint32_t mask, bool pred) { return __ballot_sync(mask, pred); }
    __device__ __forceinline__ static uint32_t shfl(uint32_t mask, uint32_t v, int src) { return __shfl_sync(mask, v, src); }
    __device__ __forceinline__ static uint32_t shfl_up(uint32_t mask, uint32_t v, int d) { return __shfl_up_sync(mask, v, d); }
    __device__ __forceinline__ static uint32_t atomic_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
    __device__ __forceinline__ static uint32_t atomic_min(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
    __device__ __forceinline__ static uint32_t atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
    __device__ __forceinline__ static unsigned long long atomic_add64(unsigned long long *p, unsigned long long v) {
        return atomicAdd(p, v);
    }
};
#endif

// One lane per call, no collectives: the serial policy of the host-side table builders and of code paths that never
// reach a pooled stage (every collective degenerates to its single-lane meaning).
struct SoloWarp {
    static constexpr bool kDevice = false;
    AT_HD static int lane() { return 0; }
    AT_HD static void sync(uint32_t) {}
    AT_HD static uint32_t ballot(uint32_t, bool pred) { return pred ? 1u : 0u; }
    AT_HD static uint32_t shfl(uint32_t, uint32_t v, int) { return v; }
    AT_HD static uint32_t shfl_up(uint32_t, uint32_t v, int) { return v; }
    AT_HD static uint32_t atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
    AT_HD static uint32_t atomic_min(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = v < o ? v : o; return o; }
    AT_HD static uint32_t atomic_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
    AT_HD static unsigned long long atomic_add64(unsigned long long *p, unsigned long long v) {
        unsigned long long o = *p; *p = o + v; return o;
    }
};

// inclusive prefix sum of a small per-lane count over the lanes of a full-warp mask
template <class WP>
AT_HD uint32_t warp_inclusive_sum(uint32_t mask, uint32_t v) {
    const int lane = WP::lane();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = WP::shfl_up(mask, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

}  // namespace at
