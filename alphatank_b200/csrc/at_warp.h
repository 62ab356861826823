// at_warp.h - the warp-collective operations the simulation uses, behind a policy class.
//
// The pooled stage of the simulation (at_sim.h: pool_los) is written once against a policy `WP`: DevWarp (below) maps
// every operation onto the sm_100a warp intrinsics; the test-only SIMT emulator (tests/cpu_emul/simt.h) provides the
// same interface on top of 32 fibers per warp, so that the very same cooperative source runs - and is compared with
// the CPU restatement of the reference - in the GPU-less build container.
//
//   WP::sync(mask)                   __syncwarp: execution + memory barrier among the lanes of `mask`
//   WP::ballot(mask, pred)           __ballot_sync
//   WP::shfl(mask, v, src)           __shfl_sync (32-bit)
#pragma once
#include <stdint.h>

#include "at_types.h"

namespace at {

#if defined(__CUDACC__)
struct DevWarp {
    __device__ __forceinline__ static void sync(uint32_t mask) { __syncwarp(mask); }
    __device__ __forceinline__ static uint32_t ballot(uint32_t mask, bool pred) { return __ballot_sync(mask, pred); }
    __device__ __forceinline__ static uint32_t shfl(uint32_t mask, uint32_t v, int src) { return __shfl_sync(mask, v, src); }
};
#endif

}  // namespace at
