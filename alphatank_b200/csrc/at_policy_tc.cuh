// at_policy_tc.cuh - the policies' first layers on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// models/ppo_utils.py:31-45: every policy is four MLPs (critic + three action heads) D -> 128 -> 128 -> k that share
// their input, so their first layers are ONE contraction z1[N, 512] = x[N, D] . W1[512, D]^T + b1 (D = 528 for the 2v2
// rows).  policy_first_kernel computes it per 128-row tile:
//   * TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) streams 32-column chunks of the x tile and of W1 into a
//     two-stage shared-memory ring, completion on mbarriers (transaction bytes); rows / columns past the end of the
//     tensors arrive as zeros, which is how the K tail (528 = 16.5 chunks) and a ragged last tile are handled;
//   * one elected thread issues tcgen05.mma.kind::tf32 (M = 128, N = 256, K = 8 per instruction, two instructions per
//     k-step for the 512 output columns) straight from the swizzled shared-memory tiles - fp32 in, TF32 multiply, fp32
//     accumulate, the precision of the cuBLAS TF32 path it replaces - into a 128-lane x 512-column fp32 accumulator
//     that fills the SM's tensor memory; tcgen05.commit hands each ring slot back to the producer and, after the last
//     k-step, the accumulator to the epilogue;
//   * the four warps read their 32 TMEM lanes (= rows) back with tcgen05.ld (32 columns per instruction), add the bias
//     and store the pre-activations.
// Roles: warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, all warps = epilogue.  One CTA per SM (160 KB of
// operands, all 512 TMEM columns).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace at_tc {

constexpr int TM = 128;            // rows per tile (TMEM lanes)
constexpr int TN = 512;            // the four first layers side by side
constexpr int KC = 32;             // fp32 columns per chunk = one 128-byte swizzle row
constexpr int STAGES = 2;
constexpr uint32_t A_BYTES = TM * KC * 4;          // 16 KB
constexpr uint32_t B_BYTES = TN * KC * 4;          // 64 KB
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;   // (+ alignment slack, + barriers)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
                 : "memory");
}
// shared-memory matrix descriptor of a K-major tile in the 128-byte-swizzle layout TMA writes: rows of 128 bytes,
// 8-row groups 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N / 8 at bit 17, M / 16 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {   // arrives on `bar` when every MMA issued so far has finished
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {   // 32 consecutive columns of this thread's lane
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// z1[n_rows][512] = x . W1^T + b1.  map_x: the [n_rows][D] fp32 rows (row stride arbitrary, a multiple of 16 bytes),
// map_w: W1 [512][D]; both encoded with a {32, 128} / {32, 256} box and the 128-byte swizzle.
__global__ void __launch_bounds__(128, 1) policy_first_kernel(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_w,
                                                               const float *__restrict__ b1, float *__restrict__ z1, int n_rows,
                                                               int k_chunks) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1024-byte aligned
    unsigned char *stage_a[STAGES], *stage_b[STAGES];
    for (int s = 0; s < STAGES; ++s) { stage_a[s] = smem + (size_t)s * STAGE_BYTES; stage_b[s] = stage_a[s] + A_BYTES; }
    uint64_t *bars = (uint64_t *)(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + STAGES, *acc_full = bars + 2 * STAGES;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * TM;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {   // the whole tensor memory of the SM: 128 lanes x 512 fp32 columns
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer
        for (int k = 0; k < k_chunks; ++k) {
            const int s = k % STAGES;
            if (k >= STAGES) mbar_wait(empty + s, (uint32_t)((k / STAGES - 1) & 1));
            mbar_expect_tx(full + s, STAGE_BYTES);
            tma_load_2d(stage_a[s], &map_x, full + s, k * KC, row0);
            tma_load_2d(stage_b[s], &map_w, full + s, k * KC, 0);
            tma_load_2d(stage_b[s] + 256 * KC * 4, &map_w, full + s, k * KC, 256);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer
        constexpr uint32_t idesc = umma_idesc_tf32(TM, 256);
        for (int k = 0; k < k_chunks; ++k) {
            const int s = k % STAGES;
            mbar_wait(full + s, (uint32_t)((k / STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t a0 = smem_u32(stage_a[s]), b0 = smem_u32(stage_b[s]);
#pragma unroll
            for (int kk = 0; kk < KC / 8; ++kk) {   // 8 tf32 = 32 bytes per instruction along K, inside the swizzle row
                const uint64_t da = umma_desc_sw128(a0 + kk * 32);
                const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
                umma_tf32(tmem_base, da, umma_desc_sw128(b0 + kk * 32), idesc, acc);
                umma_tf32(tmem_base + 256, da, umma_desc_sw128(b0 + 256 * KC * 4 + kk * 32), idesc, acc);
            }
            umma_commit(empty + s);                      // the slot is free once these MMAs have read it
            if (k == k_chunks - 1) umma_commit(acc_full);   // ... and the accumulator is complete
        }
    }
    __syncwarp();
    // ===== epilogue: every thread owns one row (TMEM lane 32 * warp + lane)
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int row = row0 + warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + (uint32_t)c0, v);
        if (row < n_rows) {
            float4 *dst = reinterpret_cast<float4 *>(z1 + (size_t)row * TN + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b = __ldg(reinterpret_cast<const float4 *>(b1 + c0) + q);
                dst[q] = make_float4(__uint_as_float(v[4 * q]) + b.x, __uint_as_float(v[4 * q + 1]) + b.y,
                                     __uint_as_float(v[4 * q + 2]) + b.z, __uint_as_float(v[4 * q + 3]) + b.w);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)TN) : "memory");
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// a [rows][cols] fp32 matrix with `row_stride_bytes` between rows, read in boxes of box_rows x 32 columns
inline bool make_map_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows}, strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)KC, box_rows}, estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace at_tc
