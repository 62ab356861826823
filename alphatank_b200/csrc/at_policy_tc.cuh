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

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {   // ... the caller waits (tmem_ld_wait)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// z1[n_rows][512] = x . W1^T + b1.  map_x: the [n_rows][D] fp32 rows (row stride arbitrary, a multiple of 16 bytes),
// map_w: W1 [512][D]; both encoded with a {32, 128} / {32, 256} box and the 128-byte swizzle.
__global__ void __launch_bounds__(128, 1) policy_first_kernel(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_w,
                                                               const float *__restrict__ b1, float *__restrict__ z1, int n_rows,
                                                               int k_chunks) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms: 1024-byte aligned (offset arithmetic keeps the shared address space)
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


// ====================================================================================================================
// policy_forward_kernel<HALF_IN> - a policy's whole inference pass in ONE kernel (models/ppo_utils.py:31-62, the four MLPs
// D -> 128 -> 128 -> {1, 3, 3, 2} of train_multi_ppo.py:96-116's `get_action_and_value`): TMA-loaded observation tiles ->
// first layers on tcgen05.mma (fp32 accumulators in tensor memory) -> tanh -> second layers on tcgen05.mma -> tanh ->
// output layers -> log-softmax -> Gumbel-max draw.  Neither the [N, 512] activations nor the logits go to HBM.
//   HALF_IN = true   fp16 rows (at_step_norm16) and fp16 W1: kind::f16, K chunks of 64 columns
//   HALF_IN = false  fp32 rows (at_step_norm) and fp32 W1: kind::tf32 (the tensor core reads the upper 19 bits), K chunks of 32
// Either way a chunk is one 128-byte swizzle row per matrix row, the hidden activations and W2 are fp16 (tanh's output
// lies in [-1, 1]; fp16 carries TF32's 10 mantissa bits) and everything is accumulated in fp32.
//
// A CTA (one per SM, persistent over 128-row tiles) has three roles:
//   warp 0, one lane   TMA producer: a three-slot ring of {x chunk [128 rows], W1 chunk [256 rows]} (48 KB) for the first
//                      layers of a PAIR of networks, and W2 of each network as one slot;
//   warp 1, one lane   MMA issuer: the pair's first layers (M 128, N 256) into tensor-memory columns 0-255 (X); for every
//                      network, once its activations are in shared memory, the second layer (M 128, N 128) into
//                      Y[j & 1] = columns 256 + 128 (j & 1).  X is free as soon as its two networks are converted, so the
//                      next pair's first layers run while the previous pair is read out of Y;
//   warps 2-17         epilogue, thread = (row, quarter of a network's columns): tcgen05.ld, + bias, tanh.approx, fp16,
//                      16-byte stores into the second layer's A operand in the 128-byte-swizzle layout the tensor core
//                      reads; later tcgen05.ld of the second layer, + bias, tanh, dot with the <= 3 output rows; the four
//                      quarters of a row meet in shared memory; the heads are sampled once per tile.
// K chunks that only hold columns which are constant for the handle's lifetime (the wall / maze block of a fixed map: 281 of
// 528 columns) can be left out by the caller, their contribution folded into the bias (`chunk_mask`, at_policy_forward).
constexpr int F_TABLE_FLOATS = 512 + 512 + 9 * 128 + 12;

struct PolicyIO {            // one policy's tables and outputs
    const float *b1, *b2, *w3, *b3;   // [512] (folded), [4][128], [9][128], [9]
    int32_t *actions; float *logprobs; float *values; float *out9;
    int64_t act_stride, lp_stride, val_stride;
    uint32_t stream_id;
};
struct PolicyFwd {
    PolicyIO io[2];
    int64_t n_rows;          // rows per policy
    int tiles_per_policy, n_policies;
    uint32_t chunk_mask;     // K chunks (32 columns each) of the first layers to contract
    uint64_t seed;
    long long *counter;        // [0] the tick counter the draws are keyed by; [1] finished-CTA count (advance != 0)
    int advance;               // the last CTA to finish adds 1 to counter[0] (every CTA has read it by then)
    unsigned long long *dbg;   // diagnostics: [2][512] (event id, clock) pairs of block 0, or nullptr
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {   // MUFU.TANH: relative error <= 2^-11, the TF32 operand precision
    float y;
    asm("tanh.approx.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float round_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
// Why fp16 where the caller can provide it: tcgen05.mma.kind::f16 runs at twice the TF32 rate and every operand byte counts
// half, which matters because the first layers are bound by what an SM can pull from L2 (~130 GB/s measured: each 128-row
// tile streams all of W1).
constexpr int H_KC = 64;                                  // fp16 columns per chunk
constexpr int H_STAGES = 3;
constexpr uint32_t HA_BYTES = TM * H_KC * 2;              // 16 KB: x chunk
constexpr uint32_t HB_BYTES = 256 * H_KC * 2;             // 32 KB: W1 chunk of one half, or a whole W2 (two chunks)
constexpr uint32_t H_STAGE_BYTES = HA_BYTES + HB_BYTES;
constexpr uint32_t HH1_BYTES = 2 * HA_BYTES;              // 32 KB: one network's activations (two K chunks)
constexpr int H_EPI_WARPS = 16, H_THREADS = 64 + H_EPI_WARPS * 32;
constexpr size_t H_SMEM_BYTES = 1024 + (size_t)H_STAGES * H_STAGE_BYTES + 2 * HH1_BYTES + F_TABLE_FLOATS * 4 + 2 * 3 * 384 * 4 + 24 * 8 + 16;

// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = F16 (format 0), K-major, N / 8 at bit 17, M / 16 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <bool HALF_IN>
__global__ void __launch_bounds__(H_THREADS, 1)
policy_forward_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                        const __grid_constant__ CUtensorMap map_w1_0, const __grid_constant__ CUtensorMap map_w1_1,
                        const __grid_constant__ CUtensorMap map_w2_0, const __grid_constant__ CUtensorMap map_w2_1,
                        const __grid_constant__ PolicyFwd P) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms: 1024-byte aligned (offset arithmetic keeps the shared address space)
    unsigned char *h1 = smem + (size_t)H_STAGES * H_STAGE_BYTES;            // two buffers of HH1_BYTES
    float *tab = (float *)(h1 + 2 * HH1_BYTES);
    float *b1s = tab, *b2s = tab + 512, *w3s = tab + 1024, *b3s = tab + 1024 + 9 * 128;
    float *part = tab + F_TABLE_FLOATS;                                      // [2 buffers][3 senders][3][128]: partial outputs
    uint64_t *bars = (uint64_t *)(part + 2 * 3 * 384);
    // tensor memory: X = columns 0-255, the first layers of the current pair of networks; Y[b] = columns 256 + 128 b, the
    // second layer of network j (b = j & 1).  X is free again as soon as its two networks are converted, so the next
    // pair's first layers run while the epilogue reads the previous pair's outputs out of Y.
    uint64_t *full = bars, *empty = bars + H_STAGES, *acc1_full = bars + 2 * H_STAGES, *x_free = acc1_full + 1,
             *h1_ready = x_free + 1, *l2_done = h1_ready + 2, *y_free = l2_done + 2;
    uint32_t *tmem_slot = (uint32_t *)(y_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = P.tiles_per_policy * P.n_policies;
    const int n_chunks = __popc(P.chunk_mask);
    if (threadIdx.x == 0) {
        for (int s = 0; s < H_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc1_full, 1);
        mbar_init(x_free, H_EPI_WARPS * 32);
        for (int h = 0; h < 2; ++h) {
            mbar_init(h1_ready + h, H_EPI_WARPS * 32); mbar_init(l2_done + h, 1); mbar_init(y_free + h, H_EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int m_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // diagnostics (at_policy_debug_stamps): block 0 records (event id, SM clock) pairs per role - 0: MMA issuer, 1: epilogue
    int n_stamp = 0;
    auto stamp = [&](int role, int id) {
        if (P.dbg != nullptr && blockIdx.x == 0 && n_stamp < 255) {
            P.dbg[role * 512 + 2 * n_stamp] = (unsigned long long)id;
            P.dbg[role * 512 + 2 * n_stamp + 1] = (unsigned long long)clock64();
            ++n_stamp;
        }
    };

    // tensor-core order (the producer fills the ring in the same order), local tiles 0 .. m - 1, networks j = 4 i + net:
    //   per tile i:   L1(i, h0) L2(i, n0) L2(i, n1) L1(i, h1) L2(i, n2) L2(i, n3)
    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            auto slot = [&](uint32_t bytes) -> unsigned char * {
                const uint32_t s = g % H_STAGES;
                if (g >= H_STAGES) mbar_wait(empty + s, ((g / H_STAGES) - 1) & 1);
                mbar_expect_tx(full + s, bytes);
                return smem + (size_t)s * H_STAGE_BYTES;
            };
            auto l1 = [&](int i, int h) {
                const int tile = blockIdx.x + i * gridDim.x;
                const int p = tile / P.tiles_per_policy, row0 = (tile - p * P.tiles_per_policy) * TM;
                const CUtensorMap *mx = p ? &map_x1 : &map_x0, *mw1 = p ? &map_w1_1 : &map_w1_0;
                uint32_t m = P.chunk_mask;
                while (m) {
                    const int c = __ffs(m) - 1;
                    m &= m - 1;
                    unsigned char *sa = slot(H_STAGE_BYTES);
                    tma_load_2d(sa, mx, full + g % H_STAGES, c * (HALF_IN ? H_KC : KC), row0);
                    tma_load_2d(sa + HA_BYTES, mw1, full + g % H_STAGES, c * (HALF_IN ? H_KC : KC), h * 256);
                    ++g;
                }
            };
            auto l2 = [&](int i, int net) {
                const int tile = blockIdx.x + i * gridDim.x;
                const CUtensorMap *mw2 = tile / P.tiles_per_policy ? &map_w2_1 : &map_w2_0;
                unsigned char *sb = slot(HB_BYTES) + HA_BYTES;
                tma_load_2d(sb, mw2, full + g % H_STAGES, 0, net * 128);
                tma_load_2d(sb + HA_BYTES, mw2, full + g % H_STAGES, H_KC, net * 128);
                ++g;
            };
            for (int i = 0; i < m_tiles; ++i) {
                l1(i, 0); l2(i, 0); l2(i, 1);
                l1(i, 1); l2(i, 2); l2(i, 3);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc1 = HALF_IN ? umma_idesc_f16(TM, 256) : umma_idesc_tf32(TM, 256), idesc2 = umma_idesc_f16(TM, 128);
            uint32_t g = 0;
            auto l1 = [&](int i, int h) {
                const int u = 2 * i + h;
                stamp(0, 100 + h);
                if (u > 0) {   // the previous pair of networks has been converted out of X
                    mbar_wait(x_free, (uint32_t)((u - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                }
                for (int ci = 0; ci < n_chunks; ++ci) {
                    const uint32_t s = g % H_STAGES;
                    mbar_wait(full + s, (g / H_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t a0 = smem_u32(smem + (size_t)s * H_STAGE_BYTES), b0 = a0 + HA_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {   // 16 halves / 8 floats = 32 bytes per instruction along K
                        if (HALF_IN) umma_f16(tmem_base, umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(b0 + kk * 32), idesc1, (ci > 0 || kk > 0) ? 1u : 0u);
                        else umma_tf32(tmem_base, umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(b0 + kk * 32), idesc1, (ci > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(empty + s);
                    ++g;
                }
                umma_commit(acc1_full);
                stamp(0, 110 + h);
            };
            auto l2 = [&](int i, int net) {
                const int j = 4 * i + net, b = j & 1;
                stamp(0, 200 + net);
                mbar_wait(h1_ready + b, (uint32_t)((j >> 1) & 1));
                if (j >= 2) mbar_wait(y_free + b, (uint32_t)(((j >> 1) - 1) & 1));   // network j - 2 has been read out of Y[b]
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t s = g % H_STAGES;
                mbar_wait(full + s, (g / H_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a0 = smem_u32(h1 + (size_t)b * HH1_BYTES), b0 = smem_u32(smem + (size_t)s * H_STAGE_BYTES) + HA_BYTES;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                    for (int kk = 0; kk < H_KC / 16; ++kk)
                        umma_f16(tmem_base + 256 + b * 128, umma_desc_sw128(a0 + cc * HA_BYTES + kk * 32),
                                 umma_desc_sw128(b0 + cc * HA_BYTES + kk * 32), idesc2, (cc > 0 || kk > 0) ? 1u : 0u);
                umma_commit(empty + s);
                ++g;
                umma_commit(l2_done + b);
                stamp(0, 210 + net);
            };
            for (int i = 0; i < m_tiles; ++i) {
                l1(i, 0); l2(i, 0); l2(i, 1);
                l1(i, 1); l2(i, 2); l2(i, 3);
            }
        }
    } else {
        // ===== epilogue: thread = (row, column quarter); TMEM lane quarter = warp % 4, 32 of a network's 128 columns each
        const int q = warp & 3, cq = (warp - 2) >> 2, r = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint64_t ctr = P.counter ? (uint64_t)*P.counter : 0ull;
        int cur_p = -1, it = 0;
        // first-layer columns [32 cq, 32 cq + 32) of `net` -> + bias -> tanh -> fp16 -> half of K chunk cq / 2 of the second
        // layer's A operand (128-byte swizzle: the 16-byte unit u of row r lives at unit u ^ (r & 7))
        auto conv = [&](int j, int net) {
            unsigned char *dst = h1 + (size_t)(j & 1) * HH1_BYTES + (size_t)(cq >> 1) * HA_BYTES + (size_t)r * 128;
            uint32_t v[32];
            tmem_ld32(lane_addr + (uint32_t)((net & 1) * 128 + cq * 32), v);
            const float *bb = b1s + net * 128 + cq * 32;
#pragma unroll
            for (int u = 0; u < 4; ++u) {   // 8 columns = one 16-byte unit
                uint4 w;
                w.x = pack_half2(tanh_approx(__uint_as_float(v[8 * u]) + bb[8 * u]), tanh_approx(__uint_as_float(v[8 * u + 1]) + bb[8 * u + 1]));
                w.y = pack_half2(tanh_approx(__uint_as_float(v[8 * u + 2]) + bb[8 * u + 2]), tanh_approx(__uint_as_float(v[8 * u + 3]) + bb[8 * u + 3]));
                w.z = pack_half2(tanh_approx(__uint_as_float(v[8 * u + 4]) + bb[8 * u + 4]), tanh_approx(__uint_as_float(v[8 * u + 5]) + bb[8 * u + 5]));
                w.w = pack_half2(tanh_approx(__uint_as_float(v[8 * u + 6]) + bb[8 * u + 6]), tanh_approx(__uint_as_float(v[8 * u + 7]) + bb[8 * u + 7]));
                *reinterpret_cast<uint4 *>(dst + ((((cq & 1) * 4 + u) ^ (r & 7)) << 4)) = w;
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(h1_ready + (j & 1));
        };
        // second-layer columns of `net` -> its <= 3 outputs.  The four column quarters of a row meet in `part` (two buffers,
        // one barrier per network): network 0 (value) and 1 (head 0) are completed by quarter 0's threads, network 2 (head
        // 1) by quarter 1's, network 3 (head 2) by quarter 2's, and each group samples its head once per tile (`finish`).
        float lg[4] = {0.f, 0.f, 0.f, 0.f};   // quarter 0: value + head 0; quarter 1 / 2: the head's logits
        auto readout = [&](int j, int net) {
            mbar_wait(l2_done + (j & 1), (uint32_t)((j >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (threadIdx.x == 64) stamp(1, 400 + net);
            const int first = net == 0 ? 0 : (net == 1 ? 1 : (net == 2 ? 4 : 7)), nout = net == 0 ? 1 : (net == 3 ? 2 : 3);
            float o[3] = {0.f, 0.f, 0.f};
            uint32_t v[32];
            tmem_ld32(lane_addr + (uint32_t)(256 + (j & 1) * 128 + cq * 32), v);
            const float *bb = b2s + net * 128 + cq * 32, *ww = w3s + first * 128 + cq * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float hv = tanh_approx(__uint_as_float(v[i]) + bb[i]);
                o[0] = fmaf(hv, ww[i], o[0]);
                if (nout > 1) o[1] = fmaf(hv, ww[128 + i], o[1]);
                if (nout > 2) o[2] = fmaf(hv, ww[256 + i], o[2]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(y_free + (j & 1));   // (this thread's part of Y[b] is in registers)
            float *pb = part + (j & 1) * (3 * 384);
            const int recv = net <= 1 ? 0 : net - 1;
            if (cq != recv) {
                float *ps = pb + (cq < recv ? cq : cq - 1) * 384;
                ps[r] = o[0]; ps[128 + r] = o[1]; ps[256 + r] = o[2];
            }
            asm volatile("bar.sync 2, 512;\n" ::: "memory");
            if (cq == recv) {
#pragma unroll
                for (int q3 = 0; q3 < 3; ++q3)
                    if (q3 < nout) {
                        const float t = o[q3] + pb[q3 * 128 + r] + pb[384 + q3 * 128 + r] + pb[768 + q3 * 128 + r] + b3s[first + q3];
                        if (net == 0) lg[0] = t;
                        else if (net == 1) lg[1 + q3] = t;
                        else lg[q3] = t;
                    }
            }
        };
        // one head's draw from its logits (the Philox stream of sample_out9_kernel: head k uses u[draw .. draw + d))
        auto draw_head = [&](const PolicyIO &io, int64_t row, int head, const float *l, int d, const uint32_t *u, int draw) {
            float mx = -INFINITY;
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3)
                if (q3 < d) mx = fmaxf(mx, l[q3]);
            float se = 0.f;
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3)
                if (q3 < d) se += __expf(l[q3] - mx);
            const float lse = mx + __logf(se);   // (hardware ex2 / lg2: ~1e-7 absolute on the log-probabilities)
            int best = 0;
            float best_s = -INFINITY, best_lp = 0.f;
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3) {
                if (q3 < d) {
                    const float lp = l[q3] - lse;
                    const float un = ((float)u[(draw + q3) & 7] + 0.5f) * 2.3283064365386963e-10f;
                    const float gm = -__logf(fmaxf(-__logf(fminf(un, 0.99999994f)), 1e-20f));
                    if (lp + gm > best_s) { best_s = lp + gm; best = q3; best_lp = lp; }
                }
            }
            io.actions[row * io.act_stride + head] = best;
            io.logprobs[row * io.lp_stride + head] = best_lp;
        };
        auto finish = [&](const PolicyIO &io, int64_t row) {
            if (cq == 3) return;
            uint32_t u[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int bk = 0; bk < 2; ++bk) {
                if ((cq == 0 && bk == 1) || (cq == 2 && bk == 0)) continue;   // head 0: block 0, head 1: both, head 2: block 1
                uint32_t c0 = (uint32_t)row, c1 = (uint32_t)((uint64_t)row >> 32) | ((uint32_t)bk << 31), cc = (uint32_t)ctr,
                         c3 = (uint32_t)(ctr >> 32) ^ (io.stream_id * 0x9E3779B9u);
                at::philox4x32_10(c0, c1, cc, c3, (uint32_t)(P.seed & 0xFFFFFFFFu), (uint32_t)(P.seed >> 32));
                u[4 * bk] = c0; u[4 * bk + 1] = c1; u[4 * bk + 2] = cc; u[4 * bk + 3] = c3;
            }
            if (cq == 0) {
                io.values[row * io.val_stride] = lg[0];
                if (io.out9)
                    for (int q3 = 0; q3 < 4; ++q3) io.out9[row * 9 + q3] = lg[q3];
                draw_head(io, row, 0, lg + 1, 3, u, 0);
            } else if (cq == 1) {
                if (io.out9)
                    for (int q3 = 0; q3 < 3; ++q3) io.out9[row * 9 + 4 + q3] = lg[q3];
                draw_head(io, row, 1, lg, 3, u, 3);
            } else {
                if (io.out9) { io.out9[row * 9 + 7] = lg[0]; io.out9[row * 9 + 8] = lg[1]; }
                draw_head(io, row, 2, lg, 2, u, 6);
            }
        };
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int p = tile / P.tiles_per_policy;
            const int64_t row = (int64_t)(tile - p * P.tiles_per_policy) * TM + r;
            const bool in = row < P.n_rows;
            const PolicyIO &io = P.io[p];
            if (p != cur_p) {
                asm volatile("bar.sync 1, 512;\n" ::: "memory");
                const int t = threadIdx.x - 64;
                for (int i = t; i < 512; i += 512) { b1s[i] = __ldg(io.b1 + i); b2s[i] = __ldg(io.b2 + i); }
                for (int i = t; i < 9 * 128; i += 512) w3s[i] = __ldg(io.w3 + i);
                if (t < 12) b3s[t] = t < 9 ? __ldg(io.b3 + t) : 0.f;
                asm volatile("bar.sync 1, 512;\n" ::: "memory");
                cur_p = p;
            }
            const int j0 = 4 * it;
            // conv(j) writes h1 buffer j & 1, free once readout(j - 2) has seen the second layer of j - 2 finish
            for (int h = 0; h < 2; ++h) {
                if (threadIdx.x == 64) stamp(1, 300 + h);
                mbar_wait(acc1_full, (uint32_t)h);   // unit 2 it + h
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (threadIdx.x == 64) stamp(1, 310 + h);
                conv(j0 + 2 * h, 2 * h);
                conv(j0 + 2 * h + 1, 2 * h + 1);
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                mbar_arrive(x_free);
                if (threadIdx.x == 64) stamp(1, 320 + h);
                readout(j0 + 2 * h, 2 * h);
                if (threadIdx.x == 64) stamp(1, 330 + h);
                readout(j0 + 2 * h + 1, 2 * h + 1);
                if (threadIdx.x == 64) stamp(1, 340 + h);
            }
            if (in) finish(io, row);
            if (threadIdx.x == 64) stamp(1, 350);
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && P.advance && P.counter) {   // the tick counter moves on once the whole grid has used it
        __threadfence();
        if (atomicAdd(reinterpret_cast<unsigned long long *>(P.counter + 1), 1ull) == (unsigned long long)(gridDim.x - 1)) {
            P.counter[1] = 0;
            P.counter[0] += 1;
        }
    }
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
}

// a [rows][cols] fp16 matrix with `row_stride_bytes` between rows, read in boxes of box_rows x 64 columns
inline bool make_map_2d_f16(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows}, strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)H_KC, box_rows}, estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace at_tc
