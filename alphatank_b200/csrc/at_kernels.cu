// at_kernels.cu - sm_100a kernels and the C ABI (include/alphatank_b200.h) of the batched alphaTank step.
//
// One tick = env_kernel (the simulation: state in, state / rewards / done out) + rows_kernel (the observation rows,
// rebuilt from the stored state), the second launched as a programmatic dependent of the first.
//   env_kernel   One thread per env.  Tanks / bot FSMs of the block's 64 envs live in shared memory
//                ([field][tank][thread], conflict-free); bullets stay in HBM as [slot][env] columns (every warp access
//                is a contiguous run).  Sequential per-env semantics (firing order, tank order) are exact by
//                construction; the smart bots' line-of-sight queries of a warp are pooled (at_sim.h).  Rewards / done
//                are written here; done envs are reset in place.
//   rows_kernel  Every warp assembles the R x D rows of a group of envs as an image in shared memory (static wall /
//                maze columns once per warp, dynamic columns by independent jobs, at_rows.h) and ships the image with
//                ONE bulk-async (TMA, SASS UBLKCP) copy.
// The step is HBM-bound integer / fp64 work: no tensor cores anywhere.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/alphatank_b200.h"
#include <cuda_fp16.h>
#include "at_host.h"
#include "at_sim.h"
#include "at_rows.h"
#include "at_policy_tc.cuh"

using namespace at;

// ------------------------------------------------------------------------------------------------ device
namespace {

struct SmemLayout {
    int n;
    size_t off_trig, off_tx, off_ty, off_trew, off_bf0, off_bf1, off_tpack, off_nearlo, off_nearhi, off_bpack, off_tshot,
        off_acts, off_tlive, off_udelta, off_losq, total;
};
__host__ __device__ inline SmemLayout smem_layout(int n, int n_bots, int act_rows) {
    const int nb = n_bots > 0 ? n_bots : 1;
    SmemLayout L;
    L.n = n;
    size_t o = 0;
    L.off_trig = o; o += 722 * 8;
    L.off_tx = o; o += (size_t)n * BS * 8;
    L.off_ty = o; o += (size_t)n * BS * 8;
    L.off_trew = o; o += (size_t)n * BS * 8;
    L.off_bf0 = o; o += (size_t)nb * BS * 8;
    L.off_bf1 = o; o += (size_t)nb * BS * 8;
    L.off_tpack = o; o += (size_t)n * BS * 4;
    L.off_nearlo = o; o += (size_t)n * BS * 4;
    L.off_nearhi = o; o += n * SLOTS_PER_TANK > 32 ? (size_t)n * BS * 4 : 0;
    L.off_bpack = o; o += (size_t)nb * BS * 4;
    L.off_tshot = o; o += (size_t)n * BS * 4;
    L.off_acts = o; o += (size_t)BS * act_rows * 3 * 4;      // staging area of the action rows
    L.off_tlive = o; o += (size_t)n * BS;
    L.off_udelta = o; o += 368;
    L.off_losq = o; o += (size_t)(BS / 32) * 512;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}
__device__ inline void bind_block_mem(BlockMem &bm, unsigned char *smem, const SmemLayout &L) {
    bm.trig = (double *)(smem + L.off_trig);
    bm.tx = (double *)(smem + L.off_tx); bm.ty = (double *)(smem + L.off_ty);
    bm.trew = (double *)(smem + L.off_trew);
    bm.bf0 = (double *)(smem + L.off_bf0); bm.bf1 = (double *)(smem + L.off_bf1);
    bm.tpack = (uint32_t *)(smem + L.off_tpack); bm.nearlo = (uint32_t *)(smem + L.off_nearlo);
    bm.nearhi = (uint32_t *)(smem + L.off_nearhi);
    bm.bpack = (uint32_t *)(smem + L.off_bpack); bm.tlive = (uint8_t *)(smem + L.off_tlive);
    bm.tshot = (int32_t *)(smem + L.off_tshot);
    bm.udelta = (uint8_t *)(smem + L.off_udelta);
    bm.los_q = (uint8_t *)(smem + L.off_losq);
    bm.los_r = bm.los_q + (BS / 32) * 256;
}

#ifndef OBS_EVICT_FIRST
#define OBS_EVICT_FIRST 1
#endif
// Observation rows are write-once streaming data (277 MB per tick at 65536 envs, more than the whole L2): they are
// stored with an L2 evict-first policy so that they do not push the 53 MB of env state out of the cache.
__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_store(float *gdst, const float *ssrc, uint32_t bytes, uint64_t pol) {  // TMA 1-D bulk copy
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
#if OBS_EVICT_FIRST
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n" ::"l"(gdst), "r"(s),
                 "r"(bytes), "l"(pol)
                 : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void stg_stream(float *p, float v, uint64_t pol) {
#if OBS_EVICT_FIRST
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;\n" ::"l"(p), "f"(v), "l"(pol) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// at_step_host: how the host learns that rewards / done have landed in its memory without synchronising the stream
// (the rows kernel is still running then): the last block to finish posts `epoch` into a mapped host word.
struct HostSignal {
    unsigned int *counter;          // device: blocks that have finished (reset by the last one)
    volatile uint32_t *host_flag;   // mapped host memory; nullptr: no signalling
    uint32_t epoch;
    unsigned long long *dbg_times;  // diagnostics (at_debug_block_times): [2 * grid] globaltimer at block start / end
    int epw;                        // envs per warp (1 .. 32, lanes 0 .. epw - 1 own one): see sim_envs_per_warp
};

// at_step_host: the last block to finish posts `epoch` into a mapped host word (block-uniform call)
__device__ __forceinline__ void signal_host(const HostSignal &sig, int tid) {
    if (sig.host_flag == nullptr) return;
    __threadfence_system();   // rewards / done of this block are in host memory before it counts
    __syncthreads();
    if (tid == 0) {
        const unsigned int prev = atomicAdd(sig.counter, 1u);
        if (prev == gridDim.x - 1) {
            *sig.counter = 0u;
            __threadfence_system();
            *sig.host_flag = sig.epoch;
        }
    }
}

// The simulation of one tick (IS_STEP) or a masked reset.  The observation rows are built from the stored state by
// rows_kernel, launched behind this kernel as a programmatic dependent.
template <bool IS_STEP, int SPEC = 0>
__global__ void __launch_bounds__(BS, 7) env_kernel(const __grid_constant__ KCfg c, const DevState st,
                                                 const int32_t *__restrict__ actions,
                                                 const uint8_t *__restrict__ mask,
                                                 float *__restrict__ rewards, uint8_t *__restrict__ done, HostSignal sig) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout L = smem_layout(c.n_tanks, c.n_bots, c.act_rows);
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");  // rows_kernel may take freed SM slots
    if (sig.dbg_times != nullptr && threadIdx.x == 0) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t0));
        sig.dbg_times[2 * blockIdx.x] = t0;
    }
    BlockMem bm;
    bind_block_mem(bm, smem, L);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {   // table copies: every load is issued before the first store (one round trip instead of a dozen)
        uint8_t *udl = (uint8_t *)(smem + L.off_udelta);
        constexpr int TRIG_IT = (722 + BS - 1) / BS, UDL_IT = (361 + BS - 1) / BS;
        double tv[TRIG_IT];
        uint8_t uv[UDL_IT];
#pragma unroll
        for (int k = 0; k < TRIG_IT; ++k) tv[k] = tid + k * BS < 722 ? __ldg(st.trig + tid + k * BS) : 0.0;
#pragma unroll
        for (int k = 0; k < UDL_IT; ++k) uv[k] = tid + k * BS < 361 ? __ldg(st.udelta + tid + k * BS) : (uint8_t)0;
#pragma unroll
        for (int k = 0; k < TRIG_IT; ++k) if (tid + k * BS < 722) bm.trig[tid + k * BS] = tv[k];
#pragma unroll
        for (int k = 0; k < UDL_IT; ++k) if (tid + k * BS < 361) udl[tid + k * BS] = uv[k];
    }
    __syncthreads();

    // Small batches leave most SMs' warp slots empty, and a warp's tick lasts as long as the UNION of its lanes' paths:
    // with fewer envs per warp (the other lanes idle) the same envs spread over more warps with shorter paths.
    const int64_t e_warp0 = ((int64_t)blockIdx.x * (BS / 32) + warp) * sig.epw;
    const int64_t e = e_warp0 + lane;
    const bool valid = lane < sig.epw && e < st.N;
    const unsigned valid_mask = __ballot_sync(0xffffffffu, valid);

    if (IS_STEP) {
        // This env's action rows are staged in shared memory by asynchronous copies: the first read comes after the
        // first bullet pass, so the copy latency is off the critical path.  Each thread copies and later reads only
        // its own rows.
        const int32_t *my_actions = nullptr;
        if (actions && valid_mask) {
            const int row_ints = c.act_rows * 3;
            int32_t *wstage = (int32_t *)(smem + L.off_acts) + (size_t)warp * 32 * row_ints;
            const int32_t *wsrc = actions + e_warp0 * row_ints;
            if (valid)
                for (int q = 0; q < row_ints; ++q)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((uint32_t)__cvta_generic_to_shared(wstage + lane * row_ints + q)),
                                 "l"(wsrc + lane * row_ints + q)
                                 : "memory");
            my_actions = wstage + lane * row_ints;
        }
        if (valid) step_lane<DevWarp, SPEC>(c, st, bm, e_warp0, tid, valid_mask, my_actions, rewards, done);
    } else if (valid && (!mask || mask[e])) {
        reset_lane(c, st, bm, e, tid);
    }

    if (sig.dbg_times != nullptr && tid == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t1));
        sig.dbg_times[2 * blockIdx.x + 1] = t1;
    }
    signal_host(sig, tid);
}

// The trainers' per-tick observation normaliser on ONE column of an env's [rows][dim] block (train_multi_ppo.py:89-91
// with models/ppo_utils.py:13-29: a FRESH RunningMeanStd(shape=(A, D), epsilon) is updated with the block and applied
// to it):  bm = mean_a x, bv = mean_a (x - bm)^2, total = eps + A,
//   mean = bm A / total,  var = 1 + (bv A + bm^2 eps A / total) / total,  y = (x - mean) / (sqrt(var) + 1e-8).
// Shared by tick_normalize_kernel and the normalised-rows mode of rows_kernel (same float operations, same order).
// Float arithmetic: bm and mean with IEEE divisions in the reference's order (x - mean cancels to ~1e-4 x on
// static columns, so mean must be as accurate as torch's); the variance path folds bv A = ss and uses a reciprocal
// and rsqrt (var >= 1, and the 1e-8 is absorbed: sqrt(var) + 1e-8 == sqrt(var) in fp32) - a 2e-7 relative change of y.
struct TickNormK {   // per-launch constants
    float fa, total, k_var, inv_total;
    __device__ __forceinline__ TickNormK(int rows, float eps) {
        fa = (float)rows; total = eps + fa; k_var = eps * fa / total; inv_total = 1.0f / total;
    }
};
template <int RMAX>
__device__ __forceinline__ void tick_norm_column(const float (&v)[RMAX], int rows, const TickNormK &k, float &mean, float &inv) {
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < RMAX; ++a)
        if (a < rows) sum += v[a];
    const float bm = sum / k.fa;
    float ss = 0.f;
#pragma unroll
    for (int a = 0; a < RMAX; ++a)
        if (a < rows) { const float d = v[a] - bm; ss += d * d; }
    // bm * fa / total, correctly rounded without the division: total is a launch constant, inv_total = RN(1 / total), and
    // q = RN(n inv_total), q' = RN(q + RN(n - q total) inv_total) is RN(n / total) (Markstein; checked against IEEE division on
    // 28 M samples for A = 1 .. 8 rows) - the same bits as the reference's operation order, a third of the instructions
    const float nn = bm * k.fa, q0 = nn * k.inv_total;
    mean = fmaf(fmaf(-q0, k.total, nn), k.inv_total, q0);
    const float var = 1.0f + (ss + bm * bm * k.k_var) * k.inv_total;
    inv = rsqrtf(var);
}

// ------------------------------------------------------------------------------------------------ rows_kernel
// Observation rows from the stored state (split path, the default).  Every WARP works on its own: it loops over
// groups of G consecutive envs (4 for the 2v2 layout), whose R x D rows are ONE contiguous run of the observation
// buffer, assembles them as an image in its own slice of shared memory - the static wall / maze columns are written
// once per warp, the ~45 % dynamic columns are scattered by independent jobs (at_rows.h), 32 jobs per pass, every
// pass of one kind - and hands the whole image (17 KB) to the bulk-async engine: one TMA copy per group, with an L2
// evict-first hint, no per-column stores, no transposition, and no block-wide barrier anywhere (warps only meet
// their own copy).  The next groups' state columns are requested two / one iterations ahead (bullet counts, then the
// live records), and 12 warps per SM keep the copy engine busy while each waits for its image to drain, so the
// loop is bounded by the HBM write stream.
// Launched as a programmatic dependent of env_kernel<.., false>: the prologue (tables, static columns) overlaps
// the simulation's tail; griddepcontrol.wait orders everything else behind the simulation's stores.
#ifndef AT_ROWS_SINGLE_BODY
#define AT_ROWS_SINGLE_BODY 1
#endif
constexpr int RK_MAX_WARPS = 16;   // warps per block: a launch parameter (blockDim.x / 32)
struct RowsSmem {   // per warp, after the block's trig table
    size_t off_img, off_nimg, off_rx, off_ry, off_rm, off_tx, off_ty, off_wlo, off_whi, off_tp, off_zn, off_emit, off_occ, off_jl, per_warp, total;
};
__host__ __device__ inline RowsSmem rows_smem_layout(int G, int n, int row_floats, int warps, int nimg_elem /* 0 | 4 | 2 bytes */) {
    RowsSmem L;
    size_t o = 0;
    const size_t S = (size_t)n * SLOTS_PER_TANK;
    L.off_img = o; o += ((size_t)G * row_floats * 4 + 127) & ~(size_t)127;
    L.off_nimg = o; o += ((size_t)G * row_floats * nimg_elem + 127) & ~(size_t)127;   // the normalised copy of the image (fp32 / fp16)
    L.off_rx = o; o += (size_t)G * S * 8;
    L.off_ry = o; o += (size_t)G * S * 8;
    L.off_rm = o; o += (size_t)G * S * 8;
    L.off_tx = o; o += (size_t)G * n * 8;
    L.off_ty = o; o += (size_t)G * n * 8;
    L.off_wlo = o; o += (size_t)G * 8;
    L.off_whi = o; o += (size_t)G * 8;
    L.off_tp = o; o += (size_t)G * n * 4;
    L.off_zn = o; o += (size_t)G * 4;
    L.off_emit = o; o += (size_t)G * 4;
    L.off_occ = o; o += (size_t)G * S;     // which bullet positions of the image hold a bullet
    o = (o + 1) & ~(size_t)1;
    L.off_jl = o; o += (size_t)G * S * 2;  // the group's position jobs, compacted
    L.per_warp = (o + 127) & ~(size_t)127;
    L.total = 722 * 8 + 64 + L.per_warp * warps;   // (722 * 8 + 64 is a multiple of 128)
    return L;
}

// NT / NR / TEAM: compile-time shape (0 / 0 / -1: from the config); MAZE: per-env maps; PP: (env, slot) pairs a lane
// prefetches per group (G * 6n <= 32 PP)
// NORMK: 0 = raw rows only, 1 = can also emit tick-normalised fp32 rows, 2 = tick-normalised fp16 rows (+ raw): every
// build carries only its own code (the raw-rows kernel lost 15 % when the normalisers shared its body).
template <int NT, int NR, int TEAM, bool MAZE, int PP, int NORMK>
__global__ void __launch_bounds__(RK_MAX_WARPS * 32) rows_kernel(const __grid_constant__ KCfg c, const DevState st,
                                                          const uint8_t *__restrict__ mask, float *__restrict__ obs,
                                                          const float *__restrict__ g_tmpl, int lgG, int n_groups,
                                                          float *__restrict__ obs_norm, float norm_eps, int norm_half) {
    extern __shared__ __align__(128) unsigned char smem[];
    typedef RowsShape<NT, NR, TEAM> SH;
    // obs_norm: also emit (norm2: a second, normalised image) or only emit (inplace: obs == nullptr; the image is built
    // in full every group and normalised in place) the tick-normalised rows
    // norm_half bit 0: the normalised rows leave as fp16 (always through a second image: the raw one stays incremental);
    // bit 1: persistent output buffers - the columns that never change on a fixed map are not written again
    const bool norm = NORMK != 0 && obs_norm != nullptr, half_out = NORMK == 2 && norm, norm2 = norm && (obs != nullptr || half_out),
               inplace = norm && !norm2;
    const TickNormK nk(SH::rows(c), norm_eps);
    const int G = 1 << lgG, n = SH::n(c), S = n * SLOTS_PER_TANK, nrows = SH::rows(c), RD = nrows * c.obs_dim;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const RowsSmem L = rows_smem_layout(G, n, RD, nwarps, norm2 ? (half_out ? 2 : 4) : 0);
    double *trig = (double *)smem;
    unsigned char *ws = smem + 722 * 8 + 64 + (size_t)warp * L.per_warp;
    float *img = (float *)(ws + L.off_img), *nimg = norm2 ? (float *)(ws + L.off_nimg) : img;
    __half *nimg16 = (__half *)(ws + L.off_nimg);
    double *rx = (double *)(ws + L.off_rx), *ry = (double *)(ws + L.off_ry);
    uint64_t *rm = (uint64_t *)(ws + L.off_rm);
    double *tx = (double *)(ws + L.off_tx), *ty = (double *)(ws + L.off_ty);
    uint32_t *tp = (uint32_t *)(ws + L.off_tp), *zn = (uint32_t *)(ws + L.off_zn);
    uint64_t *wl = (uint64_t *)(ws + L.off_wlo), *wh = (uint64_t *)(ws + L.off_whi);
    int *em_s = (int *)(ws + L.off_emit);
    uint8_t *occ = (uint8_t *)(ws + L.off_occ);
    uint16_t *jl = (uint16_t *)(ws + L.off_jl);
    const int N = (int)st.N;   // (at_create guarantees 6 n N < 2^31: 32-bit column indices)

    // ---- prologue: nothing here depends on the simulation kernel's stores
    for (int i = tid; i < 722; i += (int)blockDim.x) trig[i] = __ldg(st.trig + i);
    for (int p = lane; p < G * S; p += 32) { rm[p] = ROWS_EMPTY; occ[p] = 0; }
    for (int i = lane; i < G * RD / 4; i += 32) reinterpret_cast<float4 *>(img)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = ((G * RD) & ~3) + lane; i < G * RD; i += 32) img[i] = 0.f;
    if (half_out)   // the fp16 image starts out as zeros too (the static columns are filled in below)
        for (int i = lane; i < (G * RD + 1) / 2; i += 32) reinterpret_cast<uint32_t *>(nimg16)[i] = 0u;
    __syncwarp();
    constexpr int RMAX = NR ? NR : MAX_TANKS;
    if (!MAZE) {
        const int sl = 4 * c.n_walls + CELLS;
        for (int q = lane; q < sl; q += 32) {
            const float v = __ldg(g_tmpl + q);
            float vn = 0.f;
            if (norm) {   // a static column holds the same value in every row: its normalised value is static too
                float vv[RMAX], mean, inv;
#pragma unroll
                for (int a = 0; a < RMAX; ++a) vv[a] = v;
                tick_norm_column<RMAX>(vv, nrows, nk, mean, inv);
                vn = (v - mean) * inv;
            }
            for (int gr = 0; gr < G * nrows; ++gr) {   // gr = env * rows + row
                if (!inplace) img[gr * c.obs_dim + c.off_walls + q] = v;
                if (half_out) nimg16[gr * c.obs_dim + c.off_walls + q] = __float2half_rn(vn);
                else if (norm) nimg[gr * c.obs_dim + c.off_walls + q] = vn;
            }
        }
    }
    // columns the normaliser visits per group: [0, off_walls) and [dyn2, D) (everything for per-env maps)
    const int dyn2 = MAZE ? c.off_walls : c.off_zones, ndyn = c.off_walls + (c.obs_dim - dyn2);
    const bool bulk_ok = (((size_t)RD * (half_out ? 2 : 4)) % 16 == 0) && ((((uintptr_t)obs) & 15u) == 0) && ((((uintptr_t)obs_norm) & 15u) == 0);
    const uint64_t pol = make_evict_first_policy();
    // persistent-buffer mode: the dynamic runs of a row, widened to 16-byte multiples for fp32 (4 floats) and fp16 (8 halves)
    const int cA4 = (c.off_walls + 3) & ~3, cB4 = dyn2 & ~3, cA8 = (c.off_walls + 7) & ~7, cB8 = dyn2 & ~7;
    const bool need4 = obs != nullptr || (norm && !half_out), need8 = half_out;   // fp32 / fp16 images to ship
    const bool skip_static = (norm_half & 2) != 0 && !MAZE && (!need4 || ((c.obs_dim & 3) == 0 && cA4 < cB4)) &&
                             (!need8 || ((c.obs_dim & 7) == 0 && cA8 < cB8));
    __syncthreads();   // the trig table (the only block-wide barrier)
    asm volatile("griddepcontrol.wait;\n" ::: "memory");

    // ---- prefetch registers.  Stage A: bullet count + emit flag of the lane's (env, slot) pairs (two groups
    // ahead, two register sets used alternately - a register move of an in-flight load would wait for it); stage B:
    // the live records of those pairs, the lane's (env, tank) pair and env scalars (one group ahead).
    uint32_t x_cnt[PP], x_em[PP], y_cnt[PP], y_em[PP];
    uint64_t r_m[PP];
    double r_x[PP], r_y[PP];
    double t_x = 0.0, t_y = 0.0;
    uint32_t t_p = 0, e_zn = 0, e_em = 0;
    uint64_t e_wl = 0, e_wh = 0;
    const int my_g = lane & (G - 1);   // (32 is a multiple of G: a lane's pairs, tank pair and env scalars share the env)
    auto load_a = [&](int grp, uint32_t *cn, uint32_t *em) {
        const int e = grp * G + my_g;
        const bool in = grp < n_groups && e < N;
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            cn[pp] = 0; em[pp] = 0;
            if (in && lane + pp * 32 < G * S) {
                cn[pp] = st.cnt[e];
                em[pp] = mask ? (uint32_t)mask[e] : 1u;
            }
        }
    };
    auto load_b = [&](int grp, const uint32_t *cn, const uint32_t *em) {
        const int e = grp * G + my_g;
        const bool in = grp < n_groups && e < N;
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            r_m[pp] = ROWS_EMPTY; r_x[pp] = 0.0; r_y[pp] = 0.0;
            const uint32_t s = (uint32_t)((lane + pp * 32) >> lgG);
            if (em[pp] && s < cn[pp]) {   // (cn == 0 outside the batch)
                const int g = (int)s * N + e;
                r_m[pp] = st.bmeta[g]; r_x[pp] = st.bx[g]; r_y[pp] = st.by[g];
            }
        }
        t_x = 0.0; t_y = 0.0; t_p = 0;
        if (in && lane < G * n) {
            const int g = (lane >> lgG) * N + e;
            t_x = st.tx[g]; t_y = st.ty[g]; t_p = st.tpack[g];
        }
        e_zn = 0; e_em = 0; e_wl = 0; e_wh = 0;
        if (in && lane >= 32 - G) {   // env scalars: the warp's last G lanes (lane 32 - G + g has my_g == g)
            e_zn = st.zones[e];
            e_em = mask ? (uint32_t)mask[e] : 1u;
            if (MAZE) { e_wl = st.wall_lo[e]; e_wh = st.wall_hi[e]; }
        }
    };
    const int gs = gridDim.x * nwarps, g0 = blockIdx.x * nwarps + warp;
    const int jobs_pos = G * S, jobs_tank = G * nrows * n, jobs = jobs_pos + jobs_tank + (MAZE ? G * CELLS : 0);

    uint32_t nz_prev = 0xFFFFFFFFu;   // batches of the normalised image that may hold non-zero values (warp-uniform)
    // one group: `use` = counts of group grp + 1 (loaded one iteration ago), `fill` = receives those of grp + 2
    auto group = [&](int grp, const uint32_t *use_cnt, const uint32_t *use_em, uint32_t *fill_cnt, uint32_t *fill_em) {
        // ---- phase 1: this group's prefetched columns -> the warp's tables
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            const uint64_t m = r_m[pp];
            if (m != ROWS_EMPTY) {
                const int q = ((int)(m & 15u) * SLOTS_PER_TANK + (int)((m >> BM_RANK) & 7u)) * G + my_g;
                rm[q] = m; rx[q] = r_x[pp]; ry[q] = r_y[pp];
            }
        }
        if (lane < G * n) { tx[lane] = t_x; ty[lane] = t_y; tp[lane] = t_p; }   // [tank][env]
        bool my_em = true;
        if (lane >= 32 - G) {
            zn[my_g] = e_zn; em_s[my_g] = (int)e_em;
            if (MAZE) { wl[my_g] = e_wl; wh[my_g] = e_wh; }
            my_em = e_em != 0;
        }
        // ---- prefetch: records of the next group, counts of the one after
        load_b(grp + gs, use_cnt, use_em);
        load_a(grp + 2 * gs, fill_cnt, fill_em);
        // the previous copy must be done reading the image
        if (skip_static || lane == 0) bulk_wait_read_all();   // (persistent-buffer mode: several lanes issue copies)
        const bool all_emit = __all_sync(0xffffffffu, my_em);   // (also the warp barrier between phase 1 and 2)
        // ---- phase 2: 32 jobs per pass.  Position jobs first: a position that holds no bullet now and held none when
        // the image was last filled needs no work (it is still zero) - about 60 % of them - so the positions that do are
        // compacted into a list (ballot + prefix) and the lanes work through that list densely instead of visiting every
        // position with two lanes in three idle.  (An image that is normalised in place is rebuilt in full.)
        int npos = jobs_pos;
        if (!inplace) {
            npos = 0;
            for (int j0 = 0; j0 < jobs_pos; j0 += 32) {
                const int j = j0 + lane;
                bool need = false;
                if (j < jobs_pos) {
                    const int jp = j >> lgG;
                    const int k = jp / n, o = jp - k * n;
                    const int q = (o * SLOTS_PER_TANK + k) * G + (j & (G - 1));
                    const bool has = rm[q] != ROWS_EMPTY;
                    need = has || occ[q];
                    occ[q] = has ? 1 : 0;
                }
                const unsigned nb = __ballot_sync(0xffffffffu, need);
                if (need) jl[npos + __popc(nb & ((1u << lane) - 1u))] = (uint16_t)j;
                npos += __popc(nb);
            }
            __syncwarp();
        }
        for (int i = lane; i < npos; i += 32) {
            const int j = inplace ? i : (int)jl[i];
            const int g = j & (G - 1);
            const int jp = j >> lgG;
            const int k = jp / n, o = jp - k * n;
            const int q = (o * SLOTS_PER_TANK + k) * G + g;
            const uint64_t m = rm[q];
            rows_position_job<NT, NR, TEAM>(c, trig, o, k, m, rx[q], ry[q], tx + g, ty + g, G, img + g * RD, g + 4 * (o >> 1));
            if (m != ROWS_EMPTY) rm[q] = ROWS_EMPTY;   // the table is empty again for the next group
        }
        for (int j = jobs_pos + lane; j < jobs; j += 32) {
            const int g = j & (G - 1);
            if (j < jobs_pos + jobs_tank) {   // the rows' own blocks first, then the other-tank blocks
                const int j2 = (j - jobs_pos) >> lgG;
                int r, i;
                if (j2 < nrows) {
                    r = j2;
                    i = SH::team(c) ? c.agent_tank[r] : r;
                } else {
                    const int ro = j2 - nrows;
                    r = ro / (n - 1);
                    const int oi = ro - r * (n - 1), t = SH::team(c) ? c.agent_tank[r] : r;
                    i = oi < t ? oi : oi + 1;
                }
                rows_tank_job<NT, NR, TEAM>(c, trig, st.corn, r, i, tx + g, ty + g, tp + g, G, zn[g], img + g * RD);
            } else {
                rows_maze_job(c, wl[g], wh[g], (j - jobs_pos - jobs_tank) >> lgG, img + g * RD);
            }
        }
        if (norm && norm2 && (c.obs_dim & 1) == 0) {
            // the tick normaliser over the dynamic columns, raw image -> second image, two adjacent columns per lane
            // (64-bit loads, one half2 / float2 store per row): slots = pairs of [0, off_walls) and [dyn2, D), then the
            // <= 2 single columns an odd boundary leaves over
            __syncwarp();
            const int np1 = c.off_walls >> 1, d2 = (dyn2 + 1) & ~1, np2 = (c.obs_dim - d2) >> 1;
            const int ns1 = c.off_walls & 1, n_slots = np1 + np2 + ns1 + (dyn2 & 1);
            uint32_t nz_now = 0;   // bit (g * 16 + pass): the pass's 32 slots hold a non-zero value in this group
            for (int g = 0; g < G; ++g)
                for (int s0 = 0, jb = 0; s0 < n_slots; s0 += 32, ++jb) {
                    const int sl = s0 + lane;
                    const bool in = sl < n_slots, pair = sl < np1 + np2;
                    int col = 2 * sl;
                    if (sl >= np1) col = d2 + 2 * (sl - np1);
                    if (!pair) col = (sl - np1 - np2 == 0 && ns1) ? c.off_walls - 1 : dyn2;
                    const float *src = img + g * RD + col;
                    float vx[RMAX], vy[RMAX];
                    bool zero = true;
#pragma unroll
                    for (int a = 0; a < RMAX; ++a) {
                        vx[a] = 0.f; vy[a] = 0.f;
                        if (in && a < nrows) {
                            if (pair) { const float2 t = *reinterpret_cast<const float2 *>(src + a * c.obs_dim); vx[a] = t.x; vy[a] = t.y; }
                            else vx[a] = src[a * c.obs_dim];
                        }
                        zero = zero && vx[a] == 0.f && vy[a] == 0.f;
                    }
                    const uint32_t bit = (G <= 2 && jb < 16) ? 1u << (g * 16 + jb) : 0u;
                    const bool all_zero = __all_sync(0xffffffffu, zero);
                    if (all_zero && bit && !(nz_prev & bit)) continue;   // zeros, and the last group left zeros there
                    if (!all_zero) nz_now |= bit;
                    float mx = 0.f, ix = 0.f, my = 0.f, iy = 0.f;
                    if (!all_zero) {
                        tick_norm_column<RMAX>(vx, nrows, nk, mx, ix);
                        tick_norm_column<RMAX>(vy, nrows, nk, my, iy);
                    }
                    if (in) {
#pragma unroll
                        for (int a = 0; a < RMAX; ++a)
                            if (a < nrows) {
                                const float yx = all_zero ? 0.f : (vx[a] - mx) * ix, yy = all_zero ? 0.f : (vy[a] - my) * iy;
                                const int o = g * RD + a * c.obs_dim + col;
                                if (half_out) {
                                    if (pair) *reinterpret_cast<__half2 *>(nimg16 + o) = __floats2half2_rn(yx, yy);
                                    else nimg16[o] = __float2half_rn(yx);
                                } else {
                                    if (pair) *reinterpret_cast<float2 *>(nimg + o) = make_float2(yx, yy);
                                    else nimg[o] = yx;
                                }
                            }
                    }
                }
            nz_prev = (G <= 2) ? nz_now : 0xFFFFFFFFu;
        } else if (norm) {   // one column per lane (odd row lengths, and the image that is normalised in place)
            __syncwarp();
            for (int g = 0; g < G; ++g)
                for (int j0 = 0; j0 < ndyn; j0 += 32) {
                    const int j = j0 + lane;
                    const bool in = j < ndyn;
                    const int col = j < c.off_walls ? j : j - c.off_walls + dyn2;
                    const float *src = img + g * RD + col;
                    float *dn = nimg + g * RD + col;
                    __half *dh = nimg16 + g * RD + col;
                    float vv[RMAX];
                    bool zero = true;
#pragma unroll
                    for (int a = 0; a < RMAX; ++a) {
                        vv[a] = in && a < nrows ? src[a * c.obs_dim] : 0.f;
                        zero = zero && vv[a] == 0.f;
                    }
                    if (__all_sync(0xffffffffu, zero)) {   // 32 all-zero columns (empty bullet positions) normalise to zero
                        if (norm2 && in) {
                            if (half_out) for (int a = 0; a < nrows; ++a) dh[a * c.obs_dim] = __float2half_rn(0.f);
                            else for (int a = 0; a < nrows; ++a) dn[a * c.obs_dim] = 0.f;
                        }
                        continue;
                    }
                    float mean, inv;
                    tick_norm_column<RMAX>(vv, nrows, nk, mean, inv);
                    if (in) {
#pragma unroll
                        for (int a = 0; a < RMAX; ++a)
                            if (a < nrows) {
                                const float y = (vv[a] - mean) * inv;
                                if (half_out) dh[a * c.obs_dim] = __float2half_rn(y);
                                else dn[a * c.obs_dim] = y;
                            }
                    }
                }
        }
        fence_async_smem();   // the images are read by the async proxy
        __syncwarp();
        // ---- phase 3: ship the image(s)
        float *dst = obs + (int64_t)grp * G * RD, *dstn = obs_norm + (int64_t)grp * G * RD;
        __half *dsth = reinterpret_cast<__half *>(obs_norm) + (int64_t)grp * G * RD;
        if (all_emit && bulk_ok && skip_static) {
            // persistent buffers: per row the two dynamic runs [0, cA) and [cB, D) only (16-byte multiples; the few static
            // columns inside them are rewritten with their own values), one bulk copy per lane and run
            const int n_pieces = 2 * G * nrows;
            bool issued = false;
            for (int pc = lane; pc < n_pieces; pc += 32) {
                const int gr = pc >> 1, second = pc & 1;
                if (obs) {
                    const int off = gr * c.obs_dim + (second ? cB4 : 0), len = second ? c.obs_dim - cB4 : cA4;
                    bulk_store(dst + off, img + off, (uint32_t)(len * 4), pol);
                }
                if (half_out) {
                    const int off = gr * c.obs_dim + (second ? cB8 : 0), len = second ? c.obs_dim - cB8 : cA8;
                    bulk_store(reinterpret_cast<float *>(dsth + off), reinterpret_cast<const float *>(nimg16 + off), (uint32_t)(len * 2), pol);
                } else if (norm) {
                    const int off = gr * c.obs_dim + (second ? cB4 : 0), len = second ? c.obs_dim - cB4 : cA4;
                    bulk_store(dstn + off, nimg + off, (uint32_t)(len * 4), pol);
                }
                issued = true;
            }
            if (issued) bulk_commit();
        } else if (all_emit && bulk_ok) {
            if (lane == 0) {
                if (obs) bulk_store(dst, img, (uint32_t)(G * RD * 4), pol);
                if (half_out) bulk_store(reinterpret_cast<float *>(dsth), reinterpret_cast<const float *>(nimg16), (uint32_t)(G * RD * 2), pol);
                else if (norm) bulk_store(dstn, nimg, (uint32_t)(G * RD * 4), pol);   // (nimg == img when in place)
                bulk_commit();
            }
        } else {   // masked reset / ragged tail / rows that are not 16-byte multiples: coalesced plain stores
            for (int g = 0; g < G; ++g) {
                if (!em_s[g]) continue;
                if (obs) for (int q = lane; q < RD; q += 32) stg_stream(dst + g * RD + q, img[g * RD + q], pol);
                if (half_out) for (int q = lane; q < RD; q += 32) dsth[g * RD + q] = nimg16[g * RD + q];
                else if (norm) for (int q = lane; q < RD; q += 32) stg_stream(dstn + g * RD + q, nimg[g * RD + q], pol);
            }
            __syncwarp();   // the tables (em_s) are rewritten by the next group's phase 1
        }
    };

    load_a(g0, x_cnt, x_em);
    load_a(g0 + gs, y_cnt, y_em);
    load_b(g0, x_cnt, x_em);
#if AT_ROWS_SINGLE_BODY
    // one copy of the group body: y = counts of the next group, x = being filled, handed over at the end of the
    // iteration (the moves wait for loads that were issued a whole group of work earlier)
#pragma unroll 1
    for (int grp = g0; grp < n_groups; grp += gs) {
        group(grp, y_cnt, y_em, x_cnt, x_em);
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) { y_cnt[pp] = x_cnt[pp]; y_em[pp] = x_em[pp]; }
    }
#else
    for (int grp = g0;;) {   // x / y alternate between "counts of the next group" and "being filled"
        if (grp >= n_groups) break;
        group(grp, y_cnt, y_em, x_cnt, x_em);
        grp += gs;
        if (grp >= n_groups) break;
        group(grp, x_cnt, x_em, y_cnt, y_em);
        grp += gs;
    }
#endif
    if (skip_static || lane == 0) bulk_wait_read_all();   // shared memory must outlive the copies' reads
}
typedef void (*rows_kernel_fn)(const KCfg, const DevState, const uint8_t *, float *, const float *, int, int, float *, float, int);

__global__ void gather_kernel(const __grid_constant__ KCfg c, const DevState st, int64_t first, int64_t count,
                              PackedEnv *out) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < count) gather_env(c, st, first + q, out[q]);
}
__global__ void scatter_kernel(const __grid_constant__ KCfg c, const DevState st, int64_t first, int64_t count,
                               const PackedEnv *in) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < count) scatter_env(c, st, first + q, in[q]);
}

// TERMINAL: envs whose episode ended in the latest step and was replaced by the auto-reset report the latched info of
// that episode (Sim::latch_terminal_info) - what the reference's step() returned before its caller reset.
template <bool TERMINAL>
__global__ void info_kernel(const __grid_constant__ KCfg c, const DevState st, int32_t *winner, int32_t *hits,
                            int32_t *be_hits, int32_t *change_time, uint8_t *latched) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    const bool lt = TERMINAL && st.term_tick[e] == st.tick[e];
    if (TERMINAL && latched) latched[e] = lt ? 1 : 0;
    int w = -1;
    for (int i = 0; i < c.n_tanks; ++i) {
        int64_t g = (int64_t)i * st.N + e;
        if (w < 0 && (st.tpack[g] & TP_ALIVE)) w = i;
        uint32_t h = lt ? st.term_hits[g] : st.thits[g];
        if (hits) hits[e * c.n_tanks + i] = (int32_t)(h & 0xFFFFu);
        if (be_hits) be_hits[e * c.n_tanks + i] = (int32_t)(h >> 16);
        if (change_time) change_time[e * c.n_tanks + i] = st.change_time[g];
    }
    if (winner) winner[e] = lt ? (int)st.term_winner[e] - 1 : w;
}

__global__ void prev_rew_kernel(const DevState st, int row, int tank) {   // at_set_reward_mode: start of the increments
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < st.N) st.prev_rew[(int64_t)row * st.N + e] = (float)st.trew[(int64_t)tank * st.N + e];
}
// per-tank columns for env.game_env.tanks[i].<attr> readers: out[c][e], c = alive, x, y, angle, reward, num_hit, num_be_hit
__global__ void tank_columns_kernel(const DevState st, int tank, double *__restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    const int64_t g = (int64_t)tank * st.N + e;
    const uint32_t pk = st.tpack[g], th = st.thits[g];
    out[e] = (pk & TP_ALIVE) ? 1.0 : 0.0;
    out[st.N + e] = st.tx[g];
    out[2 * st.N + e] = st.ty[g];
    out[3 * st.N + e] = (double)(pk & 511u);
    out[4 * st.N + e] = st.trew[g];
    out[5 * st.N + e] = (double)(th & 0xFFFFu);
    out[6 * st.N + e] = (double)(th >> 16);
}

__global__ void synthetic_actions_kernel(uint64_t seed, int64_t env_id_base, int64_t n_envs, int rows, uint32_t tick,
                                         int32_t *out) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_envs * rows) return;
    int64_t e = k / rows;
    int a = (int)(k % rows);
    int64_t env_id = env_id_base + e;
    uint32_t c0 = tick, c1 = 1u + (uint32_t)a, c2 = (uint32_t)((uint64_t)env_id & 0xFFFFFFFFu),
             c3 = (uint32_t)((uint64_t)env_id >> 32);
    philox4x32_10(c0, c1, c2, c3, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32));
    out[k * 3 + 0] = (int32_t)(((uint64_t)c0 * 3u) >> 32);
    out[k * 3 + 1] = (int32_t)(((uint64_t)c1 * 3u) >> 32);
    out[k * 3 + 2] = (int32_t)(((uint64_t)c2 * 2u) >> 32);
}

__global__ void bfs_batch_kernel(const uint8_t *maze, const int32_t *query, int64_t n, int32_t *out) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int32_t *q = query + 5 * k;
    const uint8_t *m = maze + (int64_t)q[0] * CELLS;
    uint64_t lo = 0, hi = 0;
    for (int i = 0; i < CELLS; ++i)
        if (m[i]) { if (i < 64) lo |= 1ull << i; else hi |= 1ull << (i - 64); }
    int nr, nc;
    int len = bfs_first_step(lo, hi, q[1], q[2], q[3], q[4], &nr, &nc);
    out[3 * k] = len; out[3 * k + 1] = len > 1 ? nr : -1; out[3 * k + 2] = len > 1 ? nc : -1;
}
__global__ void maze_kernel(uint64_t seed, int64_t env_id_base, int64_t n, uint8_t *maze, uint32_t *ctr_out) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const MazeOut mz = generate_maze_mask(seed, env_id_base + k, 0u);
    const uint64_t lo = mz.lo, hi = mz.hi;
    const uint32_t ctr = mz.ctr;
    for (int i = 0; i < CELLS; ++i) maze[k * CELLS + i] = (uint8_t)((i < 64 ? lo >> i : hi >> (i - 64)) & 1ull);
    if (ctr_out) ctr_out[k] = ctr;
}


// Per-tick observation normaliser of the reference's trainers (train_multi_ppo.py:89-91 with models/ppo_utils.py:
// 13-29): a FRESH RunningMeanStd(shape=(A, D), epsilon) is updated with the env's own [A, D] block and applied to
// it, i.e. per env and column  bm = mean_a x, bv = mean_a (x - bm)^2, total = eps + A,
//   mean = bm A / total,  var = 1 + (bv A + bm^2 eps A / total) / total,  y = (x - mean) / (sqrt(var) + 1e-8).
// One pass: each thread owns 4 consecutive columns of one env and its A rows (16-byte loads / stores) - the torch
// formulation costs ~10 elementwise / reduction passes over the 277 MB observation tensor per tick.
// ROWS > 0: compile-time row count (the env's rows stay in registers; a runtime-indexed array lives in local memory
// and the kernel then runs at a quarter of the HBM rate); ROWS == 0: any row count up to AT_MAX_TANKS.
template <int VEC, int ROWS>
__global__ void __launch_bounds__(256) tick_normalize_kernel(const float *__restrict__ x, float *__restrict__ y,
                                                             int64_t n_envs, int rows_rt, int dim, float eps) {
    const int rows = ROWS ? ROWS : rows_rt;
    constexpr int RMAX = ROWS ? ROWS : AT_MAX_TANKS;
    constexpr int UN = ROWS ? (ROWS <= 2 ? 4 : 2) : 1;   // column chunks per thread: all their loads are issued first
    const int per_env = dim / VEC;
    const int64_t total_k = n_envs * per_env;
    const int64_t k0 = (int64_t)blockIdx.x * (blockDim.x * UN) + threadIdx.x;
    float v[UN][RMAX][VEC];
    const float *src[UN];
    bool ok[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        const int64_t k = k0 + (int64_t)u * blockDim.x;
        ok[u] = k < total_k;
        const int64_t e = ok[u] ? k / per_env : 0;
        const int c0 = ok[u] ? (int)(k - e * per_env) * VEC : 0;
        src[u] = x + (e * rows) * dim + c0;
#pragma unroll
        for (int a = 0; a < RMAX; ++a) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) v[u][a][q] = 0.f;
            if (ok[u] && a < rows) {
                if (VEC == 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(src[u] + (int64_t)a * dim);
                    v[u][a][0] = t.x; v[u][a][1 % VEC] = t.y; v[u][a][2 % VEC] = t.z; v[u][a][3 % VEC] = t.w;
                } else {
                    v[u][a][0] = src[u][(int64_t)a * dim];
                }
            }
        }
    }
    const TickNormK nk(rows, eps);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        if (!ok[u]) continue;
        float *dst = y + (src[u] - x);
        float mean[VEC], inv[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            float col[RMAX];
#pragma unroll
            for (int a = 0; a < RMAX; ++a) col[a] = v[u][a][q];
            tick_norm_column<RMAX>(col, rows, nk, mean[q], inv[q]);
        }
#pragma unroll
        for (int a = 0; a < RMAX; ++a) {
            if (a < rows) {
                if (VEC == 4) {
                    float4 t;
                    t.x = (v[u][a][0] - mean[0]) * inv[0]; t.y = (v[u][a][1 % VEC] - mean[1 % VEC]) * inv[1 % VEC];
                    t.z = (v[u][a][2 % VEC] - mean[2 % VEC]) * inv[2 % VEC]; t.w = (v[u][a][3 % VEC] - mean[3 % VEC]) * inv[3 % VEC];
                    *reinterpret_cast<float4 *>(dst + (int64_t)a * dim) = t;
                } else {
                    dst[(int64_t)a * dim] = (v[u][a][0] - mean[0]) * inv[0];
                }
            }
        }
    }
}

// Categorical sampling of the policies' action heads (models/ppo_utils.py:53-61: Categorical(logits=..).sample()
// per head + log_prob) as ONE pass: log-softmax, Gumbel-max draw and the chosen log-probability of up to 4 heads per
// row.  The torch formulation is ~15 small kernels per policy and tick.  Uniforms come from Philox4x32-10 keyed by
// `seed`, counter = (row, draw block, *d_counter, stream) - the caller bumps *d_counter once per tick on the device,
// so a CUDA-graph replay draws fresh numbers.
struct SampleHeads {
    const float *logits[4];
    int dims[4];
    int n_heads;
};
__global__ void __launch_bounds__(256) sample_heads_kernel(const SampleHeads h, int64_t n_rows, uint64_t seed,
                                                           const int64_t *__restrict__ d_counter, uint32_t stream_id,
                                                           int32_t *__restrict__ actions, int64_t act_stride,
                                                           float *__restrict__ logprobs, int64_t lp_stride) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t ctr = d_counter ? (uint64_t)*d_counter : 0ull;
    uint32_t u[8];
    for (int b = 0; b < 2; ++b) {
        uint32_t c0 = (uint32_t)row, c1 = (uint32_t)((uint64_t)row >> 32) | ((uint32_t)b << 31), c2 = (uint32_t)ctr,
                 c3 = (uint32_t)(ctr >> 32) ^ (stream_id * 0x9E3779B9u);
        philox4x32_10(c0, c1, c2, c3, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32));
        u[4 * b] = c0; u[4 * b + 1] = c1; u[4 * b + 2] = c2; u[4 * b + 3] = c3;
    }
    int draw = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= h.n_heads) break;
        const int d = h.dims[k];
        const float *l = h.logits[k] + row * d;
        float v[4];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = j < d ? l[j] : -INFINITY;
            mx = fmaxf(mx, v[j]);
        }
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) se += j < d ? expf(v[j] - mx) : 0.f;
        const float lse = mx + logf(se);
        int best = 0;
        float best_s = -INFINITY, best_lp = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < d) {
                const float lp = v[j] - lse;
                const float un = ((float)u[(draw + j) & 7] + 0.5f) * 2.3283064365386963e-10f;   // (0, 1]
                const float g = -logf(fmaxf(-logf(fminf(un, 0.99999994f)), 1e-20f));           // Gumbel(0, 1)
                if (lp + g > best_s) { best_s = lp + g; best = j; best_lp = lp; }
            }
        }
        draw += d;
        actions[row * act_stride + k] = best;
        logprobs[row * lp_stride + k] = best_lp;
    }
}

// The tail of a policy's inference pass (models/ppo_utils.py:31-62: per network Linear(128, 128) -> Tanh -> Linear(128,
// out), then Categorical(logits).sample() / log_prob per head and the critic's value) as ONE pass over the second
// layers' pre-activations: tanh, the 4 small output layers (128 x {1, 3, 3, 2}), log-softmax, Gumbel-max draw, chosen
// log-probability.  Half a warp per row, the 9 dot products reduced with shuffles.  Replaces 8 tanh kernels, 8 skinny GEMMs / GEMVs and the
// sampling kernel per policy and tick.
struct PolicyTail {
    const float *z2[4];     // [n_rows][128] pre-activations of layer two: critic, head 0, head 1, head 2
    const float *w3;        // [9][128] output-layer rows in that order (critic; 3 + 3 + 2 logits)
    const float *b3;        // [9]
};
// tanh(x) = 1 - 2 / (exp(2x) + 1) with the hardware exponential and reciprocal: ~1e-6 absolute error (the GEMMs in
// front of it run in TF32, ~1e-3 relative), a quarter of the instructions of tanhf
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(2.0f * fminf(fmaxf(x, -15.0f), 15.0f));
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__global__ void __launch_bounds__(256) policy_tail_kernel(const PolicyTail p, int64_t n_rows, uint64_t seed,
                                                          const int64_t *__restrict__ d_counter, uint32_t stream_id,
                                                          int32_t *__restrict__ actions, int64_t act_stride,
                                                          float *__restrict__ logprobs, int64_t lp_stride,
                                                          float *__restrict__ values, int64_t val_stride) {
    // half a warp per row: lane h (0..15) holds elements 8h .. 8h+7 of each network's 128-vector (two 16-byte loads
    // per network, eight in flight); the 9 dot products of both rows are reduced together with 4 shuffle levels
    __shared__ float4 w_s[9][32];
    __shared__ float b_s[9];
    for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) w_s[i / 32][i % 32] = reinterpret_cast<const float4 *>(p.w3)[i];
    if (threadIdx.x < 9) b_s[threadIdx.x] = p.b3[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, h = lane & 15;
    const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
    const uint64_t ctr = d_counter ? (uint64_t)*d_counter : 0ull;
    constexpr int first[4] = {0, 1, 4, 7}, dims[4] = {1, 3, 3, 2};
    const int64_t rounds = (n_rows + nhw - 1) / nhw;   // (uniform trip count: the shuffles need the whole warp)
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t row = hw0 + it * nhw;
        const bool in = row < n_rows;
        float4 v[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < 2; ++q)
                v[k][q] = in ? __ldcs(reinterpret_cast<const float4 *>(p.z2[k] + row * 128) + 2 * h + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float out[9];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float t[8] = {fast_tanh(v[k][0].x), fast_tanh(v[k][0].y), fast_tanh(v[k][0].z), fast_tanh(v[k][0].w),
                          fast_tanh(v[k][1].x), fast_tanh(v[k][1].y), fast_tanh(v[k][1].z), fast_tanh(v[k][1].w)};
#pragma unroll
            for (int j = 0; j < dims[k]; ++j) {
                const float4 w0 = w_s[first[k] + j][2 * h], w1 = w_s[first[k] + j][2 * h + 1];
                out[first[k] + j] = ((t[0] * w0.x + t[1] * w0.y) + (t[2] * w0.z + t[3] * w0.w)) +
                                    ((t[4] * w1.x + t[5] * w1.y) + (t[6] * w1.z + t[7] * w1.w));
            }
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) out[j] += __shfl_xor_sync(0xffffffffu, out[j], d);
            out[j] += b_s[j];
        }
        if (h == 0 && in) {
            values[row * val_stride] = out[0];
            uint32_t u[8];
            for (int b = 0; b < 2; ++b) {
                uint32_t c0 = (uint32_t)row, c1 = (uint32_t)((uint64_t)row >> 32) | ((uint32_t)b << 31), cc = (uint32_t)ctr,
                         c3 = (uint32_t)(ctr >> 32) ^ (stream_id * 0x9E3779B9u);
                philox4x32_10(c0, c1, cc, c3, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32));
                u[4 * b] = c0; u[4 * b + 1] = c1; u[4 * b + 2] = cc; u[4 * b + 3] = c3;
            }
            int draw = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                const int d = dims[k];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < d) mx = fmaxf(mx, out[first[k] + j]);
                float se = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < d) se += expf(out[first[k] + j] - mx);
                const float lse = mx + logf(se);
                int best = 0;
                float best_s = -INFINITY, best_lp = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (j < d) {
                        const float lp = out[first[k] + j] - lse;
                        const float un = ((float)u[(draw + j) & 7] + 0.5f) * 2.3283064365386963e-10f;
                        const float g = -logf(fmaxf(-logf(fminf(un, 0.99999994f)), 1e-20f));
                        if (lp + g > best_s) { best_s = lp + g; best = j; best_lp = lp; }
                    }
                }
                draw += d;
                actions[row * act_stride + (k - 1)] = best;
                logprobs[row * lp_stride + (k - 1)] = best_lp;
            }
        }
    }
}

// The middle of a policy's inference pass on the tensor cores: for network k (blockIdx.y: critic, three actor heads)
//   out9[row][first_k ..] = W3_k . tanh(W2_k . tanh(z1[row][128 k .. 128 k + 128)) + b2_k) + b3_k
// i.e. tanh of the first layers' pre-activations, the second layers (4 x 128 x 128, TF32 mma.sync m16n8k8 with fp32
// accumulation - the precision of the cuBLAS TF32 GEMMs it replaces), their tanh and the tiny output layers, without
// the [N, 512] activations ever going back to HBM.  A block keeps W2_k in shared memory (padded: conflict-free
// fragment loads) and streams 128-row tiles of z1 through it; compute warp w owns rows 16 w .. 16 w + 15 of the tile.
constexpr int PM_ROWS = 128, PM_LD = 132, PM_CWARPS = 8, PM_LWARPS = 4, PM_THREADS = (PM_CWARPS + PM_LWARPS) * 32;
struct PolicyMid {
    const float *z1;   // [n_rows][512] pre-activations of the four first layers (bias included)
    const float *w2;   // [4][128 out][128 in]
    const float *b2;   // [4][128]
    const float *w3;   // [9][128] (critic; 3 + 3 + 2 logits)
    const float *b3;   // [9]
    float *out9;       // [n_rows][9]
};
__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(v));
    return r;
}
// One block per SM: 8 compute warps (16 rows of a 128-row tile each) and 4 loader warps that fill the other of two
// A buffers with the next tile (z1 -> tanh -> tf32) meanwhile; the roles meet once per tile at a block barrier.
__global__ void __launch_bounds__(PM_THREADS, 1) policy_mid_kernel(const PolicyMid p, int64_t n_rows) {
    extern __shared__ __align__(16) unsigned char pm_smem[];
    uint32_t *Ws = (uint32_t *)pm_smem;                 // [128 n][PM_LD]  tf32 bits of W2_k[n][k]
    uint32_t *As0 = Ws + 128 * PM_LD;                   // 2 x [128 rows][PM_LD] tf32 bits of tanh(z1)
    float *b2s = (float *)(As0 + 2 * PM_ROWS * PM_LD);  // [128]
    float *w3s = b2s + 128;                             // [3][128]
    const int k = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int first = k == 0 ? 0 : (k == 1 ? 1 : (k == 2 ? 4 : 7)), nout = k == 0 ? 1 : (k == 3 ? 2 : 3);
    for (int i = tid; i < 128 * 32; i += PM_THREADS) {  // W2_k, 16 bytes at a time
        const int n = i >> 5, c4 = (i & 31) * 4;
        const float4 w = __ldg(reinterpret_cast<const float4 *>(p.w2 + ((int64_t)k * 128 + n) * 128 + c4));
        uint4 t;
        t.x = to_tf32(w.x); t.y = to_tf32(w.y); t.z = to_tf32(w.z); t.w = to_tf32(w.w);
        *reinterpret_cast<uint4 *>(Ws + n * PM_LD + c4) = t;
    }
    if (tid < 128) {
        b2s[tid] = __ldg(p.b2 + k * 128 + tid);
        for (int j = 0; j < 3; ++j) w3s[j * 128 + tid] = j < nout ? __ldg(p.w3 + (first + j) * 128 + tid) : 0.f;
    }
    const float b3v[3] = {__ldg(p.b3 + first), nout > 1 ? __ldg(p.b3 + first + 1) : 0.f, nout > 2 ? __ldg(p.b3 + first + 2) : 0.f};
    const int64_t n_tiles = (n_rows + PM_ROWS - 1) / PM_ROWS;
    const bool loader = warp >= PM_CWARPS;
    const int ltid = tid - PM_CWARPS * 32;              // 0 .. 127 among the loader threads
    // a tile's slice of z1 -> tanh -> tf32 (loader warps): 16 loads of a thread in flight at a time
    auto fill = [&](int64_t tile, uint32_t *As) {
        const int64_t row0 = tile * PM_ROWS;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float4 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int i = ltid + (half * 16 + q) * (PM_LWARPS * 32), r = i >> 5, c4 = (i & 31) * 4;
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tile < n_tiles && row0 + r < n_rows)
                    v[q] = __ldcs(reinterpret_cast<const float4 *>(p.z1 + (row0 + r) * 512 + k * 128 + c4));
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int i = ltid + (half * 16 + q) * (PM_LWARPS * 32), r = i >> 5, c4 = (i & 31) * 4;
                uint4 t;
                t.x = to_tf32(fast_tanh(v[q].x)); t.y = to_tf32(fast_tanh(v[q].y)); t.z = to_tf32(fast_tanh(v[q].z)); t.w = to_tf32(fast_tanh(v[q].w));
                *reinterpret_cast<uint4 *>(As + r * PM_LD + c4) = t;
            }
        }
    };
    if (loader) fill(blockIdx.x, As0);
    __syncthreads();   // the tables and the first tile
    int buf = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        uint32_t *As = As0 + buf * PM_ROWS * PM_LD;
        if (loader) {
            fill(tile + gridDim.x, As0 + (buf ^ 1) * PM_ROWS * PM_LD);
        } else {
            const int64_t row0 = tile * PM_ROWS;
            float acc[16][4];
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
            // fragments by ldmatrix: a tf32 8 x 4 block is an 8 x 8 b16 matrix (row = 16 bytes), and thread (gid, tig) of
            // an ldmatrix result holds row gid, 32-bit column tig - exactly the m16n8k8 A (row, k) / B (n, k) fragment order.
            // lane l supplies the row address of matrix l / 8: A = {rows 0-7 | 8-15} x {k 0-3 | 4-7}, B = two n-tiles x {k 0-3 | 4-7}
            const uint32_t a_addr = (uint32_t)__cvta_generic_to_shared(As + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PM_LD + (lane >> 4) * 4);
            const uint32_t b_addr = (uint32_t)__cvta_generic_to_shared(Ws + ((lane & 7) + (lane >> 4) * 8) * PM_LD + ((lane >> 3) & 1) * 4);
#pragma unroll 2
            for (int ks = 0; ks < 16; ++ks) {
                uint32_t a0, a1, a2, a3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
                             : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(a_addr + ks * 32));
#pragma unroll
                for (int np = 0; np < 8; ++np) {   // two n-tiles per ldmatrix
                    uint32_t b0, b1, b2, b3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
                                 : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(b_addr + (np * 16 * PM_LD + ks * 8) * 4));
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
                                 : "+f"(acc[2 * np][0]), "+f"(acc[2 * np][1]), "+f"(acc[2 * np][2]), "+f"(acc[2 * np][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
                                 : "+f"(acc[2 * np + 1][0]), "+f"(acc[2 * np + 1][1]), "+f"(acc[2 * np + 1][2]), "+f"(acc[2 * np + 1][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b2), "r"(b3));
                }
            }
            // epilogue: this thread holds columns nt * 8 + 2 tig (+1) of rows gid (acc[.][0..1]) and gid + 8 (acc[.][2..3])
            float o_lo[3] = {0.f, 0.f, 0.f}, o_hi[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int col = nt * 8 + 2 * tig + q;
                    const float bb = b2s[col];
                    const float h_lo = fast_tanh(acc[nt][q] + bb), h_hi = fast_tanh(acc[nt][2 + q] + bb);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float w = w3s[j * 128 + col];
                        o_lo[j] += h_lo * w;
                        o_hi[j] += h_hi * w;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                o_lo[j] += __shfl_xor_sync(0xffffffffu, o_lo[j], 1); o_lo[j] += __shfl_xor_sync(0xffffffffu, o_lo[j], 2);
                o_hi[j] += __shfl_xor_sync(0xffffffffu, o_hi[j], 1); o_hi[j] += __shfl_xor_sync(0xffffffffu, o_hi[j], 2);
            }
            if (tig == 0) {
                const int64_t r_lo = row0 + warp * 16 + gid, r_hi = r_lo + 8;
                for (int j = 0; j < nout; ++j) {
                    if (r_lo < n_rows) p.out9[r_lo * 9 + first + j] = o_lo[j] + b3v[j];
                    if (r_hi < n_rows) p.out9[r_hi * 9 + first + j] = o_hi[j] + b3v[j];
                }
            }
        }
        __syncthreads();   // the next tile is in the other buffer; this one may be overwritten
    }
}

// sampling from the [n_rows][9] outputs of policy_mid_kernel (value, 3 + 3 + 2 logits): as sample_heads_kernel
__global__ void __launch_bounds__(256) sample_out9_kernel(const float *__restrict__ out9, int64_t n_rows, uint64_t seed,
                                                          const int64_t *__restrict__ d_counter, uint32_t stream_id,
                                                          int32_t *__restrict__ actions, int64_t act_stride,
                                                          float *__restrict__ logprobs, int64_t lp_stride,
                                                          float *__restrict__ values, int64_t val_stride) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t ctr = d_counter ? (uint64_t)*d_counter : 0ull;
    float out[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) out[j] = out9[row * 9 + j];
    values[row * val_stride] = out[0];
    uint32_t u[8];
    for (int b = 0; b < 2; ++b) {
        uint32_t c0 = (uint32_t)row, c1 = (uint32_t)((uint64_t)row >> 32) | ((uint32_t)b << 31), cc = (uint32_t)ctr,
                 c3 = (uint32_t)(ctr >> 32) ^ (stream_id * 0x9E3779B9u);
        philox4x32_10(c0, c1, cc, c3, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32));
        u[4 * b] = c0; u[4 * b + 1] = c1; u[4 * b + 2] = cc; u[4 * b + 3] = c3;
    }
    constexpr int first[4] = {0, 1, 4, 7}, dims[4] = {1, 3, 3, 2};
    int draw = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        const int d = dims[k];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (j < d) mx = fmaxf(mx, out[first[k] + j]);
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (j < d) se += expf(out[first[k] + j] - mx);
        const float lse = mx + logf(se);
        int best = 0;
        float best_s = -INFINITY, best_lp = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (j < d) {
                const float lp = out[first[k] + j] - lse;
                const float un = ((float)u[(draw + j) & 7] + 0.5f) * 2.3283064365386963e-10f;
                const float g = -logf(fmaxf(-logf(fminf(un, 0.99999994f)), 1e-20f));
                if (lp + g > best_s) { best_s = lp + g; best = j; best_lp = lp; }
            }
        }
        draw += d;
        actions[row * act_stride + (k - 1)] = best;
        logprobs[row * lp_stride + (k - 1)] = best_lp;
    }
}

}  // namespace

// -------------------------------------------------------------------------------------------------- host
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(x)                                                                                        \
    do {                                                                                                   \
        cudaError_t err__ = (x);                                                                           \
        if (err__ != cudaSuccess)                                                                          \
            return fail(AT_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(err__));                  \
    } while (0)

// Every entry point runs on the handle's device and leaves the calling thread's current device as it found it.
struct DeviceGuard {
    int prev;
    cudaError_t err;
    explicit DeviceGuard(int device) : prev(-1), err(cudaSuccess) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != device) err = cudaSetDevice(device);
        else prev = -1;   // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    cudaError_t status() const { return err; }
};

struct at_env {
    at_config user_cfg;
    KCfg k;
    DevState st;
    int device;
    int64_t n_envs;
    void *arena;  // one allocation, carved into the SoA columns (at_host.h arena_layout)
    size_t arena_bytes;
    int32_t *d_actions_scratch;
    uint8_t *d_actions_u8;    // at_step_host_u8: the packed actions as they arrive (allocated on first use)
    float *d_rewards_scratch;
    uint8_t *d_done_scratch;
    HostLuts luts;
    const float *d_tmpl;
    double *d_dyn;            // st.dyn: parameters that may change between steps (at_set_weakness)
    size_t smem_bytes_sim;    // env_kernel
    int sim_epw;              // envs per warp of the simulation kernel (sim_envs_per_warp; AT_SIM_EPW overrides)
    int sim_spec;             // SimT<SPEC> specialisation of the step's simulation kernel (0: generic; AT_NO_SPEC=1 forces it)
    bool pdl;                 // rows_kernel is launched as a programmatic dependent (AT_NO_PDL=1: plain launch)
    int rows_lgG, rows_grid, rows_warps;  // rows_kernel: log2(envs per group), persistent grid, warps per block
    rows_kernel_fn rows_fn, rows_fn_norm, rows_fn_half;   // this config's shape: raw rows only / + normalised fp32 / + normalised fp16
    size_t rows_smem, rows_smem_norm;   // (norm: with the second, normalised image)
    int rows_grid_norm, rows_grid_inplace;   // (in place: normalised fp32 rows only, one image)
    bool rows_skip_static;                // at_set_rows_mode: the steps' output buffers are persistent (static columns not rewritten)
    int rows_warps_h, rows_grid_h;        // fp16 normalised rows (at_step_norm16): warps per block, grid
    size_t rows_smem_h;
    int64_t launches;
    unsigned int *d_sig_counter;          // at_step_host: finished-block counter of the simulation kernel
    volatile uint32_t *h_sig_flag;        // ... and the mapped host word its last block posts the epoch to
    volatile uint32_t *h_sig_flag_dev;    // (device alias of h_sig_flag)
    uint32_t sig_epoch;
    unsigned long long *dbg_times;        // diagnostics: see at_debug_block_times
    cudaStream_t copy_stream;             // at_step_host: the actions' DMA runs beside the previous step's rows kernel
    cudaEvent_t ev_actions;
    std::vector<std::pair<const void *, void *>> mapped;  // host buffers seen by at_step_host -> device alias (null: not mapped)
};

template <class... Args>
static cudaError_t launch_pdl(void (*fn)(Args...), int grid, int block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, fn, args...);
}

static int launch_rows_kernel(at_env *env, const uint8_t *d_mask, float *d_obs, cudaStream_t s, float *d_obs_norm, float norm_eps,
                              bool pdl, bool norm_half = false, bool allow_skip = false);
static int launch_env_kernel(at_env *env, bool is_step, const int32_t *d_actions, const uint8_t *d_mask,
                             float *d_obs, float *d_rewards, uint8_t *d_done, cudaStream_t s, float *d_obs_norm = nullptr,
                             float norm_eps = 0.f, bool post_to_host = false, bool norm_half = false) {
    const int epw = env->dbg_times ? 32 : env->sim_epw;   // (the block stamps are sized for full warps)
    const int per_block = (BS / 32) * epw;
    const int grid = (int)((env->n_envs + per_block - 1) / per_block);
    HostSignal sig = {nullptr, nullptr, 0u, env->dbg_times, epw};
    if (post_to_host) { sig.counter = env->d_sig_counter; sig.host_flag = env->h_sig_flag_dev; sig.epoch = ++env->sig_epoch; }
    const bool rows = (d_obs != nullptr || d_obs_norm != nullptr) && env->k.rows > 0;
    if (is_step && env->sim_spec == 5)
        env_kernel<true, 5><<<grid, BS, env->smem_bytes_sim, s>>>(env->k, env->st, d_actions, d_mask, d_rewards, d_done, sig);
    else if (is_step && env->sim_spec == 4)
        env_kernel<true, 4><<<grid, BS, env->smem_bytes_sim, s>>>(env->k, env->st, d_actions, d_mask, d_rewards, d_done, sig);
    else if (is_step && env->sim_spec == 3)
        env_kernel<true, 3><<<grid, BS, env->smem_bytes_sim, s>>>(env->k, env->st, d_actions, d_mask, d_rewards, d_done, sig);
    else if (is_step && env->sim_spec == 2)
        env_kernel<true, 2><<<grid, BS, env->smem_bytes_sim, s>>>(env->k, env->st, d_actions, d_mask, d_rewards, d_done, sig);
    else if (is_step && env->sim_spec == 1)
        env_kernel<true, 1><<<grid, BS, env->smem_bytes_sim, s>>>(env->k, env->st, d_actions, d_mask, d_rewards, d_done, sig);
    else if (is_step)
        env_kernel<true><<<grid, BS, env->smem_bytes_sim, s>>>(env->k, env->st, d_actions, d_mask, d_rewards, d_done, sig);
    else
        env_kernel<false><<<grid, BS, env->smem_bytes_sim, s>>>(env->k, env->st, d_actions, d_mask, d_rewards, d_done, sig);
    env->launches += 1;
    CUDA_TRY(cudaGetLastError());
    if (!rows) return AT_OK;
    return launch_rows_kernel(env, d_mask, d_obs, s, d_obs_norm, norm_eps, env->pdl, norm_half, is_step && d_mask == nullptr);
}

// rows_kernel: the observation rows (raw and / or tick-normalised) of the stored state
static int launch_rows_kernel(at_env *env, const uint8_t *d_mask, float *d_obs, cudaStream_t s, float *d_obs_norm, float norm_eps,
                              bool pdl, bool norm_half, bool allow_skip) {
    const int G = 1 << env->rows_lgG;
    const int n_groups = (int)((env->n_envs + G - 1) / G);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    norm_half = norm_half && d_obs_norm;
    const int warps = norm_half ? env->rows_warps_h : env->rows_warps;
    const int want = (n_groups + warps - 1) / warps;
    const int rgrid = norm_half ? env->rows_grid_h : ((d_obs_norm && d_obs) ? env->rows_grid_norm : (d_obs_norm ? env->rows_grid_inplace : env->rows_grid));
    cfg.gridDim = dim3((unsigned)(want < rgrid ? want : rgrid));
    cfg.blockDim = dim3(warps * 32);
    cfg.dynamicSmemBytes = norm_half ? env->rows_smem_h : ((d_obs_norm && d_obs) ? env->rows_smem_norm : env->rows_smem);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, norm_half ? env->rows_fn_half : (d_obs_norm ? env->rows_fn_norm : env->rows_fn), env->k, env->st, d_mask, (float *)d_obs, env->d_tmpl, env->rows_lgG, n_groups,
                                (float *)d_obs_norm, norm_eps, (norm_half ? 1 : 0) | (allow_skip && env->rows_skip_static ? 2 : 0)));
    env->launches += 1;
    return AT_OK;
}

extern "C" {

const char *at_last_error(void) { return g_err.c_str(); }
const char *at_version(void) { return "alphatank_b200 0.1 (sm_100a)"; }

int at_create(const at_config *cfg, int64_t n_envs, int64_t env_id_base, uint64_t seed, int device, at_env **out) {
    if (!cfg || !out || n_envs <= 0) return fail(AT_ERR_ARG, "at_create: null config / handle or n_envs <= 0");
    if (n_envs * AT_MAX_TANKS * SLOTS_PER_TANK >= ((int64_t)1 << 31))
        return fail(AT_ERR_ARG, "at_create: more than 2^31 / 48 envs per handle (32-bit column indices); shard the batch");
    KCfg k;
    std::string msg = build_kcfg(cfg, seed, env_id_base, k);
    if (!msg.empty()) return fail(AT_ERR_ARG, "at_create: " + msg);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(AT_ERR_NOGPU, "at_create: no usable CUDA device (the batched env has no CPU fallback)");
    }
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(AT_ERR_NOGPU, "at_create: device is not sm_100 class; this library ships sm_100a code only");

    {
        const M128 nm = near_wall_mask(k.wall_lo, k.wall_hi);
        k.near_lo = nm.lo;
        k.near_hi = nm.hi;
    }
    at_env *env = new at_env();
    env->user_cfg = *cfg;
    env->k = k;
    env->device = device;
    env->n_envs = n_envs;
    env->launches = 0;
    build_luts(env->luts);
    if (!env->luts.udelta_ok) { delete env; return fail(AT_ERR_CUDA, "at_create: host libm gives a normalised direction more than 1 ulp from (cos, -sin)"); }
    const ArenaLayout L = arena_layout(k, n_envs);
    env->arena_bytes = L.bytes;
    if (cudaMalloc(&env->arena, L.bytes) != cudaSuccess) {
        std::string m = std::string("at_create: cudaMalloc of ") + std::to_string(L.bytes) + " bytes failed";
        cudaGetLastError();
        delete env;
        return fail(AT_ERR_CUDA, m);
    }
    cudaMemset(env->arena, 0, L.bytes);
    char *b = (char *)env->arena;
    cudaMemset(b + L.o_ttick, 0xFF, 4 * (size_t)n_envs);   // no terminal info latched yet (tick starts at 0)
    bind_state(env->st, b, L, n_envs);
    env->d_actions_scratch = (int32_t *)(b + L.o_act);
    env->d_actions_u8 = nullptr;
    env->d_rewards_scratch = (float *)(b + L.o_rw);
    env->d_done_scratch = (uint8_t *)(b + L.o_dn);
    cudaMemcpy(b + L.o_trig, env->luts.trig.data(), 722 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(b + L.o_corn, env->luts.corn.data(), 361 * 4 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(b + L.o_pend, env->luts.pend.data(), PEND_LUT * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(b + L.o_udir, env->luts.udelta.data(), 364, cudaMemcpyHostToDevice);
    if (k.map != AT_MAP_MAZE) {  // fixed map: path-finding becomes one table lookup
        std::vector<uint16_t> tab((size_t)CELLS * CELLS);
        fill_bfs_table(k.wall_lo, k.wall_hi, tab.data());
        cudaMemcpy(b + L.o_bfs, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice);
        env->st.bfs_tab = (const uint16_t *)(b + L.o_bfs);
    }

    env->dbg_times = nullptr;
    env->rows_skip_static = false;
    env->d_sig_counter = nullptr; env->h_sig_flag = nullptr; env->h_sig_flag_dev = nullptr; env->sig_epoch = 0;
    {
        void *hp = nullptr, *dp = nullptr;
        if (cudaMalloc(&dp, 64) == cudaSuccess) { cudaMemset(dp, 0, 64); env->d_sig_counter = (unsigned int *)dp; }
        if (cudaHostAlloc(&hp, 64, cudaHostAllocMapped) == cudaSuccess) {
            memset(hp, 0, 64);
            void *alias = nullptr;
            if (cudaHostGetDevicePointer(&alias, hp, 0) == cudaSuccess) {
                env->h_sig_flag = (volatile uint32_t *)hp;
                env->h_sig_flag_dev = (volatile uint32_t *)alias;
            } else {
                cudaFreeHost(hp);
            }
        }
        env->copy_stream = nullptr; env->ev_actions = nullptr;
        if (cudaStreamCreateWithFlags(&env->copy_stream, cudaStreamNonBlocking) != cudaSuccess) env->copy_stream = nullptr;
        if (cudaEventCreateWithFlags(&env->ev_actions, cudaEventDisableTiming) != cudaSuccess) env->ev_actions = nullptr;
        cudaGetLastError();
    }
    {
        std::vector<double> at((size_t)361 * ATAB_STRIDE);
        fill_atan_table(at.data());
        cudaMemcpy(b + L.o_atab, at.data(), at.size() * 8, cudaMemcpyHostToDevice);
    }
    env->d_dyn = (double *)(b + L.o_dyn);
    cudaMemcpy(env->d_dyn, &k.weakness, 8, cudaMemcpyHostToDevice);
    env->d_tmpl = (const float *)(b + L.o_tmpl);
    if (k.map != AT_MAP_MAZE) {
        std::vector<float> tm((size_t)4 * k.n_walls + CELLS);
        fill_static_template(k.wall_lo, k.wall_hi, tm.data());
        cudaMemcpy(b + L.o_tmpl, tm.data(), tm.size() * 4, cudaMemcpyHostToDevice);
    }
    env->smem_bytes_sim = smem_layout(k.n_tanks, k.n_bots, k.act_rows).total;
    {   // The dynamic shared-memory limit is a process-wide attribute of each kernel instantiation: never set it to this
        // handle's own need (a second handle with a smaller image would lower it under the first one) - raise every
        // kernel to the device's opt-in maximum once and leave it there.
        const int lim = (int)prop.sharedMemPerBlockOptin;
        cudaFuncSetAttribute(env_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        cudaFuncSetAttribute(env_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        cudaFuncSetAttribute(env_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        cudaFuncSetAttribute(env_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        cudaFuncSetAttribute(env_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        cudaFuncSetAttribute(env_kernel<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        cudaFuncSetAttribute(env_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        if (env->smem_bytes_sim > (size_t)lim) {
            cudaFree(env->arena);
            delete env;
            return fail(AT_ERR_ARG, "at_create: the simulation's working set does not fit shared memory");
        }
    }
    {   // rows_kernel: the largest group (power of two, <= 4 envs) whose image keeps 12 warps on an SM, else whatever
        // fits at all; the G * 6n (env, slot) pairs must fit the lanes' prefetch registers (<= 3 per lane)
        const char *np = getenv("AT_NO_PDL"), *gg = getenv("AT_ROWS_LGG");
        env->pdl = !(np && np[0] == '1');
        {   // AT_NO_SPEC=1: generic kernel; 2 / 3: stop below that specialisation level
            const char *ns = getenv("AT_NO_SPEC");
            env->sim_spec = sim_spec_for(k, ns && ns[0] >= '1' && ns[0] <= '3' ? ns[0] - '1' : 3);
        }
        {   // envs per warp: spread a small batch until every warp scheduler (4 per SM) has one warp - measured on a
            // B200 (profiles/r06_small_batches.md): 256 envs 0.104 -> 0.059 ms, 4096 envs 0.112 -> 0.092 ms per tick;
            // past one warp per scheduler the spread costs more (the warps contend) than the shorter paths save
            const char *ep = getenv("AT_SIM_EPW");
            const int64_t slots = (int64_t)prop.multiProcessorCount * 4;
            int epw = 1;
            while (epw < 32 && (n_envs + epw - 1) / epw > slots) epw *= 2;
            if (ep && atoi(ep) >= 1 && atoi(ep) <= 32) epw = atoi(ep);
            env->sim_epw = epw;
        }
        const int RD = k.rows * k.obs_dim;
        const char *ww = getenv("AT_ROWS_WARPS");
        int W = ww ? atoi(ww) : 8;
        if (W < 1) W = 1;
        if (W > RK_MAX_WARPS) W = RK_MAX_WARPS;
        env->rows_warps = W;
        int lg = gg ? atoi(gg) : 1;
        if (lg < 0) lg = 0;
        if (lg > 3) lg = 3;
        while (lg > 0 && (rows_smem_layout(1 << lg, k.n_tanks, RD, W, 4).total > (size_t)prop.sharedMemPerBlockOptin ||
                          (1 << lg) * k.n_tanks * SLOTS_PER_TANK > 96 || (1 << lg) * k.n_tanks > 32))
            --lg;
        const int pairs = (1 << lg) * k.n_tanks * SLOTS_PER_TANK;
#define AT_ROWS_PICK(...) do { env->rows_fn = rows_kernel<__VA_ARGS__, 0>; env->rows_fn_norm = rows_kernel<__VA_ARGS__, 1>; \
                               env->rows_fn_half = rows_kernel<__VA_ARGS__, 2>; } while (0)
        if (k.map == AT_MAP_MAZE) AT_ROWS_PICK(0, 0, -1, true, 3);
        else if (pairs <= 64 && k.n_tanks == 4 && k.rows == 2 && k.team_layout) AT_ROWS_PICK(4, 2, 1, false, 2);
        else if (pairs <= 96 && k.n_tanks == 4 && k.rows == 2 && k.team_layout) AT_ROWS_PICK(4, 2, 1, false, 3);
        else if (pairs <= 64 && k.n_tanks == 2 && k.rows == 2 && !k.team_layout) AT_ROWS_PICK(2, 2, 0, false, 2);
        else AT_ROWS_PICK(0, 0, -1, false, 3);
#undef AT_ROWS_PICK
        env->rows_lgG = lg;
        env->rows_smem = rows_smem_layout(1 << lg, k.n_tanks, RD, W, 0).total;
        env->rows_smem_norm = rows_smem_layout(1 << lg, k.n_tanks, RD, W, 4).total;
        env->rows_grid_norm = prop.multiProcessorCount;
        env->rows_grid_inplace = prop.multiProcessorCount;
        env->rows_grid = prop.multiProcessorCount;
        if (k.rows > 0) {
            if (env->rows_smem_norm > (size_t)prop.sharedMemPerBlockOptin) {
                cudaFree(env->arena);
                delete env;
                return fail(AT_ERR_ARG, "at_create: an observation image does not fit shared memory");
            }
            // (never this handle's own size: see the note at the simulation kernels above)
            cudaFuncSetAttribute((const void *)env->rows_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
            cudaFuncSetAttribute((const void *)env->rows_fn_norm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
            cudaFuncSetAttribute((const void *)env->rows_fn_half, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
            int bps = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, (const void *)env->rows_fn, W * 32, env->rows_smem) != cudaSuccess || bps < 1) {
                cudaGetLastError();
                bps = 1;
            }
            const char *bp = getenv("AT_ROWS_BPS");
            if (bp && atoi(bp) > 0 && atoi(bp) < bps) bps = atoi(bp);
            env->rows_grid = prop.multiProcessorCount * bps;
            int bpn = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bpn, (const void *)env->rows_fn_norm, W * 32, env->rows_smem_norm) != cudaSuccess || bpn < 1) {
                cudaGetLastError();
                bpn = 1;
            }
            env->rows_grid_norm = prop.multiProcessorCount * bpn;
            int bpi = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bpi, (const void *)env->rows_fn_norm, W * 32, env->rows_smem) != cudaSuccess || bpi < 1) {
                cudaGetLastError();
                bpi = 1;
            }
            env->rows_grid_inplace = prop.multiProcessorCount * bpi;
            // fp16 normalised rows: raw fp32 image + fp16 image per warp; the warps per block that put most warps on an SM
            int best_w = 1, best_tot = 0, best_b = 1;
            for (int w = RK_MAX_WARPS; w >= 1; --w) {
                const size_t sm = rows_smem_layout(1 << lg, k.n_tanks, RD, w, 2).total;
                if (sm > (size_t)prop.sharedMemPerBlockOptin) continue;
                int b = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, (const void *)env->rows_fn_half, w * 32, sm) != cudaSuccess) { cudaGetLastError(); b = 0; }
                if (b * w > best_tot) { best_tot = b * w; best_w = w; best_b = b; }
            }
            const char *wh = getenv("AT_ROWS_WARPS_H");
            if (wh && atoi(wh) >= 1 && atoi(wh) <= RK_MAX_WARPS) {
                best_w = atoi(wh);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&best_b, (const void *)env->rows_fn_half, best_w * 32,
                                                                  rows_smem_layout(1 << lg, k.n_tanks, RD, best_w, 2).total) != cudaSuccess || best_b < 1) {
                    cudaGetLastError();
                    best_b = 1;
                }
            }
            env->rows_warps_h = best_w;
            env->rows_smem_h = rows_smem_layout(1 << lg, k.n_tanks, RD, best_w, 2).total;
            env->rows_grid_h = prop.multiProcessorCount * (best_b < 1 ? 1 : best_b);
        }
    }
    // GamingENV.__init__ -> self.reset() (gaming_env.py:43): the constructor's own reset consumes draws
    int rc = launch_env_kernel(env, false, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
    if (rc == AT_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(AT_ERR_CUDA, "at_create: reset kernel failed");
    if (rc != AT_OK) {
        cudaFree(env->arena);
        delete env;
        return rc;
    }
    *out = env;
    return AT_OK;
}

int at_destroy(at_env *env) {
    if (!env) return AT_OK;
    DeviceGuard guard(env->device);
    cudaDeviceSynchronize();
    cudaFree(env->arena);
    if (env->d_actions_u8) cudaFree(env->d_actions_u8);
    if (env->d_sig_counter) cudaFree(env->d_sig_counter);
    if (env->h_sig_flag) cudaFreeHost((void *)env->h_sig_flag);
    if (env->copy_stream) cudaStreamDestroy(env->copy_stream);
    if (env->ev_actions) cudaEventDestroy(env->ev_actions);
    delete env;
    return AT_OK;
}

int at_obs_dim(const at_env *env) { return env ? env->k.obs_dim : AT_ERR_ARG; }
int at_obs_rows(const at_env *env) { return env ? env->k.rows : AT_ERR_ARG; }
int at_action_rows(const at_env *env) { return env ? env->k.act_rows : AT_ERR_ARG; }
int at_num_walls(const at_env *env) { return env ? env->k.n_walls : AT_ERR_ARG; }
int64_t at_num_envs(const at_env *env) { return env ? env->n_envs : AT_ERR_ARG; }
int64_t at_launch_count(const at_env *env) { return env ? env->launches : AT_ERR_ARG; }

int at_reset(at_env *env, const uint8_t *d_mask, float *d_obs, void *stream) {
    if (!env) return fail(AT_ERR_ARG, "at_reset: null handle");
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    return launch_env_kernel(env, false, nullptr, d_mask, d_obs, nullptr, nullptr, (cudaStream_t)stream);
}

int at_step(at_env *env, const int32_t *d_actions, float *d_obs, float *d_rewards, uint8_t *d_done, void *stream) {
    if (!env) return fail(AT_ERR_ARG, "at_step: null handle");
    if (env->k.act_rows > 0 && !d_actions) return fail(AT_ERR_ARG, "at_step: this mode needs an actions array");
    if ((env->k.rows > 0 && !d_rewards) || !d_done) return fail(AT_ERR_ARG, "at_step: rewards / done must not be null");
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    return launch_env_kernel(env, true, d_actions, nullptr, d_obs, d_rewards, d_done, (cudaStream_t)stream);
}

int at_step_norm(at_env *env, const int32_t *d_actions, float *d_obs, float *d_obs_norm, float epsilon, float *d_rewards,
                 uint8_t *d_done, void *stream) {
    if (!env) return fail(AT_ERR_ARG, "at_step_norm: null handle");
    if (env->k.act_rows > 0 && !d_actions) return fail(AT_ERR_ARG, "at_step_norm: this mode needs an actions array");
    if ((env->k.rows > 0 && !d_rewards) || !d_done) return fail(AT_ERR_ARG, "at_step_norm: rewards / done must not be null");
    if (!d_obs_norm) return fail(AT_ERR_ARG, "at_step_norm: the normalised observation buffer must not be null");
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    return launch_env_kernel(env, true, d_actions, nullptr, d_obs, d_rewards, d_done, (cudaStream_t)stream, d_obs_norm, epsilon);
}

int at_step_norm16(at_env *env, const int32_t *d_actions, float *d_obs, uint16_t *d_obs_norm16, float epsilon, float *d_rewards,
                   uint8_t *d_done, void *stream) {
    if (!env) return fail(AT_ERR_ARG, "at_step_norm16: null handle");
    if (env->k.act_rows > 0 && !d_actions) return fail(AT_ERR_ARG, "at_step_norm16: this mode needs an actions array");
    if ((env->k.rows > 0 && !d_rewards) || !d_done) return fail(AT_ERR_ARG, "at_step_norm16: rewards / done must not be null");
    if (!d_obs_norm16) return fail(AT_ERR_ARG, "at_step_norm16: the normalised observation buffer must not be null");
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    return launch_env_kernel(env, true, d_actions, nullptr, d_obs, d_rewards, d_done, (cudaStream_t)stream, (float *)d_obs_norm16, epsilon,
                             false, true);
}

// device-visible alias of a host buffer (pinned + mapped memory under unified addressing), or nullptr
static void *mapped_alias(const void *h) {
    if (!h) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged) return at.devicePointer;
    return nullptr;
}

// u8 -> i32 expansion of packed host actions (at_step_host_u8): runs on the copy stream right behind the DMA copy
__global__ void expand_actions_kernel(const uint8_t *__restrict__ src, int32_t *__restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (int32_t)src[i];
}

static int step_host_impl(at_env *env, const void *h_actions, int action_bytes, float *d_obs, float *h_rewards, uint8_t *h_done,
                          void *stream) {
    if (!env) return fail(AT_ERR_ARG, "at_step_host: null handle");
    if ((env->k.rows > 0 && !h_rewards) || !h_done) return fail(AT_ERR_ARG, "at_step_host: rewards / done must not be null");
    if (env->k.act_rows > 0 && !h_actions) return fail(AT_ERR_ARG, "at_step_host: this mode needs an actions array");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    const int64_t N = env->n_envs;
    // Pinned (mapped) host buffers are handed to the kernel as they are: it fetches each env's 24 bytes of actions
    // with asynchronous copies that are hidden behind the first bullet pass and posts rewards / done straight into
    // host memory, so the step costs one launch + one synchronisation instead of three DMA copies around the
    // kernel.  (Cutting the batch into pipelined slices was tried and is slower: the kernel is latency-bound, a
    // slice takes as long as the whole batch.)  Pageable buffers take the copy path.
    void *dev[3];
    const void *host[3] = {h_actions, h_rewards, h_done};
    for (int q = 0; q < 3; ++q) {   // the attribute query is a driver call: remember the buffers already seen
        dev[q] = nullptr;
        bool hit = false;
        for (const auto &m : env->mapped)
            if (m.first == host[q]) { dev[q] = m.second; hit = true; break; }
        if (!hit) {
            dev[q] = mapped_alias(host[q]);
            if (env->mapped.size() >= 256) env->mapped.clear();
            env->mapped.emplace_back(host[q], dev[q]);
        }
    }
    static const bool allow_direct = !getenv("AT_STEP_HOST_COPY");  // diagnostic switch: force the copy path
    const bool direct = allow_direct && (env->k.act_rows == 0 || dev[0]) && (env->k.rows == 0 || dev[1]) && dev[2] &&
                        (action_bytes == 4 || env->k.act_rows == 0 || (env->copy_stream && env->ev_actions));   // (packed actions always go by DMA)
    if (direct) {
        // The call returns as soon as rewards / done are in host memory: the simulation kernel's last block posts the
        // step's epoch into a mapped host word and the host spins on it - no driver synchronisation, and the rows
        // kernel (whose output stays in HBM, valid in stream order) overlaps the host's work for the next step.
        static const bool spin = !getenv("AT_STEP_HOST_SYNC");   // diagnostic switch: synchronise the stream instead
        const bool sigok = spin && env->d_sig_counter && env->h_sig_flag_dev;
        // Actions: a DMA copy on the handle's copy stream - the host calls in as soon as the previous step's simulation
        // is done, so the copy runs beside that step's rows kernel and the simulation reads HBM (letting the kernel fetch
        // the rows from mapped host memory itself costs it 65 us: 24-byte reads across PCIe are latency-bound).
        // Rewards / done: posted writes from the kernel straight into the mapped host buffers.
        static const bool dma = !getenv("AT_STEP_HOST_MAPPED_ACTIONS");   // diagnostic switch: the kernel fetches them
        const int32_t *acts = (const int32_t *)dev[0];
        const int64_t n_act = N * env->k.act_rows * 3;
        if (action_bytes == 1 && env->k.act_rows > 0 && !env->d_actions_u8) CUDA_TRY(cudaMalloc((void **)&env->d_actions_u8, (size_t)n_act));
        if ((dma || action_bytes == 1) && env->k.act_rows > 0 && env->copy_stream && env->ev_actions) {
            if (action_bytes == 1) {   // a quarter of the bytes across PCIe, widened on the device
                CUDA_TRY(cudaMemcpyAsync(env->d_actions_u8, h_actions, (size_t)n_act, cudaMemcpyHostToDevice, env->copy_stream));
                expand_actions_kernel<<<(int)((n_act + 255) / 256), 256, 0, env->copy_stream>>>(env->d_actions_u8, env->d_actions_scratch, n_act);
                env->launches += 1;
            } else {
                CUDA_TRY(cudaMemcpyAsync(env->d_actions_scratch, h_actions, (size_t)n_act * 4, cudaMemcpyHostToDevice, env->copy_stream));
            }
            CUDA_TRY(cudaEventRecord(env->ev_actions, env->copy_stream));
            CUDA_TRY(cudaStreamWaitEvent(s, env->ev_actions, 0));
            acts = env->d_actions_scratch;
        }
        int rc = launch_env_kernel(env, true, acts, nullptr, d_obs, (float *)dev[1], (uint8_t *)dev[2], s,
                                   nullptr, 0.f, sigok);
        if (rc != AT_OK) return rc;
        if (!sigok) {
            CUDA_TRY(cudaStreamSynchronize(s));
            return AT_OK;
        }
        const uint32_t want = env->sig_epoch;
        for (uint64_t spins = 0; *env->h_sig_flag != want; ++spins) {
            if ((spins & 0xFFFFu) == 0xFFFFu) {   // every ~65 k polls: has the stream died?
                const cudaError_t q = cudaStreamQuery(s);
                if (q != cudaSuccess && q != cudaErrorNotReady)
                    return fail(AT_ERR_CUDA, std::string("at_step_host: ") + cudaGetErrorString(q));
                if (q == cudaSuccess && *env->h_sig_flag != want)
                    return fail(AT_ERR_CUDA, "at_step_host: the stream drained without the completion signal");
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        __sync_synchronize();
        return AT_OK;
    }
    if (env->k.act_rows > 0) {
        const int64_t n_act = N * env->k.act_rows * 3;
        if (action_bytes == 1) {
            if (!env->d_actions_u8) CUDA_TRY(cudaMalloc((void **)&env->d_actions_u8, (size_t)n_act));
            CUDA_TRY(cudaMemcpyAsync(env->d_actions_u8, h_actions, (size_t)n_act, cudaMemcpyHostToDevice, s));
            expand_actions_kernel<<<(int)((n_act + 255) / 256), 256, 0, s>>>(env->d_actions_u8, env->d_actions_scratch, n_act);
            env->launches += 1;
        } else {
            CUDA_TRY(cudaMemcpyAsync(env->d_actions_scratch, h_actions, (size_t)n_act * 4, cudaMemcpyHostToDevice, s));
        }
    }
    int rc = launch_env_kernel(env, true, env->d_actions_scratch, nullptr, d_obs, env->d_rewards_scratch,
                               env->d_done_scratch, s);
    if (rc != AT_OK) return rc;
    if (env->k.rows > 0)
        CUDA_TRY(cudaMemcpyAsync(h_rewards, env->d_rewards_scratch, (size_t)N * env->k.rows * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h_done, env->d_done_scratch, (size_t)N, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return AT_OK;
}

int at_step_host(at_env *env, const int32_t *h_actions, float *d_obs, float *h_rewards, uint8_t *h_done, void *stream) {
    return step_host_impl(env, h_actions, 4, d_obs, h_rewards, h_done, stream);
}

int at_step_host_u8(at_env *env, const uint8_t *h_actions, float *d_obs, float *h_rewards, uint8_t *h_done, void *stream) {
    return step_host_impl(env, h_actions, 1, d_obs, h_rewards, h_done, stream);
}

int at_observe(at_env *env, float *d_obs, void *stream) {
    if (!env || !d_obs) return fail(AT_ERR_ARG, "at_observe: null handle or buffer");
    if (env->k.rows == 0) return AT_OK;
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    return launch_rows_kernel(env, nullptr, d_obs, (cudaStream_t)stream, nullptr, 0.f, false);
}

int at_set_reward_mode(at_env *env, int mode) {
    if (!env || (mode != AT_REWARD_CUMULATIVE && mode != AT_REWARD_DELTA))
        return fail(AT_ERR_ARG, "at_set_reward_mode: null handle or unknown mode");
    if (mode == AT_REWARD_DELTA && !env->k.reward_delta) {   // the increments start from the values as they are now
        DeviceGuard guard(env->device);
        CUDA_TRY(guard.status());
        const int n = env->k.n_tanks;
        for (int r = 0; r < env->k.rows; ++r) {   // prev_rew[r] = (float)trew[tank of row r]: one small kernel per row
            const int t = env->k.team_layout ? env->k.agent_tank[r] : r;
            (void)n;
            prev_rew_kernel<<<(int)((env->n_envs + 255) / 256), 256>>>(env->st, r, t);
        }
        CUDA_TRY(cudaDeviceSynchronize());
    }
    env->k.reward_delta = mode == AT_REWARD_DELTA;
    return AT_OK;
}

int at_forget_host_buffers(at_env *env) {
    if (!env) return fail(AT_ERR_ARG, "at_forget_host_buffers: null handle");
    env->mapped.clear();
    return AT_OK;
}

int at_get_tank_columns(at_env *env, int tank, double *d_out, void *stream) {
    if (!env || !d_out || tank < 0 || tank >= env->k.n_tanks) return fail(AT_ERR_ARG, "at_get_tank_columns: bad argument");
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    tank_columns_kernel<<<(int)((env->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(env->st, tank, d_out);
    env->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_get_info(at_env *env, int32_t *d_winner, int32_t *d_hits, int32_t *d_be_hits, int32_t *d_change_time,
                void *stream) {
    if (!env) return fail(AT_ERR_ARG, "at_get_info: null handle");
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    const int grid = (int)((env->n_envs + 127) / 128);
    info_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(env->k, env->st, d_winner, d_hits, d_be_hits, d_change_time, nullptr);
    env->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_get_terminal_info(at_env *env, int32_t *d_winner, int32_t *d_hits, int32_t *d_be_hits, int32_t *d_change_time,
                         uint8_t *d_latched, void *stream) {
    if (!env) return fail(AT_ERR_ARG, "at_get_terminal_info: null handle");
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    const int grid = (int)((env->n_envs + 127) / 128);
    info_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(env->k, env->st, d_winner, d_hits, d_be_hits, d_change_time, d_latched);
    env->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_synthetic_actions(at_env *env, int64_t tick, int32_t *d_actions, void *stream) {
    if (!env || !d_actions) return fail(AT_ERR_ARG, "at_synthetic_actions: null argument");
    if (env->k.act_rows == 0) return AT_OK;
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    const int64_t total = env->n_envs * env->k.act_rows;
    synthetic_actions_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        env->k.seed, env->k.env_id_base, env->n_envs, env->k.act_rows, (uint32_t)tick, d_actions);
    env->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_get_state(at_env *env, int64_t first, int64_t count, at_env_state *h_out) {
    if (!env || !h_out || first < 0 || count <= 0 || first + count > env->n_envs)
        return fail(AT_ERR_ARG, "at_get_state: bad range");
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    PackedEnv *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(PackedEnv) * count));
    gather_kernel<<<(int)((count + 63) / 64), 64>>>(env->k, env->st, first, count, d);
    env->launches += 1;
    std::vector<PackedEnv> h((size_t)count);
    cudaError_t err = cudaMemcpy(h.data(), d, sizeof(PackedEnv) * count, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (err != cudaSuccess) return fail(AT_ERR_CUDA, std::string("at_get_state: ") + cudaGetErrorString(err));
    for (int64_t q = 0; q < count; ++q) unpack_env(env->k, env->luts, h[(size_t)q], h_out[q]);
    return AT_OK;
}

int at_set_state(at_env *env, int64_t first, int64_t count, const at_env_state *h_in) {
    if (!env || !h_in || first < 0 || count <= 0 || first + count > env->n_envs)
        return fail(AT_ERR_ARG, "at_set_state: bad range");
    std::vector<PackedEnv> h((size_t)count);
    for (int64_t q = 0; q < count; ++q) {
        std::string msg = pack_env(env->k, env->luts, h_in[q], h[(size_t)q]);
        if (!msg.empty()) return fail(AT_ERR_ARG, msg);
    }
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    PackedEnv *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(PackedEnv) * count));
    cudaError_t err = cudaMemcpy(d, h.data(), sizeof(PackedEnv) * count, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) {
        scatter_kernel<<<(int)((count + 63) / 64), 64>>>(env->k, env->st, first, count, d);
        env->launches += 1;
        err = cudaDeviceSynchronize();
    }
    cudaFree(d);
    if (err != cudaSuccess) return fail(AT_ERR_CUDA, std::string("at_set_state: ") + cudaGetErrorString(err));
    return AT_OK;
}

int at_bfs_batch(int device, const uint8_t *d_maze, const int32_t *d_query, int64_t n, int32_t *d_out, void *stream) {
    if (!d_maze || !d_query || !d_out || n <= 0) return fail(AT_ERR_ARG, "at_bfs_batch: bad argument");
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    bfs_batch_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_maze, d_query, n, d_out);
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}
int at_generate_mazes(int device, uint64_t seed, int64_t env_id_base, int64_t n, uint8_t *d_maze, uint32_t *d_ctr,
                      void *stream) {
    if (!d_maze || n <= 0) return fail(AT_ERR_ARG, "at_generate_mazes: bad argument");
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    maze_kernel<<<(int)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(seed, env_id_base, n, d_maze, d_ctr);
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_set_weakness(at_env *env, double weakness) {
    if (!env || !(weakness >= 0.0 && weakness <= 1.0)) return fail(AT_ERR_ARG, "at_set_weakness: null handle or weakness outside [0, 1]");
    // The kernels read the weakness from device memory (DevState::dyn), not from the by-value kernel config, so the
    // change also reaches steps that replay a captured CUDA graph.  Synchronous copy: it has landed when the call
    // returns, i.e. before any step enqueued afterwards.
    DeviceGuard guard(env->device);
    CUDA_TRY(guard.status());
    CUDA_TRY(cudaMemcpy(env->d_dyn, &weakness, 8, cudaMemcpyHostToDevice));
    env->user_cfg.weakness = weakness;
    env->k.weakness = weakness;
    return AT_OK;
}

int at_tick_normalize(int device, const float *d_obs, int64_t n_envs, int rows, int dim, float epsilon, float *d_out,
                      void *stream) {
    if (!d_obs || !d_out || n_envs <= 0 || rows <= 0 || rows > AT_MAX_TANKS || dim <= 0)
        return fail(AT_ERR_ARG, "at_tick_normalize: bad argument");
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    const bool vec = dim % 4 == 0 && ((((uintptr_t)d_obs) | ((uintptr_t)d_out)) & 15u) == 0;
    const int64_t threads = n_envs * (vec ? dim / 4 : dim);
    const int un = !vec || rows > 4 ? 1 : (rows <= 2 ? 4 : 2);   // tick_normalize_kernel: UN chunks per thread
    const int grid = (int)((threads + 256 * un - 1) / (256 * un));
    cudaStream_t s = (cudaStream_t)stream;
    if (vec && rows == 2) tick_normalize_kernel<4, 2><<<grid, 256, 0, s>>>(d_obs, d_out, n_envs, rows, dim, epsilon);
    else if (vec && rows == 1) tick_normalize_kernel<4, 1><<<grid, 256, 0, s>>>(d_obs, d_out, n_envs, rows, dim, epsilon);
    else if (vec && rows == 3) tick_normalize_kernel<4, 3><<<grid, 256, 0, s>>>(d_obs, d_out, n_envs, rows, dim, epsilon);
    else if (vec && rows == 4) tick_normalize_kernel<4, 4><<<grid, 256, 0, s>>>(d_obs, d_out, n_envs, rows, dim, epsilon);
    else if (vec) tick_normalize_kernel<4, 0><<<grid, 256, 0, s>>>(d_obs, d_out, n_envs, rows, dim, epsilon);
    else tick_normalize_kernel<1, 0><<<grid, 256, 0, s>>>(d_obs, d_out, n_envs, rows, dim, epsilon);
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_sample_heads(int device, const float *const *d_logits, const int *dims, int n_heads, int64_t n_rows, uint64_t seed,
                    const int64_t *d_counter, uint32_t stream_id, int32_t *d_actions, int64_t act_stride,
                    float *d_logprobs, int64_t lp_stride, void *stream) {
    if (!d_logits || !dims || n_heads < 1 || n_heads > 4 || n_rows <= 0 || !d_actions || !d_logprobs)
        return fail(AT_ERR_ARG, "at_sample_heads: bad argument");
    SampleHeads h;
    memset(&h, 0, sizeof h);
    h.n_heads = n_heads;
    int total = 0;
    for (int k = 0; k < n_heads; ++k) {
        if (!d_logits[k] || dims[k] < 1 || dims[k] > 4) return fail(AT_ERR_ARG, "at_sample_heads: a head needs 1..4 logits");
        h.logits[k] = d_logits[k];
        h.dims[k] = dims[k];
        total += dims[k];
    }
    if (total > 8) return fail(AT_ERR_ARG, "at_sample_heads: more than 8 logits per row");
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    sample_heads_kernel<<<(int)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h, n_rows, seed, d_counter, stream_id,
                                                                                   d_actions, act_stride, d_logprobs, lp_stride);
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_policy_tail(int device, const float *const *d_z2, const float *d_w3, const float *d_b3, int64_t n_rows, uint64_t seed,
                   const int64_t *d_counter, uint32_t stream_id, int32_t *d_actions, int64_t act_stride, float *d_logprobs,
                   int64_t lp_stride, float *d_values, int64_t val_stride, void *stream) {
    if (!d_z2 || !d_w3 || !d_b3 || n_rows <= 0 || !d_actions || !d_logprobs || !d_values)
        return fail(AT_ERR_ARG, "at_policy_tail: bad argument");
    PolicyTail p;
    for (int k = 0; k < 4; ++k) {
        if (!d_z2[k] || (((uintptr_t)d_z2[k]) & 15u)) return fail(AT_ERR_ARG, "at_policy_tail: layer-two buffers must be 16-byte aligned");
        p.z2[k] = d_z2[k];
    }
    if (((uintptr_t)d_w3) & 15u) return fail(AT_ERR_ARG, "at_policy_tail: the output-layer matrix must be 16-byte aligned");
    p.w3 = d_w3;
    p.b3 = d_b3;
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    int64_t blocks = (n_rows * 16 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    policy_tail_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, n_rows, seed, d_counter, stream_id, d_actions, act_stride,
                                                                   d_logprobs, lp_stride, d_values, val_stride);
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_policy_mid_tail(int device, const float *d_z1, const float *d_w2, const float *d_b2, const float *d_w3, const float *d_b3,
                       float *d_out9, int64_t n_rows, uint64_t seed, const int64_t *d_counter, uint32_t stream_id,
                       int32_t *d_actions, int64_t act_stride, float *d_logprobs, int64_t lp_stride, float *d_values,
                       int64_t val_stride, void *stream) {
    if (!d_z1 || !d_w2 || !d_b2 || !d_w3 || !d_b3 || !d_out9 || n_rows <= 0 || !d_actions || !d_logprobs || !d_values)
        return fail(AT_ERR_ARG, "at_policy_mid_tail: bad argument");
    if ((((uintptr_t)d_z1) | ((uintptr_t)d_w2)) & 15u) return fail(AT_ERR_ARG, "at_policy_mid_tail: z1 / w2 must be 16-byte aligned");
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    static bool attr_set[64] = {false};
    const size_t smem = (size_t)(128 + 2 * PM_ROWS) * PM_LD * 4 + (128 + 3 * 128) * 4;
    if (device < 64 && !attr_set[device]) {
        CUDA_TRY(cudaFuncSetAttribute(policy_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[device] = true;
    }
    PolicyMid p = {d_z1, d_w2, d_b2, d_w3, d_b3, d_out9};
    const int64_t n_tiles = (n_rows + PM_ROWS - 1) / PM_ROWS;
    int gx = 37;                       // 37 x 4 networks = 148 blocks = one per SM
    if (n_tiles < gx) gx = (int)n_tiles;
    cudaStream_t s = (cudaStream_t)stream;
    policy_mid_kernel<<<dim3(gx, 4), PM_THREADS, smem, s>>>(p, n_rows);
    CUDA_TRY(cudaGetLastError());
    sample_out9_kernel<<<(int)((n_rows + 255) / 256), 256, 0, s>>>(d_out9, n_rows, seed, d_counter, stream_id, d_actions, act_stride,
                                                                  d_logprobs, lp_stride, d_values, val_stride);
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_policy_first(int device, const float *d_x, int64_t x_row_stride, int64_t n_rows, int dim, const float *d_w1, const float *d_b1,
                    float *d_z1, void *stream) {
    if (!d_x || !d_w1 || !d_b1 || !d_z1 || n_rows <= 0 || dim <= 0 || dim % 4 != 0 || x_row_stride % 4 != 0 || x_row_stride < dim)
        return fail(AT_ERR_ARG, "at_policy_first: bad argument (dim and the row stride must be multiples of 4 floats)");
    if ((((uintptr_t)d_x) | ((uintptr_t)d_w1) | ((uintptr_t)d_z1) | ((uintptr_t)d_b1)) & 15u)
        return fail(AT_ERR_ARG, "at_policy_first: buffers must be 16-byte aligned");
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    CUtensorMap mx, mw;
    if (!at_tc::make_map_2d(&mx, d_x, (uint64_t)n_rows, (uint64_t)dim, (uint64_t)x_row_stride * 4, at_tc::TM) ||
        !at_tc::make_map_2d(&mw, d_w1, (uint64_t)at_tc::TN, (uint64_t)dim, (uint64_t)dim * 4, 256))
        return fail(AT_ERR_CUDA, "at_policy_first: cuTensorMapEncodeTiled is not available or rejected the tensors");
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(at_tc::policy_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_tc::SMEM_BYTES));
        attr_set = true;
    }
    const int grid = (int)((n_rows + at_tc::TM - 1) / at_tc::TM), kc = (dim + at_tc::KC - 1) / at_tc::KC;
    at_tc::policy_first_kernel<<<grid, 128, at_tc::SMEM_BYTES, (cudaStream_t)stream>>>(mx, mw, d_b1, d_z1, (int)n_rows, kc);
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

static unsigned long long *g_policy_dbg = nullptr;   // at_policy_debug_stamps

static int policy_forward_impl(bool half, int device, const at_policy *pol, int n_policies, int64_t n_rows, int dim, uint32_t chunk_mask,
                               uint64_t seed, int64_t *d_counter, int advance_counter, void *stream) {
    if (!pol || n_policies < 1 || n_policies > 2 || n_rows <= 0 || n_rows > (int64_t)1 << 30 || dim <= 0 || dim % (half ? 8 : 4) != 0 || dim > (half ? 2048 : 1024))
        return fail(AT_ERR_ARG, "at_policy_forward: 1 or 2 policies, dim a multiple of 4 up to 1024 (fp16: of 8 up to 2048)");
    const int kc = half ? at_tc::H_KC : at_tc::KC, k_chunks = (dim + kc - 1) / kc;
    const uint32_t all = k_chunks >= 32 ? 0xFFFFFFFFu : ((1u << k_chunks) - 1u);
    if (chunk_mask == 0) chunk_mask = all;
    if (chunk_mask & ~all) return fail(AT_ERR_ARG, "at_policy_forward: chunk_mask names columns past dim");
    DeviceGuard guard(device);
    CUDA_TRY(guard.status());
    at_tc::PolicyFwd P;
    memset(&P, 0, sizeof P);
    CUtensorMap maps[6];
    for (int i = 0; i < 2; ++i) {
        const at_policy &q = pol[i < n_policies ? i : 0];
        if (!q.x || !q.w1 || !q.b1 || !q.w2 || !q.b2 || !q.w3 || !q.b3 || !q.actions || !q.logprobs || !q.values ||
            q.x_row_stride % (half ? 8 : 4) != 0 || q.x_row_stride < dim)
            return fail(AT_ERR_ARG, "at_policy_forward: null buffer or a row stride that is not a multiple of 4 floats");
        if ((((uintptr_t)q.x) | ((uintptr_t)q.w1) | ((uintptr_t)q.w2)) & 15u)
            return fail(AT_ERR_ARG, "at_policy_forward: x / w1 / w2 must be 16-byte aligned");
        const bool ok = half ? (at_tc::make_map_2d_f16(&maps[i], q.x, (uint64_t)n_rows, (uint64_t)dim, (uint64_t)q.x_row_stride * 2, at_tc::TM) &&
                                at_tc::make_map_2d_f16(&maps[2 + i], q.w1, 512, (uint64_t)dim, (uint64_t)dim * 2, 256) &&
                                at_tc::make_map_2d_f16(&maps[4 + i], q.w2, 512, 128, 128 * 2, 128))
                             : (at_tc::make_map_2d(&maps[i], q.x, (uint64_t)n_rows, (uint64_t)dim, (uint64_t)q.x_row_stride * 4, at_tc::TM) &&
                                at_tc::make_map_2d(&maps[2 + i], q.w1, 512, (uint64_t)dim, (uint64_t)dim * 4, 256) &&
                                at_tc::make_map_2d_f16(&maps[4 + i], q.w2, 512, 128, 128 * 2, 128));   // (W2 is fp16 in both)
        if (!ok)
            return fail(AT_ERR_CUDA, "at_policy_forward: cuTensorMapEncodeTiled is not available or rejected the tensors");
        at_tc::PolicyIO &io = P.io[i];
        io.b1 = q.b1; io.b2 = q.b2; io.w3 = q.w3; io.b3 = q.b3;
        io.actions = q.actions; io.logprobs = q.logprobs; io.values = q.values; io.out9 = q.out9;
        io.act_stride = q.act_stride; io.lp_stride = q.lp_stride; io.val_stride = q.val_stride;
        io.stream_id = q.stream_id;
    }
    P.n_rows = n_rows;
    P.tiles_per_policy = (int)((n_rows + at_tc::TM - 1) / at_tc::TM);
    P.n_policies = n_policies;
    P.chunk_mask = chunk_mask;
    P.seed = seed;
    P.counter = (long long *)d_counter;
    P.advance = advance_counter;
    P.dbg = g_policy_dbg;
    static bool attr_set[64] = {false};
    static int sms[64] = {0};
    if (device < 0 || device >= 64) return fail(AT_ERR_ARG, "at_policy_forward: device index out of range");
    if (!attr_set[device]) {
        CUDA_TRY(cudaFuncSetAttribute(at_tc::policy_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_tc::H_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(at_tc::policy_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_tc::H_SMEM_BYTES));
        CUDA_TRY(cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device));
        attr_set[device] = true;
    }
    const int n_tiles = P.tiles_per_policy * n_policies;
    const int grid = n_tiles < sms[device] ? n_tiles : sms[device];
    if (half)
        at_tc::policy_forward_kernel<true><<<grid, at_tc::H_THREADS, at_tc::H_SMEM_BYTES, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3],
                                                                                                           maps[4], maps[5], P);
    else
        at_tc::policy_forward_kernel<false><<<grid, at_tc::H_THREADS, at_tc::H_SMEM_BYTES, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3],
                                                                                                         maps[4], maps[5], P);
    CUDA_TRY(cudaGetLastError());
    return AT_OK;
}

int at_policy_forward(int device, const at_policy *pol, int n_policies, int64_t n_rows, int dim, uint32_t chunk_mask, uint64_t seed,
                      int64_t *d_counter, int advance_counter, void *stream) {
    return policy_forward_impl(false, device, pol, n_policies, n_rows, dim, chunk_mask, seed, d_counter, advance_counter, stream);
}

int at_policy_forward16(int device, const at_policy *pol, int n_policies, int64_t n_rows, int dim, uint32_t chunk_mask, uint64_t seed,
                        int64_t *d_counter, int advance_counter, void *stream) {
    return policy_forward_impl(true, device, pol, n_policies, n_rows, dim, chunk_mask, seed, d_counter, advance_counter, stream);
}

int at_set_rows_mode(at_env *env, int persistent_buffers) {
    if (!env) return fail(AT_ERR_ARG, "at_set_rows_mode: null handle");
    env->rows_skip_static = persistent_buffers != 0;
    return AT_OK;
}

int at_policy_debug_stamps(unsigned long long *d_stamps) {
    g_policy_dbg = d_stamps;   // device uint64[2][512] or NULL: block 0 of at_policy_forward16 records (event, SM clock) pairs
    return AT_OK;
}

int at_obs_static_range(const at_env *env, int32_t *lo, int32_t *hi) {
    if (!env || !lo || !hi) return fail(AT_ERR_ARG, "at_obs_static_range: null argument");
    // walls + maze columns of a fixed map never change (gym_env_multi.py:283-292); per-env mazes do
    *lo = env->k.off_walls;
    *hi = env->k.map == AT_MAP_MAZE ? env->k.off_walls : env->k.off_zones;
    return AT_OK;
}

int at_debug_block_times(at_env *env, unsigned long long *d_times) {
    if (!env) return fail(AT_ERR_ARG, "at_debug_block_times: null handle");
    env->dbg_times = d_times;   // [2 * ceil(n_envs / 64)] device words, or NULL to switch the stamps off
    return AT_OK;
}

int at_get_luts(const at_env *env, double *h_trig, double *h_corner) {
    if (!env) return fail(AT_ERR_ARG, "at_get_luts: null handle");
    if (h_trig) memcpy(h_trig, env->luts.trig.data(), 722 * 8);
    if (h_corner) memcpy(h_corner, env->luts.corn.data(), 361 * 4 * 8);
    return AT_OK;
}

}  // extern "C"
