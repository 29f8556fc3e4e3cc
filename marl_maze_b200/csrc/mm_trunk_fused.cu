// mm_trunk_fused.cu -- K4 trunk, third generation: the three trunk layers, the heads AND the sampling in ONE persistent kernel.
// Replaces Actor.forward's `layers` + move_head / mark_head (networks.py:36-41) and PPO.get_action (PPO.py:170-186) for the rollout.  sm_100a.
//
// Why: the per-layer 3xFP16 kernel (mm_linear16.cu) is bound by L2 -> SM traffic, not by the tensor pipe (ncu, profiles/r02h: tensor pipe
// 29 %, lts 63 %): every 128 x 136 output tile re-reads its 128 x K fp32 activation tile and its weight half from L2, and the activations
// make an HBM round trip between the layers.  Here one CTA owns a 128-row tile through all layers (x ~ hi + lo, both fp16, three MMAs
// hi.hi + lo.hi + hi.lo per k-step into one fp32 accumulator, as in mm_linear16.cu):
//   layer 0 : x0 tile (fp32, TMA, 128-byte swizzle) -> splitter warps -> fp16 hi / lo in a TENSOR MEMORY ring (TS-form MMAs); weights
//             streamed through an 8-stage ring of [hi | lo] half-N tiles (N = 128 | 144); D = 128 x 272 fp32 in TMEM columns 0..271
//   epilogue: tcgen05.ld -> scale, bias, ReLU -> fp16 hi / lo: the hi plane goes to TENSOR MEMORY (columns 272..407, read by two of the
//             three MMAs of the next layer), the lo plane to SHARED MEMORY in the canonical K-major SWIZZLE_64B layout (9 k-blocks of 32
//             columns, 72 KB): the next layer's A operand never leaves the SM.  (Both planes in shared memory made the SS-form MMAs read
//             ~120 B/clk of shared memory beside the TMA's incoming weights -- the first version of this kernel was bound by that.)
//   layer 1, layer 2 : A = (TMEM hi plane, shared-memory lo plane), weights through the same ring
//   heads   : a fourth, 16-row "layer" ([move_head; mark_head] padded) on the same path; its epilogue reads 6 accumulator columns per
//             row, masks, samples with Philox (or evaluates the given actions) and writes actions + joint log-probs (one shuffle between
//             the two agents of an env): nothing but actions / log-probs is stored.
// Per 128-row tile the SM reads 235 KB of activations and 1.15 MB of weights from L2.  One CTA per SM (217 KB of shared memory, 512 TMEM
// columns), persistent over tiles.  Warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = splitter (layer 0) and
// epilogue.  The fp32 landing slots of layer 0 alias the lo plane (dead while layer 0's main loop runs).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include "mm_env.cuh"
#include "mm_policy_heads.cuh"
#include "mm_tc.cuh"

namespace mm {

constexpr int TF_BM = 128, TF_BK = 32, TF_N = 264, TF_N0 = 128, TF_N1 = 144, TF_NH = 16;
#ifndef MM_TF_WSTAGES
#define MM_TF_WSTAGES 8
#endif
// MM_TF_CLUSTER = 2: the CTAs of a cluster pair walk the same weight stream (different row tiles), each loads HALF of every weight tile and
// multicasts it to both (cp.async.bulk.tensor ... .multicast::cluster): the L2 -> SM weight traffic per SM halves.  Measured need: with one CTA
// per cluster the main loops ran at 27.5 B/clk/SM of L2 reads in every layer (7.5 TB/s over the chip, the L2's limit) and waited for weights.
#ifndef MM_TF_CLUSTER
#define MM_TF_CLUSTER 2
#endif
constexpr int TF_CLUSTER = MM_TF_CLUSTER;
static_assert(TF_CLUSTER == 1 || TF_CLUSTER == 2, "cluster of one or two CTAs");
// MM_TF_ADEDICATED = 1: the fp32 landing slots of layer 0 get shared memory of their own (taken from the weight ring: build with MM_TF_WSTAGES 6 and
// MM_TF_ASLOTS 2, or 7 and 1) instead of aliasing the lo plane, so the activation producer loads the NEXT tile's first blocks while this tile's layers
// 1-3 run instead of waiting for the head MMAs (5.2 of 48 kclk per tile were that wait).
#ifndef MM_TF_STAGGER_CLK
#define MM_TF_STAGGER_CLK 0
#endif
#ifndef MM_TF_ADEDICATED
#define MM_TF_ADEDICATED 0
#endif
#ifndef MM_TF_ASLOTS
#define MM_TF_ASLOTS 4
#endif
constexpr int TF_WSTAGES = MM_TF_WSTAGES, TF_ASLOTS = MM_TF_ASLOTS, TF_OPS = 3;
constexpr uint32_t TF_WROW = 64;                                  // fp16 weight rows of one k-block: 32 x 2 bytes, SWIZZLE_64B
constexpr uint32_t TF_WHALF = TF_N1 * TF_WROW;                    // 9216: one of (hi, lo) of a stage, sized for the wider half
constexpr uint32_t TF_W_BYTES = 2 * TF_WHALF;                     // 18432
constexpr uint32_t TF_H_OFF = TF_WSTAGES * TF_W_BYTES;            // 147456 (a multiple of 1024)
constexpr int TF_HKB = 9;                                         // 264 columns -> 9 k-blocks of 32 (288, columns 264.. are zero)
constexpr uint32_t TF_H_KB_BYTES = TF_BM * 64;                    // 8192
constexpr uint32_t TF_H_PLANE = TF_HKB * TF_H_KB_BYTES;           // 73728: the lo plane
constexpr uint32_t TF_A_BYTES = TF_BM * TF_BK * 4;                // 16384: fp32 landing slot (layer 0), 128-byte rows, SWIZZLE_128B
static_assert(MM_TF_ADEDICATED || TF_ASLOTS * TF_A_BYTES <= TF_H_PLANE, "landing slots alias the lo plane");
static_assert(TF_H_OFF % 1024 == 0, "swizzle atoms need 1024-byte alignment");
constexpr uint32_t TF_A_OFF = MM_TF_ADEDICATED ? TF_H_OFF + TF_H_PLANE : TF_H_OFF;   // landing slots: behind the lo plane, or aliasing it
constexpr uint32_t TF_RING_BYTES = TF_H_OFF + TF_H_PLANE + (MM_TF_ADEDICATED ? TF_ASLOTS * TF_A_BYTES : 0u);   // 221184
constexpr uint32_t TF_BAR_OFF = TF_RING_BYTES;
constexpr uint32_t TF_BIAS_OFF = TF_BAR_OFF + 512;
static_assert((2 * TF_WSTAGES + 2 * TF_ASLOTS + 2 * TF_OPS + 3) * 8 + 4 <= 512, "barriers fit in front of the biases");
constexpr int TF_BIAS_FLOATS = 3 * TF_N + 8;                      // three trunk biases + the six head biases
constexpr uint32_t TF_SMEM_BYTES = TF_BIAS_OFF + TF_BIAS_FLOATS * 4 + 1024 /*alignment slack*/;
static_assert(TF_SMEM_BYTES <= 232448, "one CTA per SM");
// splitter / epilogue warps: 2 or 4 per TMEM lane quarter.  The three epilogues are serial with the MMAs (the next layer needs the whole operand and reuses
// the accumulator columns): with 8 warps an epilogue took ~3.1 kclk of ~48 kclk per tile, bound by its own latencies (two warps per scheduler)
#ifndef MM_TF_EPI_WARPS
#define MM_TF_EPI_WARPS 8
#endif
constexpr int TF_EPI_WARPS = MM_TF_EPI_WARPS, TF_WPQ = TF_EPI_WARPS / 4, TF_THREADS = 64 + 32 * TF_EPI_WARPS + 32;   // + the activation-tile producer warp
static_assert(TF_EPI_WARPS == 8 || TF_EPI_WARPS == 16, "two or four epilogue warps per lane quarter");
// TMEM: D columns 0..271 | hi plane of the activations 272..407 (136 columns = 272 fp16 k-elements) | layer-0 operand ring 408 + 32 * stage
constexpr uint32_t TF_TMEM_COLS = 512, TF_TMEM_H_COL = 272, TF_TMEM_A_COL = 408;
static_assert(TF_TMEM_A_COL + 32 * TF_OPS <= TF_TMEM_COLS, "operand ring fits");

__device__ __forceinline__ uint64_t tf_desc(uint32_t saddr) {   // K-major SWIZZLE_64B: 8-row atoms of 512 bytes (UMMA layout type 4)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8 * TF_WROW) >> 4) << 32) | ((uint64_t)1 << 46) | (4ull << 61);
}
__host__ __device__ constexpr uint32_t tf_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void tf_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tf_umma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tf_tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                 "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tf_tmem_st4(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tf_tmem_ld32_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
          "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tf_tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// this CTA's part of a weight tile, delivered to the same shared-memory offset (and signalled on the same barrier offset) in every CTA of the mask
__device__ __forceinline__ void tf_tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
                 : "memory");
}
// ask the L2 for a box of a tensor map ahead of its use (no shared memory, no completion to wait for)
__device__ __forceinline__ void tf_tma_prefetch_l2(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tf_commit_mc(uint64_t* bar, uint16_t mask) {   // arrive on the barrier at this offset in every CTA of the mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t tf_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tf_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tf_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// -DMM_TF_PROFILE: per-CTA cycle counters of where each role waits (written to args.prof [blocks][16] u64); tools/trunk_profile.py
#ifdef MM_TF_PROFILE
#define TF_PROF_DECL long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt0
#define TF_PROF_WAIT(i, stmt) do { pt0 = clock64(); stmt; pf[i] += clock64() - pt0; } while (0)
#define TF_PROF_T0() (pt0 = clock64())
#define TF_PROF_ADD(i) (pf[i] += clock64() - pt0)
#else
#define TF_PROF_DECL
#define TF_PROF_WAIT(i, stmt) stmt
#define TF_PROF_T0()
#define TF_PROF_ADD(i)
#endif

struct TFMaps {
    CUtensorMap a;               // x0 [M][460] fp32, {32 x 128} boxes, 128-byte swizzle
    CUtensorMap w[3][2][2];      // [layer][column half][hi, lo]: fp16 [264][kpad], {32 x (128|144) / TF_CLUSTER} boxes, 64-byte swizzle
    CUtensorMap wh[2];           // heads [hi, lo]: fp16 [16][288], {32 x 16 / TF_CLUSTER} boxes
};
struct TFArgs {
    const float* bias[4];        // three trunk biases [264], head biases [6]
    const float* acc_scale[4];   // device scalars 2^-e (the weights are stored as 2^e W)
    HeadArgs heads;
    int M, n_tiles;
    unsigned long long* prof;    // MM_TF_PROFILE builds only
};

// fp32 -> (fp16 hi, fp16 lo) pairs, both round-to-nearest; element 2j in the low half of the 32-bit word
__device__ __forceinline__ void tf_split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

constexpr int kTfK0 = 460, kTfNKB0 = (kTfK0 + TF_BK - 1) / TF_BK;   // 15 k-blocks in layer 0

// one 32-column chunk of a trunk layer's accumulator -> relu(acc * scale + bias) -> hi to the TMEM plane, lo to the shared-memory plane
__device__ __forceinline__ void tf_epilogue_chunk(const uint32_t* v, int c, float scale, const float* s_bias, uint32_t t_lane, uint8_t* hlo, int arow) {
    uint32_t hi[16];
    uint8_t* dst_lo = hlo + c * TF_H_KB_BYTES + arow * 64;
    const int col0 = c * 32;
#pragma unroll
    for (int q = 0; q < 4; q++) {   // 8 columns = one 16-byte chunk of the 64-byte row
        uint32_t l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = col0 + 8 * q + 2 * j;
            float y0 = 0.f, y1 = 0.f;
            if (col < TF_N) {   // TF_N is even: a pair is in or out
                const float2 b2 = *reinterpret_cast<const float2*>(s_bias + col);   // same address in every lane: broadcast
                y0 = fmaxf(fmaf(__uint_as_float(v[8 * q + 2 * j]), scale, b2.x), 0.f);
                y1 = fmaxf(fmaf(__uint_as_float(v[8 * q + 2 * j + 1]), scale, b2.y), 0.f);
            }
            tf_split2(y0, y1, hi[4 * q + j], l[j]);
        }
        const int pq = (q ^ ((arow >> 1) & 3)) << 4;   // SWIZZLE_64B: 16-byte chunk index ^= bits 7-8 of the byte offset
        *reinterpret_cast<uint4*>(dst_lo + pq) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    const uint32_t th = t_lane + TF_TMEM_H_COL + (uint32_t)(c * 16);
    tf_tmem_st8(th, hi);
    if (c < TF_HKB - 1) tf_tmem_st8(th + 8, hi + 8);   // the last chunk holds 8 real columns: its second K = 16 step is never read
}

__global__ void __cluster_dims__(TF_CLUSTER, 1, 1) __launch_bounds__(TF_THREADS, 1) k_trunk_fused(const __grid_constant__ TFMaps maps, const TFArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* hlo = smem + TF_H_OFF;
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + TF_BAR_OFF);
    uint64_t* w_empty = w_full + TF_WSTAGES;
    uint64_t* a_full = w_empty + TF_WSTAGES;
    uint64_t* a_free = a_full + TF_ASLOTS;
    uint64_t* op_full = a_free + TF_ASLOTS;      // splitter -> MMA: operand columns of this ring stage written
    uint64_t* op_free = op_full + TF_OPS;        // MMA -> splitter: the MMAs that read this ring stage have completed
    uint64_t* d_full = op_free + TF_OPS;         // MMA -> epilogue (and producer): all MMAs of a layer have completed (4 per tile)
    uint64_t* h_ready = d_full + 1;              // epilogue -> MMA: D has been read and the next layer's A operand is written (3 per tile)
    uint64_t* d_free = h_ready + 1;              // heads epilogue -> MMA: D has been read, the next tile may overwrite it (1 per tile)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_free + 1);
    float* s_bias = reinterpret_cast<float*>(smem + TF_BIAS_OFF);   // [3][264] + [8]

    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TF_WSTAGES; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], TF_CLUSTER); }   // every CTA of the cluster frees a stage
        for (int s = 0; s < TF_ASLOTS; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_free[s], 32 * TF_EPI_WARPS); }
        for (int s = 0; s < TF_OPS; s++) { mbar_init(&op_full[s], 32 * TF_EPI_WARPS); mbar_init(&op_free[s], 1); }
        mbar_init(d_full, 1); mbar_init(h_ready, 32 * TF_EPI_WARPS); mbar_init(d_free, 32 * TF_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TF_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 3 * TF_N; i += TF_THREADS) s_bias[i] = args.bias[i / TF_N][i % TF_N];
    if (threadIdx.x < 8) s_bias[3 * TF_N + threadIdx.x] = threadIdx.x < 6 ? args.bias[3][threadIdx.x] : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    // The CTAs of a cluster execute the same sequence of weight stages: tile pair p = (cluster index + i * clusters), tile = TF_CLUSTER * p + rank.  A CTA
    // whose tile lies beyond the last one still runs (TMA zero-fills its rows, nothing is written): the weight ring is a shared protocol.
    const uint32_t crank = TF_CLUSTER > 1 ? tf_cluster_rank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << TF_CLUSTER) - 1u);
    const int pair0 = blockIdx.x / TF_CLUSTER, pair_stride = gridDim.x / TF_CLUSTER;
    const int n_pairs = (args.n_tiles + TF_CLUSTER - 1) / TF_CLUSTER;
    if (TF_CLUSTER > 1) tf_cluster_sync();   // the peer's barriers are initialised before anything of ours can land on them
#if MM_TF_STAGGER_CLK > 0
    // All CTAs walk the same layer sequence in lockstep, so the chip's L2 sees every layer-0 phase (weights + the fp32 activation stream: ~31 B/clk/SM)
    // at once and every layer-1/2 phase (20 B/clk/SM) at once.  Every other cluster starts MM_TF_STAGGER_CLK clocks late to interleave the phases.
    if ((blockIdx.x / TF_CLUSTER) & 1) { const long long t0 = clock64(); while (clock64() - t0 < (long long)MM_TF_STAGGER_CLK) { } }
#endif

    if (warp == 0) {
        {  // ===== weight producer (whole warp loops, one elected lane issues)
            uint32_t wit = 0, t_local = 0;
            TF_PROF_DECL;
            for (int pair = pair0; pair < n_pairs; pair += pair_stride, t_local++) {
                // the weight stream does not depend on the tile: this warp runs ahead across layer and tile boundaries, bounded only by the ring
                for (int layer = 0; layer < 4; layer++) {
                    const int nkb = layer == 0 ? kTfNKB0 : TF_HKB;
                    for (int kb = 0; kb < nkb; kb++) {
                        const int k0 = kb * TF_BK;
                        for (int half = 0; half < (layer == 3 ? 1 : 2); half++) {
                            const int s = wit % TF_WSTAGES;
                            TF_PROF_WAIT(2, mbar_wait(&w_empty[s], ((wit / TF_WSTAGES) & 1) ^ 1));
                            uint8_t* st = smem + s * TF_W_BYTES;
                            const int n_rows = layer == 3 ? TF_NH : half ? TF_N1 : TF_N0;
                            const CUtensorMap* mh = layer == 3 ? &maps.wh[0] : &maps.w[layer][half][0];
                            const CUtensorMap* ml = layer == 3 ? &maps.wh[1] : &maps.w[layer][half][1];
                            const int n0 = layer == 3 ? 0 : half * TF_N0;
                            if (elect_one()) {
                                mbar_expect_tx(&w_full[s], 2u * (uint32_t)n_rows * TF_WROW);   // the whole tile: this CTA's part + the peer's
                                if (TF_CLUSTER > 1) {
                                    const int part = n_rows / TF_CLUSTER, r0 = (int)crank * part;
                                    tf_tma_load_2d_mc(st + r0 * TF_WROW, mh, k0, n0 + r0, &w_full[s], cmask);
                                    tf_tma_load_2d_mc(st + TF_WHALF + r0 * TF_WROW, ml, k0, n0 + r0, &w_full[s], cmask);
                                } else {
                                    tma_load_2d(st, mh, k0, n0, &w_full[s]);
                                    tma_load_2d(st + TF_WHALF, ml, k0, n0, &w_full[s]);
                                }
                            }
                            wit++;
                        }
                    }
                }
            }
#ifdef MM_TF_PROFILE
            if (lane == 0) args.prof[blockIdx.x * 16 + 2] = (unsigned long long)pf[2];
#endif
        }
    } else if (warp == 2 + TF_EPI_WARPS) {
        {  // ===== activation-tile producer (layer 0): its own warp, so that a full landing ring never holds the weight stream back
            uint32_t ait = 0, t_local = 0;
            TF_PROF_DECL;
            for (int pair = pair0; pair < n_pairs; pair += pair_stride, t_local++) {
                const int m0 = (pair * TF_CLUSTER + (int)crank) * TF_BM;
                // The landing slots alias the lo plane: the previous tile's head MMAs (the 4th d_full completion of that tile) must be done.  A parity
                // wait only tells two consecutive phases apart and this warp can be several phases behind: it follows all four completions in order.
                if (!MM_TF_ADEDICATED && t_local)
                    for (int i = 0; i < 4; i++) TF_PROF_WAIT(0, mbar_wait(d_full, (4 * (t_local - 1) + i) & 1));
                for (int kb = 0; kb < kTfNKB0; kb++, ait++) {
                    const int sa = ait % TF_ASLOTS;
                    TF_PROF_WAIT(1, mbar_wait(&a_free[sa], ((ait / TF_ASLOTS) & 1) ^ 1));
                    if (elect_one()) {
                        mbar_expect_tx(&a_full[sa], TF_A_BYTES);
                        tma_load_2d(smem + TF_A_OFF + sa * TF_A_BYTES, &maps.a, kb * TF_BK, m0, &a_full[sa]);
                    }
                }
                // The next tile's activation blocks can only be LOADED once this tile's head MMAs are done (the landing slots alias the lo plane), which
                // exposes their HBM latency at every tile start: ask the L2 for them now, while layers 1-2 of this tile run.
                if (pair + pair_stride < n_pairs && elect_one()) {
                    const int m1 = ((pair + pair_stride) * TF_CLUSTER + (int)crank) * TF_BM;
                    for (int kb = 0; kb < kTfNKB0; kb++) tf_tma_prefetch_l2(&maps.a, kb * TF_BK, m1);
                }
            }
#ifdef MM_TF_PROFILE
            if (lane == 0) for (int i = 0; i < 2; i++) args.prof[blockIdx.x * 16 + i] = (unsigned long long)pf[i];
#endif
        }
    } else if (warp == 1) {
        {  // ===== MMA issuer (whole warp loops, one elected lane issues)
            const uint32_t idesc0 = tf_idesc(TF_BM, TF_N0), idesc1 = tf_idesc(TF_BM, TF_N1), idesch = tf_idesc(TF_BM, TF_NH);
            uint32_t wit = 0, oit = 0, hcnt = 0, t_local = 0;
            TF_PROF_DECL;
#ifdef MM_TF_PROFILE
            const long long t_start = clock64();
#endif
            for (int pair = pair0; pair < n_pairs; pair += pair_stride, t_local++) {
                if (t_local) { TF_PROF_WAIT(0, mbar_wait(d_free, (t_local - 1) & 1)); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                for (int layer = 0; layer < 4; layer++) {
                    const int nkb = layer == 0 ? kTfNKB0 : TF_HKB;
                    const int ktot = layer == 0 ? kTfK0 : TF_N;
                    if (layer) {  // the previous layer's epilogue has drained D and written the A operand (TMEM hi plane, shared-memory lo plane)
                        TF_PROF_WAIT(1, mbar_wait(h_ready, hcnt & 1)); hcnt++;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int kb = 0; kb < nkb; kb++) {
                        const int os = oit % TF_OPS;
                        if (layer == 0) {
                            TF_PROF_WAIT(2, mbar_wait(&op_full[os], (oit / TF_OPS) & 1));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        const int ksteps = (ktot - kb * TF_BK > 16) ? 2 : 1;   // the last k-block may hold only one K = 16 step of real columns
                        // Consecutive MMAs into the SAME accumulator columns serialise on the accumulator's read-modify-write latency (measured: ~150 clocks
                        // per N = 128 | 144 MMA issued back to back into one accumulator, 64 | 72 nominal), so the two column halves -- independent
                        // accumulators -- are issued alternately; the heads (N = 16) rotate over four accumulators that the epilogue adds up.
                        const int nh = layer == 3 ? 1 : 2;
                        uint32_t st[2];
                        for (int half = 0; half < nh; half++) {
                            const int s = (wit + half) % TF_WSTAGES;
                            TF_PROF_WAIT(3 + (layer ? 1 : 0), mbar_wait(&w_full[s], ((wit + half) / TF_WSTAGES) & 1));
                            st[half] = smem_u32(smem + s * TF_W_BYTES);
                        }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        for (int k = 0; k < ksteps; k++) {
                            const uint64_t o = (uint64_t)(k * 2);   // 32 bytes inside the 64-byte swizzle row
                            uint32_t ta_hi, ta_lo = 0;
                            uint64_t a_lo = 0;
                            if (layer == 0) { ta_hi = tmem_base + TF_TMEM_A_COL + (uint32_t)(os * 32 + k * 8); ta_lo = ta_hi + 16; }
                            else { ta_hi = tmem_base + TF_TMEM_H_COL + (uint32_t)(kb * 16 + k * 8); a_lo = tf_desc(smem_u32(hlo + kb * TF_H_KB_BYTES)) + o; }
                            const int kstep = kb * 2 + k;
                            if (elect_one())
#pragma unroll
                            for (int prod = 0; prod < 3; prod++) {   // hi.hi, lo.hi, hi.lo
                                for (int half = 0; half < nh; half++) {
                                    const uint64_t b = tf_desc(st[half] + (prod == 2 ? TF_WHALF : 0u)) + o;
                                    uint32_t d, idesc, acc;
                                    if (layer == 3) { d = tmem_base + (uint32_t)((kstep & 3) * TF_NH); idesc = idesch; acc = (kstep < 4 && prod == 0) ? 0u : 1u; }
                                    else { d = tmem_base + (half ? (uint32_t)TF_N0 : 0u); idesc = half ? idesc1 : idesc0; acc = (kstep == 0 && prod == 0) ? 0u : 1u; }
                                    if (prod == 1) { if (layer == 0) tf_umma_ts(d, ta_lo, b, idesc, acc); else tf_umma_ss(d, a_lo, b, idesc, acc); }
                                    else tf_umma_ts(d, ta_hi, b, idesc, acc);
                                }
                            }
                        }
                        for (int half = 0; half < nh; half++) {
                            const int s = wit % TF_WSTAGES;
                            if (elect_one()) { if (TF_CLUSTER > 1) tf_commit_mc(&w_empty[s], cmask); else umma_commit(&w_empty[s]); }
                            wit++;
                        }
                        if (layer == 0) { if (elect_one()) umma_commit(&op_free[os]); oit++; }
                    }
                    if (elect_one()) umma_commit(d_full);
                }
            }
#ifdef MM_TF_PROFILE
            if (lane == 0) {
                for (int i = 0; i < 5; i++) args.prof[blockIdx.x * 16 + 3 + i] = (unsigned long long)pf[i];
                args.prof[blockIdx.x * 16 + 15] = (unsigned long long)(clock64() - t_start);
            }
#endif
        }
    } else {
        // ===== splitter (layer 0) + epilogue: thread = row (its TMEM lane); the two warps of a lane quarter share the columns
        const int whalf = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int arow = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t ait = 0, oit = 0, dcnt = 0;
        TF_PROF_DECL;
        for (int pair = pair0; pair < n_pairs; pair += pair_stride) {
            const int m0 = (pair * TF_CLUSTER + (int)crank) * TF_BM;
            // ---- layer 0 main loop: fp32 landing slot -> fp16 hi / lo -> tensor memory
            for (int kb = 0; kb < kTfNKB0; kb++, ait++, oit++) {
                const int sa = ait % TF_ASLOTS, os = oit % TF_OPS;
                const float4* rowp = reinterpret_cast<const float4*>(smem + TF_A_OFF + sa * TF_A_BYTES + arow * 128);
                TF_PROF_WAIT(0, mbar_wait(&a_full[sa], (ait / TF_ASLOTS) & 1));
                constexpr int kC = 8 / TF_WPQ;   // float4 chunks of the 32-column block per warp of the quarter
                float4 av[kC];
#pragma unroll
                for (int c = 0; c < kC; c++) av[c] = rowp[(c + kC * whalf) ^ (arow & 7)];
                tf_arrive(&a_free[sa]);
                uint32_t hi[2 * kC], lo[2 * kC];
#pragma unroll
                for (int c = 0; c < kC; c++) {
                    tf_split2(av[c].x, av[c].y, hi[2 * c], lo[2 * c]);
                    tf_split2(av[c].z, av[c].w, hi[2 * c + 1], lo[2 * c + 1]);
                }
                TF_PROF_WAIT(1, mbar_wait(&op_free[os], ((oit / TF_OPS) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ta = t_lane + TF_TMEM_A_COL + (uint32_t)(os * 32);
                if (TF_WPQ == 2) { tf_tmem_st8(ta + 8 * whalf, hi); tf_tmem_st8(ta + 16 + 8 * whalf, lo); }
                else { tf_tmem_st4(ta + 4 * whalf, hi); tf_tmem_st4(ta + 16 + 4 * whalf, lo); }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tf_arrive(&op_full[os]);
            }
            // ---- three trunk epilogues: warp half 0 takes chunks 0..4, half 1 chunks 5..8 (32 columns each); loads run one chunk ahead
            for (int layer = 0; layer < 3; layer++, dcnt++) {
                TF_PROF_WAIT(2 + (layer ? 1 : 0), mbar_wait(d_full, dcnt & 1));
                TF_PROF_T0();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const float scale = __ldg(args.acc_scale[layer]);
                const float* sb = s_bias + layer * TF_N;
                // the nine 32-column chunks over the quarter's warps: 5 + 4, or 3 + 2 + 2 + 2
                const int c_lo = TF_WPQ == 2 ? (whalf ? 5 : 0) : (whalf ? 1 + 2 * whalf : 0), n_c = TF_WPQ == 2 ? (whalf ? 4 : 5) : (whalf ? 2 : 3);
                uint32_t v0[32], v1[32];
                tf_tmem_ld32_nowait(t_lane + (uint32_t)(c_lo * 32), v0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < (TF_WPQ == 2 ? 5 : 3); i++) {
                    if (i < n_c) {
                        uint32_t* cur = (i & 1) ? v1 : v0;
                        uint32_t* nxt = (i & 1) ? v0 : v1;
                        if (i + 1 < n_c) tf_tmem_ld32_nowait(t_lane + (uint32_t)((c_lo + i + 1) * 32), nxt);
                        tf_epilogue_chunk(cur, c_lo + i, scale, sb, t_lane, hlo, arow);
                        if (i + 1 < n_c) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's reads
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tf_arrive(h_ready);
                TF_PROF_ADD(4);
            }
            // ---- heads: 6 logits per row from accumulator columns 0..5, mask / sample / evaluate (PPO.get_action, PPO.py:170-186)
            TF_PROF_WAIT(3, mbar_wait(d_full, dcnt & 1)); dcnt++;
            TF_PROF_T0();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (!whalf) {
                uint32_t v[8], w[8];   // four partial accumulators (k-steps 0, 4, 8, .. | 1, 5, .. | 2, 6, .. | 3, 7, ..), 16 columns apart
                tf_tmem_ld8(t_lane, v);
#pragma unroll
                for (int a = 1; a < 4; a++) {
                    tf_tmem_ld8(t_lane + (uint32_t)(a * TF_NH), w);
#pragma unroll
                    for (int j = 0; j < 6; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tf_arrive(d_free);   // D has been read: the next tile's layer 0 may start
                const float scale = __ldg(args.acc_scale[3]);
                const long long row = (long long)m0 + arow;
                float lp = 0.f;
                if (row < args.M) {
                    float l[6];
#pragma unroll
                    for (int j = 0; j < 6; j++) l[j] = fmaf(__uint_as_float(v[j]), scale, s_bias[3 * TF_N + j]);
                    lp = head_sample_or_eval(l, row, args.heads);
                }
                const float lp_pair = lp + __shfl_xor_sync(0xffffffffu, lp, 1);   // rows 2e, 2e+1 (the two agents of an env) sit in adjacent lanes
                if (row < args.M && (row & 1) == 0) args.heads.logp[row >> 1] = lp_pair;
            } else {
                tf_arrive(d_free);
            }
            TF_PROF_ADD(5);
        }
#ifdef MM_TF_PROFILE
        if (threadIdx.x == 64) for (int i = 0; i < 6; i++) args.prof[blockIdx.x * 16 + 8 + i] = (unsigned long long)pf[i];
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (TF_CLUSTER > 1) tf_cluster_sync();   // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TF_TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
static bool tf_map_f32(CUtensorMap* m, const float* base, int rows, int cols) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TF_BK, (cuuint32_t)TF_BM};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool tf_map_f16(CUtensorMap* m, const void* base, int rows, int kpad, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)TF_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static unsigned long long* g_tf_prof = nullptr;
void trunk_fused_set_profile_buffer(unsigned long long* p) { g_tf_prof = p; }   // [148][16] u64; only read by MM_TF_PROFILE builds

// x0 [M][460] fp32 -> actions / joint log-probs.  w16[l][0|1] = fp16 hi / lo of 2^e W_l: [264][kpad_l] (kpad = 480, 288, 288) for the trunk layers,
// [16][288] for the heads (l = 3); asc[l] = device scalar 2^-e; bias[3] = the six head biases.
cudaError_t launch_trunk_fused(const float* x0, const void* const w16[4][2], const float* const asc[4], const float* const bias[4], const HeadArgs& heads, int M,
                               cudaStream_t stream) {
    if (!x0 || M <= 0) return cudaErrorInvalidValue;
    static const int kpad[3] = {480, 288, 288};
    struct Entry { const void* x; const void* w0; int M; TFMaps maps; };
    static thread_local Entry cache[4];
    static thread_local int next_slot = 0;
    const TFMaps* found = nullptr;
    for (int i = 0; i < 4; i++)
        if (cache[i].x == x0 && cache[i].w0 == w16[0][0] && cache[i].M == M) { found = &cache[i].maps; break; }
    if (!found) {
        Entry& e = cache[next_slot];
        next_slot = (next_slot + 1) % 4;
        e.x = nullptr;
        if (!tf_map_f32(&e.maps.a, x0, M, 460)) return cudaErrorInvalidValue;
        for (int l = 0; l < 3; l++)
            for (int h = 0; h < 2; h++)
                for (int p = 0; p < 2; p++)
                    if (!tf_map_f16(&e.maps.w[l][h][p], w16[l][p], TF_N, kpad[l], (h ? TF_N1 : TF_N0) / TF_CLUSTER)) return cudaErrorInvalidValue;
        for (int p = 0; p < 2; p++)
            if (!tf_map_f16(&e.maps.wh[p], w16[3][p], TF_NH, 288, TF_NH / TF_CLUSTER)) return cudaErrorInvalidValue;
        e.x = x0; e.w0 = w16[0][0]; e.M = M;
        found = &e.maps;
    }
    static PerDeviceFlag configured;
    static int sm_count[kMaxDevices] = {};
    const int dslot = current_device_slot();
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_trunk_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF_SMEM_BYTES);
        int dev = 0, n = 0;
        if (e == cudaSuccess) e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) { configured.retract(); return e; }
        if (TF_CLUSTER > 1) {   // how many cluster pairs are resident at once (a GPC with an odd number of free SMs strands one)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(n / TF_CLUSTER * TF_CLUSTER)); cfg.blockDim = dim3(TF_THREADS); cfg.dynamicSmemBytes = TF_SMEM_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = TF_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, k_trunk_fused, &cfg) == cudaSuccess && nc > 0 && nc * TF_CLUSTER < n) n = nc * TF_CLUSTER;
            (void)cudaGetLastError();
        }
        sm_count[dslot] = n;
    }
    TFArgs a{};
    for (int l = 0; l < 4; l++) { a.bias[l] = bias[l]; a.acc_scale[l] = asc[l]; }
    a.heads = heads; a.M = M; a.n_tiles = (M + TF_BM - 1) / TF_BM; a.prof = g_tf_prof;
    const int sms = sm_count[dslot] > 0 ? sm_count[dslot] : 148;
    const int pairs = (a.n_tiles + TF_CLUSTER - 1) / TF_CLUSTER;
    const int blocks = TF_CLUSTER * (pairs < sms / TF_CLUSTER ? pairs : sms / TF_CLUSTER);
    k_trunk_fused<<<blocks, TF_THREADS, TF_SMEM_BYTES, stream>>>(*found, a);
    return cudaGetLastError();
}

}  // namespace mm
