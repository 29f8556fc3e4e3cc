// mm_policy_tc.cu -- K4 tensor-core path: Y = relu(X W^T + b) on tcgen05 with error-compensated TF32 (3xTF32), sm_100a.
//
// The parity bar of the policy path is 1e-5 relative on values / log-probs (BASELINE north_star), which single-pass TF32/BF16
// tensor-core math cannot meet.  Every operand is therefore used as a two-term split x ~ x_hi + x_lo with x_hi = tf32(x) and
// x_lo = tf32(x - x_hi), both rounded to nearest (|x - x_hi - x_lo| <= 2^-22 |x|, zero-mean), and each k-step issues three MMAs
// into the same fp32 TMEM accumulator:  hi*hi + lo*hi + hi*lo  (the dropped lo*lo term is ~2^-22).
//
// Activations travel through HBM ONCE, as plain fp32: the split of the A operand happens inside the kernel.  One CTA = one 128-row
// tile of X through one layer (N = 264 outputs, padded to 272 = 144 + 128 so that each half is a legal UMMA N for M = 128).
// Warp roles (default build): warp 0 = TMA producer (one lane): pre-split hi/lo weight tiles through a 3-stage mbarrier ring, the fp32
// activation tile through a slot of its own; warp 1 = TMEM allocator + MMA issuer (one lane), TS-form MMAs: A from tensor memory, B from
// shared memory; warps 2-9 = SPLITTER during the main loop (thread = row: read the landed fp32 row, write its hi and lo halves into a
// 3-stage [hi 32 | lo 32]-column TMEM ring with tcgen05.st) and EPILOGUE afterwards (TMEM -> registers -> bias + ReLU / gate / heads ->
// 32 x 32 blocks in the TMA's swizzle -> cp.async.bulk.tensor stores).  Out-of-bounds rows / columns (M tail, K = 460 -> 480,
// N = 264 -> 272) are zero-filled by TMA loads and clipped by TMA stores.  -DMM_TC_A_TMEM=0 builds the first version (A split in shared
// memory, SS-form MMAs, 2-stage ring), -DMM_TC_SPLIT_WARPS=4 / -DMM_TC_TMA_STORE=0 the intermediate ones, for A/B measurements.
#include <cuda.h>
#include <stdio.h>
#include "mm_env.cuh"
#include "mm_policy_heads.cuh"
#include "mm_tc.cuh"

namespace mm {

#ifndef MM_TC_BK
#define MM_TC_BK 32
#endif
#ifndef MM_TC_A_TMEM
#define MM_TC_A_TMEM 1
#endif
#ifndef MM_TC_STAGES
#define MM_TC_STAGES (MM_TC_A_TMEM ? 3 : 2)
#endif
constexpr int TC_BM = 128, TC_BK = MM_TC_BK, TC_N1 = 144, TC_N2 = 128, TC_N = 264, TC_STAGES = MM_TC_STAGES;
constexpr uint32_t TC_ROW_BYTES = TC_BK * 4;                       // one K-block row: 128 B (SWIZZLE_128B) or 64 B (SWIZZLE_64B)
constexpr uint32_t TC_ATOM_BYTES = 8 * TC_ROW_BYTES;               // 8-row swizzle atom = stride byte offset of the UMMA descriptor
constexpr uint64_t TC_LAYOUT = TC_BK == 32 ? 2ull : 4ull;          // cute::UMMA::LayoutType SWIZZLE_128B / SWIZZLE_64B
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 4;    // 16384
constexpr uint32_t TC_B1_BYTES = TC_N1 * TC_BK * 4;   // 18432
constexpr uint32_t TC_B2_BYTES = TC_N2 * TC_BK * 4;   // 16384
constexpr uint32_t TC_W_BYTES = 2 * (TC_B1_BYTES + TC_B2_BYTES);   // 69632: [W1 hi | W2 hi | W1 lo | W2 lo]
#if MM_TC_A_TMEM
// TMEM-operand build: the pipeline ring holds WEIGHT tiles only (3 stages); the activation tile passes through a single 16 KB slot of its
// own -- the splitter frees it as soon as it has read its rows, long before the MMAs of that k-block run -- and through a 3-stage ring of
// [hi 32 | lo 32] TMEM columns.
constexpr uint32_t TC_W_OFF = 0, TC_STAGE_BYTES = TC_W_BYTES;
constexpr uint32_t TC_A_SLOT_OFF = TC_STAGES * TC_STAGE_BYTES;
constexpr uint32_t TC_RING_BYTES = TC_A_SLOT_OFF + TC_A_BYTES;
static_assert(TC_STAGES <= 3, "TMEM columns 288 + 64 * stages must stay below 512");
#else
constexpr uint32_t TC_W_OFF = 2 * TC_A_BYTES, TC_STAGE_BYTES = 2 * TC_A_BYTES + TC_W_BYTES;  // [A hi | A lo | weights] per stage: 102400
constexpr uint32_t TC_RING_BYTES = TC_STAGES * TC_STAGE_BYTES;
#endif
constexpr uint32_t TC_SMEM_BYTES = TC_RING_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
static_assert(TC_SMEM_BYTES <= 232448, "227 KB of shared memory per CTA");
#ifndef MM_TC_SPLIT_WARPS
#define MM_TC_SPLIT_WARPS 8   // splitter / epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, i.e. two per scheduler)
#endif
constexpr int TC_SPLIT_WARPS = MM_TC_SPLIT_WARPS;
static_assert(TC_SPLIT_WARPS == 4 || TC_SPLIT_WARPS == 8, "one or two warps per TMEM lane quarter");
constexpr int TC_THREADS = 64 + 32 * TC_SPLIT_WARPS;
// MM_TC_A_TMEM: the activation operand reaches the tensor core through TENSOR MEMORY instead of shared memory.  An SS-form tf32 MMA of
// M = 128, N = 144 reads (128 + 144) * 32 bytes of shared memory for 78 clocks of math -- 111 of the SM's 128 bytes/clock -- and the 3xTF32
// scheme issues three of them per k-step on top of the TMA writes and the splitter's own traffic: the SS kernel is shared-memory-bandwidth
// bound at ~42 % tensor activity.  With A in TMEM the splitter reads each landed fp32 row once and writes hi / lo with tcgen05.st, and the
// MMAs read only the weight tiles from shared memory.
static_assert(!MM_TC_A_TMEM || MM_TC_BK == 32, "the TMEM-A splitter addresses the 128-byte swizzle");
constexpr uint32_t TC_TMEM_A_COL = 288;  // D occupies columns 0..271; A stages: [hi 32 | lo 32] per pipeline stage from column 288
constexpr uint32_t TC_TMEM_COLS = 512;

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row atoms of 1024 bytes.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(TC_ATOM_BYTES >> 4) << 32) | ((uint64_t)1 << 46) | (TC_LAYOUT << 61);
}
// cute::UMMA::InstrDescriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
struct TcMaps {
    CUtensorMap a, w1_hi, w2_hi, w1_lo, w2_lo, y;
};
// MM_TC_TMA_STORE: the epilogue hands each 32 x 32 block of the result to the TMA (cp.async.bulk.tensor shared -> global) instead of
// storing it with the LSU: the warp moves on to its next block at once, the tensor map clips rows >= M and columns >= n_valid, and the
// burst of 148 CTAs x 135 KB leaves asynchronously.
#ifndef MM_TC_TMA_STORE
#define MM_TC_TMA_STORE 1
#endif

// Epilogue modes.  TC_EPI_RELU: y = relu(acc + bias).  TC_EPI_HEADS: this is the last trunk layer -- instead of storing y, the epilogue
// contracts each row with the 6 head rows (5 move logits + 1 mark logit), masks, samples (or evaluates) the action and writes actions +
// the env's joint log-prob: "sampling fused in the epilogue".  TC_EPI_GATE / TC_EPI_PLAIN are the backward (data-gradient) uses of the
// same GEMM, dH = dZ W: y = acc where the ReLU the gradient flows back through was open, else 0 -- or y = acc; n_valid columns are stored
// with row pitch ldy (the 460-wide dX of the first layer is produced as two column blocks).  The ReLU pattern travels as BITS: a forward
// launch (TC_EPI_RELU) can emit gate_out [M][9] u32 (bit b of word c = y[row][32c+b] > 0, 36 bytes per row instead of 1056), and the
// TC_EPI_GATE epilogue reads its row's 9 words before the main loop, so the gate costs neither bandwidth nor exposed latency.
enum { TC_EPI_RELU = 0, TC_EPI_HEADS = 1, TC_EPI_GATE = 2, TC_EPI_PLAIN = 3 };
template <int kMode>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_linear_tf32x3(const __grid_constant__ TcMaps maps, const float* __restrict__ bias, float* __restrict__ y, int M, int K, const float* __restrict__ head_w,
                const float* __restrict__ head_b, const HeadArgs heads, const uint32_t* __restrict__ gate, uint32_t* __restrict__ gate_out, int n_valid, int ldy) {
    constexpr bool kHeads = kMode == TC_EPI_HEADS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_RING_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* split_done = empty + TC_STAGES;
    uint64_t* tmem_full = split_done + TC_STAGES;
    uint64_t* a_full = tmem_full + 1;   // TMEM-operand build: the activation slot has landed / has been read by every splitter thread
    uint64_t* a_free = a_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_free + 1);

    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM;
    const int nkb = (K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&split_done[s], 32 * TC_SPLIT_WARPS); }
        mbar_init(tmem_full, 1);
        mbar_init(a_full, 1); mbar_init(a_free, 32 * TC_SPLIT_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

    if (warp == 0) {
        {  // ===== TMA producer (the whole warp runs the loop, one elected lane issues: see elect_one)
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % TC_STAGES;
                const int k0 = kb * TC_BK;
                uint8_t* st = smem + s * TC_STAGE_BYTES;
#if MM_TC_A_TMEM
                mbar_wait(a_free, (kb & 1) ^ 1);              // the splitter has read the previous activation tile out of the slot
                if (elect_one()) {
                    mbar_expect_tx(a_full, TC_A_BYTES);
                    tma_load_2d(smem + TC_A_SLOT_OFF, &maps.a, k0, m0, a_full);
                }
                mbar_wait(&empty[s], ((kb / TC_STAGES) & 1) ^ 1);
                if (elect_one()) {
                mbar_expect_tx(&full[s], TC_W_BYTES);
#else
                mbar_wait(&empty[s], ((kb / TC_STAGES) & 1) ^ 1);
                if (elect_one()) {
                mbar_expect_tx(&full[s], TC_STAGE_BYTES - TC_A_BYTES);  // the A_lo slot is produced in-kernel
                tma_load_2d(st, &maps.a, k0, m0, &full[s]);
#endif
                tma_load_2d(st + TC_W_OFF, &maps.w1_hi, k0, 0, &full[s]);
                tma_load_2d(st + TC_W_OFF + TC_B1_BYTES, &maps.w2_hi, k0, TC_N1, &full[s]);
                tma_load_2d(st + TC_W_OFF + TC_B1_BYTES + TC_B2_BYTES, &maps.w1_lo, k0, 0, &full[s]);
                tma_load_2d(st + TC_W_OFF + 2 * TC_B1_BYTES + TC_B2_BYTES, &maps.w2_lo, k0, TC_N1, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        {  // ===== MMA issuer (whole warp loops, one elected lane issues): D[128 x 272] (two N chunks) += A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T
            constexpr uint32_t id1 = umma_idesc_tf32(TC_BM, TC_N1), id2 = umma_idesc_tf32(TC_BM, TC_N2);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % TC_STAGES;
                mbar_wait(&full[s], (kb / TC_STAGES) & 1);        // weight tiles landed
                mbar_wait(&split_done[s], (kb / TC_STAGES) & 1);  // A tile rewritten as hi, lo tile written (fenced to the async proxy)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + s * TC_STAGE_BYTES);
                if (elect_one()) {
#if !MM_TC_A_TMEM
                const uint64_t a_hi = umma_desc(st), a_lo = umma_desc(st + TC_A_BYTES);
#endif
                const uint64_t b1_hi = umma_desc(st + TC_W_OFF), b2_hi = umma_desc(st + TC_W_OFF + TC_B1_BYTES);
                const uint64_t b1_lo = umma_desc(st + TC_W_OFF + TC_B1_BYTES + TC_B2_BYTES), b2_lo = umma_desc(st + TC_W_OFF + 2 * TC_B1_BYTES + TC_B2_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; k++) {  // UMMA_K = 8 tf32 = 32 bytes: advance the start address inside the 128-byte swizzle row
                    const uint64_t o = (uint64_t)(k * 2);
                    const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
#if MM_TC_A_TMEM
                    const uint32_t ta_hi = tmem_base + TC_TMEM_A_COL + (uint32_t)(s * 64 + k * 8), ta_lo = ta_hi + 32;
                    umma_tf32_ts(tmem_base, ta_hi, b1_hi + o, id1, first);
                    umma_tf32_ts(tmem_base, ta_lo, b1_hi + o, id1, 1u);
                    umma_tf32_ts(tmem_base, ta_hi, b1_lo + o, id1, 1u);
                    umma_tf32_ts(tmem_base + TC_N1, ta_hi, b2_hi + o, id2, first);
                    umma_tf32_ts(tmem_base + TC_N1, ta_lo, b2_hi + o, id2, 1u);
                    umma_tf32_ts(tmem_base + TC_N1, ta_hi, b2_lo + o, id2, 1u);
#else
                    umma_tf32(tmem_base, a_hi + o, b1_hi + o, id1, first);
                    umma_tf32(tmem_base, a_lo + o, b1_hi + o, id1, 1u);
                    umma_tf32(tmem_base, a_hi + o, b1_lo + o, id1, 1u);
                    umma_tf32(tmem_base + TC_N1, a_hi + o, b2_hi + o, id2, first);
                    umma_tf32(tmem_base + TC_N1, a_lo + o, b2_hi + o, id2, 1u);
                    umma_tf32(tmem_base + TC_N1, a_hi + o, b2_lo + o, id2, 1u);
#endif
                }
                umma_commit(&empty[s]);  // frees the stage when these MMAs have read it
                }
            }
            if (elect_one()) umma_commit(tmem_full);      // accumulator complete
        }
    } else {
        // ===== splitter (main loop): 128 threads turn each landed fp32 A tile into (hi in place, lo beside it)
        const int st_tid = threadIdx.x - 64;
        const int whalf = (warp - 2) >> 2;  // 0 for the first warp of a TMEM lane quarter, 1 for the second (8-warp build)
        uint32_t gbits[9];
        if (kMode == TC_EPI_GATE) {
            const long long grow = (long long)m0 + (warp & 3) * 32 + lane;  // the output row this thread owns in the epilogue
#pragma unroll
            for (int c = 0; c < 9; c++) gbits[c] = grow < M ? __ldg(&gate[grow * 9 + c]) : 0u;
        }
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % TC_STAGES;
#if MM_TC_A_TMEM
            // thread = row (its TMEM lane): logical 16-byte chunk c of row r sits at physical chunk c ^ (r & 7) of the 128-byte swizzled row.
            // With 8 warps the two warps of a quarter take k-columns 0..15 and 16..31 of the row.
            constexpr int kCh = TC_SPLIT_WARPS == 4 ? 8 : 4;          // 16-byte chunks per thread
            const int arow = (warp & 3) * 32 + lane;
            const float4* rowp = reinterpret_cast<const float4*>(smem + TC_A_SLOT_OFF + arow * 128);
            mbar_wait(a_full, kb & 1);                                // this k-block's activation tile is in the slot
            float4 av[kCh];
#pragma unroll
            for (int c = 0; c < kCh; c++) av[c] = rowp[(c + (TC_SPLIT_WARPS == 8 ? 4 * whalf : 0)) ^ (arow & 7)];
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(a_free)) : "memory");  // slot read: the next tile may land
            uint32_t hi[4 * kCh], lo[4 * kCh];
#pragma unroll
            for (int c = 0; c < kCh; c++) {
                const float e[4] = {av[c].x, av[c].y, av[c].z, av[c].w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    hi[4 * c + j] = tf32_rn_bits(e[j]);
                    lo[4 * c + j] = tf32_rn_bits(e[j] - __uint_as_float(hi[4 * c + j]));
                }
            }
            mbar_wait(&empty[s], ((kb / TC_STAGES) & 1) ^ 1);         // TMEM A stage s is free: the MMAs of k-block kb - TC_STAGES have completed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TC_TMEM_A_COL + (uint32_t)(s * 64);
            if constexpr (TC_SPLIT_WARPS == 4) {
                tmem_st_32x32(ta, hi);
                tmem_st_32x32(ta + 32, lo);
            } else {
                tmem_st_32x16(ta + 16 * whalf, hi);
                tmem_st_32x16(ta + 32 + 16 * whalf, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#else
            mbar_wait(&full[s], (kb / TC_STAGES) & 1);
            float4* raw = reinterpret_cast<float4*>(smem + s * TC_STAGE_BYTES);
            float4* lo_t = reinterpret_cast<float4*>(smem + s * TC_STAGE_BYTES + TC_A_BYTES);
#pragma unroll
            for (int q = 0; q < (int)(TC_A_BYTES / 16 / (32 * TC_SPLIT_WARPS)); q++) {
                const int j = st_tid + 32 * TC_SPLIT_WARPS * q;
                const float4 v = raw[j];
                float4 h, l;
                h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
                l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y); l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
                raw[j] = h; lo_t[j] = l;
            }
#endif
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma's operand reads
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split_done[s])) : "memory");
        }
        // ===== epilogue: warp w may only touch TMEM lanes 32*(w%4) .. +31; one thread = one output row
        const int quarter = warp & 3;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // The accumulator row of a lane is scattered over rows of Y (stride 1056 B): storing straight from registers costs 32
        // half-written sectors per instruction (measured: half of the kernel's time).  Each warp therefore transposes 32x32 blocks
        // through shared memory -- the pipeline stages are free once tmem_full has fired -- and writes whole 128-byte row segments.
        constexpr int kTP = 36;  // padded row pitch (floats): 16-byte aligned, conflict-free for the quarter-warp float4 patterns below
#if MM_TC_TMA_STORE
        // two 4 KB blocks per warp in the TMA's 128-byte swizzle (row r, 16-byte chunk q at r*128 + ((q ^ (r&7)) << 4): conflict-free writes)
        uint8_t* t_blk = smem + (size_t)(warp - 2) * 8192;
        float* t_y = nullptr; (void)t_y;
#else
        float* t_y = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * kTP;
#endif
        const int row0 = m0 + quarter * 32;
        float hacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int c_lo = (TC_SPLIT_WARPS == 8 && whalf) ? 5 : 0, c_hi = (TC_SPLIT_WARPS == 8 && !whalf) ? 5 : 9;  // a quarter's two warps share the chunks
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c++) {  // 9 x 32 columns >= 264
            uint32_t gword = 0;
            if (kMode == TC_EPI_GATE) {  // gbits[c] with a rolled loop: select without dynamic register indexing
#pragma unroll
                for (int i = 0; i < 9; i++) gword = (i == c) ? gbits[i] : gword;
            }
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                  "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                  "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + (uint32_t)(c * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t word = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                float yv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int col = c * 32 + 4 * q + j;
                    if (kMode == TC_EPI_RELU || kMode == TC_EPI_HEADS) yv[j] = fmaxf(__uint_as_float(v[4 * q + j]) + (col < n_valid ? __ldg(&bias[col]) : 0.f), 0.f);
                    else if (kMode == TC_EPI_GATE) yv[j] = ((gword >> (4 * q + j)) & 1u) ? __uint_as_float(v[4 * q + j]) : 0.f;
                    else yv[j] = __uint_as_float(v[4 * q + j]);
                    if (kMode == TC_EPI_RELU) word |= (yv[j] > 0.f ? 1u : 0u) << (4 * q + j);
                }
                if (kHeads) {
                    if (c * 32 + 4 * q < TC_N) {  // TC_N is a multiple of 4: whole float4 groups are in or out
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            const float4 wv = __ldg(reinterpret_cast<const float4*>(head_w + j * TC_N + c * 32 + 4 * q));  // same address in every lane: broadcast
                            hacc[j] = fmaf(yv[3], wv.w, fmaf(yv[2], wv.z, fmaf(yv[1], wv.y, fmaf(yv[0], wv.x, hacc[j]))));
                        }
                    }
                } else {
#if MM_TC_TMA_STORE
                    *reinterpret_cast<float4*>(t_blk + ((c - c_lo) & 1) * 4096 + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_float4(yv[0], yv[1], yv[2], yv[3]);
#else
                    *reinterpret_cast<float4*>(&t_y[lane * kTP + 4 * q]) = make_float4(yv[0], yv[1], yv[2], yv[3]);
#endif
                }
            }
            if (kHeads) continue;
            if (kMode == TC_EPI_RELU && gate_out && row0 + lane < M) gate_out[(size_t)(row0 + lane) * 9 + c] = word;
#if MM_TC_TMA_STORE
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's block writes -> visible to the TMA
            __syncwarp();
            if (lane == 0) {
                if (row0 < M && c * 32 < n_valid) tma_store_2d(&maps.y, t_blk + ((c - c_lo) & 1) * 4096, c * 32, row0);
                tma_store_commit();
                tma_store_wait_read<1>();  // the block written two iterations from now is the one whose store was committed before this one
            }
            __syncwarp();
            continue;
#endif
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; it++) {  // 4 rows x 128 contiguous bytes per instruction
                const int r = it * 4 + (lane >> 3), c4 = lane & 7;
                const int col = c * 32 + 4 * c4;
                if (row0 + r < M && col < n_valid) {
                    float4 o = *reinterpret_cast<const float4*>(&t_y[r * kTP + 4 * c4]);
                    *reinterpret_cast<float4*>(y + (size_t)(row0 + r) * ldy + col) = o;
                }
            }
            __syncwarp();
        }
#if MM_TC_TMA_STORE
        if (!kHeads && lane == 0) tma_store_wait_read<0>();  // shared memory must outlive the TMA's reads
#endif
        if (kHeads && TC_SPLIT_WARPS == 8) {  // add the partner warp's partial head dot products (same rows, the other column chunks)
            float* xch = reinterpret_cast<float*>(smem + (size_t)TC_SPLIT_WARPS * 8192) + (size_t)(quarter * 32 + lane) * 6;
            if (whalf) {
#pragma unroll
                for (int j = 0; j < 6; j++) xch[j] = hacc[j];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (!whalf) {
#pragma unroll
                for (int j = 0; j < 6; j++) hacc[j] += xch[j];
            }
        }
        if (kHeads && !(TC_SPLIT_WARPS == 8 && whalf)) {  // one thread = one agent row; rows 2e and 2e+1 (the two agents of env e) sit in adjacent lanes
            const long long row = (long long)row0 + lane;
            float lp = 0.f;
            if (row < M) {
                float l[6];
#pragma unroll
                for (int j = 0; j < 6; j++) l[j] = hacc[j] + __ldg(&head_b[j]);
                lp = head_sample_or_eval(l, row, heads);
            }
            const float lp_pair = lp + __shfl_xor_sync(0xffffffffu, lp, 1);
            if (row < M && (lane & 1) == 0) heads.logp[row >> 1] = lp_pair;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
// fp32 row-major [rows][cols] with row pitch ld floats, {32 cols x 32 rows} box, 128-byte swizzle: the epilogue's store map
static bool make_store_map(CUtensorMap* m, float* base, int rows, int cols, int ld) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// fp32 row-major [rows][cols] with a {32 cols x box_rows} box, 128-byte swizzle, zero fill out of bounds
static bool make_map(CUtensorMap* m, const float* base, int rows, int cols, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               TC_BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Y = epi(X W^T) with X [M][K] plain fp32 and W given as the split (w_hi, w_lo) [n_rows_w][K], n_rows_w <= 264.  mode TC_EPI_RELU writes
// y = relu(. + bias) [M][264]; TC_EPI_HEADS (last layer, `heads` given) runs heads + sampling in the epilogue and writes actions / log-probs
// instead of y; TC_EPI_GATE / TC_EPI_PLAIN write n_rows_w columns with row pitch ldy (gate = ReLU bit words [M][9], as emitted through gate_out by a TC_EPI_RELU launch).
cudaError_t launch_linear_tc_ex(const float* x, const float* w_hi, const float* w_lo, int n_rows_w, const float* bias, float* y, int ldy, int M, int K, int mode,
                                const uint32_t* gate, uint32_t* gate_out, const float* head_w, const float* head_b, const HeadArgs* heads, cudaStream_t stream) {
    // The rollout calls this with the same scratch / weight pointers every step: keep the encoded maps (a tensor map depends only on
    // base pointer, extents and box) in a small per-thread cache instead of re-encoding 15 of them per policy step.
    struct Entry { const float *x, *wh, *wl, *y; int M, K, nw, ldy; TcMaps maps; };
    static thread_local Entry cache[8];
    static thread_local int next_slot = 0;
    if (n_rows_w <= 0 || n_rows_w > TC_N || (n_rows_w & 3) || (ldy & 3) || (K & 3)) return cudaErrorInvalidValue;
    const TcMaps* found = nullptr;
    for (int i = 0; i < 8; i++)
        if (cache[i].x == x && cache[i].wh == w_hi && cache[i].wl == w_lo && cache[i].y == y && cache[i].ldy == ldy && cache[i].M == M && cache[i].K == K && cache[i].nw == n_rows_w) { found = &cache[i].maps; break; }
    if (!found) {
        Entry& e = cache[next_slot];
        next_slot = (next_slot + 1) % 8;
        e.x = nullptr;
        if (!make_map(&e.maps.a, x, M, K, TC_BM) || !make_map(&e.maps.w1_hi, w_hi, n_rows_w, K, TC_N1) || !make_map(&e.maps.w2_hi, w_hi, n_rows_w, K, TC_N2) ||
            !make_map(&e.maps.w1_lo, w_lo, n_rows_w, K, TC_N1) || !make_map(&e.maps.w2_lo, w_lo, n_rows_w, K, TC_N2))
            return cudaErrorInvalidValue;
        if (y && !make_store_map(&e.maps.y, y, M, n_rows_w, ldy)) return cudaErrorInvalidValue;
        e.x = x; e.wh = w_hi; e.wl = w_lo; e.y = y; e.ldy = ldy; e.M = M; e.K = K; e.nw = n_rows_w;
        found = &e.maps;
    }
    const TcMaps& maps = *found;
    static PerDeviceFlag configured;
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_linear_tf32x3<TC_EPI_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tf32x3<TC_EPI_HEADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tf32x3<TC_EPI_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tf32x3<TC_EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
        if (e != cudaSuccess) { configured.retract(); return e; }
    }
    const int blocks = (M + TC_BM - 1) / TC_BM;
    switch (mode) {
    case TC_EPI_HEADS:
        if (!heads) return cudaErrorInvalidValue;
        k_linear_tf32x3<TC_EPI_HEADS><<<blocks, TC_THREADS, TC_SMEM_BYTES, stream>>>(maps, bias, nullptr, M, K, head_w, head_b, *heads, nullptr, nullptr, TC_N, TC_N);
        break;
    case TC_EPI_RELU:
        k_linear_tf32x3<TC_EPI_RELU><<<blocks, TC_THREADS, TC_SMEM_BYTES, stream>>>(maps, bias, y, M, K, nullptr, nullptr, HeadArgs{}, nullptr, gate_out, n_rows_w, ldy);
        break;
    case TC_EPI_GATE:
        if (!gate) return cudaErrorInvalidValue;
        k_linear_tf32x3<TC_EPI_GATE><<<blocks, TC_THREADS, TC_SMEM_BYTES, stream>>>(maps, nullptr, y, M, K, nullptr, nullptr, HeadArgs{}, gate, nullptr, n_rows_w, ldy);
        break;
    case TC_EPI_PLAIN:
        k_linear_tf32x3<TC_EPI_PLAIN><<<blocks, TC_THREADS, TC_SMEM_BYTES, stream>>>(maps, nullptr, y, M, K, nullptr, nullptr, HeadArgs{}, nullptr, nullptr, n_rows_w, ldy);
        break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_linear_tc(const float* x, const float* w_hi, const float* w_lo, const float* bias, float* y, int M, int K, const float* head_w,
                             const float* head_b, const HeadArgs* heads, cudaStream_t stream) {
    return launch_linear_tc_ex(x, w_hi, w_lo, TC_N, bias, y, TC_N, M, K, heads ? TC_EPI_HEADS : TC_EPI_RELU, nullptr, nullptr, head_w, head_b, heads, stream);
}

}  // namespace mm
