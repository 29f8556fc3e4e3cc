// mm_linear16.cu -- K4/K5 forward GEMM, second generation: Y = relu(X W^T + b) on tcgen05 with an error-compensated FP16 split
// (3xFP16), N-split tiles, two CTAs per SM.  sm_100a.
//
// Why a second kernel (mm_policy_tc.cu's k_linear_tf32x3 stays for the backward GEMMs, whose operands need TF32's 8-bit exponent):
//   * FP16 and TF32 carry the same 11 significant bits, so x ~ x_hi + x_lo with x_hi = fp16(x), x_lo = fp16(x - x_hi) (both round to
//     nearest) is as exact as the TF32 split (|x - x_hi - x_lo| <= 2^-22 |x|) as long as the operands sit inside FP16's exponent range --
//     activations of this MLP do (|x| < 100), weights are pre-multiplied by a power of two on the host (exact) and the accumulator is
//     multiplied back in the epilogue.  kind::f16 runs at twice the kind::tf32 rate (K = 16 instead of 8 per 32-byte step) and its
//     weight tiles are half the bytes: the TF32 kernel was bound by re-streaming 68 KB of hi/lo weights per k-block from L2
//     (profiles/r01h: 2.95 GB per launch, 8.5 TB/s of the ~10.7 TB/s the L2 delivers).
//   * One CTA = 128 rows x ONE HALF of the 264 output columns (half 0: columns 0..127, N = 128; half 1: columns 128..263, N = 144):
//     144 accumulator columns + a 3-stage ring of [hi 16 | lo 16] operand columns = 240 of 256 TMEM columns, 87 KB of shared memory
//     -> TWO CTAs per SM.  While one CTA drains its accumulator (tcgen05.ld -> bias + ReLU -> TMA store) and starts up (barrier
//     init, TMEM allocation, first TMA round trip), the other one's MMAs keep the tensor pipe busy: the serial prologue / epilogue
//     that cost the 1-CTA/SM kernel ~40 % of its time (tensor pipe active 60 %) is overlapped without a persistent scheduler.
//     Price: the activation tile is fetched and split once per half (L2 hit for the second).
// Warp roles as in the first kernel: warp 0 = TMA producer (fp32 activation tile through a 2-slot ring, fp16 hi/lo weight tiles through
// a 3-stage ring), warp 1 = TMEM allocator + MMA issuer (TS-form: A from tensor memory, B from shared memory), warps 2-9 = splitter
// (thread = row: fp32 -> packed fp16 hi / lo -> tcgen05.st) then epilogue.  In the last trunk layer (HEADS) the epilogue contracts its
// half of the row with the six head rows and writes PARTIAL head sums; k_heads_finish adds the two halves, masks, samples and writes
// actions + joint log-probs (PPO.get_action, PPO.py:170-186).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include "mm_env.cuh"
#include "mm_policy_heads.cuh"
#include "mm_tc.cuh"

namespace mm {

constexpr int L16_BM = 128, L16_BK = 32, L16_NMAX = 144, L16_STAGES = 3, L16_NSPLIT0 = 128;
constexpr uint32_t L16_A_BYTES = L16_BM * L16_BK * 4;          // 16384: fp32 activation k-block, 128-byte rows, SWIZZLE_128B
constexpr uint32_t L16_WROW = L16_BK * 2;                      // 64-byte fp16 weight rows, SWIZZLE_64B
constexpr uint32_t L16_WHALF = L16_NMAX * L16_WROW;            // 9216: one of (hi, lo), sized for the wider half
constexpr uint32_t L16_W_BYTES = 2 * L16_WHALF;                // 18432 per stage: [hi | lo]
constexpr uint32_t L16_A_OFF = L16_STAGES * L16_W_BYTES;       // 55296
constexpr uint32_t L16_RING_BYTES = L16_A_OFF + 2 * L16_A_BYTES;  // 88064
constexpr uint32_t L16_SMEM_BYTES = L16_RING_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
static_assert(2 * (L16_SMEM_BYTES + 1024) <= 232448, "two CTAs per SM");
constexpr int L16_SPLIT_WARPS = 8, L16_THREADS = 64 + 32 * L16_SPLIT_WARPS;
constexpr uint32_t L16_TMEM_COLS = 256, L16_TMEM_A_COL = 160;  // D: columns 0..143; operand ring: 160 + 32 * stage, [hi 16 | lo 16]
static_assert(L16_TMEM_A_COL + 32 * L16_STAGES <= L16_TMEM_COLS && L16_TMEM_A_COL >= 160, "chunk 4 of the wide half reads D columns 128..159");
constexpr uint32_t L16_EPI_BYTES = L16_SPLIT_WARPS * 8192;     // epilogue staging (two 4 KB blocks per warp) reuses the ring
static_assert(L16_EPI_BYTES + L16_BM * 8 * 4 <= L16_RING_BYTES, "epilogue staging + head exchange fit in the ring");

// K-major SWIZZLE_64B shared-memory descriptor: 8-row atoms of 512 bytes (cute::UMMA::SmemDescriptor, layout type 4)
__device__ __forceinline__ uint64_t l16_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8 * L16_WROW) >> 4) << 32) | ((uint64_t)1 << 46) | (4ull << 61);
}
// cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A = B = F16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t l16_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                 "r"(v[6]), "r"(v[7])
                 : "memory");
}

struct L16Maps {
    CUtensorMap a, w_hi[2], w_lo[2], y;   // weight maps per column half (their boxes differ in rows)
};
struct L16Args {
    const float* bias;          // [n_valid]
    const float* acc_scale;     // device scalar: 1 / (the power of two the weights were multiplied by)
    uint32_t* gate_out;         // RELU: [M][9] ReLU bit words, or nullptr
    const float* head_w;        // HEADS: [6][264]
    float* heads_part;          // HEADS: [M][2][8] partial head sums (6 used)
    int M, K, n_valid, n_split; // n_split = 1 (n_valid <= 128) or 2
};

enum { L16_RELU = 0, L16_HEADS = 1 };
template <int kMode>
__global__ void __launch_bounds__(L16_THREADS, 2) k_linear_f16x3(const __grid_constant__ L16Maps maps, const L16Args args) {
    constexpr bool kHeads = kMode == L16_HEADS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L16_RING_BYTES);
    uint64_t* empty = full + L16_STAGES;
    uint64_t* split_done = empty + L16_STAGES;
    uint64_t* a_full = split_done + L16_STAGES;   // [2]
    uint64_t* a_free = a_full + 2;                // [2]
    uint64_t* tmem_full = a_free + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    const int tile = blockIdx.x / args.n_split, half = blockIdx.x - tile * args.n_split;
    const int m0 = tile * L16_BM;
    const int n0 = half * L16_NSPLIT0;
    // this CTA's accumulator width: the valid columns of its half, rounded up to the UMMA granularity for M = 128 (16)
    const int n_cols = min(args.n_valid - n0, half == 0 && args.n_split == 2 ? L16_NSPLIT0 : L16_NMAX);
    const int n_mma = (n_cols + 15) & ~15;
    const int nkb = (args.K + L16_BK - 1) / L16_BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < L16_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&split_done[s], 32 * L16_SPLIT_WARPS); }
        for (int s = 0; s < 2; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_free[s], 32 * L16_SPLIT_WARPS); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(L16_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

    if (warp == 0) {
        {  // ===== TMA producer (the whole warp runs the loop, one elected lane issues: see elect_one)
            const uint32_t w_tx = 2u * (uint32_t)n_mma * L16_WROW;   // the boxes are n_mma rows tall (host: one map pair per half)
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % L16_STAGES, sa = kb & 1;
                const int k0 = kb * L16_BK;
                mbar_wait(&a_free[sa], ((kb >> 1) & 1) ^ 1);      // the splitter has read the tile that used this slot two k-blocks ago
                if (elect_one()) {
                    mbar_expect_tx(&a_full[sa], L16_A_BYTES);
                    tma_load_2d(smem + L16_A_OFF + sa * L16_A_BYTES, &maps.a, k0, m0, &a_full[sa]);
                }
                mbar_wait(&empty[s], ((kb / L16_STAGES) & 1) ^ 1);
                uint8_t* st = smem + s * L16_W_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(&full[s], w_tx);
                    tma_load_2d(st, &maps.w_hi[half], k0, n0, &full[s]);
                    tma_load_2d(st + L16_WHALF, &maps.w_lo[half], k0, n0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        {  // ===== MMA issuer (whole warp loops, one elected lane issues): D[128 x n_mma] += A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T, two K = 16 steps per k-block
            const uint32_t idesc = l16_idesc(L16_BM, n_mma);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % L16_STAGES;
                mbar_wait(&full[s], (kb / L16_STAGES) & 1);        // weight tiles landed
                mbar_wait(&split_done[s], (kb / L16_STAGES) & 1);  // operand columns of this stage written (tcgen05.st fenced)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + s * L16_W_BYTES);
                const uint64_t b_hi = l16_desc(st), b_lo = l16_desc(st + L16_WHALF);
                if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 2; k++) {  // UMMA_K = 16 halves = 32 bytes: advance the start address inside the 64-byte swizzle row
                    const uint64_t o = (uint64_t)(k * 2);
                    const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
                    const uint32_t ta_hi = tmem_base + L16_TMEM_A_COL + (uint32_t)(s * 32 + k * 8), ta_lo = ta_hi + 16;
                    umma_f16_ts(tmem_base, ta_hi, b_hi + o, idesc, first);
                    umma_f16_ts(tmem_base, ta_lo, b_hi + o, idesc, 1u);
                    umma_f16_ts(tmem_base, ta_hi, b_lo + o, idesc, 1u);
                }
                umma_commit(&empty[s]);  // frees the weight stage AND the operand columns when these MMAs have read them
                }
            }
            if (elect_one()) umma_commit(tmem_full);
        }
    } else {
        // ===== splitter: thread = row (its TMEM lane); the two warps of a lane quarter take k-columns 0..15 and 16..31 of the k-block
        const int whalf = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int arow = quarter * 32 + lane;
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % L16_STAGES, sa = kb & 1;
            // logical 16-byte chunk c of row r sits at physical chunk c ^ (r & 7) of the 128-byte swizzled row
            const float4* rowp = reinterpret_cast<const float4*>(smem + L16_A_OFF + sa * L16_A_BYTES + arow * 128);
            mbar_wait(&a_full[sa], (kb >> 1) & 1);
            float4 av[4];
#pragma unroll
            for (int c = 0; c < 4; c++) av[c] = rowp[(c + 4 * whalf) ^ (arow & 7)];
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_free[sa])) : "memory");  // slot read: the next tile may land
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float e[4] = {av[c].x, av[c].y, av[c].z, av[c].w};
#pragma unroll
                for (int j = 0; j < 2; j++) {  // element 2j in the low half of the 32-bit column, 2j+1 in the high half
                    const __half2 h2 = __floats2half2_rn(e[2 * j], e[2 * j + 1]);
                    const float2 hf = __half22float2(h2);
                    const __half2 l2 = __floats2half2_rn(e[2 * j] - hf.x, e[2 * j + 1] - hf.y);
                    hi[2 * c + j] = *reinterpret_cast<const uint32_t*>(&h2);
                    lo[2 * c + j] = *reinterpret_cast<const uint32_t*>(&l2);
                }
            }
            mbar_wait(&empty[s], ((kb / L16_STAGES) & 1) ^ 1);   // operand stage s is free: the MMAs of k-block kb - 3 have completed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + L16_TMEM_A_COL + (uint32_t)(s * 32);
            tmem_st_32x8(ta + 8 * whalf, hi);
            tmem_st_32x8(ta + 16 + 8 * whalf, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split_done[s])) : "memory");
        }
        // ===== epilogue: warp w may only touch TMEM lanes 32 * (w % 4) .. +31; one thread = one output row
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint8_t* t_blk = smem + (size_t)(warp - 2) * 8192;   // two 4 KB blocks per warp in the TMA's 128-byte swizzle
        const int row0 = m0 + quarter * 32;
        const float scale = __ldg(args.acc_scale);
        float hacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int nchunks = (n_cols + 31) >> 5;
        const int c_mid = (nchunks + 1) >> 1;
        const int c_lo = whalf ? c_mid : 0, c_hi = whalf ? nchunks : c_mid;   // a quarter's two warps share the 32-column chunks
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c++) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                  "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                  "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + (uint32_t)(c * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int gcol0 = n0 + c * 32;   // global output column of this chunk
            uint32_t word = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                float yv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int col = gcol0 + 4 * q + j;
                    const bool ok = col < args.n_valid && c * 32 + 4 * q + j < n_cols;   // columns past this half's share are junk (operand ring / padding)
                    yv[j] = ok ? fmaxf(fmaf(__uint_as_float(v[4 * q + j]), scale, __ldg(&args.bias[col])), 0.f) : 0.f;
                    word |= (yv[j] > 0.f ? 1u : 0u) << (4 * q + j);
                }
                if (kHeads) {
                    if (gcol0 + 4 * q < args.n_valid) {  // n_valid is a multiple of 4: whole float4 groups are in or out
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            const float4 wv = __ldg(reinterpret_cast<const float4*>(args.head_w + j * args.n_valid + gcol0 + 4 * q));  // same address in every lane
                            hacc[j] = fmaf(yv[3], wv.w, fmaf(yv[2], wv.z, fmaf(yv[1], wv.y, fmaf(yv[0], wv.x, hacc[j]))));
                        }
                    }
                } else {
                    *reinterpret_cast<float4*>(t_blk + ((c - c_lo) & 1) * 4096 + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_float4(yv[0], yv[1], yv[2], yv[3]);
                }
            }
            if (kHeads) continue;
            if (args.gate_out && row0 + lane < args.M) args.gate_out[(size_t)(row0 + lane) * 9 + (gcol0 >> 5)] = word;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's block writes -> visible to the TMA
            __syncwarp();
            if (lane == 0) {
                if (row0 < args.M && gcol0 < args.n_valid) tma_store_2d(&maps.y, t_blk + ((c - c_lo) & 1) * 4096, gcol0, row0);
                tma_store_commit();
                tma_store_wait_read<1>();  // the block written two iterations from now is the one whose store was committed before this one
            }
            __syncwarp();
        }
        if (!kHeads && lane == 0) tma_store_wait_read<0>();  // shared memory must outlive the TMA's reads
        if (kHeads) {  // add the partner warp's partial head sums (same rows, the other chunks), then write this half's partial
            float* xch = reinterpret_cast<float*>(smem + L16_EPI_BYTES) + (size_t)(quarter * 32 + lane) * 8;
            if (whalf) {
#pragma unroll
                for (int j = 0; j < 6; j++) xch[j] = hacc[j];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (!whalf) {
                const long long row = (long long)row0 + lane;
                if (row < args.M) {
                    float* dst = args.heads_part + ((size_t)row * 2 + half) * 8;
                    *reinterpret_cast<float4*>(dst) = make_float4(hacc[0] + xch[0], hacc[1] + xch[1], hacc[2] + xch[2], hacc[3] + xch[3]);
                    *reinterpret_cast<float2*>(dst + 4) = make_float2(hacc[4] + xch[4], hacc[5] + xch[5]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L16_TMEM_COLS) : "memory");
    }
}

// logits = the two halves' partial head sums + bias; mask, sample (or evaluate), joint log-prob of the env's two agents
__global__ void __launch_bounds__(256) k_heads_finish(const float* __restrict__ part, const float* __restrict__ head_b, const HeadArgs ha, int R) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // R is even: rows 2e, 2e+1 sit in adjacent lanes
    float lp = 0.f;
    if (row < R) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(part + row * 16)), b0 = __ldg(reinterpret_cast<const float4*>(part + row * 16 + 8));
        const float2 a1 = __ldg(reinterpret_cast<const float2*>(part + row * 16 + 4)), b1 = __ldg(reinterpret_cast<const float2*>(part + row * 16 + 12));
        const float l[6] = {a0.x + b0.x + __ldg(&head_b[0]), a0.y + b0.y + __ldg(&head_b[1]), a0.z + b0.z + __ldg(&head_b[2]),
                            a0.w + b0.w + __ldg(&head_b[3]), a1.x + b1.x + __ldg(&head_b[4]), a1.y + b1.y + __ldg(&head_b[5])};
        lp = head_sample_or_eval(l, row, ha);
    }
    const float lp_pair = lp + __shfl_xor_sync(0xffffffffu, lp, 1);
    if (row < R && (row & 1) == 0) ha.logp[row >> 1] = lp_pair;
}

// ------------------------------------------------------------------------------------------------ host side
static bool l16_map_f32(CUtensorMap* m, const float* base, int rows, int cols, int ld) {   // {32 cols x 128 rows} boxes, 128-byte swizzle, zero fill
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)L16_BK, (cuuint32_t)L16_BM};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool l16_map_f16(CUtensorMap* m, const void* base, int rows, int kpad, int box_rows) {  // fp16 [rows][kpad], {32 x box_rows} boxes, 64-byte swizzle
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)L16_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool l16_map_store(CUtensorMap* m, float* base, int rows, int cols, int ld) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Y = relu(acc_scale * X W^T + bias): X [M][K] fp32 (row pitch K), W given as the fp16 split (w_hi, w_lo) [n_rows_w][kpad] of 2^e * W with
// kpad = K rounded up to 32 (zero padded) and *acc_scale = 2^-e.  heads != nullptr: last trunk layer -- partial head sums into
// heads_part [M][2][8] and k_heads_finish instead of a store (n_rows_w must be 264).
cudaError_t launch_linear_f16x3(const float* x, const void* w_hi, const void* w_lo, int n_rows_w, int kpad, const float* acc_scale, const float* bias, float* y, int ldy,
                                int M, int K, uint32_t* gate_out, const float* head_w, const float* head_b, const HeadArgs* heads, float* heads_part,
                                cudaStream_t stream) {
    if (n_rows_w <= 0 || n_rows_w > 264 || (n_rows_w & 3) || (ldy & 3) || (K & 3) || kpad < K || (kpad & 31) || !acc_scale || !bias) return cudaErrorInvalidValue;
    if (heads && (n_rows_w != 264 || !heads_part || !head_w || !head_b)) return cudaErrorInvalidValue;
    if (!heads && !y) return cudaErrorInvalidValue;
    const int n_split = n_rows_w > L16_NSPLIT0 ? 2 : 1;
    struct Entry { const void *x, *wh, *wl, *y; int M, K, nw, ldy, kpad; L16Maps maps; };
    static thread_local Entry cache[12];
    static thread_local int next_slot = 0;
    const L16Maps* found = nullptr;
    for (int i = 0; i < 12; i++)
        if (cache[i].x == x && cache[i].wh == w_hi && cache[i].wl == w_lo && cache[i].y == y && cache[i].ldy == ldy && cache[i].M == M && cache[i].K == K &&
            cache[i].nw == n_rows_w && cache[i].kpad == kpad) { found = &cache[i].maps; break; }
    if (!found) {
        Entry& e = cache[next_slot];
        next_slot = (next_slot + 1) % 12;
        e.x = nullptr;
        if (!l16_map_f32(&e.maps.a, x, M, K, K)) return cudaErrorInvalidValue;
        for (int h = 0; h < n_split; h++) {
            const int n0 = h * L16_NSPLIT0;
            const int n_cols = n_rows_w - n0 < (h == 0 && n_split == 2 ? L16_NSPLIT0 : L16_NMAX) ? n_rows_w - n0 : (h == 0 && n_split == 2 ? L16_NSPLIT0 : L16_NMAX);
            const int n_mma = (n_cols + 15) & ~15;
            if (!l16_map_f16(&e.maps.w_hi[h], w_hi, n_rows_w, kpad, n_mma) || !l16_map_f16(&e.maps.w_lo[h], w_lo, n_rows_w, kpad, n_mma)) return cudaErrorInvalidValue;
        }
        if (n_split == 1) { e.maps.w_hi[1] = e.maps.w_hi[0]; e.maps.w_lo[1] = e.maps.w_lo[0]; }
        if (y) { if (!l16_map_store(&e.maps.y, y, M, n_rows_w, ldy)) return cudaErrorInvalidValue; }
        else e.maps.y = e.maps.a;
        e.x = x; e.wh = w_hi; e.wl = w_lo; e.y = y; e.ldy = ldy; e.M = M; e.K = K; e.nw = n_rows_w; e.kpad = kpad;
        found = &e.maps;
    }
    static PerDeviceFlag configured;
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_linear_f16x3<L16_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L16_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_f16x3<L16_HEADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L16_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_f16x3<L16_RELU>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_f16x3<L16_HEADS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { configured.retract(); return e; }
    }
    L16Args a{};
    a.bias = bias; a.acc_scale = acc_scale; a.gate_out = gate_out; a.head_w = head_w; a.heads_part = heads_part; a.M = M; a.K = K; a.n_valid = n_rows_w; a.n_split = n_split;
    const int blocks = ((M + L16_BM - 1) / L16_BM) * n_split;
    if (heads) {
        k_linear_f16x3<L16_HEADS><<<blocks, L16_THREADS, L16_SMEM_BYTES, stream>>>(*found, a);
        k_heads_finish<<<(M + 255) / 256, 256, 0, stream>>>(heads_part, head_b, *heads, M);
    } else {
        k_linear_f16x3<L16_RELU><<<blocks, L16_THREADS, L16_SMEM_BYTES, stream>>>(*found, a);
    }
    return cudaGetLastError();
}

}  // namespace mm
