// mm_update_tc.cu -- weight-gradient GEMM of the PPO update (PPO.py:58-85, loss.backward() through the actor's Linear layers) on
// tcgen05 with error-compensated TF32 (3xTF32), sm_100a.
//
//     dW[n][k] = sum_r dZ[r][n] * H[r][k]          dZ [R][264] (gradient at a layer's pre-activation), H [R][Kin] (the layer's input)
//     db[n]    = sum_r dZ[r][n]                    (comes out of the same MMAs as column Kin: the H tile carries a column of ones)
//
// The reduction runs over the ROWS of two row-major activations, so both operands are "MN-major" for the tensor core: a TMA box of
// {32 columns x 32 rows} with the 32-byte-chunk 128-byte swizzle lands exactly as canonical MN-major SWIZZLE_128B_BASE32B atoms, no
// transposition anywhere.  One CTA owns a [128 x 288] tile of dW (m-tile of dZ columns x n-tile of H columns) and a contiguous slab of
// rows; it accumulates in TMEM over its slab and writes ONE partial tile (plain stores -- the sum over slabs is a tiny second pass, so
// the result is deterministic).  Warp roles as in mm_policy_tc.cu: warp 0 = TMA, warp 1 = MMA issue, warps 2-5 = splitter, then epilogue.
// Both operands are activations, so both are split in the kernel: the H tile in shared memory (hi in place, lo beside it); the dZ tile
// goes to TENSOR MEMORY -- thread m reads column m of the landed boxes (conflict-free: a warp reads one 128-byte row per step), and
// writes its hi / lo rows with tcgen05.st, which makes dZ^T the K-major TMEM A operand of a TS-form MMA and keeps its three reads per
// k-step off the shared-memory port (the SS form of this kernel was shared-memory-bandwidth bound).
#include <cuda.h>
#include <stdio.h>
#include "mm_env.cuh"
#include "mm_tc.cuh"

namespace mm {

constexpr int WG_BK = 32;                 // rows per pipeline stage (4 UMMA k-steps of 8)
constexpr int WG_M = 128, WG_N = 288;     // dW tile: 128 dZ-columns x 288 H-columns (9 boxes of 32); MMA N = 160 + 128
constexpr int WG_N1 = 160, WG_N2 = 128;
constexpr int WG_STAGES = 2;
constexpr uint32_t WG_BOX_BYTES = 32 * WG_BK * 4;                          // 4096: one {32 col x 32 row} box
constexpr uint32_t WG_A_BYTES = (WG_M / 32) * WG_BOX_BYTES;                // 16384
constexpr uint32_t WG_B_BYTES = (WG_N / 32) * WG_BOX_BYTES;                // 36864
constexpr uint32_t WG_STAGE_BYTES = WG_A_BYTES + 2 * WG_B_BYTES;           // dZ boxes (fp32, read once by the splitter) | H hi (in place) | H lo
constexpr uint32_t WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;
constexpr uint32_t WG_TMEM_A_COL = 288;   // D occupies TMEM columns 0..287; dZ^T stages [hi 32 | lo 32] per pipeline stage from column 288
constexpr int WG_SPLIT_WARPS = 8;                         // two per scheduler: one warp alone cannot cover its own shared-memory / TMEM latencies
constexpr int WG_THREADS = 64 + 32 * WG_SPLIT_WARPS;
constexpr uint32_t WG_TMEM_COLS = 512;

// MN-major operands of 32-bit elements have exactly one legal shared-memory layout, SWIZZLE_128B_BASE32B (cute::UMMA::LayoutType 1,
// Swizzle<2,5,2>: the 32-byte chunk index of a 128-byte row is XORed with the row index mod 4); its TMA counterpart is
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Canonical form ((8,n),(4,k)) in 16-byte units: 32-element groups along M/N are LBO = one box
// apart, groups of 4 k-rows are SBO = 512 bytes apart (one UMMA_K = 8 step spans two of them).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(WG_BOX_BYTES >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) | ((uint64_t)1 << 46) | (1ull << 61);
}
// as umma_idesc_tf32 (mm_policy_tc.cu) with b_major = MN (bit 16); A comes from tensor memory, K-major by construction
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float tf32_rn_i(float x) {  // cvt.rna.tf32.f32 on finite values, on the integer pipe
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

struct WgMaps {
    CUtensorMap a, b;  // A: the operand whose columns become rows of the result (through TMEM); B: columns of the result (shared memory)
};

__global__ void __launch_bounds__(WG_THREADS, 1)
k_wgrad_tf32x3(const __grid_constant__ WgMaps maps, float* __restrict__ part, int R, int a_cols, int b_cols, int ones_on_a, int n_mt, int n_nt, int kb_per, int ld) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* split_done = empty + WG_STAGES;
    uint64_t* tmem_full = split_done + WG_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    const int tiles = n_mt * n_nt;
    const int tile = blockIdx.x % tiles, slab = blockIdx.x / tiles;  // the tiles of one slab are neighbours: they share their H / dZ boxes in L2
    const int mt = tile % n_mt, nt = tile / n_mt;
    const int nkb_total = (R + WG_BK - 1) / WG_BK;
    const int kb0 = slab * kb_per;
    const int nkb = min(kb_per, nkb_total - kb0);  // >= 1 by construction of the grid
    // the appended column of ones (bias gradient) lives on whichever operand is H: as local column ones_col of this n-tile's B boxes, or as
    // row a_cols of the A operand
    const int ones_col = ones_on_a ? -1 : b_cols - nt * WG_N;
    const int out_rows = a_cols + (ones_on_a ? 1 : 0);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < WG_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&split_done[s], 32 * WG_SPLIT_WARPS); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(WG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

    if (warp == 0) {
        {  // ===== TMA producer (the whole warp runs the loop, one elected lane issues: see elect_one): 4 dZ boxes + 9 H boxes per stage, out-of-range columns / rows arrive as zeros
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % WG_STAGES;
                mbar_wait(&empty[s], ((kb / WG_STAGES) & 1) ^ 1);
                uint8_t* st = smem + s * WG_STAGE_BYTES;
                const int r0 = (kb0 + kb) * WG_BK;
                if (elect_one()) {
                    mbar_expect_tx(&full[s], WG_A_BYTES + WG_B_BYTES);
#pragma unroll
                    for (int j = 0; j < WG_M / 32; j++) tma_load_2d(st + j * WG_BOX_BYTES, &maps.a, mt * WG_M + 32 * j, r0, &full[s]);
#pragma unroll
                    for (int j = 0; j < WG_N / 32; j++) tma_load_2d(st + WG_A_BYTES + j * WG_BOX_BYTES, &maps.b, nt * WG_N + 32 * j, r0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        {  // ===== MMA issuer (whole warp loops, one elected lane issues): D[128 x 288] += dZ_hi^T H_hi + dZ_lo^T H_hi + dZ_hi^T H_lo
            constexpr uint32_t id1 = umma_idesc_tf32_mn(WG_M, WG_N1), id2 = umma_idesc_tf32_mn(WG_M, WG_N2);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % WG_STAGES;
                mbar_wait(&split_done[s], (kb / WG_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + s * WG_STAGE_BYTES);
                if (elect_one()) {
#pragma unroll
                for (int g = 0; g < WG_BK / 8; g++) {  // UMMA_K = 8 rows = 1024 bytes down every box
                    const uint32_t o = g * 1024u;
                    const uint32_t ta_hi = tmem_base + WG_TMEM_A_COL + (uint32_t)(s * 64 + g * 8), ta_lo = ta_hi + 32;
                    const uint64_t b1_hi = umma_desc_mn(st + WG_A_BYTES + o), b1_lo = umma_desc_mn(st + WG_A_BYTES + WG_B_BYTES + o);
                    const uint64_t b2_hi = umma_desc_mn(st + WG_A_BYTES + (WG_N1 / 32) * WG_BOX_BYTES + o);
                    const uint64_t b2_lo = umma_desc_mn(st + WG_A_BYTES + WG_B_BYTES + (WG_N1 / 32) * WG_BOX_BYTES + o);
                    const uint32_t first = (kb == 0 && g == 0) ? 0u : 1u;
                    umma_tf32_ts(tmem_base, ta_hi, b1_hi, id1, first);
                    umma_tf32_ts(tmem_base, ta_lo, b1_hi, id1, 1u);
                    umma_tf32_ts(tmem_base, ta_hi, b1_lo, id1, 1u);
                    umma_tf32_ts(tmem_base + WG_N1, ta_hi, b2_hi, id2, first);
                    umma_tf32_ts(tmem_base + WG_N1, ta_lo, b2_hi, id2, 1u);
                    umma_tf32_ts(tmem_base + WG_N1, ta_hi, b2_lo, id2, 1u);
                }
                umma_commit(&empty[s]);
                }
            }
            if (elect_one()) umma_commit(tmem_full);
        }
    } else {
        // ===== splitter, 8 warps.  H boxes: 2304 float4 per stage over 256 threads; thread t owns float4 t + 256 q, i.e. box q, row t/8,
        // physical 16-byte slot t%8 of that row; logical column c of row r sits in slot (((c>>3) ^ (r&3)) << 1) | ((c>>2)&1).
        const int st_tid = threadIdx.x - 64;
        // the ones column: box ones_col/32 (= its q); row r of it belongs to the thread with t/8 == r and t%8 == slot(r)
        const bool has_ones = ones_col >= 0 && ones_col < WG_N;
        const int ones_cb = ones_col & 31, ones_e = ones_col & 3;
        const bool ones_owner = has_ones && ((st_tid & 7) == ((((ones_cb >> 3) ^ ((st_tid >> 3) & 3)) << 1) | ((ones_cb >> 2) & 1)));
        const int ones_q = has_ones ? (ones_col >> 5) : -1;
        const int quarter_a = warp & 3;           // this warp's TMEM lane quarter = dZ box: thread = dZ column quarter_a*32 + lane
        const int khalf = ((warp - 2) >> 2) * 16;  // the two warps of a quarter take k-rows 0..15 and 16..31 of the stage
        const bool a_is_ones = ones_on_a && (mt * WG_M + quarter_a * 32 + lane == a_cols);
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % WG_STAGES;
            mbar_wait(&full[s], (kb / WG_STAGES) & 1);
            {   // dZ^T -> TMEM: element (row k, column lane) of box quarter_a is at k*128 + (((lane>>3) ^ (k&3)) << 5) + (lane&7)*4
                const uint8_t* box = smem + s * WG_STAGE_BYTES + quarter_a * WG_BOX_BYTES + khalf * 128;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int k = 0; k < 16; k++) {  // khalf is a multiple of 4: (khalf + k) & 3 == k & 3
                    const float v = *reinterpret_cast<const float*>(box + k * 128 + ((((lane >> 3) ^ (k & 3)) << 5) | ((lane & 7) << 2)));
                    hi[k] = a_is_ones ? 0x3F800000u : tf32_rn_bits(v);
                    lo[k] = a_is_ones ? 0u : tf32_rn_bits(v - __uint_as_float(hi[k]));
                }
                const uint32_t ta = tmem_base + ((uint32_t)(quarter_a * 32) << 16) + WG_TMEM_A_COL + (uint32_t)(s * 64 + khalf);
                tmem_st_32x16(ta, hi);
                tmem_st_32x16(ta + 32, lo);
            }
            float4* raw = reinterpret_cast<float4*>(smem + s * WG_STAGE_BYTES + WG_A_BYTES);
            float4* lo_t = reinterpret_cast<float4*>(smem + s * WG_STAGE_BYTES + WG_A_BYTES + WG_B_BYTES);
#pragma unroll
            for (int q = 0; q < (int)(WG_B_BYTES / 16 / (32 * WG_SPLIT_WARPS)); q++) {
                const int j = st_tid + 32 * WG_SPLIT_WARPS * q;
                const float4 v = raw[j];
                float4 h, l;
                h.x = tf32_rn_i(v.x); h.y = tf32_rn_i(v.y); h.z = tf32_rn_i(v.z); h.w = tf32_rn_i(v.w);
                l.x = tf32_rn_i(v.x - h.x); l.y = tf32_rn_i(v.y - h.y); l.z = tf32_rn_i(v.z - h.z); l.w = tf32_rn_i(v.w - h.w);
                if (ones_owner && q == ones_q) {
                    if (ones_e == 0) { h.x = 1.0f; l.x = 0.0f; } else if (ones_e == 1) { h.y = 1.0f; l.y = 0.0f; }
                    else if (ones_e == 2) { h.z = 1.0f; l.z = 0.0f; } else { h.w = 1.0f; l.w = 0.0f; }
                }
                raw[j] = h; lo_t[j] = l;
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split_done[s])) : "memory");
        }
        // ===== epilogue: TMEM lane = dZ column (row of dW), TMEM column = H column
        const int quarter = warp & 3;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int n = mt * WG_M + quarter * 32 + lane;
        float* dst = part + ((size_t)slab * out_rows + n) * ld + nt * WG_N;
#pragma unroll 1
        for (int c = (khalf ? 5 : 0); c < (khalf ? WG_N / 32 : 5); c++) {  // the two warps of a quarter share the 9 column chunks
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                  "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                  "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + (uint32_t)(c * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (n < out_rows) {
#pragma unroll
                for (int q = 0; q < 8; q++)
                    *reinterpret_cast<float4*>(dst + c * 32 + 4 * q) =
                        make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(WG_TMEM_COLS) : "memory");
    }
}

// fp32 row-major [rows][cols], box = {32 cols x 32 rows}, 128-byte swizzle of 32-byte chunks, zero fill out of bounds
static bool make_map_mn(CUtensorMap* m, const float* base, int rows, int cols) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)WG_BK};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Geometry for (R, n_out, k_in).  The result tile grid is 128 rows (A operand columns) x 288 columns (B operand columns); dZ^T [H | 1] can be
// laid out either way round, and the orientation with fewer tiles wins: 264 x 461 is 3 x 2 tiles as dZ^T [H|1] but 4 x 1 as [H|1]^T dZ.
// transposed = 0: part [slabs][n_out][ld], column k_in = db.   transposed = 1: part [slabs][k_in + 1][ld], row k_in = db.
void wgrad_geometry(int R, int n_out, int k_in, int* slabs, int* ld, int* kb_per, int* out_rows, int* transposed) {
    const int t0 = ((n_out + WG_M - 1) / WG_M) * ((k_in + 1 + WG_N - 1) / WG_N), t1 = ((k_in + 1 + WG_M - 1) / WG_M) * ((n_out + WG_N - 1) / WG_N);
    const int tr = t1 < t0 ? 1 : 0;
    const int tiles = tr ? t1 : t0;
    const int n_nt = tr ? (n_out + WG_N - 1) / WG_N : (k_in + 1 + WG_N - 1) / WG_N;
    const int nkb_total = (R + WG_BK - 1) / WG_BK;
    int want = 148 / tiles;  // one CTA per SM
    if (want < 1) want = 1;
    int per = (nkb_total + want - 1) / want;
    if (per < 1) per = 1;
    *slabs = (nkb_total + per - 1) / per;
    *ld = n_nt * WG_N;
    *kb_per = per;
    *out_rows = tr ? k_in + 1 : n_out;
    *transposed = tr;
}

// per-slab partial sums of dZ^T [H | 1] (or its transpose, see wgrad_geometry)
cudaError_t launch_wgrad_tc(const float* dz, const float* h, int R, int n_out, int k_in, float* part, cudaStream_t stream) {
    if (R <= 0 || (n_out & 3) || (k_in & 3)) return cudaErrorInvalidValue;
    int slabs, ld, kb_per, out_rows, tr;
    wgrad_geometry(R, n_out, k_in, &slabs, &ld, &kb_per, &out_rows, &tr);
    const int a_cols = tr ? k_in : n_out, b_cols = tr ? n_out : k_in;
    const int n_mt = (out_rows + WG_M - 1) / WG_M, n_nt = ld / WG_N;
    WgMaps maps;
    if (!make_map_mn(&maps.a, tr ? h : dz, R, a_cols) || !make_map_mn(&maps.b, tr ? dz : h, R, b_cols)) return cudaErrorInvalidValue;
    static PerDeviceFlag configured;
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM_BYTES);
        if (e != cudaSuccess) { configured.retract(); return e; }
    }
    k_wgrad_tf32x3<<<slabs * n_mt * n_nt, WG_THREADS, WG_SMEM_BYTES, stream>>>(maps, part, R, a_cols, b_cols, tr, n_mt, n_nt, kb_per, ld);
    return cudaGetLastError();
}

}  // namespace mm
