// mm_tokens_mma.cu -- K4 stage 1, second generation: the 23-token projection + self-attention + residual of one agent row (networks.py:58-65,
// 75-82) with the attention's two small matrix products on the warp-level tensor path.  sm_100a.
//
// Why: the SIMT version (k_tokens_r, mm_policy.cu) spends ~930 FMA warp-instructions per row, 690 of them in the attention (23 x 23 scores of length
// 10, 23 x 23 x 20 context sums) on 23 of 32 lanes; the FMA pipe issues one warp-instruction per 2 clocks per scheduler, which bounds the kernel at
// ~0.21 ms for 131 072 rows (measured 0.34 ms; FFMA2 has the same lane throughput, profiles/r02g).  Batched over rows these products have no shared
// operand, so tcgen05 (one B operand per instruction) does not fit; mma.sync does: per row S = Q K^T is 32 x 24 x 16 (padded) and O = P V is
// 32 x 24 x 24, 54 HMMA instructions with the error-compensated fp16 split (x ~ hi + lo, hi.hi + lo.hi + hi.lo, fp32 accumulation -- 1e-6 relative,
// like the trunk), ~1000 MAC/clk/SM on this path (profiles/r02i).
// One warp = one row at a time.  Phase A (SIMT, lane = token): the token's (token | key | query | value) = M_a x_a + b_a from the per-token affine
// maps in shared memory (as k_tokens); keys / queries / values go to the warp's shared tiles as fp16 hi / lo rows (48-byte pitch: conflict-free for
// 16-byte stores and for ldmatrix), the token itself as fp32.  Phase B (whole warp): ldmatrix -> HMMA scores -> softmax on the accumulator fragments
// (rows live in lane quads: two shuffles per reduction) -> the probabilities become the A fragments of the second product in registers (the
// m16n8 C layout IS the m16n8k16 A layout) -> HMMA context -> + token (residual) -> x0.
#include <cuda_fp16.h>
#include "mm_env.cuh"

namespace mm {

constexpr int TM_TOK = 23, TM_EMB = 20, TM_KQ = 10, TM_X0 = TM_TOK * TM_EMB;
#ifndef MM_TOKM_WARPS
#define MM_TOKM_WARPS 8   // 8 warps share one copy of the maps: 104 KB per block, two blocks per SM (0.295 -> 0.276 ms against 4 warps x 3 blocks)
#endif
constexpr int TM_WARPS = MM_TOKM_WARPS;
constexpr int TM_PITCH = 48;                                    // bytes per token row of the fp16 tiles (16 or 24 halves used)
constexpr int TM_Q_BYTES = 32 * TM_PITCH, TM_KV_BYTES = 24 * TM_PITCH;
// per-warp tile: Q hi, Q lo (32 rows: the M dimension is padded to two m16 tiles), K hi, K lo, V hi, V lo (24 rows), token fp32 [23][20]
constexpr int TM_OFF_QH = 0, TM_OFF_QL = TM_Q_BYTES, TM_OFF_KH = 2 * TM_Q_BYTES, TM_OFF_KL = TM_OFF_KH + TM_KV_BYTES, TM_OFF_VH = TM_OFF_KL + TM_KV_BYTES,
              TM_OFF_VL = TM_OFF_VH + TM_KV_BYTES, TM_OFF_TOK = TM_OFF_VL + TM_KV_BYTES, TM_TILE_BYTES = TM_OFF_TOK + 24 * TM_EMB * 4;
static_assert(TM_TILE_BYTES % 16 == 0, "16-byte aligned tiles");
constexpr int TM_MAP_FLOATS = 60 * TM_TOK * 5;                  // [60][23][4] + [60][23]
// Rows a warp carries through phase A together: every map element fetched from shared memory serves TM_ROWS rows (240 of the ~445 shared-memory
// wavefronts per row are those fetches); phase B runs one row at a time on the row's own tile.  Measured (profiles/r03d_token_rows_variants.jsonl): 2 rows
// x 8 warps 0.305 ms, 2 x 4 warps x 2 blocks 0.298, 3 x 6 warps 0.366, 4 x 4 warps 0.452 against 0.290 for one row x 8 warps x 2 blocks -- the second
// tile per warp halves the resident warps and the lost latency hiding costs more than the saved wavefronts.  Default 1.
#ifndef MM_TOKM_ROWS
#define MM_TOKM_ROWS 1
#endif
constexpr int TM_ROWS = MM_TOKM_ROWS;
constexpr int TM_SMEM_BYTES = TM_MAP_FLOATS * 4 + TM_WARPS * TM_ROWS * TM_TILE_BYTES;

struct TokOffsets { int tokm, tokb, proj_col, proj_dim; };

__device__ __forceinline__ void tm_split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}
__device__ __forceinline__ void tm_ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void tm_ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void tm_mma16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tm_mma8(float* c, uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}

// Default: expf.  -DMM_TOKM_FAST_EXP=1: exp(x) for finite x <= 0 as 2^t (1 + e ln 2) with t = fl(x log2 e) and e = x log2 e - t recovered by an FMA (the
// argument's rounding error put back; what is left is ex2.approx's own ~2^-22) -- measured no faster (0.305 vs 0.294 ms): the kernel is bound by
// shared-memory wavefronts (the per-token maps), not by issue slots.
__device__ __forceinline__ float tm_exp(float x) {
#if !defined(MM_TOKM_FAST_EXP) || !MM_TOKM_FAST_EXP
    return expf(x);
#else
    const float t = x * 1.4426950408889634f;
    const float e = fmaf(x, 1.4426950408889634f, -t) + x * 1.925963033500649e-8f;   // + x * (log2 e - fl(log2 e))
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return fmaf(r * e, 0.6931471805599453f, r);
#endif
}

#ifndef MM_TOKM_MINBLOCKS
#define MM_TOKM_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(TM_WARPS * 32, MM_TOKM_MINBLOCKS) k_tokens_mma(const float* __restrict__ obs, const float* __restrict__ wts, const TokOffsets o,
                                                                                   float* __restrict__ x0, int nrows) {
    extern __shared__ __align__(16) uint8_t tm_smem[];
    float (*s_m)[TM_TOK][4] = reinterpret_cast<float (*)[TM_TOK][4]>(tm_smem);                     // [60][23][4]
    float (*s_b)[TM_TOK] = reinterpret_cast<float (*)[TM_TOK]>(tm_smem + 60 * TM_TOK * 16);       // [60][23]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tiles = tm_smem + TM_MAP_FLOATS * 4 + w * TM_ROWS * TM_TILE_BYTES;
    for (int i = threadIdx.x; i < 60 * TM_TOK * 4; i += blockDim.x) (&s_m[0][0][0])[i] = wts[o.tokm + i];
    for (int i = threadIdx.x; i < 60 * TM_TOK; i += blockDim.x) (&s_b[0][0])[i] = wts[o.tokb + i];
    for (int i = lane; i < TM_ROWS * TM_TILE_BYTES / 16; i += 32) reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0u, 0u, 0u, 0u);   // padding rows / columns stay zero
    __syncthreads();
    const bool on = lane < TM_TOK;
    const int a = on ? lane : 0;
    const int c0 = (int)wts[o.proj_col + a], nd = (int)wts[o.proj_dim + a];
    const int g = lane >> 2, t4 = lane & 3;
    // ldmatrix row addresses.  A (queries, x4): matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15) of an m16 tile.
    const uint32_t q_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * TM_PITCH + (lane >> 4) * 16);
    // B of the scores (keys, x4 = two n8 tiles): matrices (n 0-7, k 0-7), (n 0-7, k 8-15), (n 8-15, k 0-7), (n 8-15, k 8-15)
    const uint32_t k_off = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * TM_PITCH + ((lane >> 3) & 1) * 16);
    // B of the context (values, x4.trans, one n8 tile of the embedding): matrices (tokens 0-7), (8-15), (16-23), (16-23 again, unused)
    const uint32_t v_off = (uint32_t)(((lane & 7) + min(lane >> 3, 2) * 8) * TM_PITCH);
#pragma unroll 1
    for (int row0 = (blockIdx.x * TM_WARPS + w) * TM_ROWS; row0 < nrows; row0 += gridDim.x * TM_WARPS * TM_ROWS) {
        // ---------------- phase A: lane = token, TM_ROWS rows per map fetch
        float x[TM_ROWS][4];
#pragma unroll
        for (int r = 0; r < TM_ROWS; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) x[r][c] = (c < nd && row0 + r < nrows) ? obs[(size_t)(row0 + r) * kObs + c0 + c] : 0.f;
        auto affine = [&](int j, float* out) {   // out[r] = (M_a x_a + b_a)[j] of row r
            const float4 m = *reinterpret_cast<const float4*>(&s_m[j][a][0]);
            const float b = s_b[j][a];
#pragma unroll
            for (int r = 0; r < TM_ROWS; r++) out[r] = fmaf(x[r][3], m.w, fmaf(x[r][2], m.z, fmaf(x[r][1], m.y, fmaf(x[r][0], m.x, b))));
        };
#pragma unroll
        for (int part = 0; part < 2; part++) {   // keys (map rows 20-29) -> K tiles, queries (30-39) -> Q tiles
            uint32_t hi[TM_ROWS][8], lo[TM_ROWS][8];
#pragma unroll
            for (int d = 0; d < 8; d++) {
                if (d < TM_KQ / 2) {
                    float v0[TM_ROWS], v1[TM_ROWS];
                    affine(20 + 10 * part + 2 * d, v0); affine(21 + 10 * part + 2 * d, v1);
#pragma unroll
                    for (int r = 0; r < TM_ROWS; r++) tm_split2(v0[r], v1[r], hi[r][d], lo[r][d]);
                } else {
#pragma unroll
                    for (int r = 0; r < TM_ROWS; r++) hi[r][d] = lo[r][d] = 0u;
                }
            }
            if (on) {
#pragma unroll
                for (int r = 0; r < TM_ROWS; r++) {
                    uint8_t* tile = tiles + r * TM_TILE_BYTES;
                    uint4* ph = reinterpret_cast<uint4*>(tile + (part ? TM_OFF_QH : TM_OFF_KH) + a * TM_PITCH);
                    uint4* pl = reinterpret_cast<uint4*>(tile + (part ? TM_OFF_QL : TM_OFF_KL) + a * TM_PITCH);
                    ph[0] = make_uint4(hi[r][0], hi[r][1], hi[r][2], hi[r][3]); ph[1] = make_uint4(hi[r][4], hi[r][5], hi[r][6], hi[r][7]);
                    pl[0] = make_uint4(lo[r][0], lo[r][1], lo[r][2], lo[r][3]); pl[1] = make_uint4(lo[r][4], lo[r][5], lo[r][6], lo[r][7]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {   // values (map rows 40-59), 8 columns = one 16-byte chunk at a time; the last chunk holds 4 values + padding
            uint32_t hi[TM_ROWS][4], lo[TM_ROWS][4];
#pragma unroll
            for (int d = 0; d < 4; d++) {
                if (4 * i + d < TM_EMB / 2) {
                    float v0[TM_ROWS], v1[TM_ROWS];
                    affine(40 + 8 * i + 2 * d, v0); affine(41 + 8 * i + 2 * d, v1);
#pragma unroll
                    for (int r = 0; r < TM_ROWS; r++) tm_split2(v0[r], v1[r], hi[r][d], lo[r][d]);
                } else {
#pragma unroll
                    for (int r = 0; r < TM_ROWS; r++) hi[r][d] = lo[r][d] = 0u;
                }
            }
            if (on) {
#pragma unroll
                for (int r = 0; r < TM_ROWS; r++) {
                    uint8_t* tile = tiles + r * TM_TILE_BYTES;
                    reinterpret_cast<uint4*>(tile + TM_OFF_VH + a * TM_PITCH)[i] = make_uint4(hi[r][0], hi[r][1], hi[r][2], hi[r][3]);
                    reinterpret_cast<uint4*>(tile + TM_OFF_VL + a * TM_PITCH)[i] = make_uint4(lo[r][0], lo[r][1], lo[r][2], lo[r][3]);
                }
            }
        }
#pragma unroll
        for (int d4 = 0; d4 < TM_EMB / 4; d4++) {   // the token itself (map rows 0-19), fp32: the residual
            float t0[TM_ROWS], t1[TM_ROWS], t2[TM_ROWS], t3[TM_ROWS];
            affine(4 * d4, t0); affine(4 * d4 + 1, t1); affine(4 * d4 + 2, t2); affine(4 * d4 + 3, t3);
            if (on) {
#pragma unroll
                for (int r = 0; r < TM_ROWS; r++) reinterpret_cast<float4*>(tiles + r * TM_TILE_BYTES + TM_OFF_TOK + a * TM_EMB * 4)[d4] = make_float4(t0[r], t1[r], t2[r], t3[r]);
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int r = 0; r < TM_ROWS; r++) {
        const int row = row0 + r;
        if (row >= nrows) break;
        uint8_t* tile = tiles + r * TM_TILE_BYTES;
        const uint32_t t_u32 = (uint32_t)__cvta_generic_to_shared(tile);
        // ---------------- phase B: scores S[32 x 24] = Q K^T (three fp16 products), rows = query tokens
        float sc[2][3][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++)
#pragma unroll
                for (int i = 0; i < 4; i++) sc[mt][nt][i] = 0.f;
        {
            uint32_t kh[3][2], kl[3][2], dummy0, dummy1;
            tm_ldsm4(t_u32 + TM_OFF_KH + k_off, kh[0][0], kh[0][1], kh[1][0], kh[1][1]);
            tm_ldsm4(t_u32 + TM_OFF_KL + k_off, kl[0][0], kl[0][1], kl[1][0], kl[1][1]);
            // third n8 tile (tokens 16-23): matrices (n 16-23, k 0-7), (n 16-23, k 8-15); lanes 16-31 repeat those addresses
            const uint32_t k2_off = (uint32_t)((16 + (lane & 7)) * TM_PITCH + ((lane >> 3) & 1) * 16);
            tm_ldsm4(t_u32 + TM_OFF_KH + k2_off, kh[2][0], kh[2][1], dummy0, dummy1);
            tm_ldsm4(t_u32 + TM_OFF_KL + k2_off, kl[2][0], kl[2][1], dummy0, dummy1);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                uint32_t qh[4], ql[4];
                tm_ldsm4(t_u32 + TM_OFF_QH + mt * 16 * TM_PITCH + q_off, qh[0], qh[1], qh[2], qh[3]);
                tm_ldsm4(t_u32 + TM_OFF_QL + mt * 16 * TM_PITCH + q_off, ql[0], ql[1], ql[2], ql[3]);
#pragma unroll
                for (int nt = 0; nt < 3; nt++) {
                    tm_mma16(sc[mt][nt], qh[0], qh[1], qh[2], qh[3], kh[nt][0], kh[nt][1]);
                    tm_mma16(sc[mt][nt], ql[0], ql[1], ql[2], ql[3], kh[nt][0], kh[nt][1]);
                    tm_mma16(sc[mt][nt], qh[0], qh[1], qh[2], qh[3], kl[nt][0], kl[nt][1]);
                }
            }
        }
        // softmax over the 23 key tokens of every query row: accumulator registers 0,1 = row g, 2,3 = row g + 8; columns 8 nt + 2 t4 + {0, 1}
        uint32_t ph16[2][4], pl16[2][4], ph8[2][2], pl8[2][2];   // probabilities as A fragments: k16 step (tokens 0-15), k8 step (tokens 16-23)
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {   // h = 0: row g, h = 1: row g + 8
                float v[6];
#pragma unroll
                for (int nt = 0; nt < 3; nt++) { v[2 * nt] = sc[mt][nt][2 * h] * 0.31622776601683794f; v[2 * nt + 1] = sc[mt][nt][2 * h + 1] * 0.31622776601683794f; }
                if (t4 == 3) v[5] = -1e30f;   // column 23 is padding (finite: tm_exp's residual term must not see an infinity); exp underflows to 0
                float m = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(v[4], v[5]));
                m = fmaxf(m, __shfl_xor_sync(kFull, m, 1)); m = fmaxf(m, __shfl_xor_sync(kFull, m, 2));
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 6; i++) { v[i] = tm_exp(v[i] - m); s += v[i]; }
                s += __shfl_xor_sync(kFull, s, 1); s += __shfl_xor_sync(kFull, s, 2);
                const float inv = 1.f / s;
#pragma unroll
                for (int i = 0; i < 6; i++) v[i] *= inv;
                tm_split2(v[0], v[1], ph16[mt][h], pl16[mt][h]);           // a0 / a1: k 2 t4 .. of tokens 0-7
                tm_split2(v[2], v[3], ph16[mt][2 + h], pl16[mt][2 + h]);   // a2 / a3: tokens 8-15
                tm_split2(v[4], v[5], ph8[mt][h], pl8[mt][h]);             // tokens 16-23
            }
        }
        // context O[32 x 24] = P V, then the residual and the store; one n8 tile of the embedding at a time
#pragma unroll
        for (int nt = 0; nt < 3; nt++) {
            uint32_t vh[4], vl[4];
            tm_ldsm4t(t_u32 + TM_OFF_VH + v_off + nt * 16, vh[0], vh[1], vh[2], vh[3]);
            tm_ldsm4t(t_u32 + TM_OFF_VL + v_off + nt * 16, vl[0], vl[1], vl[2], vl[3]);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                float oc[4] = {0.f, 0.f, 0.f, 0.f};
                tm_mma16(oc, ph16[mt][0], ph16[mt][1], ph16[mt][2], ph16[mt][3], vh[0], vh[1]);
                tm_mma16(oc, pl16[mt][0], pl16[mt][1], pl16[mt][2], pl16[mt][3], vh[0], vh[1]);
                tm_mma16(oc, ph16[mt][0], ph16[mt][1], ph16[mt][2], ph16[mt][3], vl[0], vl[1]);
                tm_mma8(oc, ph8[mt][0], ph8[mt][1], vh[2]);
                tm_mma8(oc, pl8[mt][0], pl8[mt][1], vh[2]);
                tm_mma8(oc, ph8[mt][0], ph8[mt][1], vl[2]);
                const int d = nt * 8 + 2 * t4;
                if (d < TM_EMB) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int tok = mt * 16 + g + 8 * h;
                        if (tok < TM_TOK) {
                            const float2 tk = *reinterpret_cast<const float2*>(tile + TM_OFF_TOK + (tok * TM_EMB + d) * 4);
                            *reinterpret_cast<float2*>(x0 + (size_t)row * TM_X0 + tok * TM_EMB + d) = make_float2(tk.x + oc[2 * h], tk.y + oc[2 * h + 1]);
                        }
                    }
                }
            }
        }
        }
        __syncwarp();  // the tiles are rewritten by the warp's next rows
    }
}

cudaError_t launch_tokens_mma(const float* wts, const float* obs, int R, float* x0, int off_tokm, int off_tokb, int off_col, int off_dim, cudaStream_t stream) {
    static PerDeviceFlag configured;
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_tokens_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES);
        if (e != cudaSuccess) { configured.retract(); return e; }
    }
    TokOffsets o{off_tokm, off_tokb, off_col, off_dim};
    const int blocks = (R + TM_WARPS * TM_ROWS - 1) / (TM_WARPS * TM_ROWS);
    k_tokens_mma<<<blocks < 148 * MM_TOKM_MINBLOCKS ? blocks : 148 * MM_TOKM_MINBLOCKS, TM_WARPS * 32, TM_SMEM_BYTES, stream>>>(obs, wts, o, x0, R);
    return cudaGetLastError();
}

}  // namespace mm
