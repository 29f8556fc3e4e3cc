// mm_tokens_proj.cu -- K4 stage 1, third generation: the 23-token projection + self-attention + residual of one agent row (networks.py:58-65,
// 75-82) with EVERY matrix product of the attention on the warp-level tensor path.  sm_100a.
//
// The second generation (k_tokens_mma, mm_tokens_mma.cu) evaluated the per-token affine maps [token | key | query | value] = M_a x_a + b_a (60 outputs
// of <= 4 inputs, the attention weights folded in on the host) on the FMA pipe and only the two attention products on mma.sync.  ncu
// (profiles/r03b_k4_tokens_mma_summary.txt): bound by shared-memory wavefronts -- 83 % of peak, 443 per row, 240 of them the map fetches (each lane reads
// its own 60 x 5 floats for every row), 56 more to pass keys / queries / values through shared tiles.  Here only the 20 token outputs are evaluated from
// the maps; keys, queries and values are what the reference says they are -- products of the token matrix with Wk, Wq, Wv (networks.py:76-78) -- on the
// tensor path, and they never touch shared memory:
//   T [32 x 32] (tokens x embedding, zero padded)  -> fp16 hi / lo tiles (the only shared-memory round trip)
//   Q = T Wq^T / sqrt(10), K = T Wk^T               : A = T by ldmatrix, B = weight fragments (block prologue); the m16n8 accumulator layout of Q IS the
//                                                     A layout of the score product, that of K its B layout (token g, dims 2t, 2t+1)
//   S = Q K^T -> softmax on the accumulator fragments -> P as A fragments (as in the second generation)
//   C = P T                                          : B = T by ldmatrix.trans (the value product commutes: sum_j p_ij Wv t_j = Wv sum_j p_ij t_j)
//   O = C Wv^T ; x0 = T + O                          : A = the accumulators of C, B = weight fragments; T of the residual from an fp32 copy (exact)
// Every product is the error-compensated fp16 split (x ~ hi + lo: hi.hi + lo.hi + hi.lo, fp32 accumulation), operands pre-scaled by powers of two so
// that the lo halves stay normal fp16 numbers.  Per row: ~200 shared-memory wavefronts instead of 443, ~140 HMMA instead of 54, ~300 fewer FMA-pipe
// instructions.
#include <cuda_fp16.h>
#include "mm_env.cuh"

namespace mm {

constexpr int TP_TOK = 23, TP_EMB = 20, TP_KQ = 10, TP_X0 = TP_TOK * TP_EMB;
#ifndef MM_TOKP_WARPS
#define MM_TOKP_WARPS 8
#endif
#ifndef MM_TOKP_HOIST
#define MM_TOKP_HOIST 0   // (1 =) query / key weight fragments held in registers across rows (24 words) instead of re-read from shared memory per row
#endif
#ifndef MM_TOKP_MINBLOCKS
#define MM_TOKP_MINBLOCKS 2
#endif
constexpr int TP_WARPS = MM_TOKP_WARPS;
constexpr int TP_PITCH = 80;                                    // bytes per token row of the T tiles: 32 halves + 16 bytes (conflict-free ldmatrix)
constexpr int TP_T_BYTES = 32 * TP_PITCH;                       // one of (hi, lo)
constexpr int TP_TOKF_OFF = 2 * TP_T_BYTES;                     // the tokens again as fp32 [23][20]: the residual of x0 = T + O is exact
constexpr int TP_TILE_BYTES = TP_TOKF_OFF + 24 * TP_EMB * 4;
constexpr int TP_MAP_FLOATS = TP_EMB * TP_TOK * 5;              // rows 0-19 of tokm [60][23][4] and of tokb [60][23]
constexpr int TP_NFRAG = 42;                                    // Wq 2 n-tiles, Wk 2, Wv 3; 6 words each: hi {k 0-15 lo half, k 0-15 hi half, k 16-23}, lo {..}
constexpr int TP_SMEM_BYTES = TP_MAP_FLOATS * 4 + TP_NFRAG * 32 * 4 + TP_WARPS * TP_TILE_BYTES;
// powers of two that keep hi AND lo of every operand in fp16's normal range: tokens x 16, weights x 64, queries / keys x 16, probabilities x 256
constexpr float TP_TS = 16.f, TP_WS = 64.f, TP_QS = 16.f, TP_PS = 256.f;

struct TokPOffsets { int tokm, tokb, proj_col, proj_dim, att_q, att_k, att_v; };

// x ~ hi + lo, both fp16.  MM_TOKP_SPLIT_RZ = 1: hi = x truncated to fp16's 11 significant bits -- packed by a round-toward-zero conversion, and taken as
// a float by masking the fp32 mantissa (an ALU op) instead of converting the half back (a conversion-pipe op: with ~50 splits per row the kernel was
// bound by that pipe, ncu: math-pipe throttle the top stall); lo = fp16(x - hi) is then at most 2^-10 |x| instead of 2^-11 |x|, its own rounding 2^-21 |x|.
// Exact for every x in fp16's normal range (the operands are pre-scaled into it); below it the two forms of hi differ by < 2^-24 absolute.
#ifndef MM_TOKP_SPLIT_RZ
#define MM_TOKP_SPLIT_RZ 0   // measured: 0.267 vs 0.261 ms -- the half -> float conversions were not the limiter; kept as an option
#endif
__device__ __forceinline__ void tp_split2(float x, float y, uint32_t& hi, uint32_t& lo) {
#if MM_TOKP_SPLIT_RZ
    asm("cvt.rz.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));   // first source -> upper half
    const float hx = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u), hy = __uint_as_float(__float_as_uint(y) & 0xFFFFE000u);
    const __half2 l2 = __floats2half2_rn(x - hx, y - hy);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
#else
    const __half2 h2 = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
#endif
}
__device__ __forceinline__ void tp_ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void tp_ldsm2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void tp_ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void tp_mma16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tp_mma8(float* c, const uint32_t* a, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b0));
}
// The mma wrappers are plain (non-volatile) asm: pure functions of their operands, so the compiler is free to interleave the independent accumulator
// chains of the two m16 tiles (six dependent HMMAs per chain otherwise issue back to back and wait out each other's latency).
// c += A B for A = (a16 | a8) over k = 0-15 | 16-23 given as hi / lo fragments and B = one weight n8 tile given as its six fragment words
__device__ __forceinline__ void tp_mma_split(float* c, const uint32_t* ah16, const uint32_t* al16, const uint32_t* ah8, const uint32_t* al8, const uint32_t* f) {
    tp_mma16(c, al16, f[0], f[1]);   // lo . hi
    tp_mma16(c, ah16, f[3], f[4]);   // hi . lo
    tp_mma16(c, ah16, f[0], f[1]);   // hi . hi
    tp_mma8(c, al8, f[2]);
    tp_mma8(c, ah8, f[5]);
    tp_mma8(c, ah8, f[2]);
}

__global__ void __launch_bounds__(TP_WARPS * 32, MM_TOKP_MINBLOCKS) k_tokens_proj(const float* __restrict__ obs, const float* __restrict__ wts, const TokPOffsets o,
                                                                                    float* __restrict__ x0, int nrows) {
    extern __shared__ __align__(16) uint8_t tp_smem[];
    float (*s_m)[TP_TOK][4] = reinterpret_cast<float (*)[TP_TOK][4]>(tp_smem);                        // [20][23][4]
    float (*s_b)[TP_TOK] = reinterpret_cast<float (*)[TP_TOK]>(tp_smem + TP_EMB * TP_TOK * 16);       // [20][23]
    uint32_t* s_f = reinterpret_cast<uint32_t*>(tp_smem + TP_MAP_FLOATS * 4);                          // [42][32]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tile = tp_smem + TP_MAP_FLOATS * 4 + TP_NFRAG * 32 * 4 + w * TP_TILE_BYTES;
    for (int i = threadIdx.x; i < TP_EMB * TP_TOK * 4; i += blockDim.x) (&s_m[0][0][0])[i] = wts[o.tokm + i];   // the first 20 of the 60 map rows
    for (int i = threadIdx.x; i < TP_EMB * TP_TOK; i += blockDim.x) (&s_b[0][0])[i] = wts[o.tokb + i];
    // weight fragments.  B[k = e][n = d] = W[d][e] ("col" operand): word 0 = (e = 2t, 2t+1), 1 = (e = 2t+8, 2t+9), 2 = (e = 16+2t, 17+2t) of row d = 8 nt + g
    for (int i = threadIdx.x; i < 7 * 3 * 32; i += blockDim.x) {
        const int ln = i & 31, j = (i >> 5) % 3, tl = (i >> 5) / 3;           // tl: 0-1 Wq, 2-3 Wk, 4-6 Wv
        const int mat = tl < 2 ? 0 : tl < 4 ? 1 : 2, nt = tl < 2 ? tl : tl < 4 ? tl - 2 : tl - 4;
        const int d = 8 * nt + (ln >> 2), e = (j == 0 ? 0 : j == 1 ? 8 : 16) + 2 * (ln & 3);
        const int rows = mat == 2 ? TP_EMB : TP_KQ;
        const float* W = wts + (mat == 0 ? o.att_q : mat == 1 ? o.att_k : o.att_v);
        const float sc = (mat == 0 ? 0.31622776601683794f : 1.f) * TP_WS;      // queries carry the 1 / sqrt(kq_dim) of networks.py:79
        const float w0 = (d < rows && e < TP_EMB) ? W[d * TP_EMB + e] * sc : 0.f, w1 = (d < rows && e + 1 < TP_EMB) ? W[d * TP_EMB + e + 1] * sc : 0.f;
        uint32_t hi, lo;
        tp_split2(w0, w1, hi, lo);
        s_f[(tl * 6 + j) * 32 + ln] = hi;
        s_f[(tl * 6 + 3 + j) * 32 + ln] = lo;
    }
    for (int i = lane; i < TP_TILE_BYTES / 16; i += 32) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0u, 0u, 0u, 0u);   // padding rows / columns stay zero
    __syncthreads();
    const bool on = lane < TP_TOK;
    const int a = on ? lane : 0;
    const int c0 = (int)wts[o.proj_col + a], nd = (int)wts[o.proj_dim + a];
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t t_hi = (uint32_t)__cvta_generic_to_shared(tile), t_lo = t_hi + TP_T_BYTES;
    // ldmatrix row addresses.  A (x4): matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15) of an m16 tile
    const uint32_t a16_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * TP_PITCH + (lane >> 4) * 16);
    // A (x2, the k = 16-23 step): matrices (rows 0-7), (rows 8-15) at columns 16-23; lanes 16-31 repeat valid addresses
    const uint32_t a8_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * TP_PITCH + 32);
    // B of the context product (x4.trans): matrices (tokens 0-7), (8-15), (16-23), (16-23 again, unused), one n8 tile of the embedding per load
    const uint32_t bt_off = (uint32_t)(((lane & 7) + min(lane >> 3, 2) * 8) * TP_PITCH);
    auto frag6 = [&](int tl, uint32_t* f) {
#pragma unroll
        for (int j = 0; j < 6; j++) f[j] = s_f[(tl * 6 + j) * 32 + lane];
    };
    // the query / key weight fragments stay in registers for the whole kernel (24 words); the value fragments are re-read per row
#if MM_TOKP_HOIST
    uint32_t fq[2][6], fk[2][6];
#pragma unroll
    for (int nt = 0; nt < 2; nt++) { frag6(nt, fq[nt]); frag6(2 + nt, fk[nt]); }
#endif
    // the observation columns of a row are fetched one row ahead (a warp has nothing else in flight while it waits for them)
    const int row_stride = gridDim.x * TP_WARPS;
    float xn[4];
    {
        const int row = blockIdx.x * TP_WARPS + w;
#pragma unroll
        for (int c = 0; c < 4; c++) xn[c] = (on && c < nd && row < nrows) ? obs[(size_t)row * kObs + c0 + c] : 0.f;
    }
#pragma unroll 1
    for (int row = blockIdx.x * TP_WARPS + w; row < nrows; row += row_stride) {
        // ---------------- tokens: lane = token, t = P_a x_a + b_a (networks.py:58-65), written as fp16 hi / lo of 16 t
        {
            float x[4];
#pragma unroll
            for (int c = 0; c < 4; c++) x[c] = xn[c];
#pragma unroll
            for (int c = 0; c < 4; c++) xn[c] = (on && c < nd && row + row_stride < nrows) ? obs[(size_t)(row + row_stride) * kObs + c0 + c] : 0.f;
            uint32_t hi[10], lo[10];
            float tf[TP_EMB];
#pragma unroll
            for (int e2 = 0; e2 < TP_EMB / 2; e2++) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const float4 m = *reinterpret_cast<const float4*>(&s_m[2 * e2 + q][a][0]);
                    tf[2 * e2 + q] = fmaf(x[3], m.w, fmaf(x[2], m.z, fmaf(x[1], m.y, fmaf(x[0], m.x, s_b[2 * e2 + q][a]))));
                }
                tp_split2(tf[2 * e2] * TP_TS, tf[2 * e2 + 1] * TP_TS, hi[e2], lo[e2]);
            }
            if (on) {
                float4* rf = reinterpret_cast<float4*>(tile + TP_TOKF_OFF + a * TP_EMB * 4);
#pragma unroll
                for (int q = 0; q < TP_EMB / 4; q++) rf[q] = make_float4(tf[4 * q], tf[4 * q + 1], tf[4 * q + 2], tf[4 * q + 3]);
                uint8_t* rh = tile + a * TP_PITCH;
                uint8_t* rl = rh + TP_T_BYTES;
                reinterpret_cast<uint4*>(rh)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); reinterpret_cast<uint4*>(rh)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                reinterpret_cast<uint2*>(rh)[4] = make_uint2(hi[8], hi[9]);
                reinterpret_cast<uint4*>(rl)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); reinterpret_cast<uint4*>(rl)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                reinterpret_cast<uint2*>(rl)[4] = make_uint2(lo[8], lo[9]);
            }
        }
        __syncwarp();
        // ---------------- T as A fragments (both m16 tiles: tokens 0-15, 16-31)
        uint32_t th16[2][4], tl16[2][4], th8[2][2], tl8[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            tp_ldsm4(t_hi + mt * 16 * TP_PITCH + a16_off, th16[mt][0], th16[mt][1], th16[mt][2], th16[mt][3]);
            tp_ldsm4(t_lo + mt * 16 * TP_PITCH + a16_off, tl16[mt][0], tl16[mt][1], tl16[mt][2], tl16[mt][3]);
            tp_ldsm2(t_hi + mt * 16 * TP_PITCH + a8_off, th8[mt][0], th8[mt][1]);
            tp_ldsm2(t_lo + mt * 16 * TP_PITCH + a8_off, tl8[mt][0], tl8[mt][1]);
        }
        // ---------------- Q = T Wq^T / sqrt(10) -> A fragments of the score product; K = T Wk^T -> its B fragments.  Accumulators hold 16 * 64 * value.
        uint32_t qh[2][4], ql[2][4];     // [m tile of query tokens][a0..a3]
        uint32_t kh[4][2], kl[4][2];     // [n8 tile of key tokens = 2 mt + h][b0, b1]
        constexpr float kProjToOp = TP_QS / (TP_TS * TP_WS);
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
#if !MM_TOKP_HOIST
            uint32_t fq[2][6], fk[2][6];
            frag6(nt, fq[nt]); frag6(2 + nt, fk[nt]);
#endif
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                tp_mma_split(c, th16[mt], tl16[mt], th8[mt], tl8[mt], fq[nt]);
                tp_split2(c[0] * kProjToOp, c[1] * kProjToOp, qh[mt][2 * nt], ql[mt][2 * nt]);           // row g:     a0 (k 0-7) | a2 (k 8-15)
                tp_split2(c[2] * kProjToOp, c[3] * kProjToOp, qh[mt][2 * nt + 1], ql[mt][2 * nt + 1]);   // row g + 8: a1 | a3
            }
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                tp_mma_split(c, th16[mt], tl16[mt], th8[mt], tl8[mt], fk[nt]);
                tp_split2(c[0] * kProjToOp, c[1] * kProjToOp, kh[2 * mt][nt], kl[2 * mt][nt]);           // key tokens 16 mt + g: b0 (k 0-7) | b1 (k 8-15)
                tp_split2(c[2] * kProjToOp, c[3] * kProjToOp, kh[2 * mt + 1][nt], kl[2 * mt + 1][nt]);   // key tokens 16 mt + 8 + g
            }
        }
        // ---------------- scores S[32 x 24] = Q K^T (accumulators hold 256 S), rows = query tokens
        float sc[2][3][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) {
#pragma unroll
                for (int i = 0; i < 4; i++) sc[mt][nt][i] = 0.f;
                tp_mma16(sc[mt][nt], ql[mt], kh[nt][0], kh[nt][1]);
                tp_mma16(sc[mt][nt], qh[mt], kl[nt][0], kl[nt][1]);
                tp_mma16(sc[mt][nt], qh[mt], kh[nt][0], kh[nt][1]);
            }
        // softmax over the 23 key tokens of every query row: accumulator registers 0,1 = row g, 2,3 = row g + 8; columns 8 nt + 2 t4 + {0, 1}
        uint32_t ph16[2][4], pl16[2][4], ph8[2][2], pl8[2][2];   // 256 P as A fragments: k16 step (tokens 0-15), k8 step (tokens 16-23)
        // Scores leave the accumulators already in the log2 domain (x log2 e folded into the power-of-two unscaling) and the exponential is ONE ex2.approx:
        // expf's careful range reduction cost ~10 instructions x 24 per lane per row, a fifth of the kernel.  The argument's rounding (2^-24 |arg|) puts a
        // relative error of |arg| 4e-8 on exp(arg), i.e. an ABSOLUTE error <= 4e-8 |arg| e^arg <= 1.5e-8 on a probability -- below the split's own 2^-22.
        constexpr float kScoreScale = 1.4426950408889634f / (TP_QS * TP_QS);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {   // h = 0: row g, h = 1: row g + 8
                float v[6];
#pragma unroll
                for (int nt = 0; nt < 3; nt++) { v[2 * nt] = sc[mt][nt][2 * h] * kScoreScale; v[2 * nt + 1] = sc[mt][nt][2 * h + 1] * kScoreScale; }
                if (t4 == 3) v[5] = -1e30f;   // column 23 is padding; exp underflows to 0
                float m = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(v[4], v[5]));
                m = fmaxf(m, __shfl_xor_sync(kFull, m, 1)); m = fmaxf(m, __shfl_xor_sync(kFull, m, 2));
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 6; i++) { asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[i]) : "f"(v[i] - m)); s += v[i]; }
                s += __shfl_xor_sync(kFull, s, 1); s += __shfl_xor_sync(kFull, s, 2);
                const float inv = TP_PS / s;
#pragma unroll
                for (int i = 0; i < 6; i++) v[i] *= inv;
                tp_split2(v[0], v[1], ph16[mt][h], pl16[mt][h]);           // a0 / a1: key tokens 0-7
                tp_split2(v[2], v[3], ph16[mt][2 + h], pl16[mt][2 + h]);   // a2 / a3: key tokens 8-15
                tp_split2(v[4], v[5], ph8[mt][h], pl8[mt][h]);             // key tokens 16-23
            }
        }
        // ---------------- C[32 x 24] = P T (accumulators hold 256 * 16 C) -> A fragments of 16 C
        uint32_t ch16[2][4], cl16[2][4], ch8[2][2], cl8[2][2];
        constexpr float kCtxToOp = 1.f / TP_PS;
#pragma unroll
        for (int nt = 0; nt < 3; nt++) {
            uint32_t bh[4], bl[4];
            tp_ldsm4t(t_hi + bt_off + nt * 16, bh[0], bh[1], bh[2], bh[3]);
            tp_ldsm4t(t_lo + bt_off + nt * 16, bl[0], bl[1], bl[2], bl[3]);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                tp_mma16(c, pl16[mt], bh[0], bh[1]);
                tp_mma16(c, ph16[mt], bl[0], bl[1]);
                tp_mma16(c, ph16[mt], bh[0], bh[1]);
                tp_mma8(c, pl8[mt], bh[2]);
                tp_mma8(c, ph8[mt], bl[2]);
                tp_mma8(c, ph8[mt], bh[2]);
                if (nt < 2) {
                    tp_split2(c[0] * kCtxToOp, c[1] * kCtxToOp, ch16[mt][2 * nt], cl16[mt][2 * nt]);
                    tp_split2(c[2] * kCtxToOp, c[3] * kCtxToOp, ch16[mt][2 * nt + 1], cl16[mt][2 * nt + 1]);
                } else {
                    tp_split2(c[0] * kCtxToOp, c[1] * kCtxToOp, ch8[mt][0], cl8[mt][0]);
                    tp_split2(c[2] * kCtxToOp, c[3] * kCtxToOp, ch8[mt][1], cl8[mt][1]);
                }
            }
        }
        // ---------------- O = C Wv^T (accumulators hold 16 * 64 O); x0 = T + O with T from the fp32 tile, one n8 tile of the embedding at a time.
        // (The residual as T . 64 I on the same accumulator -- the token fragments are at hand -- saved ~30 instructions per row, but a token then enters
        // x0 as hi + lo, 2^-22 relative off: 3.6e-7 instead of 2.4e-7 worst-case absolute error of x0, enough to show in a 10-step Adam trajectory test.)
        constexpr float kOutScale = 1.f / (TP_TS * TP_WS);
#pragma unroll
        for (int nt = 0; nt < 3; nt++) {
            uint32_t f[6];
            frag6(4 + nt, f);
            const int d = nt * 8 + 2 * t4;
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                float oc[4] = {0.f, 0.f, 0.f, 0.f};
                tp_mma_split(oc, ch16[mt], cl16[mt], ch8[mt], cl8[mt], f);
                if (d < TP_EMB) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int tok = mt * 16 + g + 8 * h;
                        if (tok < TP_TOK) {
                            const float2 tk = *reinterpret_cast<const float2*>(tile + TP_TOKF_OFF + (tok * TP_EMB + d) * 4);
                            *reinterpret_cast<float2*>(x0 + (size_t)row * TP_X0 + tok * TP_EMB + d) =
                                make_float2(fmaf(oc[2 * h], kOutScale, tk.x), fmaf(oc[2 * h + 1], kOutScale, tk.y));
                        }
                    }
                }
            }
        }
        __syncwarp();  // the tile is rewritten by the warp's next row
    }
}

cudaError_t launch_tokens_proj(const float* wts, const float* obs, int R, float* x0, int off_tokm, int off_tokb, int off_col, int off_dim, int off_q, int off_k, int off_v,
                               cudaStream_t stream) {
    static PerDeviceFlag configured;
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_tokens_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES);
        if (e != cudaSuccess) { configured.retract(); return e; }
    }
    TokPOffsets o{off_tokm, off_tokb, off_col, off_dim, off_q, off_k, off_v};
    const int blocks = (R + TP_WARPS - 1) / TP_WARPS;
    k_tokens_proj<<<blocks < 148 * MM_TOKP_MINBLOCKS ? blocks : 148 * MM_TOKP_MINBLOCKS, TP_WARPS * 32, TP_SMEM_BYTES, stream>>>(obs, wts, o, x0, R);
    return cudaGetLastError();
}

}  // namespace mm
