// mm_policy.cu -- K4: actor / critic forward over the whole environment batch with action sampling fused behind it.
//
// Replaces Actor.forward (+Projection, m_Attention), Critic.forward (networks.py:31-41,58-65,75-82,96-102) and
// PPO.get_action (PPO.py:170-186) for the rollout.  fp32 throughout (parity bar: 1e-5 relative on logits, values and
// log-probs, which single-pass bf16/tf32 tensor-core math cannot meet; the error-compensated tcgen05 path is
// DESIGN.md section 9 "next").  Stages, all on the caller's stream:
//   k_tokens   : obs [R,65] -> projection (23 tokens x 20) -> self-attention + residual -> x0 [R,460]
//   k_linear   : Y = relu(X W^T + b), smem-tiled SGEMM (460->264, 264->264, 264->264)
//   k_heads    : 5 move logits + 1 mark logit, mask, Categorical / Bernoulli sample (Philox) or evaluate given actions,
//                joint log-prob per env (sum over the two agents, PPO.py:118,121)
//   k_critic   : centralised critic [E,130] -> 64 -> 64 -> 1
// Weights arrive as ONE flat fp32 buffer laid out by mm_policy_offsets() (host packs it from the state_dict).
#include <cuda_fp16.h>
#include "mm_env.cuh"
#include "mm_policy_heads.cuh"

namespace mm {

constexpr int kTok = 23, kEmb = 20, kKQ = 10, kX0 = kTok * kEmb;  // 460
constexpr int kHid = 264, kCH = 64;

// flat weight buffer layout (floats)
struct PolicyOffsets {
    int proj_w, proj_b, proj_col, proj_dim;  // [23][20][4], [23][20], [23] (as float), [23] (as float)
    int att_k, att_q, att_v;                 // [10][20], [10][20], [20][20]
    int l0_w, l0_b, l1_w, l1_b, l2_w, l2_b;  // [264][460],[264] ; [264][264],[264] x2
    int head_w, head_b;                      // [6][264] (5 move rows then the mark row), [6]
    int c0_w, c0_b, c1_w, c1_b, c2_w, c2_b;  // critic [64][130],[64],[64][64],[64],[1][64],[1]
    int l0_whi, l0_wlo, l1_whi, l1_wlo, l2_whi, l2_wlo;  // TF32 hi/lo splits of the three trunk weights (tensor-core path)
    int c0_wt, c1_wt;                                    // critic weights transposed: [130][64], [64][64] (coalesced lane = neuron reads)
    int tokm, tokb;                                      // per-token affine maps [60][23][4], [60][23]: rows 0-19 token, 20-29 key, 30-39 query, 40-59 value
    // FP16 hi/lo splits of 2^e * W for the three trunk layers (mm_linear16.cu): fp16 [264][kpad] each, kpad = K rounded up to 32 (480, 288,
    // 288; zero padded), stored in this fp32 buffer (two halves per float slot); l*_asc = the device scalar 2^-e
    int l0_h16, l0_l16, l1_h16, l1_l16, l2_h16, l2_l16, l0_asc, l1_asc, l2_asc;
    // the six head rows (5 move + 1 mark) as a fourth "layer" of the fused trunk kernel: fp16 hi / lo of 2^e [move_head; mark_head], [16][288] (rows 6.. and
    // columns 264.. zero), and the scalar 2^-e
    int lh_h16, lh_l16, lh_asc;
    int total;
};
constexpr int kPad0 = 480, kPad1 = 288;   // K = 460 / 264 rounded up to the 32-element k-block
__host__ __device__ inline PolicyOffsets policy_offsets() {
    PolicyOffsets o; int p = 0;
    auto take = [&](int n) { int r = p; p += (n + 3) & ~3; return r; };  // keep every block 16-byte aligned
    o.proj_w = take(kTok * kEmb * 4); o.proj_b = take(kTok * kEmb); o.proj_col = take(kTok); o.proj_dim = take(kTok);
    o.att_k = take(kKQ * kEmb); o.att_q = take(kKQ * kEmb); o.att_v = take(kEmb * kEmb);
    o.l0_w = take(kHid * kX0); o.l0_b = take(kHid); o.l1_w = take(kHid * kHid); o.l1_b = take(kHid); o.l2_w = take(kHid * kHid); o.l2_b = take(kHid);
    o.head_w = take(6 * kHid); o.head_b = take(6);
    o.c0_w = take(kCH * 130); o.c0_b = take(kCH); o.c1_w = take(kCH * kCH); o.c1_b = take(kCH); o.c2_w = take(kCH); o.c2_b = take(1);
    o.l0_whi = take(kHid * kX0); o.l0_wlo = take(kHid * kX0); o.l1_whi = take(kHid * kHid); o.l1_wlo = take(kHid * kHid);
    o.l2_whi = take(kHid * kHid); o.l2_wlo = take(kHid * kHid);
    o.c0_wt = take(130 * kCH); o.c1_wt = take(kCH * kCH);
    o.tokm = take(60 * kTok * 4); o.tokb = take(60 * kTok);
    o.l0_h16 = take(kHid * kPad0 / 2); o.l0_l16 = take(kHid * kPad0 / 2); o.l1_h16 = take(kHid * kPad1 / 2); o.l1_l16 = take(kHid * kPad1 / 2);
    o.l2_h16 = take(kHid * kPad1 / 2); o.l2_l16 = take(kHid * kPad1 / 2); o.l0_asc = take(1); o.l1_asc = take(1); o.l2_asc = take(1);
    o.lh_h16 = take(16 * kPad1 / 2); o.lh_l16 = take(16 * kPad1 / 2); o.lh_asc = take(1);
    o.total = p;
    return o;
}

// ------------------------------------------------------------------------------------------------ tokens + attention
// One warp per row, one LANE per feature token (23 of 32 lanes active).  A token is an affine map of <= 4 observation columns
// (networks.py:58-65), so its key, query and value are too: the host folds Wk, Wq, Wv into per-token [60 x 4] maps (rows 0-19
// token, 20-29 key, 30-39 query, 40-59 value) and the lane evaluates all 60 outputs with 240 FMAs instead of 80 + 800.  Keys and
// values of the row's 23 tokens are exchanged through a small per-warp shared-memory tile (broadcast reads).
constexpr int kTokWarps = 4;
#ifndef MM_TOK_MINBLOCKS
#define MM_TOK_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(kTokWarps * 32, MM_TOK_MINBLOCKS) k_tokens(const float* __restrict__ obs, const float* __restrict__ wts, float* __restrict__ x0, int R) {
    const PolicyOffsets o = policy_offsets();
    __shared__ __align__(16) float s_m[60][kTok][4];
    __shared__ float s_b[60][kTok];
    __shared__ __align__(16) float s_k[kTokWarps][kTok][12], s_v[kTokWarps][kTok][kEmb];
    for (int i = threadIdx.x; i < 60 * kTok * 4; i += blockDim.x) (&s_m[0][0][0])[i] = wts[o.tokm + i];
    for (int i = threadIdx.x; i < 60 * kTok; i += blockDim.x) (&s_b[0][0])[i] = wts[o.tokb + i];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool on = lane < kTok;
    const int a = on ? lane : 0;
    const int c0 = (int)wts[o.proj_col + a], nd = (int)wts[o.proj_dim + a];
    // persistent blocks: the 27 KB of per-token maps are staged once per block, then the block's warps stride over the rows
#pragma unroll 1
    for (int row = blockIdx.x * kTokWarps + w; row < R; row += gridDim.x * kTokWarps) {
    float x[4];
#pragma unroll
    for (int c = 0; c < 4; c++) x[c] = (c < nd) ? obs[(size_t)row * kObs + c0 + c] : 0.f;
    auto affine = [&](int j) {
        const float4 m = *reinterpret_cast<const float4*>(&s_m[j][a][0]);
        return fmaf(x[3], m.w, fmaf(x[2], m.z, fmaf(x[1], m.y, fmaf(x[0], m.x, s_b[j][a]))));
    };
    {   // keys and values go straight to the row's shared tile
        float kk[kKQ], vv[kEmb];
#pragma unroll
        for (int d = 0; d < kKQ; d++) kk[d] = affine(20 + d);
#pragma unroll
        for (int d = 0; d < kEmb; d++) vv[d] = affine(40 + d);
        if (on) {
#pragma unroll
            for (int d = 0; d < kKQ; d++) s_k[w][a][d] = kk[d];
#pragma unroll
            for (int d4 = 0; d4 < kEmb / 4; d4++) *reinterpret_cast<float4*>(&s_v[w][a][4 * d4]) = make_float4(vv[4 * d4], vv[4 * d4 + 1], vv[4 * d4 + 2], vv[4 * d4 + 3]);
        }
    }
    float q[kKQ];
#pragma unroll
    for (int d = 0; d < kKQ; d++) q[d] = affine(30 + d);
    __syncwarp();
    float p[kTok];
    float m = -INFINITY;
#pragma unroll
    for (int b = 0; b < kTok; b++) {  // scores of this token's query against every key of the row (networks.py:79)
        const float4 k0 = *reinterpret_cast<const float4*>(&s_k[w][b][0]), k1 = *reinterpret_cast<const float4*>(&s_k[w][b][4]);
        const float2 k2 = *reinterpret_cast<const float2*>(&s_k[w][b][8]);
        float acc = q[0] * k0.x;
        acc = fmaf(q[1], k0.y, acc); acc = fmaf(q[2], k0.z, acc); acc = fmaf(q[3], k0.w, acc);
        acc = fmaf(q[4], k1.x, acc); acc = fmaf(q[5], k1.y, acc); acc = fmaf(q[6], k1.z, acc); acc = fmaf(q[7], k1.w, acc);
        acc = fmaf(q[8], k2.x, acc); acc = fmaf(q[9], k2.y, acc);
        p[b] = acc * 0.31622776601683794f;  // 1/sqrt(10)
        m = fmaxf(m, p[b]);
    }
    float sum = 0.f;
#pragma unroll
#ifdef MM_TOK_FAST_EXP
    for (int b = 0; b < kTok; b++) { p[b] = __expf(p[b] - m); sum += p[b]; }
#else
    for (int b = 0; b < kTok; b++) { p[b] = expf(p[b] - m); sum += p[b]; }
#endif
    const float inv = 1.f / sum;
    float ctx[kEmb];
#pragma unroll
    for (int d = 0; d < kEmb; d++) ctx[d] = 0.f;
#pragma unroll
    for (int b = 0; b < kTok; b++) {
        const float pb = p[b] * inv;
#pragma unroll
        for (int d4 = 0; d4 < kEmb / 4; d4++) {
            const float4 vv = *reinterpret_cast<const float4*>(&s_v[w][b][4 * d4]);
            ctx[4 * d4] = fmaf(pb, vv.x, ctx[4 * d4]); ctx[4 * d4 + 1] = fmaf(pb, vv.y, ctx[4 * d4 + 1]);
            ctx[4 * d4 + 2] = fmaf(pb, vv.z, ctx[4 * d4 + 2]); ctx[4 * d4 + 3] = fmaf(pb, vv.w, ctx[4 * d4 + 3]);
        }
    }
    if (on) {  // residual (networks.py:82); lane a owns columns 20a .. 20a+19 of the row
        float4* oh = reinterpret_cast<float4*>(x0 + (size_t)row * kX0 + a * kEmb);
#pragma unroll
        for (int d4 = 0; d4 < kEmb / 4; d4++)  // the token itself is evaluated last to keep it out of the register budget above
            oh[d4] = make_float4(affine(4 * d4) + ctx[4 * d4], affine(4 * d4 + 1) + ctx[4 * d4 + 1], affine(4 * d4 + 2) + ctx[4 * d4 + 2], affine(4 * d4 + 3) + ctx[4 * d4 + 3]);
    }
    __syncwarp();  // the row's key / value tile is reused by the next row of this warp
    }
}

// ------------------------------------------------------------------------------------------------ tokens + attention, R rows per warp
// k_tokens is bound by the shared-memory data pipe (ncu, profiles/r01h: LSU wavefronts 88.5 % of peak, 620 per row): 240 of them fetch the
// per-token affine maps (23 lanes x 16 B = 3 wavefronts per float4), 184 are the broadcast key / value reads of the attention, ~90 the key /
// value tile stores (scalar key stores at a 12-word lane stride conflict 3-way).  This version carries R rows through the AFFINE phase together --
// every map element fetched from shared memory serves R rows -- while the attention phase still runs one row at a time (its registers, p[23] and
// ctx[20], are not duplicated), and stores keys as float4 / float4 / float2 (conflict-free at that stride): 184 + 240 / R + ~24 wavefronts per row.
#ifndef MM_TOK_ROWS
#define MM_TOK_ROWS 2
#endif
#ifndef MM_TOK_R_MINBLOCKS
#define MM_TOK_R_MINBLOCKS (MM_TOK_ROWS <= 2 ? 4 : 3)
#endif
// floats per token in the per-warp key / value tile: 12 key (10 used) + 20 value = 32, padded to 36 -- at a 128-byte lane stride the
// float4 stores of the 23 lanes all fell into the same four banks (ncu, profiles/r02h: 21 M bank conflicts, 186 store wavefronts per row
// instead of 24); 144 bytes puts the 8 lanes of a quarter-warp on 8 different 16-byte bank groups
#ifndef MM_TOK_TSTRIDE
#define MM_TOK_TSTRIDE 36
#endif
constexpr int kTokTS = MM_TOK_TSTRIDE;
template <int R>
__global__ void __launch_bounds__(kTokWarps * 32, MM_TOK_R_MINBLOCKS) k_tokens_r(const float* __restrict__ obs, const float* __restrict__ wts, float* __restrict__ x0, int nrows) {
    const PolicyOffsets o = policy_offsets();
    extern __shared__ __align__(16) float tk_smem[];
    float (*s_m)[kTok][4] = reinterpret_cast<float (*)[kTok][4]>(tk_smem);                        // [60][23][4]
    float (*s_b)[kTok] = reinterpret_cast<float (*)[kTok]>(tk_smem + 60 * kTok * 4);               // [60][23]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tk_smem + 60 * kTok * 5 + w * (R * kTok * kTokTS);                                   // per warp: R rows x 23 tokens x (12 key + 20 value)
    for (int i = threadIdx.x; i < 60 * kTok * 4; i += blockDim.x) (&s_m[0][0][0])[i] = wts[o.tokm + i];
    for (int i = threadIdx.x; i < 60 * kTok; i += blockDim.x) (&s_b[0][0])[i] = wts[o.tokb + i];
    __syncthreads();
    const bool on = lane < kTok;
    const int a = on ? lane : 0;
    const int c0 = (int)wts[o.proj_col + a], nd = (int)wts[o.proj_dim + a];
#pragma unroll 1
    for (int row0 = (blockIdx.x * kTokWarps + w) * R; row0 < nrows; row0 += gridDim.x * kTokWarps * R) {
        float x[R][4];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) x[r][c] = (c < nd && row0 + r < nrows) ? obs[(size_t)(row0 + r) * kObs + c0 + c] : 0.f;
        // ---- affine phase: one shared-memory fetch of map row j serves all R rows
        float q[R][kKQ], tok[R][kEmb];
        {
            float kk[R][12];
#pragma unroll
            for (int d = 0; d < kKQ; d++) {
                const float4 m = *reinterpret_cast<const float4*>(&s_m[20 + d][a][0]);
                const float b = s_b[20 + d][a];
#pragma unroll
                for (int r = 0; r < R; r++) kk[r][d] = fmaf(x[r][3], m.w, fmaf(x[r][2], m.z, fmaf(x[r][1], m.y, fmaf(x[r][0], m.x, b))));
            }
            if (on) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float* kt = tile + (r * kTok + a) * kTokTS;
                    *reinterpret_cast<float4*>(kt) = make_float4(kk[r][0], kk[r][1], kk[r][2], kk[r][3]);
                    *reinterpret_cast<float4*>(kt + 4) = make_float4(kk[r][4], kk[r][5], kk[r][6], kk[r][7]);
                    *reinterpret_cast<float2*>(kt + 8) = make_float2(kk[r][8], kk[r][9]);
                }
            }
        }
#pragma unroll
        for (int d4 = 0; d4 < kEmb / 4; d4++) {
            float vv[R][4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 m = *reinterpret_cast<const float4*>(&s_m[40 + 4 * d4 + i][a][0]);
                const float b = s_b[40 + 4 * d4 + i][a];
#pragma unroll
                for (int r = 0; r < R; r++) vv[r][i] = fmaf(x[r][3], m.w, fmaf(x[r][2], m.z, fmaf(x[r][1], m.y, fmaf(x[r][0], m.x, b))));
            }
            if (on) {
#pragma unroll
                for (int r = 0; r < R; r++) *reinterpret_cast<float4*>(tile + (r * kTok + a) * kTokTS + 12 + 4 * d4) = make_float4(vv[r][0], vv[r][1], vv[r][2], vv[r][3]);
            }
        }
#pragma unroll
        for (int d = 0; d < kKQ; d++) {
            const float4 m = *reinterpret_cast<const float4*>(&s_m[30 + d][a][0]);
            const float b = s_b[30 + d][a];
#pragma unroll
            for (int r = 0; r < R; r++) q[r][d] = fmaf(x[r][3], m.w, fmaf(x[r][2], m.z, fmaf(x[r][1], m.y, fmaf(x[r][0], m.x, b))));
        }
#pragma unroll
        for (int d = 0; d < kEmb; d++) {
            const float4 m = *reinterpret_cast<const float4*>(&s_m[d][a][0]);
            const float b = s_b[d][a];
#pragma unroll
            for (int r = 0; r < R; r++) tok[r][d] = fmaf(x[r][3], m.w, fmaf(x[r][2], m.z, fmaf(x[r][1], m.y, fmaf(x[r][0], m.x, b))));
        }
        __syncwarp();
        // ---- attention phase, one row at a time (networks.py:75-82)
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float* kt = tile + r * kTok * kTokTS;
            float p[kTok];
            float mx = -INFINITY;
#pragma unroll
            for (int b = 0; b < kTok; b++) {
                const float4 k0 = *reinterpret_cast<const float4*>(kt + b * kTokTS), k1 = *reinterpret_cast<const float4*>(kt + b * kTokTS + 4);
                const float2 k2 = *reinterpret_cast<const float2*>(kt + b * kTokTS + 8);
                float acc = q[r][0] * k0.x;
                acc = fmaf(q[r][1], k0.y, acc); acc = fmaf(q[r][2], k0.z, acc); acc = fmaf(q[r][3], k0.w, acc);
                acc = fmaf(q[r][4], k1.x, acc); acc = fmaf(q[r][5], k1.y, acc); acc = fmaf(q[r][6], k1.z, acc); acc = fmaf(q[r][7], k1.w, acc);
                acc = fmaf(q[r][8], k2.x, acc); acc = fmaf(q[r][9], k2.y, acc);
                p[b] = acc * 0.31622776601683794f;  // 1/sqrt(10)
                mx = fmaxf(mx, p[b]);
            }
            float sum = 0.f;
#pragma unroll
            for (int b = 0; b < kTok; b++) { p[b] = expf(p[b] - mx); sum += p[b]; }
            const float inv = 1.f / sum;
            float ctx[kEmb];
#pragma unroll
            for (int d = 0; d < kEmb; d++) ctx[d] = 0.f;
#pragma unroll
            for (int b = 0; b < kTok; b++) {
                const float pb = p[b] * inv;
#pragma unroll
                for (int d4 = 0; d4 < kEmb / 4; d4++) {
                    const float4 vv = *reinterpret_cast<const float4*>(kt + b * kTokTS + 12 + 4 * d4);
                    ctx[4 * d4] = fmaf(pb, vv.x, ctx[4 * d4]); ctx[4 * d4 + 1] = fmaf(pb, vv.y, ctx[4 * d4 + 1]);
                    ctx[4 * d4 + 2] = fmaf(pb, vv.z, ctx[4 * d4 + 2]); ctx[4 * d4 + 3] = fmaf(pb, vv.w, ctx[4 * d4 + 3]);
                }
            }
            if (on && row0 + r < nrows) {  // residual (networks.py:82); lane a owns columns 20a .. 20a+19 of the row
                float4* oh = reinterpret_cast<float4*>(x0 + (size_t)(row0 + r) * kX0 + a * kEmb);
#pragma unroll
                for (int d4 = 0; d4 < kEmb / 4; d4++)
                    oh[d4] = make_float4(tok[r][4 * d4] + ctx[4 * d4], tok[r][4 * d4 + 1] + ctx[4 * d4 + 1], tok[r][4 * d4 + 2] + ctx[4 * d4 + 2], tok[r][4 * d4 + 3] + ctx[4 * d4 + 3]);
            }
        }
        __syncwarp();  // the tiles are rewritten by the next R rows of this warp
    }
}
constexpr int kTokRSmem = (60 * kTok * 5 + kTokWarps * MM_TOK_ROWS * kTok * kTokTS) * (int)sizeof(float);
// -DMM_TOK_MMA=0 keeps the SIMT token kernel, 1 the second generation (attention on HMMA, mm_tokens_mma.cu) for A/B runs; the default is the third
// (keys / queries / values as tensor-path products of the token tile, mm_tokens_proj.cu)
#ifndef MM_TOK_MMA
#define MM_TOK_MMA 2
#endif
cudaError_t launch_tokens_mma(const float* wts, const float* obs, int R, float* x0, int off_tokm, int off_tokb, int off_col, int off_dim, cudaStream_t stream);
cudaError_t launch_tokens_proj(const float* wts, const float* obs, int R, float* x0, int off_tokm, int off_tokb, int off_col, int off_dim, int off_q, int off_k, int off_v,
                               cudaStream_t stream);
static cudaError_t launch_tokens_any(const float* wts, const float* obs, int R, float* x0, cudaStream_t stream) {
#if MM_TOK_MMA == 2   // third generation: keys / queries / values as tensor-path products of the token tile (mm_tokens_proj.cu)
    const PolicyOffsets o = policy_offsets();
    return launch_tokens_proj(wts, obs, R, x0, o.tokm, o.tokb, o.proj_col, o.proj_dim, o.att_q, o.att_k, o.att_v, stream);
#elif MM_TOK_MMA
    const PolicyOffsets o = policy_offsets();
    return launch_tokens_mma(wts, obs, R, x0, o.tokm, o.tokb, o.proj_col, o.proj_dim, stream);
#elif MM_TOK_ROWS >= 2
    static PerDeviceFlag configured;
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_tokens_r<MM_TOK_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTokRSmem);
        if (e != cudaSuccess) { configured.retract(); return e; }
    }
    const int per_block = kTokWarps * MM_TOK_ROWS;
    k_tokens_r<MM_TOK_ROWS><<<min((R + per_block - 1) / per_block, 148 * 8), kTokWarps * 32, kTokRSmem, stream>>>(obs, wts, x0, R);
#else
    k_tokens<<<min((R + kTokWarps - 1) / kTokWarps, 148 * 8), kTokWarps * 32, 0, stream>>>(obs, wts, x0, R);
#endif
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ tokens + attention, backward
// d loss / d (per-token affine maps) from d loss / d x0 (K5: the embedding's backward for the general -- indexed -- projection, where every
// agent row has its own 23 tokens).  Same mapping as k_tokens: one warp per row, one lane per token, keys / queries / values of the row
// exchanged through a per-warp shared tile.  With y_a = M_a x_a + b_a = (token, key, query, value) of token a (60 outputs of <= 4 inputs):
//     out_a = token_a + sum_b P[a][b] value_b,  P = softmax_b(query_a . key_b / sqrt(10))            (networks.py:75-82)
//     d token_a = dO_a;   dP[a][b] = dO_a . value_b;   d value_b = sum_a P[a][b] dO_a
//     dS[a][b] = P[a][b] (dP[a][b] - sum_b' P[a][b'] dP[a][b']) / sqrt(10);   d query_a = sum_b dS[a][b] key_b;   d key_b = sum_a dS[a][b] query_a
// Lane a owns row a of P / dS (as the query) and, after a transposing pass through shared memory, column a (as key / value).  The map
// gradients dM_a[j][c] = sum_rows dy_a[j] x_a[c], db_a[j] = sum_rows dy_a[j] are a reduction over ALL rows of 300 values per token: the
// row kernel writes dy (key, query, value parts: [R][23][40]; the token part is dO itself) and k_tokens_bwd_reduce sums the outer
// products with one thread per (token, 4 outputs) and 20 register accumulators (shared-memory atomics in the row kernel -- a CAS loop
// per float -- made it 11x the forward).  The host adds the block partials and back-propagates through the folding
// [I; Wk; Wq; Wv] (P_a, b_a) with autograd.
constexpr int kTbWarps = 4;
constexpr int kTbTile = kTok * (12 + 12 + kEmb + kEmb + 24 + 24);                                  // per-warp floats: k, q, v, dO, P, dS
constexpr int kTbSmem = (60 * kTok * 4 + 60 * kTok + kTbWarps * kTbTile) * (int)sizeof(float);
__global__ void __launch_bounds__(kTbWarps * 32, 2) k_tokens_bwd(const float* __restrict__ obs, const float* __restrict__ wts, const float* __restrict__ dout,
                                                                 float* __restrict__ dy40, int R) {
    const PolicyOffsets o = policy_offsets();
    extern __shared__ __align__(16) float tb_smem[];
    float (*s_m)[kTok][4] = reinterpret_cast<float (*)[kTok][4]>(tb_smem);
    float (*s_b)[kTok] = reinterpret_cast<float (*)[kTok]>(tb_smem + 60 * kTok * 4);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tb_smem + 60 * kTok * 5 + w * kTbTile;
    float (*s_k)[12] = reinterpret_cast<float (*)[12]>(tile);
    float (*s_q)[12] = reinterpret_cast<float (*)[12]>(tile + kTok * 12);
    float (*s_v)[kEmb] = reinterpret_cast<float (*)[kEmb]>(tile + kTok * 24);
    float (*s_do)[kEmb] = reinterpret_cast<float (*)[kEmb]>(tile + kTok * (24 + kEmb));
    float (*s_p)[24] = reinterpret_cast<float (*)[24]>(tile + kTok * (24 + 2 * kEmb));
    float (*s_ds)[24] = reinterpret_cast<float (*)[24]>(tile + kTok * (48 + 2 * kEmb));
    for (int i = threadIdx.x; i < 60 * kTok * 4; i += blockDim.x) (&s_m[0][0][0])[i] = wts[o.tokm + i];
    for (int i = threadIdx.x; i < 60 * kTok; i += blockDim.x) (&s_b[0][0])[i] = wts[o.tokb + i];
    __syncthreads();
    const bool on = lane < kTok;
    const int a = on ? lane : 0;
    const int c0 = (int)wts[o.proj_col + a], nd = (int)wts[o.proj_dim + a];
#pragma unroll 1
    for (int row = blockIdx.x * kTbWarps + w; row < R; row += gridDim.x * kTbWarps) {
        float x[4];
#pragma unroll
        for (int c = 0; c < 4; c++) x[c] = (c < nd) ? obs[(size_t)row * kObs + c0 + c] : 0.f;
        auto affine = [&](int j) {
            const float4 m = *reinterpret_cast<const float4*>(&s_m[j][a][0]);
            return fmaf(x[3], m.w, fmaf(x[2], m.z, fmaf(x[1], m.y, fmaf(x[0], m.x, s_b[j][a]))));
        };
        float q[kKQ], dO[kEmb];
        {
            const float4* gp = reinterpret_cast<const float4*>(dout + (size_t)row * kX0 + a * kEmb);
#pragma unroll
            for (int d4 = 0; d4 < kEmb / 4; d4++) {
                const float4 g = on ? __ldg(gp + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
                dO[4 * d4] = g.x; dO[4 * d4 + 1] = g.y; dO[4 * d4 + 2] = g.z; dO[4 * d4 + 3] = g.w;
            }
        }
#pragma unroll
        for (int d = 0; d < kKQ; d++) q[d] = affine(30 + d);
        if (on) {
#pragma unroll
            for (int d = 0; d < kKQ; d++) { s_k[a][d] = affine(20 + d); s_q[a][d] = q[d]; }
#pragma unroll
            for (int d4 = 0; d4 < kEmb / 4; d4++) {
                *reinterpret_cast<float4*>(&s_v[a][4 * d4]) = make_float4(affine(40 + 4 * d4), affine(41 + 4 * d4), affine(42 + 4 * d4), affine(43 + 4 * d4));
                *reinterpret_cast<float4*>(&s_do[a][4 * d4]) = make_float4(dO[4 * d4], dO[4 * d4 + 1], dO[4 * d4 + 2], dO[4 * d4 + 3]);
            }
        }
        __syncwarp();
        // ---- row a of P, dP, dS; d query_a
        float p[kTok], m = -INFINITY;
#pragma unroll
        for (int b = 0; b < kTok; b++) {
            const float4 k0 = *reinterpret_cast<const float4*>(&s_k[b][0]), k1 = *reinterpret_cast<const float4*>(&s_k[b][4]);
            const float2 k2 = *reinterpret_cast<const float2*>(&s_k[b][8]);
            float acc = q[0] * k0.x;
            acc = fmaf(q[1], k0.y, acc); acc = fmaf(q[2], k0.z, acc); acc = fmaf(q[3], k0.w, acc);
            acc = fmaf(q[4], k1.x, acc); acc = fmaf(q[5], k1.y, acc); acc = fmaf(q[6], k1.z, acc); acc = fmaf(q[7], k1.w, acc);
            acc = fmaf(q[8], k2.x, acc); acc = fmaf(q[9], k2.y, acc);
            p[b] = acc * 0.31622776601683794f;
            m = fmaxf(m, p[b]);
        }
        float sum = 0.f;
#pragma unroll
        for (int b = 0; b < kTok; b++) { p[b] = expf(p[b] - m); sum += p[b]; }
        const float inv = 1.f / sum;
        float dot = 0.f, ds[kTok];
#pragma unroll
        for (int b = 0; b < kTok; b++) {
            p[b] *= inv;
            float dp = 0.f;
#pragma unroll
            for (int d4 = 0; d4 < kEmb / 4; d4++) {
                const float4 vv = *reinterpret_cast<const float4*>(&s_v[b][4 * d4]);
                dp = fmaf(dO[4 * d4], vv.x, dp); dp = fmaf(dO[4 * d4 + 1], vv.y, dp); dp = fmaf(dO[4 * d4 + 2], vv.z, dp); dp = fmaf(dO[4 * d4 + 3], vv.w, dp);
            }
            ds[b] = dp;
            dot = fmaf(p[b], dp, dot);
        }
        float* dyr = dy40 + ((size_t)row * kTok + a) * 40;  // this token's (d key 10 | d query 10 | d value 20)
        float dy_q[kKQ];
#pragma unroll
        for (int d = 0; d < kKQ; d++) dy_q[d] = 0.f;
#pragma unroll
        for (int b = 0; b < kTok; b++) {
            ds[b] = p[b] * (ds[b] - dot) * 0.31622776601683794f;
            const float4 k0 = *reinterpret_cast<const float4*>(&s_k[b][0]), k1 = *reinterpret_cast<const float4*>(&s_k[b][4]);
            const float2 k2 = *reinterpret_cast<const float2*>(&s_k[b][8]);
            dy_q[0] = fmaf(ds[b], k0.x, dy_q[0]); dy_q[1] = fmaf(ds[b], k0.y, dy_q[1]); dy_q[2] = fmaf(ds[b], k0.z, dy_q[2]); dy_q[3] = fmaf(ds[b], k0.w, dy_q[3]);
            dy_q[4] = fmaf(ds[b], k1.x, dy_q[4]); dy_q[5] = fmaf(ds[b], k1.y, dy_q[5]); dy_q[6] = fmaf(ds[b], k1.z, dy_q[6]); dy_q[7] = fmaf(ds[b], k1.w, dy_q[7]);
            dy_q[8] = fmaf(ds[b], k2.x, dy_q[8]); dy_q[9] = fmaf(ds[b], k2.y, dy_q[9]);
            if (on) { s_p[a][b] = p[b]; s_ds[a][b] = ds[b]; }
        }
        if (on) {
#pragma unroll
            for (int d = 0; d < kKQ; d += 2) *reinterpret_cast<float2*>(dyr + 10 + d) = make_float2(dy_q[d], dy_q[d + 1]);
        }
        __syncwarp();
        // ---- column a of dS and P: d key_a, then d value_a
        {
            float dy_k[kKQ];
#pragma unroll
            for (int d = 0; d < kKQ; d++) dy_k[d] = 0.f;
#pragma unroll 1
            for (int t = 0; t < kTok; t++) {  // t = the query token
                const float dst = s_ds[t][a];
                const float4 q0 = *reinterpret_cast<const float4*>(&s_q[t][0]), q1 = *reinterpret_cast<const float4*>(&s_q[t][4]);
                const float2 q2 = *reinterpret_cast<const float2*>(&s_q[t][8]);
                dy_k[0] = fmaf(dst, q0.x, dy_k[0]); dy_k[1] = fmaf(dst, q0.y, dy_k[1]); dy_k[2] = fmaf(dst, q0.z, dy_k[2]); dy_k[3] = fmaf(dst, q0.w, dy_k[3]);
                dy_k[4] = fmaf(dst, q1.x, dy_k[4]); dy_k[5] = fmaf(dst, q1.y, dy_k[5]); dy_k[6] = fmaf(dst, q1.z, dy_k[6]); dy_k[7] = fmaf(dst, q1.w, dy_k[7]);
                dy_k[8] = fmaf(dst, q2.x, dy_k[8]); dy_k[9] = fmaf(dst, q2.y, dy_k[9]);
            }
            if (on) {
#pragma unroll
                for (int d = 0; d < kKQ; d += 2) *reinterpret_cast<float2*>(dyr + d) = make_float2(dy_k[d], dy_k[d + 1]);
            }
        }
        {
            float dy_v[kEmb];
#pragma unroll
            for (int d = 0; d < kEmb; d++) dy_v[d] = 0.f;
#pragma unroll 1
            for (int t = 0; t < kTok; t++) {
                const float pt = s_p[t][a];
#pragma unroll
                for (int d4 = 0; d4 < kEmb / 4; d4++) {
                    const float4 g = *reinterpret_cast<const float4*>(&s_do[t][4 * d4]);
                    dy_v[4 * d4] = fmaf(pt, g.x, dy_v[4 * d4]); dy_v[4 * d4 + 1] = fmaf(pt, g.y, dy_v[4 * d4 + 1]);
                    dy_v[4 * d4 + 2] = fmaf(pt, g.z, dy_v[4 * d4 + 2]); dy_v[4 * d4 + 3] = fmaf(pt, g.w, dy_v[4 * d4 + 3]);
                }
            }
            if (on) {
#pragma unroll
                for (int d4 = 0; d4 < kEmb / 4; d4++) *reinterpret_cast<float4*>(dyr + 20 + 4 * d4) = make_float4(dy_v[4 * d4], dy_v[4 * d4 + 1], dy_v[4 * d4 + 2], dy_v[4 * d4 + 3]);
            }
        }
        __syncwarp();  // the row's tiles are reused by the next row of this warp
    }
}

// part[block][j][a][0..3] = sum over the block's rows of dy_a[j] x_a[c], part[block][j][a][4] = sum of dy_a[j]; thread = (token a, 4 outputs):
// groups 0-4 are the token part (dy = dO, read from dout), groups 5-14 the key | query | value parts (dy40).
constexpr int kTrThreads = kTok * 15;  // 345
__global__ void __launch_bounds__(kTrThreads) k_tokens_bwd_reduce(const float* __restrict__ obs, const float* __restrict__ wts, const float* __restrict__ dout,
                                                                  const float* __restrict__ dy40, float* __restrict__ part, int R, int rows_per_block) {
    const PolicyOffsets o = policy_offsets();
    const int a = threadIdx.x / 15, g4 = threadIdx.x % 15;
    const int c0 = (int)wts[o.proj_col + a], nd = (int)wts[o.proj_dim + a];
    const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
    float acc[4][5];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 5; c++) acc[i][c] = 0.f;
    const bool tok = g4 < 5;
    const size_t off = tok ? (size_t)a * kEmb + 4 * g4 : (size_t)a * 40 + 4 * (g4 - 5);
    const size_t pitch = tok ? (size_t)kX0 : (size_t)kTok * 40;
    const float* src = tok ? dout : dy40;
#pragma unroll 4
    for (int r = r0; r < r1; r++) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * pitch + off));
        float x[4];
#pragma unroll
        for (int c = 0; c < 4; c++) x[c] = (c < nd) ? __ldg(obs + (size_t)r * kObs + c0 + c) : 0.f;
        const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int c = 0; c < 4; c++) acc[i][c] = fmaf(dv[i], x[c], acc[i][c]);
            acc[i][4] += dv[i];
        }
    }
    const int j0 = tok ? 4 * g4 : 20 + 4 * (g4 - 5);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 5; c++) part[(((size_t)blockIdx.x * 60 + j0 + i) * kTok + a) * 5 + c] = acc[i][c];
}

int tokens_bwd_blocks() { return 148 * 4; }
size_t tokens_bwd_scratch_floats(int R) { return (size_t)R * kTok * 40; }
// the stand-alone forward is a function of the per-token maps alone (its caller, update.TokenEmbed, fills only those blocks of the buffer and its backward
// differentiates exactly this function): second-generation kernel.  _full: whatever the rollout uses, given the full weight buffer.
cudaError_t launch_tokens_fwd(const float* wts, const float* obs, int R, float* x0, cudaStream_t stream) {
#if MM_TOK_MMA
    const PolicyOffsets o = policy_offsets();
    return launch_tokens_mma(wts, obs, R, x0, o.tokm, o.tokb, o.proj_col, o.proj_dim, stream);
#else
    return launch_tokens_any(wts, obs, R, x0, stream);
#endif
}
cudaError_t launch_tokens_fwd_full(const float* wts, const float* obs, int R, float* x0, cudaStream_t stream) { return launch_tokens_any(wts, obs, R, x0, stream); }
cudaError_t launch_tokens_bwd(const float* wts, const float* obs, const float* dout, int R, float* dy40, float* part, cudaStream_t stream) {
    static PerDeviceFlag configured;
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_tokens_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kTbSmem);
        if (e != cudaSuccess) { configured.retract(); return e; }
    }
    k_tokens_bwd<<<min((R + kTbWarps - 1) / kTbWarps, 148 * 6), kTbWarps * 32, kTbSmem, stream>>>(obs, wts, dout, dy40, R);
    const int blocks = tokens_bwd_blocks();
    k_tokens_bwd_reduce<<<blocks, kTrThreads, 0, stream>>>(obs, wts, dout, dy40, part, R, (R + blocks - 1) / blocks);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ Y = relu(X W^T + b)
// 128x64 output tile per block, 256 threads, 8x4 outputs per thread, K in slabs of 16 through shared memory.
constexpr int BM = 128, BN = 64, BK = 16;
__global__ void __launch_bounds__(256) k_linear_relu(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                                                     float* __restrict__ Y, int M, int N, int K) {
    __shared__ float As[BK][BM + 4], Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tm = (tid >> 4) * 8, tn = (tid & 15) * 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += BK) {
        for (int i = tid; i < BM * BK; i += 256) {  // X tile: consecutive threads walk k (contiguous in memory)
            const int r = i / BK, c = i - r * BK;
            const int gm = m0 + r, gk = k0 + c;
            As[c][r] = (gm < M && gk < K) ? X[(size_t)gm * K + gk] : 0.f;
        }
        for (int i = tid; i < BN * BK; i += 256) {
            const int r = i / BK, c = i - r * BK;
            const int gn = n0 + r, gk = k0 + c;
            Bs[c][r] = (gn < N && gk < K) ? W[(size_t)gn * K + gk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = As[k][tm + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[k][tn + j];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int gm = m0 + tm + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gn = n0 + tn + j;
            if (gn < N) Y[(size_t)gm * N + gn] = fmaxf(acc[i][j] + bias[gn], 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------------ heads + sampling
// One warp per ENV (both agents), lanes split the 264-long dot products; lane 0 finishes the distribution math.
__global__ void __launch_bounds__(128) k_heads(const float* __restrict__ h, const float* __restrict__ wts, const HeadArgs ha) {
    const PolicyOffsets o = policy_offsets();
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= ha.E) return;
    float lp_env = 0.f;
    for (int a = 0; a < 2; a++) {
        const float* hr = h + ((size_t)w * 2 + a) * kHid;
        float l[6];
#pragma unroll
        for (int j = 0; j < 6; j++) {
            float acc = 0.f;
            for (int k = lane; k < kHid; k += 32) acc = fmaf(hr[k], wts[o.head_w + j * kHid + k], acc);
#pragma unroll
            for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(kFull, acc, s);
            l[j] = acc + wts[o.head_b + j];
        }
        if (lane == 0) lp_env += head_sample_or_eval(l, (long long)w * 2 + a, ha);
    }
    if (lane == 0) ha.logp[w] = lp_env;
}

// ------------------------------------------------------------------------------------------------ critic
// 130 -> 64 -> 64 -> 1 (networks.py:84-104).
// Third generation: both hidden layers on the warp-level tensor path (mma.sync m16n8k16, fp16 operands, fp32 accumulation) with the same
// error-compensated split as the trunk (x ~ hi + lo: hi.hi + lo.hi + hi.lo, ~2^-22 relative per product).  Why not SIMT: a warp-wide 16-byte shared-memory
// read returns 512 bytes through a 128-byte/clock pipe, so any register tiling that fits (lane = neuron pair x 4 envs: 76 us per 65 536 envs; lane = env x
// 64 neurons from broadcast weight quads: 68 us, kept under -DMM_CRITIC_SIMT) is bound by that pipe at ~1/3 of the FMA rate; on the tensor path a weight
// fragment (16 bytes per lane) feeds 6 HMMAs.
// One warp = 32 environments (two m16 tiles).  Layer 0's A fragments are read straight from the fp32 observation rows (float2 per lane: 130 is even, a
// 32-byte sector serves the four lanes of a row group), scaled by 2^8 and split in registers; the accumulator layout of an m16n8 tile IS the A layout of
// the next k16 step, so layer 1's operand (relu(acc + b), scaled by 2^4) never leaves the registers; the 64 -> 1 output layer is an fp32 dot product on the
// accumulator fragments, reduced over the lane quad.  Weights: every block builds the B fragments (hi | lo fp16 of 2^e W, e from the layer's largest
// |w|) in its shared memory once -- 52 KB, persistent blocks of 16 warps, one per SM.  Domain: |obs| < 255, hidden activations < 4094 (fp16 range after the
// scaling); the reference's observations lie in [-1, 2].
#ifdef MM_CRITIC_SIMT
constexpr int kCrIn = 130, kCrWarps = 8, kCrPitch = 131;
constexpr int kCrWFloats = kCrIn * kCH + kCH * kCH + 3 * kCH + 4;            // W0t, W1t, b0, b1, w2, b2 (+ padding)
constexpr int kCrSmemBytes = (kCrWFloats + kCrWarps * 32 * kCrPitch) * 4;

// acc[n] = b[n] + sum_k x[k] W[k][n], k ascending; x[k] read from the lane's tile row, W as 16-byte broadcast quads
template <int K>
__device__ __forceinline__ void critic_layer(float (&acc)[kCH], const float* __restrict__ xrow, const float4* __restrict__ w4, const float* __restrict__ b) {
#pragma unroll
    for (int n = 0; n < kCH; n++) acc[n] = b[n];
#pragma unroll 2
    for (int k = 0; k < K; k++) {
        const float x = xrow[k];
#pragma unroll
        for (int j = 0; j < kCH / 4; j++) {
            const float4 w = w4[k * (kCH / 4) + j];
            acc[4 * j] = fmaf(x, w.x, acc[4 * j]); acc[4 * j + 1] = fmaf(x, w.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(x, w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(x, w.w, acc[4 * j + 3]);
        }
    }
}

__global__ void __launch_bounds__(kCrWarps * 32, 1) k_critic(const float* __restrict__ obs, const float* __restrict__ wts, float* __restrict__ value, int E) {
    extern __shared__ __align__(16) float cr_smem[];
    const PolicyOffsets o = policy_offsets();
    float* s_w0 = cr_smem;                       // [130][64]
    float* s_w1 = s_w0 + kCrIn * kCH;            // [64][64]
    float* s_b0 = s_w1 + kCH * kCH;
    float* s_b1 = s_b0 + kCH;
    float* s_w2 = s_b1 + kCH;
    float* s_b2 = s_w2 + kCH;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = cr_smem + kCrWFloats + w * 32 * kCrPitch;
    for (int i = threadIdx.x; i < kCrIn * kCH / 4; i += blockDim.x) reinterpret_cast<float4*>(s_w0)[i] = reinterpret_cast<const float4*>(wts + o.c0_wt)[i];
    for (int i = threadIdx.x; i < kCH * kCH / 4; i += blockDim.x) reinterpret_cast<float4*>(s_w1)[i] = reinterpret_cast<const float4*>(wts + o.c1_wt)[i];
    if (threadIdx.x < kCH) { s_b0[threadIdx.x] = wts[o.c0_b + threadIdx.x]; s_b1[threadIdx.x] = wts[o.c1_b + threadIdx.x]; s_w2[threadIdx.x] = wts[o.c2_w + threadIdx.x]; }
    if (threadIdx.x == 0) s_b2[0] = wts[o.c2_b];
    __syncthreads();
    const int ntiles = (E + 31) >> 5;
    // tile t belongs to block t % gridDim.x; inside the block the warps take the block's tiles in turn
#pragma unroll 1
    for (int t = blockIdx.x + w * gridDim.x; t < ntiles; t += kCrWarps * gridDim.x) {
        const int e0 = t << 5;
        const int ne = min(32, E - e0);
        // stage [ne][130] contiguous floats as [32][131]: float2 granules never straddle an environment (130 is even)
        const float2* src = reinterpret_cast<const float2*>(obs + (size_t)e0 * kCrIn);
        __syncwarp();
#pragma unroll 5
        for (int i = lane; i < 32 * (kCrIn / 2); i += 32) {
            const int e = i / (kCrIn / 2), k2 = i - e * (kCrIn / 2);
            const float2 v = e < ne ? __ldg(src + i) : make_float2(0.f, 0.f);
            tile[e * kCrPitch + 2 * k2] = v.x; tile[e * kCrPitch + 2 * k2 + 1] = v.y;
        }
        __syncwarp();
        float* xrow = tile + lane * kCrPitch;
        float acc[kCH];
        critic_layer<kCrIn>(acc, xrow, reinterpret_cast<const float4*>(s_w0), s_b0);
#pragma unroll
        for (int n = 0; n < kCH; n++) xrow[n] = fmaxf(acc[n], 0.f);   // the lane's own row: no other lane reads it
        critic_layer<kCH>(acc, xrow, reinterpret_cast<const float4*>(s_w1), s_b1);
        float v = s_b2[0];
#pragma unroll
        for (int n = 0; n < kCH; n++) v = fmaf(fmaxf(acc[n], 0.f), s_w2[n], v);
        if (lane < ne) value[e0 + lane] = v;
    }
}
#else
#ifndef MM_CRITIC_WARPS
#define MM_CRITIC_WARPS 16
#endif
constexpr int kCrIn = 130, kCrWarps = MM_CRITIC_WARPS;
constexpr int kCrKS0 = (kCrIn + 15) / 16, kCrKS1 = kCH / 16, kCrNT = kCH / 8;      // 9 and 4 k16 steps, 8 n8 tiles
constexpr float kCrXScale = 256.f, kCrHScale = 16.f;
constexpr int kCrSmemBytes = (kCrKS0 + kCrKS1) * kCrNT * 32 * 16 + (3 * kCH + 4 + 2 * kCrWarps) * 4;

__device__ __forceinline__ void cr_split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}
__device__ __forceinline__ void cr_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// largest |w| of a weight matrix -> the power of two that puts it in [256, 512) (exact scaling; hi and lo stay normal fp16 numbers).
// (Staging the fp32 weights in shared memory first was measured: no faster at 65 536 envs -- the prologue is not bound by these L2 reads -- and
// 6 % slower at 1 Mi envs, where the 49 KB it takes from L1 cost observation-row hits.)
__device__ __forceinline__ float cr_block_scale(const float* __restrict__ w, int n, float* s_red) {
    float m = 0.f;
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, sft));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) m = fmaxf(m, s_red[i]);
    int ex;
    frexpf(fmaxf(m, 1e-30f), &ex);                 // m = f 2^ex, f in [0.5, 1)
    const int e = max(-14, min(24, 9 - ex));        // m 2^e in [256, 512)
    return exp2f((float)e);
}
// B fragments of W [64][K] (row = output neuron) for the k16 steps of one layer: per (step, n8 tile, lane)
// {hi b0, hi b1, lo b0, lo b1}
__device__ __forceinline__ void cr_build_frags(uint4* dst, const float* __restrict__ w, int K, int ksteps, float scale) {
    for (int i = threadIdx.x; i < ksteps * kCrNT * 32; i += blockDim.x) {
        const int lane = i & 31, nt = (i >> 5) % kCrNT, ks = (i >> 5) / kCrNT;
        const int n = 8 * nt + (lane >> 2), k0 = 16 * ks + 2 * (lane & 3);
        const float* r = w + n * K;
        const float w0 = k0 < K ? r[k0] * scale : 0.f, w1 = k0 + 1 < K ? r[k0 + 1] * scale : 0.f;
        const float w8 = k0 + 8 < K ? r[k0 + 8] * scale : 0.f, w9 = k0 + 9 < K ? r[k0 + 9] * scale : 0.f;
        uint4 f;
        cr_split2(w0, w1, f.x, f.z); cr_split2(w8, w9, f.y, f.w);
        dst[i] = f;
    }
}

__global__ void __launch_bounds__(kCrWarps * 32, 1) k_critic(const float* __restrict__ obs, const float* __restrict__ wts, float* __restrict__ value, int E) {
    extern __shared__ __align__(16) uint8_t cr_smem[];
    const PolicyOffsets o = policy_offsets();
    uint4* s_f0 = reinterpret_cast<uint4*>(cr_smem);                    // [9][8][32]
    uint4* s_f1 = s_f0 + kCrKS0 * kCrNT * 32;                           // [4][8][32]
    float* s_b0 = reinterpret_cast<float*>(s_f1 + kCrKS1 * kCrNT * 32);
    float* s_b1 = s_b0 + kCH;
    float* s_w2 = s_b1 + kCH;
    float* s_misc = s_w2 + kCH;                                          // b2, 1 / (x scale * w0 scale), 1 / (h scale * w1 scale)
    float* s_red = s_misc + 4;                                           // [2][warps]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    {
        const float sc0 = cr_block_scale(wts + o.c0_w, kCH * kCrIn, s_red);
        const float sc1 = cr_block_scale(wts + o.c1_w, kCH * kCH, s_red + kCrWarps);
        cr_build_frags(s_f0, wts + o.c0_w, kCrIn, kCrKS0, sc0);
        cr_build_frags(s_f1, wts + o.c1_w, kCH, kCrKS1, sc1);
        if (threadIdx.x < kCH) { s_b0[threadIdx.x] = wts[o.c0_b + threadIdx.x]; s_b1[threadIdx.x] = wts[o.c1_b + threadIdx.x]; s_w2[threadIdx.x] = wts[o.c2_w + threadIdx.x]; }
        if (threadIdx.x == 0) { s_misc[0] = wts[o.c2_b]; s_misc[1] = 1.f / (kCrXScale * sc0); s_misc[2] = 1.f / (kCrHScale * sc1); }
    }
    __syncthreads();
    const float inv0 = s_misc[1], inv1 = s_misc[2], b2 = s_misc[0];
    const int ntiles = (E + 31) >> 5;
#pragma unroll 1
    for (int t = blockIdx.x + w * gridDim.x; t < ntiles; t += kCrWarps * gridDim.x) {
        const int e0 = t << 5;
        float acc[2][kCrNT][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < kCrNT; nt++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc[mt][nt][i] = 0.f;
        // ---------------- layer 0: [32 x 130] x [130 x 64]; the next k16 step's observation columns are in flight while this step's HMMAs issue
        const float* xbase = obs + (size_t)(e0 + g) * kCrIn + 2 * t4;   // row e0 + 16 mt + 8 rr + g = a compile-time offset from this one
        float2 xn[2][2][2];   // [m tile][column half][row half]
        auto fetch = [&](int ks) {
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++)
#pragma unroll
                    for (int rr = 0; rr < 2; rr++) {
                        const int col = 16 * ks + 8 * hf + 2 * t4;
                        xn[mt][hf][rr] = (col < kCrIn && e0 + g + 16 * mt + 8 * rr < E)
                                             ? __ldg(reinterpret_cast<const float2*>(xbase + (16 * mt + 8 * rr) * kCrIn + 16 * ks + 8 * hf)) : make_float2(0.f, 0.f);
                    }
        };
        fetch(0);
#pragma unroll 1
        for (int ks = 0; ks < kCrKS0; ks++) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++)
#pragma unroll
                    for (int rr = 0; rr < 2; rr++)
                        cr_split2(xn[mt][hf][rr].x * kCrXScale, xn[mt][hf][rr].y * kCrXScale, ah[mt][2 * hf + rr], al[mt][2 * hf + rr]);
            if (ks + 1 < kCrKS0) fetch(ks + 1);
#pragma unroll
            for (int nt = 0; nt < kCrNT; nt++) {
                const uint4 b = s_f0[(ks * kCrNT + nt) * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    cr_mma(acc[mt][nt], al[mt], b.x, b.y);
                    cr_mma(acc[mt][nt], ah[mt], b.z, b.w);
                    cr_mma(acc[mt][nt], ah[mt], b.x, b.y);
                }
            }
        }
        // ---------------- layers 1 and 2, one m16 tile at a time: relu(acc + b0) becomes the A fragments in registers
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            asm volatile("" ::: "memory");   // the second tile re-reads the weight fragments instead of keeping all 128 words of them live across both
            uint32_t hh[kCrKS1][4], hl[kCrKS1][4];
#pragma unroll
            for (int nt = 0; nt < kCrNT; nt++) {
                const float2 bb = *reinterpret_cast<const float2*>(s_b0 + 8 * nt + 2 * t4);
                const float h0 = fmaxf(fmaf(acc[mt][nt][0], inv0, bb.x), 0.f) * kCrHScale, h1 = fmaxf(fmaf(acc[mt][nt][1], inv0, bb.y), 0.f) * kCrHScale;
                const float h2 = fmaxf(fmaf(acc[mt][nt][2], inv0, bb.x), 0.f) * kCrHScale, h3 = fmaxf(fmaf(acc[mt][nt][3], inv0, bb.y), 0.f) * kCrHScale;
                cr_split2(h0, h1, hh[nt >> 1][2 * (nt & 1)], hl[nt >> 1][2 * (nt & 1)]);            // row g
                cr_split2(h2, h3, hh[nt >> 1][2 * (nt & 1) + 1], hl[nt >> 1][2 * (nt & 1) + 1]);    // row g + 8
            }
            float a1[kCrNT][4];
#pragma unroll
            for (int nt = 0; nt < kCrNT; nt++)
#pragma unroll
                for (int i = 0; i < 4; i++) a1[nt][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < kCrKS1; ks++) {
                asm volatile("" ::: "memory");   // at most one k step's fragments in flight (the compiler otherwise hoists all 32 loads and spills the accumulators)
#pragma unroll
                for (int nt = 0; nt < kCrNT; nt++) {
                    const uint4 b = s_f1[(ks * kCrNT + nt) * 32 + lane];
                    cr_mma(a1[nt], hl[ks], b.x, b.y);
                    cr_mma(a1[nt], hh[ks], b.z, b.w);
                    cr_mma(a1[nt], hh[ks], b.x, b.y);
                }
            }
            float v0 = 0.f, v1 = 0.f;   // rows g and g + 8
#pragma unroll
            for (int nt = 0; nt < kCrNT; nt++) {
                const float2 bb = *reinterpret_cast<const float2*>(s_b1 + 8 * nt + 2 * t4);
                const float2 ww = *reinterpret_cast<const float2*>(s_w2 + 8 * nt + 2 * t4);
                v0 = fmaf(fmaxf(fmaf(a1[nt][0], inv1, bb.x), 0.f), ww.x, v0); v0 = fmaf(fmaxf(fmaf(a1[nt][1], inv1, bb.y), 0.f), ww.y, v0);
                v1 = fmaf(fmaxf(fmaf(a1[nt][2], inv1, bb.x), 0.f), ww.x, v1); v1 = fmaf(fmaxf(fmaf(a1[nt][3], inv1, bb.y), 0.f), ww.y, v1);
            }
            v0 += __shfl_xor_sync(kFull, v0, 1); v0 += __shfl_xor_sync(kFull, v0, 2);
            v1 += __shfl_xor_sync(kFull, v1, 1); v1 += __shfl_xor_sync(kFull, v1, 2);
            if (t4 == 0) {
                const int e = e0 + 16 * mt + g;
                if (e < E) value[e] = v0 + b2;
                if (e + 8 < E) value[e + 8] = v1 + b2;
            }
        }
    }
}
#endif

cudaError_t launch_critic(const float* wts, const float* obs, int E, float* value, cudaStream_t stream) {
    static PerDeviceFlag configured;
    static int n_sm[kMaxDevices] = {};
    const int dslot = current_device_slot();
    if (configured.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(k_critic, cudaFuncAttributeMaxDynamicSharedMemorySize, kCrSmemBytes);
        int dev = 0;
        if (e == cudaSuccess) e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm[dslot], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) { configured.retract(); return e; }
    }
    const int ntiles = (E + 31) / 32;
    const int blocks = (ntiles + kCrWarps - 1) / kCrWarps;
    k_critic<<<blocks < n_sm[dslot] ? blocks : n_sm[dslot], kCrWarps * 32, kCrSmemBytes, stream>>>(obs, wts, value, E);
    return cudaGetLastError();
}

int policy_offsets_host(int32_t* out) {
    const PolicyOffsets o = policy_offsets();
    const int v[MM_POLICY_N_OFFSETS] = {o.proj_w, o.proj_b, o.proj_col, o.proj_dim, o.att_k, o.att_q, o.att_v, o.l0_w, o.l0_b, o.l1_w, o.l1_b, o.l2_w, o.l2_b,
                       o.head_w, o.head_b, o.c0_w, o.c0_b, o.c1_w, o.c1_b, o.c2_w, o.c2_b, o.total, o.l0_whi, o.l0_wlo, o.l1_whi, o.l1_wlo, o.l2_whi, o.l2_wlo,
                       o.c0_wt, o.c1_wt, o.tokm, o.tokb, o.l0_h16, o.l0_l16, o.l1_h16, o.l1_l16, o.l2_h16, o.l2_l16, o.l0_asc, o.l1_asc, o.l2_asc, o.lh_h16, o.lh_l16, o.lh_asc};
    for (int i = 0; i < MM_POLICY_N_OFFSETS; i++) out[i] = v[i];
    return 0;
}

cudaError_t launch_linear_tc(const float* x, const float* w_hi, const float* w_lo, const float* bias, float* y, int M, int K, const float* head_w,
                             const float* head_b, const HeadArgs* heads, cudaStream_t stream);
cudaError_t launch_linear_f16x3(const float* x, const void* w_hi, const void* w_lo, int n_rows_w, int kpad, const float* acc_scale, const float* bias, float* y, int ldy,
                                int M, int K, uint32_t* gate_out, const float* head_w, const float* head_b, const HeadArgs* heads, float* heads_part,
                                cudaStream_t stream);

cudaError_t launch_trunk_fused(const float* x0, const void* const w16[4][2], const float* const asc[4], const float* const bias[4], const HeadArgs& heads, int M,
                               cudaStream_t stream);

// flags bit 0: trunk GEMMs on tcgen05 instead of the fp32 SIMT tiles; bit 1: critic on a forked side stream; bit 2 (with bit 0): the
// 3xFP16 N-split kernel (mm_linear16.cu, two CTAs per SM) instead of the 3xTF32 one (mm_policy_tc.cu); bit 3 (with bits 0 and 2): the three
// trunk layers + heads as ONE persistent kernel with the activations resident in shared memory (mm_trunk_fused.cu)
cudaError_t launch_policy(const float* wts, const float* obs, const uint8_t* masks, int E, float* scratch, const uint8_t* actions_in,
                          uint8_t* actions_out, float* logp, float* value, float* logits_out, int env_offset, uint64_t seed, uint64_t counter,
                          int flags, const uint64_t* counter_dev, cudaStream_t stream) {
    const PolicyOffsets o = policy_offsets();
    const int R = 2 * E;
    HeadArgs ha;
    ha.counter_dev = counter_dev;
    ha.masks = masks; ha.actions_in = actions_in; ha.actions_out = actions_out; ha.logp = logp; ha.logits_out = logits_out;
    ha.E = E; ha.env_offset = env_offset; ha.seed = seed; ha.counter = counter;
    float* x0 = scratch;                      // [R,460]
    float* h1 = scratch + (size_t)R * kX0;    // [R,264]
    float* h2 = h1 + (size_t)R * kHid;        // [R,264]
    float* hpart = h2 + (size_t)R * kHid;     // [R,2,8] partial head sums of the two column halves (3xFP16 path)
    // flags bit 1: the critic only reads obs, so it runs on a side stream forked from / joined to `stream` by events (legal inside a
    // stream capture, where it becomes a parallel branch of the graph) and overlaps the actor's token + trunk kernels.  The stream and
    // the two events are created once per host thread, on the first (eager) call.
    static thread_local cudaStream_t side_d[kMaxDevices] = {};
    static thread_local cudaEvent_t ev_fork_d[kMaxDevices] = {}, ev_join_d[kMaxDevices] = {};
    const int dslot = current_device_slot();   // streams and events belong to the device they were created on
    cudaStream_t& side = side_d[dslot];
    cudaEvent_t &ev_fork = ev_fork_d[dslot], &ev_join = ev_join_d[dslot];
    const bool overlap = value && (flags & 2);
    if (overlap) {
        if (!side) {
            cudaError_t e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaEventRecord(ev_fork, stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side, ev_fork, 0);
        if (e != cudaSuccess) return e;
        if ((e = launch_critic(wts, obs, E, value, side)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev_join, side)) != cudaSuccess) return e;
    }
    { cudaError_t e = launch_tokens_any(wts, obs, R, x0, stream); if (e != cudaSuccess) return e; }
    if ((flags & 13) == 13) {
        const void* const w16[4][2] = {{wts + o.l0_h16, wts + o.l0_l16}, {wts + o.l1_h16, wts + o.l1_l16}, {wts + o.l2_h16, wts + o.l2_l16}, {wts + o.lh_h16, wts + o.lh_l16}};
        const float* const asc[4] = {wts + o.l0_asc, wts + o.l1_asc, wts + o.l2_asc, wts + o.lh_asc};
        const float* const bias[4] = {wts + o.l0_b, wts + o.l1_b, wts + o.l2_b, wts + o.head_b};
        cudaError_t e = launch_trunk_fused(x0, w16, asc, bias, ha, R, stream);
        if (e != cudaSuccess) return e;
    } else if ((flags & 5) == 5) {
        cudaError_t e;
        if ((e = launch_linear_f16x3(x0, wts + o.l0_h16, wts + o.l0_l16, kHid, kPad0, wts + o.l0_asc, wts + o.l0_b, h1, kHid, R, kX0, nullptr, nullptr, nullptr, nullptr, nullptr, stream)) != cudaSuccess) return e;
        if ((e = launch_linear_f16x3(h1, wts + o.l1_h16, wts + o.l1_l16, kHid, kPad1, wts + o.l1_asc, wts + o.l1_b, h2, kHid, R, kHid, nullptr, nullptr, nullptr, nullptr, nullptr, stream)) != cudaSuccess) return e;
        // last layer: each column half contracts its part of the row with the six head rows; k_heads_finish adds the halves, masks, samples
        if ((e = launch_linear_f16x3(h2, wts + o.l2_h16, wts + o.l2_l16, kHid, kPad1, wts + o.l2_asc, wts + o.l2_b, nullptr, kHid, R, kHid, nullptr, wts + o.head_w, wts + o.head_b, &ha, hpart, stream)) != cudaSuccess) return e;
    } else if (flags & 1) {
        cudaError_t e;
        if ((e = launch_linear_tc(x0, wts + o.l0_whi, wts + o.l0_wlo, wts + o.l0_b, h1, R, kX0, nullptr, nullptr, nullptr, stream)) != cudaSuccess) return e;
        if ((e = launch_linear_tc(h1, wts + o.l1_whi, wts + o.l1_wlo, wts + o.l1_b, h2, R, kHid, nullptr, nullptr, nullptr, stream)) != cudaSuccess) return e;
        // last layer: heads, masking, sampling and the joint log-prob run in the epilogue; the 264-wide activation never leaves the SM
        if ((e = launch_linear_tc(h2, wts + o.l2_whi, wts + o.l2_wlo, wts + o.l2_b, nullptr, R, kHid, wts + o.head_w, wts + o.head_b, &ha, stream)) != cudaSuccess) return e;
    } else {
        dim3 grid((R + BM - 1) / BM, (kHid + BN - 1) / BN);
        k_linear_relu<<<grid, 256, 0, stream>>>(x0, wts + o.l0_w, wts + o.l0_b, h1, R, kHid, kX0);
        k_linear_relu<<<grid, 256, 0, stream>>>(h1, wts + o.l1_w, wts + o.l1_b, h2, R, kHid, kHid);
        k_linear_relu<<<grid, 256, 0, stream>>>(h2, wts + o.l2_w, wts + o.l2_b, h1, R, kHid, kHid);
        k_heads<<<(E * 32 + 127) / 128, 128, 0, stream>>>(h1, wts, ha);
    }
    if (overlap) {
        cudaError_t e = cudaStreamWaitEvent(stream, ev_join, 0);
        if (e != cudaSuccess) return e;
    } else if (value) { cudaError_t e = launch_critic(wts, obs, E, value, stream); if (e != cudaSuccess) return e; }
    return cudaGetLastError();
}

}  // namespace mm

namespace mm {
__global__ void k_add_u64(unsigned long long* p, unsigned long long v) { *p += v; }
cudaError_t launch_add_u64(unsigned long long* p, unsigned long long v, cudaStream_t stream) {
    k_add_u64<<<1, 1, 0, stream>>>(p, v);
    return cudaGetLastError();
}
}  // namespace mm
