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
    int total;
};
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
// 130 -> 64 -> 64 -> 1.  One warp handles 4 envs at a time; lane j owns hidden neurons j and j+32, weights are read from the
// TRANSPOSED copies ([k][neuron]: a warp reads 2 x 128 contiguous bytes per k, L1-resident), inputs are broadcast from shared.
constexpr int kCrEnvs = 4;
__global__ void __launch_bounds__(128) k_critic(const float* __restrict__ obs, const float* __restrict__ wts, float* __restrict__ value, int E) {
    const PolicyOffsets o = policy_offsets();
    __shared__ float s_in[4][kCrEnvs][132], s_h[4][kCrEnvs][kCH];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e0 = (blockIdx.x * 4 + wl) * kCrEnvs;
    if (e0 >= E) return;
    const int ne = min(kCrEnvs, E - e0);
    for (int i = lane; i < kCrEnvs * 130; i += 32) { const int e = i / 130, k = i - e * 130; s_in[wl][e][k] = e < ne ? obs[(size_t)(e0 + e) * 130 + k] : 0.f; }
    __syncwarp();
    float a0[kCrEnvs], a1[kCrEnvs];
#pragma unroll
    for (int e = 0; e < kCrEnvs; e++) { a0[e] = wts[o.c0_b + lane]; a1[e] = wts[o.c0_b + lane + 32]; }
    const float* w0 = wts + o.c0_wt;
#pragma unroll 2
    for (int k = 0; k < 130; k++) {
        const float wa = w0[k * kCH + lane], wb = w0[k * kCH + lane + 32];
#pragma unroll
        for (int e = 0; e < kCrEnvs; e++) { const float x = s_in[wl][e][k]; a0[e] = fmaf(x, wa, a0[e]); a1[e] = fmaf(x, wb, a1[e]); }
    }
#pragma unroll
    for (int e = 0; e < kCrEnvs; e++) { s_h[wl][e][lane] = fmaxf(a0[e], 0.f); s_h[wl][e][lane + 32] = fmaxf(a1[e], 0.f); }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < kCrEnvs; e++) { a0[e] = wts[o.c1_b + lane]; a1[e] = wts[o.c1_b + lane + 32]; }
    const float* w1 = wts + o.c1_wt;
#pragma unroll 4
    for (int k = 0; k < kCH; k++) {
        const float wa = w1[k * kCH + lane], wb = w1[k * kCH + lane + 32];
#pragma unroll
        for (int e = 0; e < kCrEnvs; e++) { const float x = s_h[wl][e][k]; a0[e] = fmaf(x, wa, a0[e]); a1[e] = fmaf(x, wb, a1[e]); }
    }
    const float v0 = wts[o.c2_w + lane], v1 = wts[o.c2_w + lane + 32], vb = wts[o.c2_b];
#pragma unroll
    for (int e = 0; e < kCrEnvs; e++) {
        float part = fmaf(fmaxf(a0[e], 0.f), v0, fmaxf(a1[e], 0.f) * v1);
#pragma unroll
        for (int sft = 16; sft; sft >>= 1) part += __shfl_xor_sync(kFull, part, sft);
        if (lane == 0 && e < ne) value[e0 + e] = part + vb;
    }
}

cudaError_t launch_critic(const float* wts, const float* obs, int E, float* value, cudaStream_t stream) {
    k_critic<<<(E + 4 * kCrEnvs - 1) / (4 * kCrEnvs), 128, 0, stream>>>(obs, wts, value, E);
    return cudaGetLastError();
}

int policy_offsets_host(int32_t* out) {
    const PolicyOffsets o = policy_offsets();
    const int v[32] = {o.proj_w, o.proj_b, o.proj_col, o.proj_dim, o.att_k, o.att_q, o.att_v, o.l0_w, o.l0_b, o.l1_w, o.l1_b, o.l2_w, o.l2_b,
                       o.head_w, o.head_b, o.c0_w, o.c0_b, o.c1_w, o.c1_b, o.c2_w, o.c2_b, o.total, o.l0_whi, o.l0_wlo, o.l1_whi, o.l1_wlo, o.l2_whi, o.l2_wlo,
                       o.c0_wt, o.c1_wt, o.tokm, o.tokb};
    for (int i = 0; i < 32; i++) out[i] = v[i];
    return 0;
}

cudaError_t launch_linear_tc(const float* x, const float* w_hi, const float* w_lo, const float* bias, float* y, int M, int K, const float* head_w,
                             const float* head_b, const HeadArgs* heads, cudaStream_t stream);

// flags bit 0: trunk GEMMs on tcgen05 (3xTF32) instead of the fp32 SIMT tiles; bit 1: critic on a forked side stream
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
    // flags bit 1: the critic only reads obs, so it runs on a side stream forked from / joined to `stream` by events (legal inside a
    // stream capture, where it becomes a parallel branch of the graph) and overlaps the actor's token + trunk kernels.  The stream and
    // the two events are created once per host thread, on the first (eager) call.
    static thread_local cudaStream_t side = nullptr;
    static thread_local cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
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
        k_critic<<<(E + 4 * kCrEnvs - 1) / (4 * kCrEnvs), 128, 0, side>>>(obs, wts, value, E);
        if ((e = cudaEventRecord(ev_join, side)) != cudaSuccess) return e;
    }
    k_tokens<<<min((R + kTokWarps - 1) / kTokWarps, 148 * 8), kTokWarps * 32, 0, stream>>>(obs, wts, x0, R);
    if (flags & 1) {
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
    } else if (value) k_critic<<<(E + 4 * kCrEnvs - 1) / (4 * kCrEnvs), 128, 0, stream>>>(obs, wts, value, E);
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
