// mm_update.cu -- K5: heads + clipped-surrogate loss of the PPO actor update, forward AND backward in one pass (PPO.py:58-76:
// get_log_probs for both agents, ratio, min(ratio*A, clip(ratio)*A), loss.backward() down to the last trunk activation).
//
// Per env e (agent rows 2e, 2e+1 of the last trunk activation H2 [2E][264]):
//     logits = H2 Wh^T + bh            (5 move logits + 1 mark logit per agent)
//     joint  = sum_agents  log softmax_masked(move logits)[move] + log(mark ? p : 1-p),  p = sigmoid(mark logit) if mask[5] else 0
//     ratio  = exp(joint - old_logp[e]);  loss_e = -min(ratio*A, clamp(ratio, 1-c, 1+c)*A) * scale
// and, with G = dloss/djoint = -[ratio within the clip range or ratio*A < clamp(ratio)*A] * A * ratio * scale (torch's minimum /
// clamp sub-gradients), the gradient at the logits is G*(onehot - softmax) on legal moves and G*(mark - p) on an enabled mark logit.
// The kernel writes dZ2 = (dlogits Wh) * (H2 > 0) -- the gradient at the last trunk layer's pre-activation -- and per-block partial sums
// of dWh = dlogits^T H2, dbh = sum dlogits and the loss; the host adds the partials (deterministic, no atomics).
// One warp walks envs; a lane owns columns lane, lane+32, ... of the 264-wide rows (coalesced 128-byte segments).
#include "mm_env.cuh"
#include "mm_update.cuh"

namespace mm {

#ifndef MM_UL_PREFETCH
#define MM_UL_PREFETCH 1
#endif
#ifndef MM_UL_MINBLOCKS
#define MM_UL_MINBLOCKS 3
#endif
constexpr int UL_HID = 264, UL_CPL = 9 /* columns per lane, the 9th only for lanes 0-7 */, UL_WARPS = 4;
constexpr int UL_PART_LD = 6 * UL_HID + 8;  // dWh [6][264], dbh [6], loss, unused

__global__ void __launch_bounds__(UL_WARPS * 32, MM_UL_MINBLOCKS) k_ppo_heads_loss(const PpoLossArgs a) {
    __shared__ float s_w[6][UL_HID];
    __shared__ float s_acc[6 * UL_HID + 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 6 * UL_HID; i += blockDim.x) s_w[i / UL_HID][i % UL_HID] = a.head_w[i];
    for (int i = threadIdx.x; i < 6 * UL_HID + 8; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    float bh[6];
#pragma unroll
    for (int j = 0; j < 6; j++) bh[j] = __ldg(&a.head_b[j]);

    float accw[6][UL_CPL], accb[6], loss_acc = 0.f;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        accb[j] = 0.f;
#pragma unroll
        for (int i = 0; i < UL_CPL; i++) accw[j][i] = 0.f;
    }
    const bool tail = lane < UL_HID - 8 * 32;  // lanes 0-7 own a 9th column
    const int wglobal = blockIdx.x * UL_WARPS + warp, wstride = gridDim.x * UL_WARPS;
    // the two activation rows of an env are fetched one env ahead: the warp has nothing else in flight while it waits for them (12 warps per SM)
    float hn[2][UL_CPL];
    auto fetch = [&](int e) {
#pragma unroll
        for (int ag = 0; ag < 2; ag++) {
            const float* hr = a.h2 + (size_t)(2 * e + ag) * UL_HID;
#pragma unroll
            for (int i = 0; i < 8; i++) hn[ag][i] = e < a.E ? __ldg(hr + lane + 32 * i) : 0.f;
            hn[ag][8] = (tail && e < a.E) ? __ldg(hr + lane + 256) : 0.f;
        }
    };
    fetch(wglobal);
    for (int e = wglobal; e < a.E; e += wstride) {
        float hv[2][UL_CPL], l[2][6];
#if !MM_UL_PREFETCH
        fetch(e);
#endif
#pragma unroll
        for (int ag = 0; ag < 2; ag++)
#pragma unroll
            for (int i = 0; i < UL_CPL; i++) hv[ag][i] = hn[ag][i];
#if MM_UL_PREFETCH
        fetch(e + wstride);
#endif
#pragma unroll
        for (int ag = 0; ag < 2; ag++)
#pragma unroll
            for (int j = 0; j < 6; j++) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) s = fmaf(hv[ag][i], s_w[j][lane + 32 * i], s);
                if (tail) s = fmaf(hv[ag][8], s_w[j][lane + 256], s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                l[ag][j] = s + bh[j];
            }
        // ---- distribution math, identical on every lane (PPO.py:154-168)
        float dl[2][6], joint = 0.f, pm[2], inv_s[2], pk[2][5];
        int mv[2], mkb[2];
        bool legal[2][6];
#pragma unroll
        for (int ag = 0; ag < 2; ag++) {
            const uint8_t* mk = a.masks + (size_t)(2 * e + ag) * 6;
#pragma unroll
            for (int j = 0; j < 6; j++) legal[ag][j] = mk[j] != 0;
            mv[ag] = a.actions[(size_t)(2 * e + ag) * 2];
            mkb[ag] = a.actions[(size_t)(2 * e + ag) * 2 + 1] != 0;
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 5; j++) if (legal[ag][j]) m = fmaxf(m, l[ag][j]);
            float s = 0.f, lsel = -INFINITY;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                pk[ag][j] = legal[ag][j] ? expf(l[ag][j] - m) : 0.f;
                s += pk[ag][j];
                if (j == mv[ag] && legal[ag][j]) lsel = l[ag][j];
            }
            inv_s[ag] = 1.f / s;
            pm[ag] = legal[ag][5] ? 1.f / (1.f + expf(-l[ag][5])) : 0.f;
            joint += ((lsel - m) - logf(s)) + logf(mkb[ag] ? pm[ag] : 1.f - pm[ag]);
        }
        const float A = __ldg(&a.adv[e]);
        const float ratio = expf(joint - __ldg(&a.old_logp[e]));
        const float lo = 1.f - a.clip, hi = 1.f + a.clip;
        const float s1 = ratio * A, s2 = fminf(fmaxf(ratio, lo), hi) * A;
        const bool pass = (ratio >= lo && ratio <= hi) || s1 < s2;
        const float G = pass ? -A * ratio * a.scale : 0.f;
        loss_acc += -fminf(s1, s2) * a.scale;
        if (lane == 0 && a.logp) a.logp[e] = joint;
#pragma unroll
        for (int ag = 0; ag < 2; ag++) {
#pragma unroll
            for (int j = 0; j < 5; j++) dl[ag][j] = legal[ag][j] ? G * ((j == mv[ag] ? 1.f : 0.f) - pk[ag][j] * inv_s[ag]) : 0.f;
            dl[ag][5] = legal[ag][5] ? G * ((mkb[ag] ? 1.f : 0.f) - pm[ag]) : 0.f;
        }
        // ---- backward through the heads and the last ReLU; head weight / bias gradient partial sums
#pragma unroll
        for (int ag = 0; ag < 2; ag++) {
            float* zr = a.dz2 + (size_t)(2 * e + ag) * UL_HID;
#pragma unroll
            for (int i = 0; i < UL_CPL; i++) {
                if (i == 8 && !tail) break;
                const int c = lane + 32 * i;
                float g = 0.f;
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    g = fmaf(dl[ag][j], s_w[j][c], g);
                    accw[j][i] = fmaf(dl[ag][j], hv[ag][i], accw[j][i]);
                }
                zr[c] = hv[ag][i] > 0.f ? g : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 6; j++) accb[j] += dl[ag][j];
        }
    }
    // ---- block reduction in a fixed warp order (deterministic), then one partial row per block
    for (int w = 0; w < UL_WARPS; w++) {
        if (warp == w) {
#pragma unroll
            for (int j = 0; j < 6; j++) {
#pragma unroll
                for (int i = 0; i < 8; i++) s_acc[j * UL_HID + lane + 32 * i] += accw[j][i];
                if (tail) s_acc[j * UL_HID + lane + 256] += accw[j][8];
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 6; j++) s_acc[6 * UL_HID + j] += accb[j];
                s_acc[6 * UL_HID + 6] += loss_acc;
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < UL_PART_LD; i += blockDim.x) a.part[(size_t)blockIdx.x * UL_PART_LD + i] = s_acc[i];
}

int ppo_loss_blocks() { return MM_UL_MINBLOCKS * 148; }
int ppo_loss_part_ld() { return UL_PART_LD; }

cudaError_t launch_ppo_heads_loss(const PpoLossArgs& a, cudaStream_t stream) {
    k_ppo_heads_loss<<<ppo_loss_blocks(), UL_WARPS * 32, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace mm

// ------------------------------------------------------------------------------------------------ segment sum
// out[s][c] = sum over rows r with seg[r] == s of x[r][c], for a handful of segments (n_seg <= 8): the backward of gathering [rows][cols]
// activations from n_seg distinct source rows (Actor.embed evaluates projection + attention once per distinct observation prefix).
// A scatter-add would funnel every row into n_seg addresses; here each block walks a contiguous row range with one float4 column group
// per thread and keeps the n_seg running sums in registers; per-block partials are added by the host (deterministic).
namespace mm {

constexpr int SS_MAX_SEG = 8, SS_THREADS = 128;

__global__ void __launch_bounds__(SS_THREADS) k_segment_sum(const float* __restrict__ x, const long long* __restrict__ seg, int rows, int cols, int n_seg,
                                                            int rows_per_block, float* __restrict__ part) {
    const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c4 = threadIdx.x; c4 * 4 < cols; c4 += SS_THREADS) {
        float4 acc[SS_MAX_SEG];
#pragma unroll
        for (int s = 0; s < SS_MAX_SEG; s++) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = r0; r < r1; r++) {
            const int sg = (int)__ldg(&seg[r]);
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * cols) + c4);
#pragma unroll
            for (int s = 0; s < SS_MAX_SEG; s++)
                if (s == sg) { acc[s].x += v.x; acc[s].y += v.y; acc[s].z += v.z; acc[s].w += v.w; }
        }
#pragma unroll
        for (int s = 0; s < SS_MAX_SEG; s++)
            if (s < n_seg) *(reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * n_seg + s) * cols) + c4) = acc[s];
    }
}

// out[r][:] = src[seg[r]][:] for a handful (<= 8) of source rows: the forward of the gather whose adjoint is k_segment_sum (the embedding of each agent row
// from the few distinct embeddings, update._ActorTrunkLoss).  The source rows sit in shared memory; a warp writes one 16-byte-per-lane stripe of a row per
// instruction with streaming stores -- the kernel is nothing but its rows x cols x 4 bytes of HBM writes (torch's index_select ran at a third of that rate).
constexpr int GR_THREADS = 256;
__global__ void __launch_bounds__(GR_THREADS) k_gather_rows(const float* __restrict__ src, const long long* __restrict__ seg, int rows, int cols, int n_src,
                                                            float* __restrict__ out) {
    extern __shared__ __align__(16) float gr_src[];   // [n_src][cols]
    for (int i = threadIdx.x; i < n_src * cols / 4; i += GR_THREADS) reinterpret_cast<float4*>(gr_src)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    __syncthreads();
    const int c4n = cols / 4, w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = GR_THREADS / 32;
    for (long long r = (long long)blockIdx.x * nw + w; r < rows; r += (long long)gridDim.x * nw) {
        int sg = (int)__ldg(&seg[r]);
        sg = sg < 0 ? 0 : sg >= n_src ? n_src - 1 : sg;   // never read outside the staged rows
        const float4* s4 = reinterpret_cast<const float4*>(gr_src + (size_t)sg * cols);
        float4* o4 = reinterpret_cast<float4*>(out + (size_t)r * cols);
        for (int c = lane; c < c4n; c += 32) __stcs(&o4[c], s4[c]);
    }
}

cudaError_t launch_gather_rows(const float* src, const long long* seg, int rows, int cols, int n_src, float* out, cudaStream_t stream) {
    if (n_src < 1 || n_src > SS_MAX_SEG || (cols & 3) || rows <= 0 || cols <= 0 || (size_t)n_src * cols * 4 > 48 * 1024) return cudaErrorInvalidValue;
    const int want = 148 * 8, nw = GR_THREADS / 32;
    const long long need = ((long long)rows + nw - 1) / nw;
    k_gather_rows<<<(int)(need < want ? need : want), GR_THREADS, (size_t)n_src * cols * 4, stream>>>(src, seg, rows, cols, n_src, out);
    return cudaGetLastError();
}

int segment_sum_blocks(int rows) {
    const int want = 148 * 8;
    return rows < want ? (rows > 0 ? rows : 1) : want;
}

cudaError_t launch_segment_sum(const float* x, const long long* seg, int rows, int cols, int n_seg, float* part, cudaStream_t stream) {
    if (n_seg < 1 || n_seg > SS_MAX_SEG || (cols & 3) || rows <= 0) return cudaErrorInvalidValue;
    const int blocks = segment_sum_blocks(rows);
    const int per = (rows + blocks - 1) / blocks;
    k_segment_sum<<<blocks, SS_THREADS, 0, stream>>>(x, seg, rows, cols, n_seg, per, part);
    return cudaGetLastError();
}

}  // namespace mm
