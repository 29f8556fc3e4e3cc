// mm_gae.cu -- K3: generalised advantage estimation as a reverse scan over [T][E] (one lane per env, coalesced rows).
//
// Replaces PPO.get_GAEs (PPO.py:193-203).  The reference runs per finished episode, in fp32 torch scalar ops in this
// exact association (python scalars meet fp32 tensors, so gamma and gamma*lam are rounded to fp32 first):
//     last step of an episode : delta = r - V_t
//     otherwise               : delta = (r + ((g*V_{t+1}) * (1 - done_{t+1}))) - V_t      <- note done of the NEXT step
//     A_t = delta + (gl * (1 - done_t)) * A_{t+1}
// Here episodes are delimited by done[t][e] inside a fixed horizon; an episode still open at t = T-1 is bootstrapped
// with v_boot[e] = V(s_T) (extension: the reference only ever sees complete episodes).  No FMA contraction anywhere,
// so advantages are bit-identical to the reference on complete episodes.
#include "mm_env.cuh"

namespace mm {

__global__ void __launch_bounds__(256) k_gae(const float* __restrict__ reward, const float* __restrict__ value, const uint8_t* __restrict__ done,
                                             const float* __restrict__ v_boot, float* __restrict__ adv, float* __restrict__ rtg,
                                             int T, int E, float g, float gl) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float a_next = 0.f;                                   // A_{t+1}
    float v_next = v_boot != nullptr ? v_boot[e] : 0.f;   // V_{t+1}
    float nd_next = 1.f;                                  // 1 - done_{t+1}; the bootstrap state is never terminal
    for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * E + e;
        const float r = reward[i], v = value[i];
        const bool d = done[i] != 0;
        float delta, a;
        if (d) {                       // t + 1 == len(ep_rew)
            delta = __fsub_rn(r, v);
            a = delta;                 // + gl * 0 * A
        } else {
            delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g, v_next), nd_next)), v);
            a = __fadd_rn(delta, __fmul_rn(gl, a_next));
        }
        adv[i] = a;
        if (rtg != nullptr) rtg[i] = __fadd_rn(a, v);
        a_next = a; v_next = v; nd_next = d ? 0.f : 1.f;
    }
}

cudaError_t launch_gae(const float* reward, const float* value, const uint8_t* done, const float* v_boot, float* adv, float* rtg,
                       int T, int E, double gamma, double lam, cudaStream_t stream) {
    if (T <= 0 || E <= 0) return cudaSuccess;
    k_gae<<<(E + 255) / 256, 256, 0, stream>>>(reward, value, done, v_boot, adv, rtg, T, E, (float)gamma, (float)(gamma * lam));
    return cudaGetLastError();
}

}  // namespace mm
