// mm_policy_heads.cuh -- masked Categorical / Bernoulli sampling and log-prob of ONE agent's action from its 6 logits.
// PPO.get_action (PPO.py:170-186) / get_log_probs (PPO.py:154-168).  Shared by the stand-alone k_heads kernel (fp32 SIMT path) and by
// the epilogue of the last tensor-core trunk layer, where the sampling is fused.
#pragma once
#include "mm_env.cuh"

namespace mm {

struct HeadArgs {
    const uint8_t* masks;        // [E][2][6]
    const uint8_t* actions_in;   // [E][2][2] or nullptr (sample)
    uint8_t* actions_out;        // [E][2][2] (when sampling)
    float* logp;                 // [E] joint log-prob of both agents (PPO.py:118,121)
    float* logits_out;           // [E][2][6] or nullptr
    int E, env_offset;
    uint64_t seed, counter;
    const uint64_t* counter_dev;  // optional: the effective counter is counter + *counter_dev (lets a captured CUDA graph advance its stream)
};

// row = global agent row (2*env + agent).  Returns log pi(action | obs) for that agent; writes the sampled action / the logits.
__device__ __forceinline__ float head_sample_or_eval(const float l[6], const long long row, const HeadArgs& h) {
    const uint8_t* mk = h.masks + row * 6;
    if (h.logits_out) for (int j = 0; j < 6; j++) h.logits_out[row * 6 + j] = l[j];
    float m = -INFINITY;  // masked Categorical over the 5 moves (PPO.py:174-176)
    for (int j = 0; j < 5; j++) if (mk[j]) m = fmaxf(m, l[j]);
    float p[5], s = 0.f;
    for (int j = 0; j < 5; j++) { p[j] = mk[j] ? expf(l[j] - m) : 0.f; s += p[j]; }
    const float p_mark = mk[5] ? 1.f / (1.f + expf(-l[5])) : 0.f;  // PPO.py:179
    int move, mark;
    if (h.actions_in) { move = h.actions_in[row * 2]; mark = h.actions_in[row * 2 + 1]; }
    else {
        uint32_t r[4];
        const uint32_t env = (uint32_t)(row >> 1), a = (uint32_t)(row & 1);
        const uint64_t ctr = h.counter + (h.counter_dev ? *h.counter_dev : 0ull);
        philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)(h.env_offset + env) * 2u + a, 0x504f4c49u, (uint32_t)h.seed, (uint32_t)(h.seed >> 32), r);
        const float u = (float)(r[0] >> 8) * (1.0f / 16777216.0f) * s;  // inverse CDF over the unnormalised masses
        float c = 0.f; move = -1; int last = 4;
        for (int j = 0; j < 5; j++) if (mk[j]) { last = j; c += p[j]; if (move < 0 && u < c) move = j; }
        if (move < 0) move = last;
        mark = ((float)(r[1] >> 8) * (1.0f / 16777216.0f) < p_mark) ? 1 : 0;  // torch.bernoulli(p)
        h.actions_out[row * 2] = (uint8_t)move; h.actions_out[row * 2 + 1] = (uint8_t)mark;
    }
    const float lp_move = (move < 5 && mk[move]) ? (l[move] - m) - logf(s) : -INFINITY;  // Categorical.log_prob
    return lp_move + logf(mark ? p_mark : 1.f - p_mark);                                   // PPO.py:181-184
}

}  // namespace mm
