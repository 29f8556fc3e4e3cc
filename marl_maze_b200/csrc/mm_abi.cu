// mm_abi.cu -- the C ABI declared in include/marl_maze_b200.h.  Thin argument checking + kernel launches; no torch types.
#include <stdio.h>
#include <string.h>
#include "mm_env.cuh"
#include "mm_update.cuh"

namespace mm {
cudaError_t launch_step_obs(const StepParams& p, bool reset_only, cudaStream_t stream);
cudaError_t launch_gae(const float*, const float*, const uint8_t*, const float*, float*, float*, int, int, double, double, cudaStream_t);
cudaError_t launch_generate(const mm_state* st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty,
                            uint64_t seed, uint32_t id_base, int id_mod, int id_mul, void* scratch, int max_blocks, const uint8_t* only, int height_cells, cudaStream_t stream);
cudaError_t launch_policy(const float*, const float*, const uint8_t*, int, float*, const uint8_t*, uint8_t*, float*, float*, float*, int, uint64_t, uint64_t,
                          int, const uint64_t*, cudaStream_t);
cudaError_t launch_tokens_fwd(const float* wts, const float* obs, int R, float* x0, cudaStream_t stream);
cudaError_t launch_tokens_fwd_full(const float* wts, const float* obs, int R, float* x0, cudaStream_t stream);
cudaError_t launch_tokens_bwd(const float* wts, const float* obs, const float* dout, int R, float* dy40, float* part, cudaStream_t stream);
size_t tokens_bwd_scratch_floats(int R);
int tokens_bwd_blocks();
cudaError_t launch_add_u64(unsigned long long* p, unsigned long long v, cudaStream_t stream);
int policy_offsets_host(int32_t* out);
cudaError_t launch_linear_f16x3(const float* x, const void* w_hi, const void* w_lo, int n_rows_w, int kpad, const float* acc_scale, const float* bias, float* y, int ldy,
                                int M, int K, uint32_t* gate_out, const float* head_w, const float* head_b, const HeadArgs* heads, float* heads_part,
                                cudaStream_t stream);
void trunk_fused_set_profile_buffer(unsigned long long* p);
cudaError_t launch_selftest_div(int amax, int bmax, unsigned long long* mismatches, cudaStream_t stream);
cudaError_t launch_critic(const float* wts, const float* obs, int E, float* value, cudaStream_t stream);
__global__ void k_load_layouts(ulonglong2*, ulonglong2*, uint4*, int, int, int, int, const uint8_t*, const int32_t*, uint16_t*);
__global__ void k_init_state(uint4*, uint32_t*, uint4*, uint32_t*, int);
__global__ void k_unpack_agents(const uint4*, const uint32_t*, const uint4*, const ulonglong2*, int, int, int32_t*);
__global__ void k_agent_place(uint4*, uint32_t*, int, int, int, int, int);
__global__ void k_unpack_envs(const uint4*, const uint32_t*, int, int32_t*);
__global__ void k_unpack_grid(const ulonglong2*, const ulonglong2*, int, uint8_t*, uint8_t*);
__global__ void k_unpack_pool_hdr(const uint4*, int, int32_t*);
}  // namespace mm

using namespace mm;

static thread_local char g_cuda_err[256] = "";

static int cuda_status(cudaError_t e) {
    if (e == cudaSuccess) return MM_OK;
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return MM_ERR_CUDA;
}
static int vision_of(const mm_state* st, int a) { const int v = (st->vision >> (8 * a)) & 0xff; return v ? v : 4; }   // 0 = the reference default
static bool state_ok(const mm_state* st) {
    return st && vision_of(st, 0) <= 4 && vision_of(st, 1) <= 4 && (st->vision >> 16) == 0 && st->n_envs > 0 && st->n_pool > 0 && st->smax >= 3 && st->smax <= MM_MAX_SIDE && st->max_timestep > 0 && st->pool_grid && st->pool_d2e &&
           st->pool_hdr && st->env_grid && st->env_hdr && st->env_episode && st->agent_a && st->agent_b;
}
static StepParams make_params(const mm_state* st) {
    StepParams p{};
    p.pool_grid = (const ulonglong2*)st->pool_grid; p.pool_d2e = (const ulonglong2*)st->pool_d2e; p.pool_hdr = (const uint4*)st->pool_hdr;
    p.env_grid = (ulonglong2*)st->env_grid; p.env_hdr = (uint4*)st->env_hdr; p.env_episode = (uint32_t*)st->env_episode;
    p.agent_a = (uint4*)st->agent_a; p.agent_b = (uint32_t*)st->agent_b;
    p.E = st->n_envs; p.P = st->n_pool; p.rows = st->smax + 2 * MM_PAD; p.smax = st->smax; p.max_t = st->max_timestep; p.env_offset = st->env_offset;
    p.inv_max_t = 1.0f / (float)st->max_timestep;
    for (int a = 0; a < 2; a++) {   // the ray features' values, in the reference's own float64 arithmetic (see StepParams)
        const int R = vision_of(st, a);
        p.vr[a] = R;
        const double distance = 1.0 / R;                                   // maze_agent.py:148
        double acc = 0.0;
        for (int c = 0; c < 5; c++) { p.de_tab[a][c] = 0.f; p.mk_tab[a][c] = 0.f; }
        for (int c = 1; c <= R; c++) {
            acc += 1.0 / R;                                                // maze_agent.py:264,267
            p.mk_tab[a][c] = (float)acc;
            p.de_tab[a][c] = c == R ? 1.0f : (float)(1.0 - (R - c) * distance);   // maze_agent.py:151,180
        }
    }
    return p;
}

extern "C" {

int mm_abi_version(void) { return 1; }
/* developer hook (not in the public header): cycle counters of the fused trunk kernel in -DMM_TF_PROFILE builds, [148][16] u64 */
int mm_debug_trunk_profile_buffer(unsigned long long* p) { trunk_fused_set_profile_buffer(p); return MM_OK; }
#ifndef MM_SRC_HASH
#define MM_SRC_HASH "unknown"
#endif
const char* mm_source_hash(void) { return MM_SRC_HASH; }
const char* mm_error_string(int code) {
    switch (code) {
        case MM_OK: return "ok";
        case MM_ERR_BAD_ARG: return "bad argument";
        case MM_ERR_CUDA: return "CUDA error (see mm_last_cuda_error)";
        case MM_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown";
    }
}
const char* mm_last_cuda_error(void) { return g_cuda_err; }

size_t mm_sizeof_pool_grid(int n_pool, int smax) { return (size_t)n_pool * (smax + 2 * MM_PAD) * 16; }
size_t mm_sizeof_pool_d2e(int n_pool, int smax) { return (size_t)n_pool * smax * 16; }
size_t mm_sizeof_pool_hdr(int n_pool) { return (size_t)n_pool * 16; }
size_t mm_sizeof_env_grid(int n_envs, int smax) { return (size_t)n_envs * (smax + 2 * MM_PAD) * 16; }
size_t mm_sizeof_env_hdr(int n_envs) { return (size_t)n_envs * 16; }
size_t mm_sizeof_env_episode(int n_envs) { return (size_t)n_envs * 4; }
size_t mm_sizeof_agent_a(int n_envs) { return (size_t)n_envs * 2 * 16; }
size_t mm_sizeof_agent_b(int n_envs) { return (size_t)n_envs * 2 * 4; }
size_t mm_sizeof_finalize_scratch(int n, int smax) { return (size_t)n * smax * smax * 2; }
#ifdef MM_K1_V1
size_t mm_sizeof_generate_scratch(int n, int smax) { return (size_t)n * smax * smax * 2; }   // DFS stack / BFS queue of the first-generation kernel
#else
size_t mm_sizeof_generate_scratch(int, int) { return 16; }   // the generator keeps all working state in shared memory (mm_generate.cu); the argument stays in the ABI
#endif

int mm_init_state(const mm_state* st, void* stream) {
    if (!state_ok(st)) return MM_ERR_BAD_ARG;
    const int n = 2 * st->n_envs;
    k_init_state<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((uint4*)st->env_hdr, (uint32_t*)st->env_episode, (uint4*)st->agent_a, (uint32_t*)st->agent_b, st->n_envs);
    return cuda_status(cudaGetLastError());
}

int mm_load_layouts(const mm_state* st, int first, int n, const uint8_t* layouts, const int32_t* hdr, void* scratch, void* stream) {
    if (!state_ok(st) || first < 0 || n < 0 || first + n > st->n_pool || (n && (!layouts || !hdr || !scratch))) return MM_ERR_BAD_ARG;
    if (n == 0) return MM_OK;
    k_load_layouts<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>((ulonglong2*)st->pool_grid, (ulonglong2*)st->pool_d2e, (uint4*)st->pool_hdr, first, n,
                                                                  st->smax + 2 * MM_PAD, st->smax, layouts, hdr, (uint16_t*)scratch);
    return cuda_status(cudaGetLastError());
}

int mm_generate_masked(const mm_state* st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty, uint64_t seed, uint32_t id_base,
                       int id_mod, int id_mul, void* scratch, int max_blocks, const uint8_t* only, int height_cells, void* stream) {
    if (!state_ok(st) || first < 0 || n < 0 || first + n > st->n_pool || side_lo < 4 || side_hi < side_lo || side_hi * 2 - 1 > st->smax || difficulty < 1 ||
        id_mod < 0 || (n && !scratch) || max_blocks < 0 || height_cells < 0 || (height_cells > 0 && (height_cells < 4 || height_cells * 2 - 1 > st->smax)))
        return MM_ERR_BAD_ARG;
    if (n == 0) return MM_OK;
    return cuda_status(launch_generate(st, first, n, side_lo, side_hi, rand_start, difficulty, seed, id_base, id_mod, id_mul, scratch, max_blocks, only,
                                       height_cells, (cudaStream_t)stream));
}
int mm_generate_ex(const mm_state* st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty, uint64_t seed, uint32_t id_base,
                   int id_mod, int id_mul, void* scratch, int max_blocks, void* stream) {
    return mm_generate_masked(st, first, n, side_lo, side_hi, rand_start, difficulty, seed, id_base, id_mod, id_mul, scratch, max_blocks, nullptr, 0, stream);
}
int mm_generate(const mm_state* st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty, uint64_t seed, uint32_t id_base,
                int id_mod, int id_mul, void* scratch, void* stream) {
    return mm_generate_ex(st, first, n, side_lo, side_hi, rand_start, difficulty, seed, id_base, id_mod, id_mul, scratch, 0, stream);
}

int mm_agent_place(const mm_state* st, int env, int agent, int x, int y, int direction, int reset, void* stream) {
    if (!state_ok(st) || env < 0 || env >= st->n_envs || agent < 0 || agent > 1 || x < 0 || y < 0 || x >= st->smax || y >= st->smax || direction < 0 || direction > 3)
        return MM_ERR_BAD_ARG;
    k_agent_place<<<1, 1, 0, (cudaStream_t)stream>>>((uint4*)st->agent_a, (uint32_t*)st->agent_b, 2 * env + agent, x, y, direction, reset);
    return cuda_status(cudaGetLastError());
}

int mm_reset(const mm_state* st, const uint8_t* reset_mask, float* obs, uint8_t* masks, void* stream) {
    if (!state_ok(st) || !obs || !masks) return MM_ERR_BAD_ARG;
    StepParams p = make_params(st);
    if (((uintptr_t)obs & 3) || ((uintptr_t)masks & 1)) return MM_ERR_BAD_ARG;
    p.reset_mask = reset_mask; p.obs = obs; p.masks = masks; p.obs_vec4 = ((uintptr_t)obs & 15) == 0;
    return cuda_status(launch_step_obs(p, true, (cudaStream_t)stream));
}

int mm_step_obs(const mm_state* st, const uint8_t* actions, float* obs, uint8_t* masks, float* reward, uint8_t* done, int auto_reset,
                uint64_t action_seed, uint8_t* actions_out, void* stream) {
    if (!state_ok(st) || !obs || !masks || !reward || !done) return MM_ERR_BAD_ARG;
    if (((uintptr_t)obs & 3) || ((uintptr_t)masks & 1) || ((uintptr_t)actions & 1) || ((uintptr_t)actions_out & 1)) return MM_ERR_BAD_ARG;
    StepParams p = make_params(st);
    p.actions = actions; p.actions_out = actions_out; p.obs = obs; p.masks = masks; p.reward = reward; p.done = done;
    p.auto_reset = auto_reset; p.action_seed = action_seed; p.obs_vec4 = ((uintptr_t)obs & 15) == 0;
    return cuda_status(launch_step_obs(p, false, (cudaStream_t)stream));
}

int mm_unpack_agents(const mm_state* st, int32_t* out, void* stream) {
    if (!state_ok(st) || !out) return MM_ERR_BAD_ARG;
    const int n = 2 * st->n_envs;
    k_unpack_agents<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const uint4*)st->agent_a, (const uint32_t*)st->agent_b, (const uint4*)st->env_hdr,
                                                                      (const ulonglong2*)st->pool_d2e, st->smax, st->n_envs, out);
    return cuda_status(cudaGetLastError());
}
int mm_unpack_envs(const mm_state* st, int32_t* out, void* stream) {
    if (!state_ok(st) || !out) return MM_ERR_BAD_ARG;
    k_unpack_envs<<<(st->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const uint4*)st->env_hdr, (const uint32_t*)st->env_episode, st->n_envs, out);
    return cuda_status(cudaGetLastError());
}
int mm_unpack_layout(const mm_state* st, int env, uint8_t* out, void* stream) {
    if (!state_ok(st) || !out || env < 0 || env >= st->n_envs) return MM_ERR_BAD_ARG;
    const int n = st->smax * st->smax, rows = st->smax + 2 * MM_PAD;
    k_unpack_grid<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const ulonglong2*)st->env_grid + (size_t)env * rows, nullptr, st->smax, out, nullptr);
    return cuda_status(cudaGetLastError());
}
int mm_unpack_pool(const mm_state* st, int p, uint8_t* out_layout, uint8_t* out_d2e, int32_t* out_hdr, void* stream) {
    if (!state_ok(st) || !out_layout || !out_d2e || !out_hdr || p < 0 || p >= st->n_pool) return MM_ERR_BAD_ARG;
    const int n = st->smax * st->smax, rows = st->smax + 2 * MM_PAD;
    k_unpack_grid<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const ulonglong2*)st->pool_grid + (size_t)p * rows,
                                                                    (const ulonglong2*)st->pool_d2e + (size_t)p * st->smax, st->smax, out_layout, out_d2e);
    k_unpack_pool_hdr<<<1, 1, 0, (cudaStream_t)stream>>>((const uint4*)st->pool_hdr, p, out_hdr);
    return cuda_status(cudaGetLastError());
}

int mm_gae(const float* reward, const float* value, const uint8_t* done, const float* v_boot, float* adv, float* rtg, int T, int E, double gamma,
           double lam, void* stream) {
    if (!reward || !value || !done || !adv || T < 0 || E < 0) return MM_ERR_BAD_ARG;
    return cuda_status(launch_gae(reward, value, done, v_boot, adv, rtg, T, E, gamma, lam, (cudaStream_t)stream));
}

int mm_selftest_div(int amax, int bmax, uint64_t* mismatches, void* stream) {
    if (amax < 0 || bmax < 1 || !mismatches) return MM_ERR_BAD_ARG;
    return cuda_status(launch_selftest_div(amax, bmax, (unsigned long long*)mismatches, (cudaStream_t)stream));
}
int mm_policy_offsets(int32_t* out) { return out ? policy_offsets_host(out) : MM_ERR_BAD_ARG; }
int mm_critic_forward(const float* weights, const float* obs, int n_envs, float* value, void* stream) {
    if (!weights || !obs || !value || n_envs <= 0) return MM_ERR_BAD_ARG;
    if (((uintptr_t)weights & 15) || ((uintptr_t)obs & 7)) return MM_ERR_BAD_ARG;   // 16-byte weight quads; the observation rows are read as float2
    return cuda_status(launch_critic(weights, obs, n_envs, value, (cudaStream_t)stream));
}
size_t mm_sizeof_policy_scratch(int n_envs) { return (size_t)n_envs * 2 * (460 + 2 * 264 + 16) * sizeof(float); }
int mm_policy_forward(const float* weights, const float* obs, const uint8_t* masks, int n_envs, void* scratch, const uint8_t* actions_in,
                      uint8_t* actions_out, float* logp, float* value, float* logits_out, int env_offset, uint64_t seed, uint64_t counter, int flags,
                      const uint64_t* counter_dev, void* stream) {
    if (!weights || !obs || !masks || n_envs <= 0 || !scratch || !logp || (!actions_in && !actions_out)) return MM_ERR_BAD_ARG;
    if (((uintptr_t)weights & 15) || ((uintptr_t)scratch & 15) || ((uintptr_t)obs & 7)) return MM_ERR_BAD_ARG;
    return cuda_status(launch_policy(weights, obs, masks, n_envs, (float*)scratch, actions_in, actions_out, logp, value, logits_out, env_offset, seed,
                                     counter, flags, counter_dev, (cudaStream_t)stream));
}
int mm_counter_add(uint64_t* counter_dev, uint64_t v, void* stream) {
    if (!counter_dev) return MM_ERR_BAD_ARG;
    return cuda_status(launch_add_u64((unsigned long long*)counter_dev, v, (cudaStream_t)stream));
}

int mm_wgrad_geometry(int rows, int n_out, int k_in, int32_t* slabs, int32_t* ld, int32_t* out_rows, int32_t* transposed) {
    if (rows <= 0 || n_out <= 0 || k_in <= 0 || !slabs || !ld || !out_rows || !transposed) return MM_ERR_BAD_ARG;
    int s, l, per, orows, tr;
    wgrad_geometry(rows, n_out, k_in, &s, &l, &per, &orows, &tr);
    *slabs = s; *ld = l; *out_rows = orows; *transposed = tr;
    return MM_OK;
}
int mm_wgrad_tf32x3(const float* dz, const float* h, int rows, int n_out, int k_in, float* part, void* stream) {
    if (!dz || !h || !part || rows <= 0 || n_out <= 0 || k_in <= 0 || (n_out & 3) || (k_in & 3)) return MM_ERR_BAD_ARG;
    if (((uintptr_t)dz & 15) || ((uintptr_t)h & 15) || ((uintptr_t)part & 15)) return MM_ERR_BAD_ARG;
    return cuda_status(launch_wgrad_tc(dz, h, rows, n_out, k_in, part, (cudaStream_t)stream));
}

int mm_linear_tf32x3(const float* x, int rows, int k, const float* w_hi, const float* w_lo, int n_rows_w, const float* bias, const uint32_t* gate_bits,
                     float* y, int ldy, int mode, uint32_t* gate_bits_out, void* stream) {
    if (!x || !w_hi || !w_lo || !y || rows <= 0 || k <= 0 || n_rows_w <= 0 || n_rows_w > 264 || ldy < n_rows_w) return MM_ERR_BAD_ARG;
    if (mode != MM_LINEAR_RELU && mode != MM_LINEAR_GATE && mode != MM_LINEAR_PLAIN) return MM_ERR_BAD_ARG;
    if ((mode == MM_LINEAR_RELU && !bias) || (mode == MM_LINEAR_GATE && !gate_bits) || (gate_bits_out && mode != MM_LINEAR_RELU))
        return MM_ERR_BAD_ARG;
    if (((uintptr_t)x & 15) || ((uintptr_t)w_hi & 15) || ((uintptr_t)w_lo & 15) || ((uintptr_t)y & 15) || ((uintptr_t)gate_bits & 3) || ((uintptr_t)gate_bits_out & 3) ||
        (k & 3) || (ldy & 3) || (n_rows_w & 3))
        return MM_ERR_BAD_ARG;
    return cuda_status(launch_linear_tc_ex(x, w_hi, w_lo, n_rows_w, bias, y, ldy, rows, k, mode, gate_bits, gate_bits_out, nullptr, nullptr, nullptr, (cudaStream_t)stream));
}
int mm_linear_f16x3(const float* x, int rows, int k, const void* w_hi16, const void* w_lo16, int n_rows_w, int kpad, const float* acc_scale, const float* bias,
                    float* y, int ldy, uint32_t* gate_bits_out, void* stream) {
    if (!x || !w_hi16 || !w_lo16 || !acc_scale || !bias || !y || rows <= 0 || k <= 0 || n_rows_w <= 0 || n_rows_w > 264 || ldy < n_rows_w || kpad < k || (kpad & 31))
        return MM_ERR_BAD_ARG;
    if (((uintptr_t)x & 15) || ((uintptr_t)w_hi16 & 15) || ((uintptr_t)w_lo16 & 15) || ((uintptr_t)y & 15) || ((uintptr_t)gate_bits_out & 3) || (k & 3) || (ldy & 3) ||
        (n_rows_w & 3))
        return MM_ERR_BAD_ARG;
    return cuda_status(launch_linear_f16x3(x, w_hi16, w_lo16, n_rows_w, kpad, acc_scale, bias, y, ldy, rows, k, gate_bits_out, nullptr, nullptr, nullptr, nullptr,
                                           (cudaStream_t)stream));
}
int mm_ppo_loss_geometry(int32_t* blocks, int32_t* ld) {
    if (!blocks || !ld) return MM_ERR_BAD_ARG;
    *blocks = ppo_loss_blocks(); *ld = ppo_loss_part_ld();
    return MM_OK;
}
int mm_ppo_heads_loss(const float* h2, const float* head_w, const float* head_b, const uint8_t* masks, const uint8_t* actions, const float* old_logp,
                      const float* adv, int n_envs, float clip, float scale, float* dz2, float* logp, float* part, void* stream) {
    if (!h2 || !head_w || !head_b || !masks || !actions || !old_logp || !adv || !dz2 || !part || n_envs <= 0) return MM_ERR_BAD_ARG;
    PpoLossArgs a{h2, head_w, head_b, masks, actions, old_logp, adv, clip, scale, dz2, logp, part, n_envs};
    return cuda_status(launch_ppo_heads_loss(a, (cudaStream_t)stream));
}

int mm_segment_sum_blocks(int rows) { return rows > 0 ? segment_sum_blocks(rows) : 0; }
int mm_segment_sum(const float* x, const int64_t* seg, int rows, int cols, int n_seg, float* part, void* stream) {
    if (!x || !seg || !part || rows <= 0 || cols <= 0 || (cols & 3) || n_seg < 1 || n_seg > 8 || ((uintptr_t)x & 15) || ((uintptr_t)part & 15)) return MM_ERR_BAD_ARG;
    return cuda_status(launch_segment_sum(x, (const long long*)seg, rows, cols, n_seg, part, (cudaStream_t)stream));
}

int mm_gather_rows(const float* src, const int64_t* seg, int rows, int cols, int n_src, float* out, void* stream) {
    if (!src || !seg || !out || rows <= 0 || cols <= 0 || (cols & 3) || n_src < 1 || n_src > 8 || ((uintptr_t)src & 15) || ((uintptr_t)out & 15)) return MM_ERR_BAD_ARG;
    return cuda_status(launch_gather_rows(src, (const long long*)seg, rows, cols, n_src, out, (cudaStream_t)stream));
}

int mm_tokens_forward(const float* weights, const float* obs, int rows, float* x0, void* stream) {
    if (!weights || !obs || !x0 || rows <= 0 || ((uintptr_t)x0 & 15)) return MM_ERR_BAD_ARG;
    return cuda_status(launch_tokens_fwd(weights, obs, rows, x0, (cudaStream_t)stream));
}
int mm_tokens_forward_full(const float* weights, const float* obs, int rows, float* x0, void* stream) {
    if (!weights || !obs || !x0 || rows <= 0 || ((uintptr_t)x0 & 15)) return MM_ERR_BAD_ARG;
    return cuda_status(launch_tokens_fwd_full(weights, obs, rows, x0, (cudaStream_t)stream));
}
int mm_tokens_backward_blocks(void) { return tokens_bwd_blocks(); }
size_t mm_sizeof_tokens_backward_scratch(int rows) { return rows > 0 ? tokens_bwd_scratch_floats(rows) * sizeof(float) : 0; }
int mm_tokens_backward(const float* weights, const float* obs, const float* d_x0, int rows, void* scratch, float* part, void* stream) {
    if (!weights || !obs || !d_x0 || !scratch || !part || rows <= 0 || ((uintptr_t)d_x0 & 15) || ((uintptr_t)scratch & 15)) return MM_ERR_BAD_ARG;
    return cuda_status(launch_tokens_bwd(weights, obs, d_x0, rows, (float*)scratch, part, (cudaStream_t)stream));
}

}  // extern "C"
