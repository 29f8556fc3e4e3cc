/*
 * marl_maze_b200.h -- C ABI of libmarl_maze_b200.so (hand-written sm_100a kernels for the MARL-Maze hot path).
 *
 * The reference (rhuangr/MARL-Maze) is pure Python and has no FFI of its own; its hot path is reached through
 * plain method calls.  This header is the boundary a maintainer would bind underneath those methods
 * (ctypes stubs are shown in INTEGRATION.md).  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it, none allocates memory;
 *   - return value: 0 = MM_OK, otherwise an MM_ERR_* code (mm_error_string() names it); nothing throws;
 *   - the caller (PyTorch on the host side) owns every buffer; sizes are given by the mm_sizeof_* helpers.
 *
 * HBM layout
 *   maze pool (read-only while stepping), P mazes of side <= smax, rows = smax + 10:
 *     pool_grid [P][rows][2] u64   bit planes (lo, hi) of the 2-bit cell value, bit (x+5) of row (y+5);
 *                                  0 path, 1 wall, 2/3 marks; everything outside W x H is wall
 *     pool_d2e  [P][smax][2] u64   bit planes of the direction-to-exit field (0 N, 1 E, 2 S, 3 W), bit (x+5) of row y
 *     pool_hdr  [P] 4 x u32        W | H<<8 | p0x<<16 | p0y<<24,  p1x | p1y<<8 | ex<<16 | ey<<24,  kx | ky<<8 | spl<<16,  id
 *   environments (E of them, two agents each):
 *     env_grid  [E][rows][2] u64   working copy of the pool grid plus the agents' marks
 *     env_hdr   [E] 4 x u32        t,  ex | ey<<8 | kx<<16 | ky<<24,  key_present | err<<1 | W<<8 | H<<16,  pool index
 *     env_episode [E] u32          episodes started so far; the next episode of env e uses pool maze (e + episode*E) mod P
 *     agent_a   [2E] 4 x u32, agent_b [2E] u32   packed per-agent state (mm_env.cuh)
 */
#ifndef MARL_MAZE_B200_H
#define MARL_MAZE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_OK 0
#define MM_ERR_BAD_ARG 1
#define MM_ERR_CUDA 2
#define MM_ERR_UNSUPPORTED 3

#define MM_OBS_DIM 65   /* networks.py:11 OBS_SPACE */
#define MM_MASK_DIM 6   /* maze_agent.py:131-140 */
#define MM_PAD 5        /* wall border kept around every stored grid */
#define MM_MAX_SIDE 54  /* side + 2*MM_PAD <= 64 bit columns */
#define MM_AGENT_FIELDS 18

typedef struct mm_state {
    /* pool */
    const void *pool_grid, *pool_d2e, *pool_hdr;
    /* environments */
    void *env_grid, *env_hdr, *env_episode, *agent_a, *agent_b;
    int32_t n_envs, n_pool, smax, max_timestep;
    int32_t env_offset;   /* global id of env 0 of this shard (keys the per-env random streams; 0 on one GPU) */
    int32_t vision;       /* Agent(..., vision_range=r), maze_agent.py:16: byte 0 = agent 0's r, byte 1 = agent 1's; 1 <= r <= 4 (the stored grids keep a
                             wall border of MM_PAD = r + 1 cells); 0 in a byte = the reference default 4 */
} mm_state;

int mm_abi_version(void);
/* SHA-256 (hex) of the sources this library was compiled from, as computed by marl_maze_b200/build.py; the Python loader compares it
 * with the sources it finds next to the library and refuses (rebuilds) a stale build.  "unknown" when built by hand. */
const char *mm_source_hash(void);
const char *mm_error_string(int code);
/* last CUDA error text seen by this library on the calling thread ("" if none) */
const char *mm_last_cuda_error(void);

/* allocation sizes in bytes */
size_t mm_sizeof_pool_grid(int n_pool, int smax);
size_t mm_sizeof_pool_d2e(int n_pool, int smax);
size_t mm_sizeof_pool_hdr(int n_pool);
size_t mm_sizeof_env_grid(int n_envs, int smax);
size_t mm_sizeof_env_hdr(int n_envs);
size_t mm_sizeof_env_episode(int n_envs);
size_t mm_sizeof_agent_a(int n_envs);
size_t mm_sizeof_agent_b(int n_envs);
size_t mm_sizeof_finalize_scratch(int n, int smax);
size_t mm_sizeof_generate_scratch(int n, int smax); /* 16: the generator keeps its working state in shared memory; the scratch argument of mm_generate* stays in the ABI (a non-NULL device pointer to that many bytes) */

/* Agent.__init__ state for every agent (maze_agent.py:16-57): x=y=0, facing south, exit_len=-1, empty memory. */
int mm_init_state(const mm_state *st, void *stream);

/*
 * Parity injection: pack byte layouts recorded from the reference into pool entries [first, first+n).
 * Replaces nothing in the reference (it exists so that identical mazes can be replayed); computes the
 * dir-to-exit field by a tree walk from the exit (the closed form of Agent.exit_route, maze.py:148-154,
 * maze_agent.py:227-260).   layouts: [n][smax][smax] u8 (0 path / 1 wall, row-major, top-left W x H used)
 * hdr: [n][11] i32 = W,H,p0x,p0y,p1x,p1y,ex,ey,kx,ky,shortest_path_len.   scratch: mm_sizeof_finalize_scratch.
 */
int mm_load_layouts(const mm_state *st, int first, int n, const uint8_t *layouts, const int32_t *hdr, void *scratch, void *stream);

/*
 * K1 -- batched maze generator.  Replaces Maze.build_maze/get_neighbors/set_start/set_end/set_key/
 * get_shortest_path (maze.py:170-273): randomized DFS carve with the corridor-length bias, exit on the left or
 * right edge (best of `difficulty` candidates), key by rejection sampling off the start->exit path.
 * Random stream: Philox4x32-10 keyed by (seed, maze id) -- see DESIGN.md for the draw mapping.  maze id of entry i (0-based in
 * this call) = id_base + i when id_mod == 0, else id_base + (i % id_mod) * id_mul + i / id_mod (pool slot e + k*E -> global
 * env slot * episodes + k, which keeps ids independent of how envs are sharded over GPUs).
 * side_lo..side_hi: the maze side is (randint(side_lo, side_hi))*2-1 like rand_range (maze.py:172); side_lo >= 4 (7x7), the
 * smallest size on which the reference's rejection loops for exit and key are guaranteed to terminate.
 */
int mm_generate(const mm_state *st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty,
                uint64_t seed, uint32_t id_base, int id_mod, int id_mul, void *scratch, void *stream);
/* as mm_generate with at most max_blocks thread blocks of 64 mazes in flight (0 = no cap): a background build on a side stream that leaves
 * most of every SM to the kernels it runs beside */
int mm_generate_ex(const mm_state *st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty,
                uint64_t seed, uint32_t id_base, int id_mod, int id_mul, void *scratch, int max_blocks, void *stream);
/* as mm_generate_ex for the slots i of [first, first + n) with only[i] != 0 (only == NULL: all): the incremental pool refill -- the reference builds one
 * maze per reset (maze.py:57), so between two rollouts only the pool slots whose mazes were consumed are built anew, the others stay.
 * height_cells > 0: rectangular mazes, Maze(default_size=[w, h]) with rand_sizes False (maze.py:26-27,171-178): every maze is 2*side_lo-1 wide and
 * 2*height_cells-1 high (both <= smax) and no size is drawn; 0: square mazes of a drawn side as in mm_generate */
int mm_generate_masked(const mm_state *st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty,
                uint64_t seed, uint32_t id_base, int id_mod, int id_mul, void *scratch, int max_blocks, const uint8_t *only, int height_cells,
                void *stream);

/*
 * Maze.reset() (maze.py:55-72) + Agent.reset() (maze_agent.py:59-79) for every env with reset_mask[e] != 0
 * (all envs when reset_mask == NULL): takes the next pool maze, places agent i on shortest_path[i], emits the
 * first observations/masks.   obs: [E][2][65] f32, masks: [E][2][6] u8.
 */
int mm_reset(const mm_state *st, const uint8_t *reset_mask, float *obs, uint8_t *masks, void *stream);
/* One agent of one env, from the host: reset != 0 = Agent.reset(x, y) (maze_agent.py:59-79: position, facing south, flags, memory, route and bounding box
 * cleared; time_from_last_seen kept), reset == 0 = Agent.move(x, y, direction) (maze_agent.py:85-87).  The agent's next observation is computed by the
 * next mm_step_obs, as in the reference (neither method observes). */
int mm_agent_place(const mm_state *st, int env, int agent, int x, int y, int direction, int reset, void *stream);

/*
 * K2 -- fused step + observation.  Replaces Maze.step / single_agent_step (maze.py:74-163) and
 * Agent.get_observations with all helpers (maze_agent.py:89-358).
 *   actions [E][2][2] u8 (move 0..4 relative to facing, mark 0/1)
 *   obs [E][2][65] f32, masks [E][2][6] u8, reward [E] f32, done [E] u8
 *   auto_reset != 0: a finished env is reset in the same launch and obs/masks are those of the new episode
 *   (the PPO.get_batch loop, PPO.py:120-130); reward/done always describe the step just taken.
 *   actions_out (may be NULL): when actions == NULL the kernel draws uniform mask-legal actions itself
 *   (Philox keyed by seed, env, agent, step) from the masks it emitted last step and records them here.
 */
int mm_step_obs(const mm_state *st, const uint8_t *actions, float *obs, uint8_t *masks, float *reward, uint8_t *done,
                int auto_reset, uint64_t action_seed, uint8_t *actions_out, void *stream);

/* Debug/test readback of the packed agent state as the fields of tools/ref_harness.py AGENT_FIELDS: out [E][2][18] i32 */
int mm_unpack_agents(const mm_state *st, int32_t *out, void *stream);
/* out [E][8] i32 = t, key_x, key_y (-1 if taken), pool index, W, H, err, episode */
int mm_unpack_envs(const mm_state *st, int32_t *out, void *stream);
/* byte layout of env e's working grid (marks included): out [smax][smax] u8 */
int mm_unpack_layout(const mm_state *st, int env, uint8_t *out, void *stream);
/* byte layout + dir-to-exit of pool maze p: out_layout [smax][smax] u8, out_d2e [smax][smax] u8, out_hdr [11] i32 */
int mm_unpack_pool(const mm_state *st, int p, uint8_t *out_layout, uint8_t *out_d2e, int32_t *out_hdr, void *stream);

/* Self-test: counts (into *mismatches, a zeroed device u64) the pairs a in [-amax, amax], b in [1, bmax] for which K2's
 * reciprocal-based integer division differs from IEEE div.rn -- must be 0 (the ratio features of the observation rely on it). */
int mm_selftest_div(int amax, int bmax, uint64_t *mismatches, void *stream);

/*
 * K3 -- GAE reverse scan.  Replaces PPO.get_GAEs (PPO.py:193-203) over fixed-horizon [T][E] buffers with
 * episode boundaries marked by done[t][e]; v_boot[e] = V(s_T) bootstraps episodes still open at t = T-1
 * (extension).  adv [T][E] f32; rtg (may be NULL) = adv + value (PPO.py:46).  gamma/lam are doubles
 * because the reference multiplies them as python floats before the product meets an fp32 tensor (PPO.py:201).
 */
int mm_gae(const float *reward, const float *value, const uint8_t *done, const float *v_boot, float *adv, float *rtg,
           int T, int E, double gamma, double lam, void *stream);

/*
 * K4 -- actor / critic forward over the whole env batch + fused action sampling.  Replaces Actor.forward (with Projection and
 * m_Attention, networks.py:31-41,58-65,75-82), Critic.forward (networks.py:96-102) and PPO.get_action (PPO.py:170-186).
 *   weights: one flat fp32 buffer; block offsets (in floats) are returned by mm_policy_offsets (order: proj_w [23][20][4],
 *            proj_b [23][20], proj_col [23], proj_dim [23], att_k [10][20], att_q [10][20], att_v [20][20], l0_w [264][460], l0_b,
 *            l1_w [264][264], l1_b, l2_w, l2_b, head_w [6][264] (5 move rows + mark row), head_b [6], c0_w [64][130], c0_b, c1_w
 *            [64][64], c1_b, c2_w [64], c2_b [1], total, then l{0,1,2}_w{hi,lo}: the trunk weights split as w = hi + lo with hi = w
 *            truncated to TF32 -- only read by the tensor-core path).  proj_col[i] is the first obs column projection i reads: 0 for every i
 *            reproduces the reference's Projection.forward (networks.py:59-63); sum(FEATURE_DIMS[:i]) is the indexed variant.
 *   obs [E][2][65] f32 (8-byte aligned: the critic reads the rows as float2), masks [E][2][6] u8, scratch: mm_sizeof_policy_scratch(E) bytes;
 *   weights and scratch 16-byte aligned (MM_ERR_BAD_ARG otherwise).
 *   actions_in == NULL: sample (Philox keyed by seed, counter, env_offset+env, agent) into actions_out [E][2][2] u8;
 *   actions_in != NULL: evaluate those actions instead (actions_out unused).
 *   logp [E] f32 = joint log-prob of both agents' actions (PPO.py:118,121); value [E] f32 (may be NULL); logits_out [E][2][6]
 *   f32 (may be NULL; 5 move logits + mark logit, unmasked).
 */
#define MM_POLICY_N_OFFSETS 44
int mm_policy_offsets(int32_t *out /* [MM_POLICY_N_OFFSETS]: the blocks listed above, then c0_wt, c1_wt (critic weights transposed), tokm, tokb
                                      (per-token affine maps), l{0,1,2}_{h16,l16} (FP16 hi / lo split of 2^e W, fp16 [264][kpad] with kpad = 480,
                                      288, 288, two halves per float slot) and l{0,1,2}_asc (the scalar 2^-e) -- read by MM_POLICY_FP16_SPLIT; lh_h16, lh_l16,
                                      lh_asc: the same split of the six head rows, fp16 [16][288] (rows 6.., columns 264.. zero) -- read by
                                      MM_POLICY_FUSED_TRUNK */);
/* critic only: value [E] = Critic(obs [E][2][65]) (networks.py:96-102); used for the bootstrap value V(s_T) */
int mm_critic_forward(const float *weights, const float *obs, int n_envs, float *value, void *stream);
size_t mm_sizeof_policy_scratch(int n_envs);
int mm_policy_forward(const float *weights, const float *obs, const uint8_t *masks, int n_envs, void *scratch, const uint8_t *actions_in,
                      uint8_t *actions_out, float *logp, float *value, float *logits_out, int env_offset, uint64_t seed, uint64_t counter,
                      int flags, const uint64_t *counter_dev, void *stream);
/* counter_dev (may be NULL): device u64 added to `counter` inside the kernel, so that a captured CUDA graph of a whole rollout
 * (counter = step index baked in, counter_dev bumped by mm_counter_add at the end of the graph) draws fresh numbers on every replay. */
int mm_counter_add(uint64_t *counter_dev, uint64_t v, void *stream);
#define MM_POLICY_TCGEN05 1 /* flags: trunk GEMMs as error-compensated 3xTF32 tcgen05.mma (TMA + TMEM); 0 = fp32 SIMT tiles */
#define MM_POLICY_OVERLAP_CRITIC 2 /* flags: run the critic on an internal side stream forked from / joined to `stream` (capturable) */
#define MM_POLICY_FP16_SPLIT 4 /* flags (with MM_POLICY_TCGEN05): the 3xFP16 kernel -- kind::f16 MMAs on the fp16 hi/lo split of the operands, 128 x 128|144
                                  tiles, two CTAs per SM (csrc/mm_linear16.cu) -- instead of the 3xTF32 one */
#define MM_POLICY_FUSED_TRUNK 8 /* flags (with MM_POLICY_TCGEN05 | MM_POLICY_FP16_SPLIT): the three trunk layers, the heads and the sampling as ONE persistent
                                   kernel, activations resident in shared memory between the layers (csrc/mm_trunk_fused.cu) */

/*
 * K5 -- building blocks of the PPO actor update (PPO.py:58-85: loss.backward() through Actor.layers), SURVEY 8(f).1.
 *
 * mm_wgrad_tf32x3: weight + bias gradient of one Linear layer.  dz [rows][n_out] f32 = gradient at the layer's pre-activation, h
 * [rows][k_in] f32 = the layer's input.  Writes `slabs` partial sums (geometry from mm_wgrad_geometry) of dW[n][k] = sum_r dz[r][n]
 * h[r][k] and db[n] = sum_r dz[r][n], in whichever orientation needs fewer 128 x 288 output tiles:
 *     transposed == 0: part [slabs][out_rows = n_out][ld]:     sum_s part[s][n][k] = dW[n][k] (k < k_in),  sum_s part[s][n][k_in] = db[n]
 *     transposed == 1: part [slabs][out_rows = k_in + 1][ld]:  sum_s part[s][k][n] = dW[n][k] (k < k_in),  sum_s part[s][k_in][n] = db[n]
 * computed as 3xTF32 tcgen05.mma with fp32 accumulation; n_out and k_in must be multiples of 4, base pointers 16-byte aligned.
 */
int mm_wgrad_geometry(int rows, int n_out, int k_in, int32_t *slabs, int32_t *ld, int32_t *out_rows, int32_t *transposed);
int mm_wgrad_tf32x3(const float *dz, const float *h, int rows, int n_out, int k_in, float *part, void *stream);

/*
 * mm_linear_tf32x3: the trunk GEMM of K4 as a stand-alone call, y = epi(x W^T) with x [rows][k] f32 and W [n_rows_w <= 264][k] given as
 * its TF32 split (w_hi = tf32(W), w_lo = tf32(W - w_hi), both round-to-nearest).  Forward of a layer: mode MM_LINEAR_RELU, W = the
 * layer's weight, y = relu(. + bias) (Actor.forward, networks.py:36-38; Critic.forward, networks.py:96-102); gate_bits_out (may be NULL)
 * receives the ReLU pattern as [rows][9] u32, bit b of word c = y[row][32c+b] > 0 (zero beyond column n_rows_w).  Data gradient of a layer: W = the layer's weight TRANSPOSED,
 * x = dZ, mode MM_LINEAR_GATE (y = acc where the bit of gate_bits -- the pattern of the ReLU below -- is set, else 0) or MM_LINEAR_PLAIN
 * (y = acc).  n_rows_w columns are written with row pitch ldy (floats), so a wider result (the 460-wide dX of layer 0) is made of column
 * blocks.
 */
#define MM_LINEAR_RELU 0
#define MM_LINEAR_GATE 2
#define MM_LINEAR_PLAIN 3
int mm_linear_tf32x3(const float *x, int rows, int k, const float *w_hi, const float *w_lo, int n_rows_w, const float *bias,
                     const uint32_t *gate_bits, float *y, int ldy, int mode, uint32_t *gate_bits_out, void *stream);

/*
 * mm_linear_f16x3: the forward GEMM as 3xFP16 (csrc/mm_linear16.cu): y[:, 0:n_rows_w] = relu(acc_scale * x W16^T + bias), x [rows][k] f32,
 * W16 = fp16 hi / lo split of 2^e W, each fp16 [n_rows_w <= 264][kpad] with kpad = k rounded up to a multiple of 32 (zero padded), and
 * *acc_scale (a DEVICE scalar) = 2^-e.  gate_bits_out (may be NULL) as in mm_linear_tf32x3.  Same result as MM_LINEAR_RELU of
 * mm_linear_tf32x3 to ~1e-6 relative, at twice the tensor-core rate and half the weight traffic.  Forward only: gradients do not fit
 * FP16's exponent range.
 */
int mm_linear_f16x3(const float *x, int rows, int k, const void *w_hi16, const void *w_lo16, int n_rows_w, int kpad, const float *acc_scale,
                    const float *bias, float *y, int ldy, uint32_t *gate_bits_out, void *stream);

/*
 * mm_ppo_heads_loss: heads + clipped surrogate, forward and backward (PPO.get_log_probs PPO.py:154-168 for both agents; ratio, clip,
 * loss PPO.py:66-72; autograd down to the last trunk activation).  h2 [2E][264] f32 = last trunk activation (agent rows 2e, 2e+1),
 * head_w [6][264] (5 move rows + mark row), head_b [6], masks [2E][6] u8, actions [2E][2] u8 (move, mark), old_logp [E], adv [E].
 *   loss   = scale * sum_e -min(ratio_e A_e, clamp(ratio_e, 1-clip, 1+clip) A_e),   ratio_e = exp(joint_e - old_logp_e)
 * Outputs: dz2 [2E][264] = dloss/d(pre-activation of the last trunk layer); logp [E] = joint_e (may be NULL); part [blocks][ld]
 * (mm_ppo_loss_geometry) per-block partial sums: [0, 6*264) dloss/dhead_w, [6*264, 6*264+6) dloss/dhead_b, [6*264+6] loss.
 */
int mm_ppo_loss_geometry(int32_t *blocks, int32_t *ld);
int mm_ppo_heads_loss(const float *h2, const float *head_w, const float *head_b, const uint8_t *masks, const uint8_t *actions,
                      const float *old_logp, const float *adv, int n_envs, float clip, float scale, float *dz2, float *logp,
                      float *part, void *stream);

/*
 * mm_segment_sum: part [blocks][n_seg][cols] <- per-block partial sums of out[s][c] = sum over rows r with seg[r] == s of x[r][c]
 * (n_seg <= 8, cols a multiple of 4, blocks = mm_segment_sum_blocks(rows)).  The backward of gathering the embedding of each agent row
 * from the few distinct (projection + attention) outputs that `Actor.embed` evaluates (networks.py:31-35 fed identical prefixes).
 */
int mm_segment_sum_blocks(int rows);
int mm_segment_sum(const float *x, const int64_t *seg, int rows, int cols, int n_seg, float *part, void *stream);
/* the gather itself: out [rows][cols] <- src[seg[r]][:] for src [n_src <= 8][cols] (cols a multiple of 4, n_src * cols * 4 <= 48 KB; seg values are
 * clamped to [0, n_src)); src and out 16-byte aligned */
int mm_gather_rows(const float *src, const int64_t *seg, int rows, int cols, int n_src, float *out, void *stream);

/*
 * The 23-token embedding of K4 as a stand-alone pair (Projection + m_Attention, networks.py:58-65,75-82, and their backward).
 * weights: the K4 buffer layout (mm_policy_offsets); only tokm [60][23][4], tokb [60][23] (the per-token affine maps: rows 0-19 token,
 * 20-29 key, 30-39 query, 40-59 value), proj_col [23] and proj_dim [23] are read.
 *   mm_tokens_forward:  x0 [rows][460] = token + attention(token) for obs [rows][65] -- a function of the maps alone (this is the forward the backward
 *                       below differentiates; second-generation kernel, mm_tokens_mma.cu).
 *   mm_tokens_forward_full: the same values through the third-generation kernel the rollout uses (mm_tokens_proj.cu: keys / queries / values as
 *                       tensor-path products of the token tile); it reads rows 0-19 of tokm / tokb AND att_q [10][20], att_k [10][20], att_v [20][20] of
 *                       the full K4 weight buffer (policy.pack_weights), which must be consistent with the maps.
 *   mm_tokens_backward: part [mm_tokens_backward_blocks()][60][23][5] <- per-block partial sums of d loss / d tokm (first four of the
 *                       five) and d loss / d tokb (the fifth), given d_x0 [rows][460] = d loss / d x0; scratch:
 *                       mm_sizeof_tokens_backward_scratch(rows) bytes (the per-row key / query / value gradients, 3680 B per row).
 */
int mm_tokens_forward(const float *weights, const float *obs, int rows, float *x0, void *stream);
int mm_tokens_forward_full(const float *weights, const float *obs, int rows, float *x0, void *stream);
int mm_tokens_backward_blocks(void);
size_t mm_sizeof_tokens_backward_scratch(int rows);
int mm_tokens_backward(const float *weights, const float *obs, const float *d_x0, int rows, void *scratch, float *part, void *stream);

#ifdef __cplusplus
}
#endif
#endif
