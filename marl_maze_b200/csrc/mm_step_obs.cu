// mm_step_obs.cu -- K2: fused environment step + observation (+ in-launch auto-reset), sm_100a.
//
// Replaces Maze.step / Maze.single_agent_step (maze.py:74-163), Maze.reset's agent placement (maze.py:55-72),
// Agent.reset (maze_agent.py:59-79) and Agent.get_observations with every helper (maze_agent.py:89-358).
//
// Mapping: one lane per AGENT, the two agents of an environment sit in adjacent lanes (2e, 2e+1), a warp owns
// 16 environments.  Everything that is a function of one agent's own position and the grid (rays, marks,
// dead ends, key / exit / other-agent sightings, bounding box) runs on both lanes at once; the reference's
// sequential coupling (agent 0 stepped and observed before agent 1, route sharing mutating the other agent,
// key pickup priority, mark overwrite order) is resolved with a handful of warp shuffles between the pair.
//
// Memory: per-agent 11-row window of the bit-plane grid is read straight from HBM (rows y-5..y+5, one 16-byte
// (lo,hi) pair per row); marks are applied in registers and the marked row written back; observations are
// staged in shared memory ([lane][65] floats, stride 65 = conflict-free) and leave as one coalesced
// 8320-byte stream per warp directly into the caller's rollout buffer.
#include "mm_env.cuh"

namespace mm {

// Block = 2 warps, 10 blocks per SM (96 registers): measured on B200 against 128-thread blocks x 5 (same 20 warps per SM): 0.1948 vs
// 0.1970 ms per 1 Mi-maze launch (profiles/r02b_k2_variants.jsonl) -- finer-grained block turnover.
#ifndef MM_K2_THREADS
#define MM_K2_THREADS 64
#endif
constexpr int kThreads = MM_K2_THREADS;
#ifndef MM_K2_MINBLOCKS
#define MM_K2_MINBLOCKS (640 / MM_K2_THREADS)
#endif
#ifndef MM_K2_MINBLOCKS2
#define MM_K2_MINBLOCKS2 3
#endif
// Groups of 32 agents a warp carries through the kernel.  2 = software-pipelined (both groups' loads in flight before the first
// observation phase).  Measured on B200 (profiles/r01_notes.md): 0.246 ms vs 0.232 ms for 1 -- at 1.1 GB of mostly sector-granular
// DRAM traffic per launch the memory system, not exposed latency, is the limiter, and the pipelined form costs a block of occupancy.
#ifndef MM_K2_GROUPS
#define MM_K2_GROUPS 1
#endif
// Window transport.  1 = each lane's 11 window rows (176 contiguous bytes) travel global -> shared as ONE cp.async.bulk (TMA bulk copy,
// completion on the warp's mbarrier) instead of 11 cp.async.cg of 16 bytes: a 16-byte request still moves a whole 32-byte sector over the
// L1 <-> L2 crossbar, so the per-row form sent 23 M sector requests per launch (1 Mi mazes) for 6.5 M distinct sectors -- the request
// path (l1tex2xbar 56 % busy in profiles/r01h) was the most loaded unit after DRAM.
#ifndef MM_K2_BULK
#define MM_K2_BULK 1
#endif

// One axis ray seen from the agent.  cw: bit j-1 = wall (or out of bounds) at distance j, j = 1..5.
// latopen: bit j-1 = a cell left or right of the ray cell at distance j is open, j = 1..4.  R = the agent's vision_range (1..4).
// Returns n = number of visible cells (maze_agent.py:218-225) and the dead-end code de in units of 1/R
// (get_dead_ends, maze_agent.py:143-181: R = wall adjacent, R-j = dead end seen at distance j, 0 = none; StepParams::de_tab maps it to the value).
__device__ __forceinline__ void ray_eval(uint32_t cw, uint32_t latopen, uint32_t own_open, int R, int& n, int& de) {
    const uint32_t RM = (1u << R) - 1u;
    n = __ffs(cw | (1u << R)) - 1;
    const uint32_t open = ~cw;
    const uint32_t FWD = (open >> 1) & RM;                  // neighbour ahead of cell j is open
    const uint32_t BACK = ((open << 1) | own_open) & RM;    // neighbour behind cell j is open
    const uint32_t LAT = latopen & RM;
    const uint32_t CNT1 = ~LAT & (BACK ^ FWD) & RM;          // exactly one open neighbour
    const uint32_t BRK0 = (LAT | (~CNT1 & ~FWD)) & RM;       // a turn, or a wall ahead without being a dead end
    const uint32_t ev = BRK0 | CNT1;
    const int j = __ffs(ev);                                // first event along the ray (0 = none within R)
    int d = (j && ((CNT1 >> (j - 1)) & 1u)) ? R - j : 0;
    de = (cw & 1u) ? R : d;
}

// Where does (dx,dy) lie relative to the agent: j = 0 same cell, 1.. distance along axis ray `dir`, 99 = off-axis.
__device__ __forceinline__ void locate(int dx, int dy, int& dir, int& j) {
    const int adx = abs(dx), ady = abs(dy);
    if (dx == 0) { dir = dy < 0 ? 0 : 2; j = ady; }
    else if (dy == 0) { dir = dx > 0 ? 1 : 3; j = adx; }
    else { dir = 0; j = 99; }
}

__device__ __forceinline__ uint32_t rev5(uint32_t v) { return __brev(v & 0x1fu) >> 27; }
__device__ __forceinline__ uint32_t rot4r(uint32_t m, int f) { return ((m | (m << 4)) >> f) & 0xfu; }  // abs-direction bits -> relative
// a / b for small integers (|a|, b <= 4096, b >= 1), correctly rounded like IEEE div.rn -- which is what makes the 7 ratio
// features bit-identical to the reference's float64-divide-then-cast.  div.rn itself takes a ~100-instruction
// slow path whenever the numerator is 0 (most of the time here); this is the Markstein sequence (correctly rounded
// reciprocal, one FMA residual, one FMA correction), checked exhaustively against div.rn by mm_selftest_div.
__device__ __forceinline__ float idiv_rcp(float fa, float fb, float rcp) {
    const float q = __fmul_rn(fa, rcp);
    const float e = __fmaf_rn(-q, fb, fa);
    return __fmaf_rn(e, rcp, q);
}
__device__ __forceinline__ float fdiv(int a, int b) { const float fb = (float)b; return idiv_rcp((float)a, fb, __frcp_rn(fb)); }

// Everything one lane carries from the step phase to the observation phase.
struct LaneCtx {
    long long g;      // global agent index
    int e, a;         // environment, agent (0 = tag 2 RED, 1 = tag 3 BLUE; main.py:18-19)
    bool valid;
    Agent me;
    uint32_t t, keyp, err, pidx;
    int W, Hh, ex, ey, kx, ky;
    float reward;
    uint32_t done;
    bool mk;          // this agent marked its (pre-move) cell this step
    int px, py;       // pre-move cell
    bool want_reset;
    bool have_d2e;    // phase 1 put this agent's dir-to-exit row in flight (an agent that knows the exit and moved ALONG its route)
    bool d2e_known;   // phase 1 already knows the dir-to-exit of the new cell without reading the field (see k2_phase1)
    float* stage;     // this warp's 32 x 65-float staging area in shared memory (doubles as the landing zone of the window rows)
    uint64_t* bar;    // this warp's mbarrier (MM_K2_BULK: completion of the window bulk copies)
};

__device__ __forceinline__ uint32_t k2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// L2 residency hints (MM_K2_L2HINT: bit 0 = state evict_last, bit 1 = streams evict_first).  Per step the kernel streams ~0.9 GB through a 126 MB L2 (window rows in, observations out: no reuse) and reads +
// writes 56 B of agent / env state per env -- 59 MB at 1 Mi envs, the only data it touches again in the next launch.  State loads and stores carry an
// evict_last policy, the window bulk copies and the observation bulk store evict_first, so that in an env-only stepping loop (BASELINE config[3]) the
// state stays resident between launches: its 2 x 59 MB leave the DRAM traffic and the FIRST of the two dependent round trips (state -> window address)
// becomes an L2 hit.
#ifndef MM_K2_PREFETCH_WARPS
#define MM_K2_PREFETCH_WARPS 0
#endif
#ifndef MM_K2_L2HINT
#define MM_K2_L2HINT 0
#endif
__device__ __forceinline__ uint64_t k2_policy_evict_last() { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); return pol; }
__device__ __forceinline__ uint64_t k2_policy_evict_first() { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); return pol; }
__device__ __forceinline__ uint4 k2_ld_keep(const uint4* a, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t k2_ld_keep(const uint32_t* a, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void k2_st_keep(uint4* a, uint4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void k2_st_keep(uint32_t* a, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory");
}

// Phase 1: load state, S1 (Maze.step / single_agent_step), reward/done, and put the window rows of the NEW position in flight.
template <bool kResetOnly>
__device__ __forceinline__ void k2_phase1(const StepParams& p, LaneCtx& c, const int lane) {
    const long long g = c.g;
    const int e = c.e, a = c.a;
    const bool valid = c.valid;
    uint4 H = make_uint4(0, 0, 0, 0), A = make_uint4(0, 0, 0, 0);
    uint32_t B = 0;
#if MM_K2_L2HINT & 1
    if (valid) { const uint64_t keep = k2_policy_evict_last(); H = k2_ld_keep(&p.env_hdr[e], keep); A = k2_ld_keep(&p.agent_a[g], keep); B = k2_ld_keep(&p.agent_b[g], keep); }
#else
    if (valid) { H = p.env_hdr[e]; A = p.agent_a[g]; B = p.agent_b[g]; }
#endif
    Agent me = unpack_agent(A, B);
    uint32_t t = H.x, keyp = H.z & 1u, err = (H.z >> 1) & 1u, pidx = H.w;
    int W = (H.z >> 8) & 0xff, Hh = (H.z >> 16) & 0xff;
    int ex = H.y & 0xff, ey = (H.y >> 8) & 0xff, kx = (H.y >> 16) & 0xff, ky = H.y >> 24;

    float reward = 0.f;
    uint32_t done = 0;
    bool mk = false;      // this agent marked its (pre-move) cell this step
    int px = 0, py = 0;   // pre-move cell
    bool want_reset = false;
    // dir-to-exit of the cell the agent ends this step on, for an agent that knows the exit (R2 in DESIGN.md: the exit route is the tree path).  It needs
    // the field only when it moved ALONG its route (the next cell's direction is new information).  Stepping off the route -- any other move, also away
    // from the exit cell itself -- leads to a cell whose path to the exit goes straight back: direction = the reverse of the move.  Not moving keeps the
    // stored direction.  (Every agent that knew the exit used to fetch a 32-byte sector of the field every step: 64 MB of DRAM reads per 1 Mi-maze launch
    // in reset steady state, where most agents know the exit.)
    bool d2e_fetch = false, d2e_known = valid && me.ke;

    if (kResetOnly) {
        want_reset = valid && (p.reset_mask == nullptr || p.reset_mask[e] != 0);
    } else {
        // ---------------------------------------------------------------- S1: Maze.step / single_agent_step
        int move, mark;
        if (p.actions != nullptr) {
            const uint32_t aw = valid ? (uint32_t)reinterpret_cast<const uint16_t*>(p.actions)[g] : 4u;
            move = aw & 0xff; mark = (aw >> 8) & 0xff;
        } else {  // uniform mask-legal action from the masks emitted by the previous launch
            uint32_t r[4];
            philox4x32_10(t, (uint32_t)(p.env_offset + e) * 2u + (uint32_t)a, 0x4d415a45u, 0u, (uint32_t)p.action_seed, (uint32_t)(p.action_seed >> 32), r);
            uint32_t legal = me.last_mask & 0x1fu;
            const int n = __popc(legal);
            move = 4;
            if (n) {
                int k = (int)(((uint64_t)r[0] * (uint32_t)n) >> 32);
                for (int i = 0; i < 4; i++) if (i < k) legal &= legal - 1;
                move = __ffs(legal) - 1;
            }
            mark = ((me.last_mask >> 5) & 1u) ? (int)(r[1] & 1u) : 0;
            if (valid && p.actions_out != nullptr) reinterpret_cast<uint16_t*>(p.actions_out)[g] = (uint16_t)(move | (mark << 8));
        }
        t += 1;  // maze.py:75
        px = me.x; py = me.y;
        if (mark == 1) { mk = true; me.lmx = px; me.lmy = py; me.mkv = 1; }  // maze.py:132-134 (grid write below)
        uint32_t cand_key = 0;
        if (move < 4) {  // maze.py:137-162
            const int nd = (move + me.dir) & 3;
            const int nx = me.x + (nd == 1) - (nd == 3), ny = me.y + (nd == 2) - (nd == 0);
            if (nx < 0 || nx >= W || ny < 0 || ny >= Hh) {
                err = 1;  // the reference prints and then indexes out of range (maze.py:141-145): treated as `stop` + error flag
            } else {
                if (me.ke) {  // exit_route pop/push == walking along / against the tree path to the exit (maze.py:148-154)
                    const bool at_end_pre = (me.x == ex && me.y == ey);
                    const bool along = !at_end_pre && nd == me.d2e;
                    me.exit_len += along ? -1 : 1;
                    if (along) { d2e_fetch = true; d2e_known = false; }
                    else me.d2e = (nd + 2) & 3;
                }
                me.x = nx; me.y = ny; me.dir = nd;
                cand_key = keyp && nx == kx && ny == ky;
                me.mem = (me.mem >> 3) | ((uint32_t)(move + 1) << 9);
            }
        } else if (move > 4) {
            err = 1;
        }
        // key pickup: agent 0 is stepped first and takes the key if both arrive together (maze.py:157-161)
        const uint32_t c0 = __shfl_sync(kFull, cand_key, lane & ~1), c1 = __shfl_sync(kFull, cand_key, lane | 1);
        const uint32_t got = a == 0 ? c0 : (c1 & ~c0 & 1u);
        if (got) { me.has = 1; me.team = 1; }
        const uint32_t got_any = c0 | c1;
        keyp &= ~got_any & 1u;
        err |= __shfl_xor_sync(kFull, err, 1);
        // reward / done (maze.py:115-121): needs only post-move positions and has_key
        const uint32_t ohas = __shfl_xor_sync(kFull, me.has, 1);
        const int oxp = __shfl_xor_sync(kFull, me.x, 1), oyp = __shfl_xor_sync(kFull, me.y, 1);
        reward = got_any ? 0.5f : 0.f;
        if ((me.has | ohas) && oxp == me.x && oyp == me.y && me.x == ex && me.y == ey) { reward = 1.f; done = 1; }
        else if ((int)t >= p.max_t) done = 1;
        want_reset = valid && done && p.auto_reset;
    }


    if (!kResetOnly) {
        // the 12 in-flight 16-byte rows per lane go global -> shared with cp.async.cg (LDGSTS.BYPASS): they never occupy L1, whose
        // capacity otherwise caps the number of outstanding window loads per SM (profiles/r01_notes.md).  Landing zone = the warp's
        // own observation staging area: [lane][11 rows] then [lane] field row.
        char* wbase_s = reinterpret_cast<char*>(c.stage);
#if MM_K2_BULK
        {   // one arrival carrying the byte count of every bulk copy this warp is about to issue
            const uint32_t vmask = __ballot_sync(kFull, valid);
            if (lane == 0 && vmask)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k2_smem_u32(c.bar)), "r"(176u * (uint32_t)__popc(vmask)) : "memory");
        }
#endif
        if (valid) {
            const ulonglong2* grid = (const ulonglong2*)(p.env_grid + (size_t)e * p.rows);
            const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(wbase_s + lane * 176), s1 = (uint32_t)__cvta_generic_to_shared(wbase_s + 32 * 176 + lane * 16);
#if MM_K2_BULK
#if MM_K2_L2HINT & 2
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], 176, [%2], %3;" ::"r"(s0), "l"(grid + me.y),
                         "r"(k2_smem_u32(c.bar)), "l"(k2_policy_evict_first())
                         : "memory");
#else
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 176, [%2];" ::"r"(s0), "l"(grid + me.y), "r"(k2_smem_u32(c.bar)) : "memory");
#endif
#else
#pragma unroll
            for (int r = 0; r < 11; r++) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 16 * r), "l"(grid + me.y + r) : "memory");
#endif
            // the dir-to-exit row is consumed only by an agent that knows the exit at the END of this step; one that does not know it yet
            // and learns it during the observation (a sighting, a shared route) fetches the row then -- rare -- instead of every
            // agent fetching a 32-byte sector every step
            if (d2e_fetch) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s1), "l"(p.pool_d2e + (size_t)pidx * p.smax + me.y) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    c.have_d2e = !kResetOnly && valid && d2e_fetch;
    c.d2e_known = !kResetOnly && d2e_known;
    c.me = me; c.t = t; c.keyp = keyp; c.err = err; c.pidx = pidx; c.W = W; c.Hh = Hh; c.ex = ex; c.ey = ey; c.kx = kx; c.ky = ky;
    c.reward = reward; c.done = done; c.mk = mk; c.px = px; c.py = py; c.want_reset = want_reset;
}

// Phase 2: S3-S5 (observations, masks, exit_ready override), optional in-launch reset, state write-back, observation copy-out.
// kPending = number of younger cp.async groups that may still be in flight when this group's rows are needed.
template <bool kResetOnly, int kPending>
__device__ __forceinline__ void k2_phase2(const StepParams& p, LaneCtx& c, const int lane) {
    const long long g = c.g;
    const int e = c.e, a = c.a;
    const bool valid = c.valid;
    Agent me = c.me;
    uint32_t t = c.t, keyp = c.keyp, err = c.err, pidx = c.pidx;
    int W = c.W, Hh = c.Hh, ex = c.ex, ey = c.ey, kx = c.kx, ky = c.ky;
    const float reward = c.reward;
    const uint32_t done = c.done;
    const bool mk = c.mk;
    const int px = c.px, py = c.py;
    const bool want_reset = c.want_reset;
    float* const stage = c.stage;

    bool wrote = false;
#pragma unroll 1
    for (int pass = kResetOnly ? 1 : 0; pass < 2; ++pass) {
        const bool act = valid && (pass == 0 || want_reset);
        if (pass == 1 && !__any_sync(kFull, act)) break;

        // What the partner sees of me.  In the reset pass agent 0 is observed while agent 1 still holds the
        // PREVIOUS episode's x, y, direction, has_key, knows_end (maze.py:64-71): lane 1 keeps
        // presenting its stale fields, lane 0 presents its freshly reset ones.
        int pres_x = me.x, pres_y = me.y, pres_dir = me.dir;
        uint32_t pres_has = me.has, pres_ke = me.ke;

        if (pass == 1) {
            uint32_t ep = (act && a == 0) ? p.env_episode[e] : 0u;
            ep = __shfl_sync(kFull, ep, lane & ~1);
            if (act) {
                pidx = (uint32_t)(((unsigned long long)e + (unsigned long long)ep * (unsigned)p.E) % (unsigned)p.P);
                if (a == 0) p.env_episode[e] = ep + 1;
                const uint4 ph = p.pool_hdr[pidx];
                W = ph.x & 0xff; Hh = (ph.x >> 8) & 0xff;
                const int sx = a ? (int)(ph.y & 0xff) : (int)((ph.x >> 16) & 0xff);
                const int sy = a ? (int)((ph.y >> 8) & 0xff) : (int)(ph.x >> 24);
                ex = (ph.y >> 16) & 0xff; ey = ph.y >> 24; kx = ph.z & 0xff; ky = (ph.z >> 8) & 0xff;
                keyp = 1; t = 0;  // maze.py:56
                // Agent.reset (maze_agent.py:59-79); time_from_last_seen is deliberately kept
                me.x = sx; me.y = sy; me.olsx = sx; me.olsy = sy;
                me.minx = me.maxx = sx; me.miny = me.maxy = sy;
                me.dir = 2; me.mkv = 0; me.mem = 0; me.ke = me.oke = 0; me.exit_len = -1; me.has = me.team = 0;
                if (a == 0) { pres_x = me.x; pres_y = me.y; pres_dir = me.dir; pres_has = 0; pres_ke = 0; }
                // fresh working copy of the grid (no marks yet); the two lanes interleave rows
                const ulonglong2* src = p.pool_grid + (size_t)pidx * p.rows;
                ulonglong2* dst = p.env_grid + (size_t)e * p.rows;
                for (int r = a; r < p.rows; r += 2) dst[r] = src[r];
            }
        }

        // ------------------------------------------------------------ window: rows y-5..y+5 of both bit planes
        const int x = me.x, y = me.y, f = me.dir;
        uint32_t l[11], h[11];  // the window shifted so that column x-5 is bit 0 (the agent's column is bit 5); lo / hi planes
        ulonglong2 dd = make_ulonglong2(0, 0);
        // Step pass: the rows were put in flight by phase 1 into this warp's staging area; they are consumed into registers before
        // any lane writes observation floats there (the __syncwarp below).  The (rare) reset pass must not touch that area -- it
        // already holds the step-pass observations of the lanes that are not resetting -- and loads directly.
        if (pass == 0 && !kResetOnly) {
            char* wbase_s = reinterpret_cast<char*>(stage);
            ulonglong2* slot = reinterpret_cast<ulonglong2*>(wbase_s + lane * 176);
            ulonglong2* dslot = reinterpret_cast<ulonglong2*>(wbase_s + 32 * 176 + lane * 16);
            // marks of this step (agent 0's first, then agent 1's: maze.py:80-90,132-133), exchanged while the rows are in flight
            const uint32_t omk = __shfl_xor_sync(kFull, (uint32_t)mk, 1);
            const int opx = __shfl_xor_sync(kFull, px, 1), opy = __shfl_xor_sync(kFull, py, 1);
            const bool m0 = a ? (omk != 0) : mk, m1 = a ? mk : (omk != 0);
            const int m0x = a ? opx : px, m0y = a ? opy : py, m1x = a ? px : opx, m1y = a ? py : opy;
            asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
#if MM_K2_BULK
            if (__ballot_sync(kFull, valid)) {  // bounded spin: a protocol bug must surface as a launch failure, never as a hung GPU
                const uint32_t b = k2_smem_u32(c.bar);
                uint32_t ok = 0;
                const long long t0 = clock64();
                while (!ok) {
                    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b) : "memory");
                    if (!ok && clock64() - t0 > 2000000000ll) __trap();
                }
            }
#endif
            if (act) {  // apply both marks to this lane's private copy of its window rows (dynamic row index = plain smem addressing)
                const int r0 = m0y - y + kPad, r1 = m1y - y + kPad;
                if (m0 && r0 >= 0 && r0 <= 10) { ulonglong2 v = slot[r0]; const unsigned long long bit = 1ull << (m0x + kPad); v.y |= bit; v.x &= ~bit; slot[r0] = v; }  // tag 2
                if (m1 && r1 >= 0 && r1 <= 10) { ulonglong2 v = slot[r1]; const unsigned long long bit = 1ull << (m1x + kPad); v.y |= bit; v.x |= bit; slot[r1] = v; }   // tag 3
                if (mk) p.env_grid[(size_t)e * p.rows + py + kPad] = slot[py - y + kPad];  // the marked row goes back to HBM with both marks applied
                if (c.have_d2e) dd = *dslot;
            }
#pragma unroll
            for (int r = 0; r < 11; r++) {
                const ulonglong2 v = act ? slot[r] : make_ulonglong2(~0ull, 0ull);
                l[r] = (uint32_t)(v.x >> x); h[r] = (uint32_t)(v.y >> x);
            }
            __syncwarp();  // every lane has its window in registers; the area may now receive observation floats
        } else {
            const ulonglong2* grid = pass == 0 ? (const ulonglong2*)(p.env_grid + (size_t)e * p.rows) : (p.pool_grid + (size_t)pidx * p.rows);
#pragma unroll
            for (int r = 0; r < 11; r++) {
                const ulonglong2 v = act ? __ldcg(&grid[y + r]) : make_ulonglong2(~0ull, 0ull);
                l[r] = (uint32_t)(v.x >> x); h[r] = (uint32_t)(v.y >> x);
            }
        }

        // ------------------------------------------------------------ per-direction bit masks (abs 0 N, 1 E, 2 S, 3 W)
        uint32_t l5, h5, wl4, wl5, wl6;
        uint32_t cw[4] = {0, 0, 0, 0}, lat[4] = {0, 0, 0, 0}, ownm[4] = {0, 0, 0, 0}, othm[4] = {0, 0, 0, 0};
        const uint32_t ua = (uint32_t)a;
        {
            l5 = l[5]; h5 = h[5]; wl4 = l[4] & ~h[4]; wl5 = l5 & ~h5; wl6 = l[6] & ~h[6];
#pragma unroll
            for (int j = 1; j <= 5; j++) {
                const int rn = 5 - j, rs = 5 + j;
                const uint32_t wn = l[rn] & ~h[rn], ws = l[rs] & ~h[rs];
                cw[0] |= ((wn >> 5) & 1u) << (j - 1);
                cw[2] |= ((ws >> 5) & 1u) << (j - 1);
                if (j <= 4) {
                    lat[0] |= (((~wn >> 4) | (~wn >> 6)) & 1u) << (j - 1);
                    lat[2] |= (((~ws >> 4) | (~ws >> 6)) & 1u) << (j - 1);
                    const uint32_t hn = (h[rn] >> 5) & 1u, ln = (l[rn] >> 5) & 1u, hs = (h[rs] >> 5) & 1u, ls = (l[rs] >> 5) & 1u;
                    ownm[0] |= (hn & (ln ^ ua ^ 1u)) << (j - 1); othm[0] |= (hn & (ln ^ ua)) << (j - 1);
                    ownm[2] |= (hs & (ls ^ ua ^ 1u)) << (j - 1); othm[2] |= (hs & (ls ^ ua)) << (j - 1);
                }
            }
        }
        const uint32_t own5 = h5 & (a ? l5 : ~l5), oth5 = h5 & ~own5;
        const uint32_t lat5 = ~wl4 | ~wl6;
        cw[1] = (wl5 >> 6) & 0x1fu; lat[1] = (lat5 >> 6) & 0xfu; ownm[1] = (own5 >> 6) & 0xfu; othm[1] = (oth5 >> 6) & 0xfu;
        cw[3] = rev5(wl5); lat[3] = rev5(lat5) & 0xfu; ownm[3] = rev5(own5) & 0xfu; othm[3] = rev5(oth5) & 0xfu;
        const uint32_t own_open = ((wl5 >> 5) & 1u) ^ 1u;
        // An agent standing INSIDE a wall got there through an illegal (masked-off) move: the reference has no wall check
        // (maze.py:137-155) and neither has this kernel -- positions, marks and observations keep following the reference -- but the
        // env's error flag is raised: the exit-route bookkeeping (R2 in DESIGN.md) is only defined on open cells.
        if (pass == 0 && !kResetOnly && act && !own_open) err = 1;
        const uint32_t cell_is_own = ((h5 >> 5) & 1u) & ((((l5 >> 5) & 1u) ^ ua) ^ 1u);  // layout[y][x] == tag

        int n[4], de[4];
        uint32_t N32 = 0, DE32 = 0, OWN32 = 0, OTH32 = 0, NB = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            ray_eval(cw[d], lat[d], own_open, p.vr[a], n[d], de[d]);
            const uint32_t vm = (1u << n[d]) - 1u;
            N32 |= (uint32_t)n[d] << (8 * d);
            DE32 |= (uint32_t)de[d] << (8 * d);
            OWN32 |= (uint32_t)__popc(ownm[d] & vm) << (8 * d);
            OTH32 |= (uint32_t)__popc(othm[d] & vm) << (8 * d);
            NB |= ((cw[d] & 1u) ^ 1u) << d;
        }
        // bounding box of seen cells, update_maze_minmax (maze_agent.py:269,313-328)
        if (act) {
            me.miny = min(me.miny, y - n[0]); me.maxx = max(me.maxx, x + n[1]);
            me.maxy = max(me.maxy, y + n[2]); me.minx = min(me.minx, x - n[3]);
        }

        // ------------------------------------------------------------ sightings: exit, key, other agent
        const int ox = __shfl_xor_sync(kFull, pres_x, 1), oy = __shfl_xor_sync(kFull, pres_y, 1), odir = __shfl_xor_sync(kFull, pres_dir, 1);
        int dE, jE, dK, jK, dO, jO;
        locate(ex - x, ey - y, dE, jE);
        locate(kx - x, ky - y, dK, jK);
        locate(ox - x, oy - y, dO, jO);
        const uint32_t at_end = jE == 0;
        const uint32_t visE = jE >= 1 && jE <= (int)((N32 >> (8 * dE)) & 0xffu);
        const uint32_t visK = keyp && jK >= 1 && jK <= (int)((N32 >> (8 * dK)) & 0xffu);
        const uint32_t sc = jO == 0;                                           // same cell (live x,y; maze_agent.py:202)
        const bool in_map = !(pass == 1 && a == 0);                            // agent_positions holds only agent 0 during its reset obs
        const uint32_t visO = in_map && jO >= 1 && jO <= (int)((N32 >> (8 * dO)) & 0xffu);
        const int rE = (dE - f) & 3, rK = (dK - f) & 3, rO = (dO - f) & 3;

        // ------------------------------------------------------------ sequential flag logic: obs(agent 0) then obs(agent 1)
        uint32_t s_ke = 0, s_oke = 0, s_team = 0, s_se = 0, s_vr = 0, s_od = 0;  // values as of this agent's own observation
#pragma unroll
        for (int rnd = 0; rnd < 2; ++rnd) {
            const uint32_t o_ke = __shfl_xor_sync(kFull, a ? pres_ke : me.ke, 1);
            const uint32_t o_has = __shfl_xor_sync(kFull, a ? pres_has : me.has, 1);
            uint32_t ke = me.ke, oke = me.oke, team = me.team, time = me.time + 1;  // maze_agent.py:195
            int olsx = me.olsx, olsy = me.olsy, el = me.exit_len;
            uint32_t se = at_end, vr = 0, od = 0, share = 0;
            if (sc) {  // maze_agent.py:202-213
                time = 0; vr = 0xfu; olsx = ox; olsy = oy; team |= o_has; oke |= o_ke; od = 1u << odir;
                if (ke && !o_ke) { oke = 1; share = 1; }
            }
            const uint32_t ke_before = ke;
            if (visE) { ke = 1; se = 1; if (el == -1) el = jE; }  // maze_agent.py:227-233
            if (visO) {                                            // maze_agent.py:239-260
                time = 0; olsx = ox; olsy = oy; oke |= o_ke; team |= o_has; od |= 1u << odir; vr = 1u << rO;
                const uint32_t ke_at = ke_before | (visE && (rE < rO || (rE == rO && jE <= 1)));
                if (jO == 1 && ke_at && !o_ke) { oke = 1; share = 1; }
            }
            const bool mine = (a == rnd);
            if (pass == 1 && rnd == 0) share = 0;  // nothing observed during agent 0's reset obs reaches the (about to be reset) agent 1
            if (mine && act) {
                me.ke = ke; me.oke = oke; me.team = team; me.time = time; me.olsx = olsx; me.olsy = olsy; me.exit_len = el;
                s_ke = ke; s_oke = oke; s_team = team; s_se = se; s_vr = vr; s_od = od;
            }
            const uint32_t sh = __shfl_xor_sync(kFull, mine ? share : 0u, 1);
            if (!mine && sh && act) { me.ke = 1; me.oke = 1; }  // route shared with me: agent.knows_end = agent.other_knows_end = True
        }

        // dir-to-exit of the current cell, for whoever knows the exit now (own sighting, or a route shared by the other agent)
        const bool d2e_from_phase1 = pass == 0 && !kResetOnly && c.d2e_known;   // stored / derived in phase 1: no field access at all
        if (act && me.ke && !d2e_from_phase1 && !(pass == 0 && !kResetOnly && c.have_d2e)) dd = __ldg(&p.pool_d2e[(size_t)pidx * p.smax + y]);
        const int d2e_here = d2e_from_phase1 ? me.d2e : (int)(((dd.y >> (x + kPad)) & 1ull) << 1 | ((dd.x >> (x + kPad)) & 1ull));
        if (act) me.d2e = d2e_here;
        // next_move_to_exit (maze_agent.py:113-118): exit_route[-1] == dir-to-exit of the current cell
        const uint32_t nm = (s_ke && !at_end) ? (1u << ((d2e_here - f) & 3)) : 0xfu;
        // exit_ready (maze.py:100-106): each term sampled right after that agent's own observation
        const uint32_t term = s_team & s_ke;
        const uint32_t exit_ready = (pass == 0) ? (term & __shfl_xor_sync(kFull, term, 1)) : 0u;

        // ------------------------------------------------------------ action mask (maze_agent.py:131-139, maze.py:107-113)
        const uint32_t DEr = __funnelshift_r(DE32, DE32, 8 * f);
        const uint32_t NBr = rot4r(NB, f);
        uint32_t mm;
        if (!s_se && !visK) {
            mm = ((DEr & 0xffu) == 0) | (((DEr >> 8) & 0xffu) == 0) << 1 | (((DEr >> 16) & 0xffu) == 0) << 2 | ((DEr >> 24) == 0) << 3;
        } else {
            mm = NBr;
        }
        if (visK) mm = 1u << rK;
        uint32_t stop = (s_vr != 0) && (x == ex) && (x == ey);  // (self.x, self.x) == maze.end -- sic, maze_agent.py:136
        if (exit_ready) {
            if (!at_end) mm = nm & (0u - nm);  // np.argmax of the one-hot / all-ones next_move_to_exit
            else { mm = 0; stop = 1; }
        }
        const uint32_t mask6 = mm | (stop << 4) | ((cell_is_own ^ 1u) << 5);

        if (act) {
            me.last_mask = mask6;
            // -------------------------------------------------------- observation vector (maze_agent.py:92-130)
            float* so = stage + lane * kObs;
            const uint32_t OWNr = __funnelshift_r(OWN32, OWN32, 8 * f), OTHr = __funnelshift_r(OTH32, OTH32, 8 * f);
            const uint32_t keyr = visK ? (1u << rK) : 0u;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                so[i] = (f == i) ? 1.f : 0.f;
                so[4 + i] = p.de_tab[a][(DEr >> (8 * i)) & 0xffu];      // 1 - j / vision_range, or 1 (wall), or 0
                so[8 + i] = p.mk_tab[a][(OWNr >> (8 * i)) & 0xffu];     // marked cells along the ray, 1 / vision_range each
                so[12 + i] = p.mk_tab[a][(OTHr >> (8 * i)) & 0xffu];
                so[16 + i] = (float)((s_vr >> i) & 1u);
                so[20 + i] = (float)((s_od >> i) & 1u);
                so[24 + i] = (float)((keyr >> i) & 1u);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {  // get_memory: oldest first, one-hot over moves 0..3
                const uint32_t mv = (me.mem >> (3 * i)) & 7u;
#pragma unroll
                for (int q = 0; q < 4; q++) so[28 + 4 * i + q] = (mv == (uint32_t)(q + 1)) ? 1.f : 0.f;
            }
            uint32_t lmr = 0;  // get_direction_from(last_mark_pos), maze_agent.py:297-311
            if (me.mkv) {
                if (me.lmx == x && me.lmy == y) lmr = 0xfu;
                else {
                    const uint32_t ab = (uint32_t)(me.lmy < y) | (uint32_t)(me.lmx > x) << 1 | (uint32_t)(me.lmy > y) << 2 | (uint32_t)(me.lmx < x) << 3;
                    lmr = rot4r(ab, f);
                }
            }
            const int we = (me.maxx - me.minx) ? (me.maxx - me.minx) : 1, he = (me.maxy - me.miny) ? (me.maxy - me.miny) : 1;
#pragma unroll
            for (int i = 0; i < 4; i++) { so[44 + i] = (float)((lmr >> i) & 1u); so[53 + i] = (float)((nm >> i) & 1u); }
            const float fwe = (float)we, fhe = (float)he, rwe = __frcp_rn(fwe), rhe = __frcp_rn(fhe);
            so[48] = idiv_rcp((float)(x - me.minx), fwe, rwe);
            so[49] = idiv_rcp((float)(me.maxy - y), fhe, rhe);
            so[50] = idiv_rcp((float)(me.olsx - me.minx), fwe, rwe);
            so[51] = idiv_rcp((float)(me.maxy - me.olsy), fhe, rhe);
            so[52] = (float)s_se;
            so[57] = me.exit_len < 40 ? idiv_rcp((float)me.exit_len, 40.f, 0.025f) : 1.f;
            so[58] = (float)s_oke;
            so[59] = (float)me.has;
            so[60] = (float)s_team;
            so[61] = me.time < 40u ? idiv_rcp((float)me.time, 40.f, 0.025f) : 1.f;
            so[62] = idiv_rcp((float)t, (float)p.max_t, p.inv_max_t);
            so[63] = a ? 0.f : 1.f;
            so[64] = a ? 1.f : 0.f;
            uint16_t* mo = reinterpret_cast<uint16_t*>(p.masks) + g * 3;
            mo[0] = (uint16_t)((mask6 & 1u) | ((mask6 >> 1) & 1u) << 8);
            mo[1] = (uint16_t)(((mask6 >> 2) & 1u) | ((mask6 >> 3) & 1u) << 8);
            mo[2] = (uint16_t)(((mask6 >> 4) & 1u) | ((mask6 >> 5) & 1u) << 8);
            wrote = true;
        }
    }

    // ---------------------------------------------------------------- state write-back
    if (!kResetOnly) err |= __shfl_xor_sync(kFull, err, 1);   // either agent's illegal move flags the env (agent 0's lane writes the header)
    if (valid && (!kResetOnly || want_reset)) {
        const uint4 hdr_new = make_uint4(t, (uint32_t)ex | ((uint32_t)ey << 8) | ((uint32_t)kx << 16) | ((uint32_t)ky << 24),
                                         keyp | (err << 1) | ((uint32_t)W << 8) | ((uint32_t)Hh << 16), pidx);
#if MM_K2_L2HINT & 1
        const uint64_t keep = k2_policy_evict_last();
        k2_st_keep(&p.agent_a[g], pack_agent(me), keep);
        k2_st_keep(&p.agent_b[g], me.time, keep);
        if (a == 0) {
            k2_st_keep(&p.env_hdr[e], hdr_new, keep);
#else
        p.agent_a[g] = pack_agent(me);
        p.agent_b[g] = me.time;
        if (a == 0) {
            p.env_hdr[e] = hdr_new;
#endif
            if (!kResetOnly) { p.reward[e] = reward; p.done[e] = (uint8_t)done; }
        }
    }

    // ---------------------------------------------------------------- observations: shared memory -> HBM, coalesced per warp
    __syncwarp();
    const uint32_t wmask = __ballot_sync(kFull, wrote);
    const long long gbase = g - lane;
    if (wmask == kFull && p.obs_vec4) {
#ifdef MM_K2_LSU_COPYOUT
        const float4* s4 = reinterpret_cast<const float4*>(stage);
        float4* o4 = reinterpret_cast<float4*>(p.obs + gbase * kObs);
#pragma unroll 4
        for (int i = lane; i < 32 * kObs / 4; i += 32) __stcs(&o4[i], s4[i]);
#else
        // the warp's 32 x 65 floats are one contiguous 8320-byte run in the rollout buffer: hand it to the bulk-copy engine
        // (cp.async.bulk shared -> global, one instruction) instead of 16 LDS.128 + STG.128 per lane
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy STS above -> visible to the async proxy
        __syncwarp();
        if (lane == 0) {
#if MM_K2_L2HINT & 2
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p.obs + gbase * kObs),
                         "r"((uint32_t)__cvta_generic_to_shared(stage)), "r"(32 * kObs * 4), "l"(k2_policy_evict_first())
                         : "memory");
#else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.obs + gbase * kObs),
                         "r"((uint32_t)__cvta_generic_to_shared(stage)), "r"(32 * kObs * 4)
                         : "memory");
#endif
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the read
        }
#endif
    } else {
        for (uint32_t m = wmask; m; m &= m - 1) {
            const int L = __ffs(m) - 1;
            const float* s = stage + L * kObs;
            float* o = p.obs + (gbase + L) * kObs;
            for (int i = lane; i < kObs; i += 32) o[i] = s[i];
        }
    }
}

// kGroups = 2: every warp carries TWO groups of 32 agents through the kernel, software-pipelined -- both groups' state loads and
// window rows are in flight before the first observation phase starts, so the second DRAM round trip of group A hides behind
// the step phase of group B, and B's behind A's whole observation phase.
template <bool kResetOnly, int kGroups>
__global__ void __launch_bounds__(kThreads, kGroups == 2 ? MM_K2_MINBLOCKS2 : MM_K2_MINBLOCKS) k_step_obs(const StepParams p) {
    extern __shared__ __align__(16) float s_obs[];  // [kGroups][kThreads][65]
    const int tid = threadIdx.x, lane = tid & 31;
    LaneCtx c[kGroups];
#if MM_K2_PREFETCH_WARPS > 0
    // A warp's first act is a DRAM round trip for its agent / env state, with nothing else to do while it waits (24 % of the kernel's stall samples sit on
    // the first use of that load).  Blocks are dispatched in index order as residency slots free up, so the warp MM_K2_PREFETCH_WARPS groups further on starts
    // about one block lifetime from now: ask the L2 for ITS state lines now (7 lines per warp: 4 of agent_a, 1 of agent_b, 2 of env_hdr).
    if (!kResetOnly) {
        const long long gp = ((long long)blockIdx.x * kGroups * kThreads + (tid & ~31)) + (long long)MM_K2_PREFETCH_WARPS * 32;   // first agent of that warp
        if (gp < 2ll * p.E) {
            const void* a = nullptr;
            if (lane < 4) a = &p.agent_a[gp + 8 * lane];
            else if (lane == 4) a = &p.agent_b[gp];
            else if (lane < 7) a = &p.env_hdr[(gp >> 1) + 8 * (lane - 5)];
            if (a) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
    }
#endif
#pragma unroll
    for (int gi = 0; gi < kGroups; gi++) {
        c[gi].g = ((long long)blockIdx.x * kGroups + gi) * kThreads + tid;
        c[gi].e = (int)(c[gi].g >> 1); c[gi].a = (int)(c[gi].g & 1); c[gi].valid = c[gi].e < p.E;
        c[gi].stage = s_obs + (gi * kThreads + (tid & ~31)) * kObs;
        c[gi].bar = reinterpret_cast<uint64_t*>(s_obs + kGroups * kThreads * kObs) + gi * (kThreads / 32) + (tid >> 5);
#if MM_K2_BULK
        if (!kResetOnly) {
            if (lane == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(k2_smem_u32(c[gi].bar)) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncwarp();
        }
#endif
        k2_phase1<kResetOnly>(p, c[gi], lane);
    }
    if (kGroups == 2) { k2_phase2<kResetOnly, 1>(p, c[0], lane); k2_phase2<kResetOnly, 0>(p, c[kGroups - 1], lane); }
    else k2_phase2<kResetOnly, 0>(p, c[0], lane);
}

__global__ void k_selftest_div(int amax, int bmax, unsigned long long* mismatches) {
    const long long n = (2ll * amax + 1) * bmax;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(i / bmax) - amax, b = (int)(i % bmax) + 1;
        const float want = __fdiv_rn((float)a, (float)b), got = fdiv(a, b);
        if (__float_as_uint(want) != __float_as_uint(got)) atomicAdd(mismatches, 1ull);
    }
}
cudaError_t launch_selftest_div(int amax, int bmax, unsigned long long* mismatches, cudaStream_t stream) {
    k_selftest_div<<<148 * 8, 256, 0, stream>>>(amax, bmax, mismatches);
    return cudaGetLastError();
}

cudaError_t launch_step_obs(const StepParams& p, bool reset_only, cudaStream_t stream) {
    const long long agents = 2ll * p.E;
    if (agents == 0) return cudaSuccess;
    // NOTE (measured, profiles/r01_notes.md): forcing the maximum shared-memory carveout halves throughput, so the driver's default is kept.
    if (reset_only) {
        const int blocks = (int)((agents + kThreads - 1) / kThreads);
        k_step_obs<true, 1><<<blocks, kThreads, (size_t)kThreads * kObs * sizeof(float) + (kThreads / 32) * sizeof(uint64_t), stream>>>(p);
    } else {
        constexpr int G = MM_K2_GROUPS;
        const size_t smem = (size_t)G * kThreads * kObs * sizeof(float) + (size_t)G * (kThreads / 32) * sizeof(uint64_t);
        static PerDeviceFlag configured;
        if (configured.first_time()) {
            cudaError_t e = cudaFuncSetAttribute(k_step_obs<false, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { configured.retract(); return e; }
        }
        const int blocks = (int)((agents + (long long)G * kThreads - 1) / ((long long)G * kThreads));
        k_step_obs<false, G><<<blocks, kThreads, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace mm
