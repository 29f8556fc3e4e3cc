// mm_env.cuh -- shared device-side definitions: packed state, Philox, launch parameters.
//
// Packed per-agent state (replaces the python fields of Agent, maze_agent.py:16-57):
//   agent_a.x : x[0:8) y[8:16) dir[16:18) knows_end[18] other_knows_end[19] has_key[20] team_has_key[21]
//               mark_valid[22] d2e_here[23:25) last_mask[26:32)
//   agent_a.y : last_mark x,y | other_last_seen x,y            (4 x u8)
//   agent_a.z : min_x, max_x, min_y, max_y                      (4 x u8)
//   agent_a.w : exit_len (i16) | memory 4 x 3 bit (move+1, oldest in the low bits) << 16
//   agent_b   : time_from_last_seen (u32; the reference never resets it, maze_agent.py:59-79 vs :195)
// `exit_route` is not stored: on a tree maze it always equals the tree path to the exit, so only the
// static dir-to-exit field of the maze and the agent's `knows_end` bit are needed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/marl_maze_b200.h"

namespace mm {

// Per-device one-time launch configuration (cudaFuncSetAttribute applies to the CURRENT device's context, so a process-wide flag would
// leave a second GPU unconfigured).  `true` the first time it is asked for the current device.
constexpr int kMaxDevices = 64;
struct PerDeviceFlag {
    bool done[kMaxDevices] = {};
    bool first_time() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
    void retract() { int d = 0; if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < kMaxDevices) done[d] = false; }
};
inline int current_device_slot() { int d = 0; return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < kMaxDevices) ? d : 0; }

constexpr int kPad = MM_PAD;
constexpr int kObs = MM_OBS_DIM;
constexpr unsigned kFull = 0xffffffffu;

struct StepParams {
    const ulonglong2* __restrict__ pool_grid;
    const ulonglong2* __restrict__ pool_d2e;
    const uint4* __restrict__ pool_hdr;
    ulonglong2* env_grid;
    uint4* env_hdr;
    uint32_t* env_episode;
    uint4* agent_a;
    uint32_t* agent_b;
    const uint8_t* __restrict__ actions;   // [E][2][2] or nullptr (kernel samples uniform legal actions)
    uint8_t* actions_out;                  // may be nullptr
    const uint8_t* __restrict__ reset_mask; // reset launch only; may be nullptr (= all)
    float* obs;                            // [E][2][65]
    uint8_t* masks;                        // [E][2][6]
    float* reward;                         // [E]
    uint8_t* done;                         // [E]
    int E, P, rows, smax, max_t, auto_reset, env_offset;
    int obs_vec4;                          // obs pointer is 16-byte aligned: whole-warp float4 copy-out allowed
    float inv_max_t;                       // correctly rounded 1/max_t (host: 1.0f / (float)max_t)
    uint64_t action_seed;
    // Agent(..., vision_range=r) per agent, 1 <= r <= 4 (maze_agent.py:16) and the ray features' values: the reference accumulates
    // 1/r per marked cell and writes 1 - j * (1/r) for a dead end at distance j, in float64 (maze_agent.py:148,180,264,267); the host builds the
    // float32 casts of exactly those sums / differences: de_tab[a][c] for dead-end code c (0 = none or at distance r, r = wall adjacent, else r - j),
    // mk_tab[a][c] for c marked cells along a ray.  r = 4 gives 0, .25, .5, .75, 1 in both.
    int vr[2];
    float de_tab[2][5], mk_tab[2][5];
};

struct Agent {
    int x, y, dir;
    uint32_t ke, oke, has, team, mkv;   // 0/1 flags
    int d2e;                            // abs direction of the first step towards the exit from (x,y)
    uint32_t last_mask;                 // 6 bits F,R,B,L,stop,mark as last emitted
    int lmx, lmy, olsx, olsy;
    int minx, maxx, miny, maxy;
    int exit_len;
    uint32_t mem;                       // 4 x 3 bits, value = move+1 (0 = empty), oldest in bits 0..2
    uint32_t time;                      // time_from_last_seen
};

__device__ __forceinline__ Agent unpack_agent(uint4 A, uint32_t B) {
    Agent g;
    g.x = A.x & 0xff; g.y = (A.x >> 8) & 0xff; g.dir = (A.x >> 16) & 3;
    g.ke = (A.x >> 18) & 1; g.oke = (A.x >> 19) & 1; g.has = (A.x >> 20) & 1; g.team = (A.x >> 21) & 1; g.mkv = (A.x >> 22) & 1;
    g.d2e = (A.x >> 23) & 3; g.last_mask = (A.x >> 26) & 0x3f;
    g.lmx = A.y & 0xff; g.lmy = (A.y >> 8) & 0xff; g.olsx = (A.y >> 16) & 0xff; g.olsy = A.y >> 24;
    g.minx = A.z & 0xff; g.maxx = (A.z >> 8) & 0xff; g.miny = (A.z >> 16) & 0xff; g.maxy = A.z >> 24;
    g.exit_len = (int)(int16_t)(A.w & 0xffff); g.mem = (A.w >> 16) & 0xfff;
    g.time = B;
    return g;
}
__device__ __forceinline__ uint4 pack_agent(const Agent& g) {
    uint4 A;
    A.x = (uint32_t)g.x | ((uint32_t)g.y << 8) | ((uint32_t)g.dir << 16) | (g.ke << 18) | (g.oke << 19) | (g.has << 20) | (g.team << 21) |
          (g.mkv << 22) | ((uint32_t)g.d2e << 23) | (g.last_mask << 26);
    A.y = (uint32_t)g.lmx | ((uint32_t)g.lmy << 8) | ((uint32_t)g.olsx << 16) | ((uint32_t)g.olsy << 24);
    A.z = (uint32_t)g.minx | ((uint32_t)g.maxx << 8) | ((uint32_t)g.miny << 16) | ((uint32_t)g.maxy << 24);
    A.w = ((uint32_t)g.exit_len & 0xffffu) | (g.mem << 16);
    return A;
}

// Philox4x32-10 (Salmon et al. 2011); identical to oracle/maze_oracle.c philox4x32_10.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// x -> nearest TF32 value (10 explicit mantissa bits), returned as fp32.  The 3xTF32 trunk carries every operand as
// hi = tf32(x), lo = tf32(x - hi): both terms are exact TF32 numbers, so the tensor core's own operand truncation is a no-op and
// the representation error |x - hi - lo| <= 2^-22 |x| is zero-mean (plain truncation would bias every post-ReLU activation down).
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

}  // namespace mm
