// mm_pool.cu -- maze-pool side kernels: layout injection (pack + dir-to-exit field), state init, debug readback.
#include "mm_env.cuh"

namespace mm {

constexpr int kMaxSide = MM_MAX_SIDE;

// Tree walk from the exit over open cells; d2e(cell) = abs direction of the first step from cell towards the exit.
// This is the closed form of Agent.exit_route (maze.py:148-154, maze_agent.py:227-260): on a perfect maze the
// reference's route stack always equals the unique tree path to the exit, so its top element is d2e(cell).
// One thread per maze; `queue` is that maze's slice of the scratch buffer (smax*smax u16).
__device__ void finalize_one(const uint8_t* lay, int lay_stride, int W, int Hh, int ex, int ey,
                             ulonglong2* grid_out, int rows, ulonglong2* d2e_out, int smax, uint16_t* queue) {
    unsigned long long open_rows[kMaxSide], seen[kMaxSide], dlo[kMaxSide], dhi[kMaxSide];
    for (int y = 0; y < smax; y++) {
        unsigned long long o = 0;
        if (y < Hh) for (int x = 0; x < W; x++) if (lay[y * lay_stride + x] == 0) o |= 1ull << (x + kPad);
        open_rows[y] = o; seen[y] = 0; dlo[y] = 0; dhi[y] = 0;
    }
    for (int r = 0; r < rows; r++) {
        const int y = r - kPad;
        const unsigned long long o = (y >= 0 && y < smax) ? open_rows[y] : 0ull;
        grid_out[r] = make_ulonglong2(~o, 0ull);  // lo plane: 1 = wall everywhere that is not open; hi plane: no marks
    }
    int head = 0, tail = 0;
    queue[tail++] = (uint16_t)(ey * 64 + ex);
    seen[ey] |= 1ull << (ex + kPad);
    while (head < tail) {
        const int c = queue[head++], x = c & 63, y = c >> 6;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int nx = x + (k == 1) - (k == 3), ny = y + (k == 2) - (k == 0);
            if (nx < 0 || nx >= W || ny < 0 || ny >= Hh) continue;
            const unsigned long long bit = 1ull << (nx + kPad);
            if (!(open_rows[ny] & bit) || (seen[ny] & bit)) continue;
            seen[ny] |= bit;
            const int back = (k + 2) & 3;
            if (back & 1) dlo[ny] |= bit;
            if (back & 2) dhi[ny] |= bit;
            queue[tail++] = (uint16_t)(ny * 64 + nx);
        }
    }
    for (int y = 0; y < smax; y++) d2e_out[y] = make_ulonglong2(dlo[y], dhi[y]);
}

__global__ void k_load_layouts(ulonglong2* pool_grid, ulonglong2* pool_d2e, uint4* pool_hdr, int first, int n, int rows, int smax,
                               const uint8_t* layouts, const int32_t* hdr, uint16_t* scratch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* h = hdr + (size_t)i * 11;
    const int p = first + i;
    finalize_one(layouts + (size_t)i * smax * smax, smax, h[0], h[1], h[6], h[7], pool_grid + (size_t)p * rows, rows,
                 pool_d2e + (size_t)p * smax, smax, scratch + (size_t)i * smax * smax);
    pool_hdr[p] = make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16) | ((uint32_t)h[3] << 24),
                             (uint32_t)h[4] | ((uint32_t)h[5] << 8) | ((uint32_t)h[6] << 16) | ((uint32_t)h[7] << 24),
                             (uint32_t)h[8] | ((uint32_t)h[9] << 8) | ((uint32_t)h[10] << 16), (uint32_t)p);
}

__global__ void k_init_state(uint4* env_hdr, uint32_t* env_episode, uint4* agent_a, uint32_t* agent_b, int E) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= 2 * E) return;
    Agent a{};
    a.dir = 2; a.exit_len = -1;  // maze_agent.py:24-57
    agent_a[g] = pack_agent(a);
    agent_b[g] = 0;
    if ((g & 1) == 0) { env_hdr[g >> 1] = make_uint4(0, 0, 0, 0); env_episode[g >> 1] = 0; }
}

__global__ void k_unpack_agents(const uint4* agent_a, const uint32_t* agent_b, const uint4* env_hdr, const ulonglong2* pool_d2e,
                                int smax, int E, int32_t* out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= 2 * E) return;
    const Agent a = unpack_agent(agent_a[g], agent_b[g]);
    const uint4 H = env_hdr[g >> 1];
    int32_t* o = out + (size_t)g * MM_AGENT_FIELDS;
    o[0] = a.x; o[1] = a.y; o[2] = a.dir; o[3] = a.ke; o[4] = a.oke; o[5] = a.has; o[6] = a.team; o[7] = a.exit_len;
    o[8] = (int32_t)a.time; o[9] = a.olsx; o[10] = a.olsy; o[11] = a.mkv ? a.lmx : -1; o[12] = a.mkv ? a.lmy : -1;
    o[13] = a.minx; o[14] = a.maxx; o[15] = a.miny; o[16] = a.maxy;
    int len = -1;
    if (a.ke) {  // len(exit_route) == tree distance to the exit: follow the field
        const int ex = H.y & 0xff, ey = (H.y >> 8) & 0xff;
        const ulonglong2* d = pool_d2e + (size_t)H.w * smax;
        int x = a.x, y = a.y; len = 0;
        while ((x != ex || y != ey) && len < 4096) {
            const ulonglong2 r = d[y];
            const int k = (int)(((r.y >> (x + kPad)) & 1ull) << 1 | ((r.x >> (x + kPad)) & 1ull));
            x += (k == 1) - (k == 3); y += (k == 2) - (k == 0); len++;
        }
    }
    o[17] = len;
}

// Agent.reset(x, y) (maze_agent.py:59-79; time_from_last_seen is deliberately kept) or Agent.move(x, y, direction) (maze_agent.py:85-87) of ONE agent: the
// host-callable form of what K2's reset pass / step pass do for every agent.  One thread.
__global__ void k_agent_place(uint4* agent_a, uint32_t* agent_b, int g, int x, int y, int dir, int reset) {
    Agent a = unpack_agent(agent_a[g], agent_b[g]);
    a.x = x; a.y = y;
    if (reset) {
        a.olsx = x; a.olsy = y; a.minx = a.maxx = x; a.miny = a.maxy = y;
        a.dir = 2; a.mkv = 0; a.mem = 0; a.ke = a.oke = 0; a.exit_len = -1; a.has = a.team = 0; a.d2e = 0;
    } else {
        a.dir = dir & 3;
    }
    agent_a[g] = pack_agent(a);
}

__global__ void k_unpack_envs(const uint4* env_hdr, const uint32_t* env_episode, int E, int32_t* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint4 H = env_hdr[e];
    int32_t* o = out + (size_t)e * 8;
    const bool kp = H.z & 1u;
    o[0] = (int32_t)H.x; o[1] = kp ? (int32_t)((H.y >> 16) & 0xff) : -1; o[2] = kp ? (int32_t)(H.y >> 24) : -1; o[3] = (int32_t)H.w;
    o[4] = (H.z >> 8) & 0xff; o[5] = (H.z >> 16) & 0xff; o[6] = (H.z >> 1) & 1u; o[7] = (int32_t)env_episode[e];
}

__global__ void k_unpack_grid(const ulonglong2* grid, const ulonglong2* d2e, int smax, uint8_t* out_layout, uint8_t* out_d2e) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= smax * smax) return;
    const int x = i % smax, y = i / smax;
    const ulonglong2 r = grid[y + kPad];
    out_layout[i] = (uint8_t)(((r.y >> (x + kPad)) & 1ull) << 1 | ((r.x >> (x + kPad)) & 1ull));
    if (d2e != nullptr && out_d2e != nullptr) {
        const ulonglong2 q = d2e[y];
        out_d2e[i] = (uint8_t)(((q.y >> (x + kPad)) & 1ull) << 1 | ((q.x >> (x + kPad)) & 1ull));
    }
}

__global__ void k_unpack_pool_hdr(const uint4* pool_hdr, int p, int32_t* out) {
    const uint4 h = pool_hdr[p];
    out[0] = h.x & 0xff; out[1] = (h.x >> 8) & 0xff; out[2] = (h.x >> 16) & 0xff; out[3] = h.x >> 24;
    out[4] = h.y & 0xff; out[5] = (h.y >> 8) & 0xff; out[6] = (h.y >> 16) & 0xff; out[7] = h.y >> 24;
    out[8] = h.z & 0xff; out[9] = (h.z >> 8) & 0xff; out[10] = h.z >> 16;
}

}  // namespace mm
