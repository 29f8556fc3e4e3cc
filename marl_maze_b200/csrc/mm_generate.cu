// mm_generate.cu -- K1: batched maze generator, one thread per maze, bit-packed rows, counter-based RNG.
//
// Replaces Maze.build_maze / get_neighbors / set_start / set_end / set_key / get_shortest_path (maze.py:170-273):
//   - all-walls grid, start on even coordinates (random when rand_start, else top middle);
//   - iterative DFS: at the stack top, if unvisited rooms exist two cells away AND random() > corridor_const,
//     carve to a random one and grow corridor_const by 1/(10*max(W,H)); otherwise pop and reset it (maze.py:180-201);
//   - `difficulty` candidate exits on a coin-flipped left/right edge; the longest start->exit path wins, ties to the
//     last drawn (dict overwrite, maze.py:203-217);
//   - key by rejection sampling: open, not start, not exit, not on the start->exit path (maze.py:252-259).
// The reference draws from Python's Mersenne Twister; here each maze owns a Philox4x32-10 stream keyed by
// (seed, maze id), consumed one 32-bit word per draw in exactly the order above.  oracle/maze_oracle.c runs the
// same algorithm under the same stream (and under MT against the reference), which is what tests compare with.
// Output goes straight to the pool in its final HBM form: bit-plane grid with wall border, dir-to-exit field, header.
#include "mm_env.cuh"

namespace mm {

constexpr int kMaxSide = MM_MAX_SIDE;

struct PhiloxStream {
    uint32_t k0, k1, ctr, buf[4];
    int have;
    __device__ void seed(uint64_t s, uint32_t id) { k0 = (uint32_t)s ^ id; k1 = (uint32_t)(s >> 32) + 0x632BE5ABu; ctr = 0; have = 0; }
    __device__ uint32_t next() {
        if (!have) { philox4x32_10(ctr++, 0, 0, 0, k0, k1, buf); have = 4; }
        const int i = 4 - have--;
        return i == 0 ? buf[0] : i == 1 ? buf[1] : i == 2 ? buf[2] : buf[3];
    }
    __device__ uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
    __device__ int randint(int a, int b) { return a + (int)below((uint32_t)(b - a + 1)); }
};

__global__ void __launch_bounds__(64) k_generate(ulonglong2* pool_grid, ulonglong2* pool_d2e, uint4* pool_hdr, int first, int n, int rows, int smax,
                                                int side_lo, int side_hi, int rand_start, int difficulty, uint64_t seed, uint32_t id_base, int id_mod, int id_mul,
                                                uint16_t* scratch, const uint8_t* __restrict__ only, int height_cells) {
    // grid-stride over mazes: a full grid for an inline build, a few blocks per SM for a background build that trickles along beside
    // other kernels (each thread's serial carve holds its residency slot for milliseconds)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (only && !only[i]) continue;   // incremental refill: this slot's maze has not been consumed, it stays
    const int p = first + i;
    uint16_t* q = scratch + (size_t)i * smax * smax;  // DFS stack, then BFS queue
    const uint32_t maze_id = id_mod ? id_base + (uint32_t)(i % id_mod) * (uint32_t)id_mul + (uint32_t)(i / id_mod) : id_base + (uint32_t)i;
    PhiloxStream rng; rng.seed(seed, maze_id);

    unsigned long long open_rows[kMaxSide], seen[kMaxSide], dlo[kMaxSide], dhi[kMaxSide];
    for (int y = 0; y < smax; y++) { open_rows[y] = 0; seen[y] = 0; dlo[y] = 0; dhi[y] = 0; }
    auto is_open = [&](int x, int y) { return (open_rows[y] >> (x + kPad)) & 1ull; };

    // rand_sizes (maze.py:171-174): one draw, square.  height_cells > 0: Maze(default_size=[w, h]) with rand_sizes False (maze.py:26-27) -- every maze
    // is 2w-1 wide and 2h-1 high and no size is drawn
    const int W = height_cells > 0 ? side_lo * 2 - 1 : rng.randint(side_lo, side_hi) * 2 - 1;  // maze.py:172
    const int Hh = height_cells > 0 ? height_cells * 2 - 1 : W;
    const int S = max(W, Hh);
    int sx, sy;
    if (rand_start) { sx = rng.randint(0, (W - 1) / 2) * 2; sy = rng.randint(0, (Hh - 1) / 2) * 2; }  // maze.py:231-234
    else { sx = ((W / 2) % 2 == 0) ? W / 2 : W / 2 - 1; sy = 0; }

    // ---- carve (maze.py:180-201)
    int sp = 0;
    q[sp++] = (uint16_t)(sy * 64 + sx);
    float cc = 0.f;
    const float inc = __fdiv_rn(1.0f, (float)(10 * S));
    while (sp) {
        const int c = q[sp - 1], x = c & 63, y = c >> 6;
        open_rows[y] |= 1ull << (x + kPad);
        int nb[4], nn = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int nx = x + 2 * ((k == 1) - (k == 3)), ny = y + 2 * ((k == 2) - (k == 0));
            if (nx >= 0 && nx < W && ny >= 0 && ny < Hh && !is_open(nx, ny)) nb[nn++] = ny * 64 + nx;
        }
        bool go = false;
        // (0,1]: with a 24-bit draw an exact 0 would (once in 2^24) lose against corridor_const == 0 and stop the carve at a
        // single cell; Python's 53-bit random() never does in practice
        if (nn) go = __fmul_rn((float)((rng.next() >> 8) + 1u), 1.0f / 16777216.0f) > cc;
        if (go) {
            const uint32_t pick = rng.below((uint32_t)nn);
            const int nc = pick == 0 ? nb[0] : pick == 1 ? nb[1] : pick == 2 ? nb[2] : nb[3];
            const int x2 = nc & 63, y2 = nc >> 6;
            open_rows[(y + y2) >> 1] |= 1ull << (((x + x2) >> 1) + kPad);
            q[sp++] = (uint16_t)nc;
            cc = __fadd_rn(cc, inc);
        } else { sp--; cc = 0.f; }
    }

    // ---- candidate exits (set_end, maze.py:239-250), all draws first: get_shortest_path consumes no randomness
    int cex[8], cey[8], clen[8];
    const int nd = difficulty < 8 ? difficulty : 8;
    for (int d = 0; d < nd; d++) {
        const int coin = rng.randint(0, 1);
        const int x = coin == 0 ? 0 : W - 1;
        bool found = false;
        for (int tries = 0; tries < 4096 && !found; tries++) {
            const int y = rng.randint(0, Hh - 1);
            if (x == sx && y == sy) continue;
            if (is_open(x, y)) { cex[d] = x; cey[d] = y; found = true; }
        }
        if (!found) {  // the reference spins forever here (an edge without an eligible cell: its generator CAN leave a partial maze).
            cex[d] = sx; cey[d] = sy;  // Way out shared with the oracle: first open cell != start in row-major order.
            for (int yy = 0; yy < Hh && !found; yy++)
                for (int xx = 0; xx < W && !found; xx++)
                    if (is_open(xx, yy) && !(xx == sx && yy == sy)) { cex[d] = xx; cey[d] = yy; found = true; }
        }
        clen[d] = 0;
    }
    int best = 0;
    if (nd > 1) {  // BFS from the start; path length (in cells) of each candidate = depth + 1
        for (int y = 0; y < smax; y++) seen[y] = 0;
        int head = 0, tail = 0, depth = 1, level_end = 1;
        q[tail++] = (uint16_t)(sy * 64 + sx); seen[sy] |= 1ull << (sx + kPad);
        while (head < tail) {
            if (head == level_end) { depth++; level_end = tail; }
            const int c = q[head++], x = c & 63, y = c >> 6;
            for (int d = 0; d < nd; d++) if (cex[d] == x && cey[d] == y) clen[d] = depth;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int nx = x + (k == 1) - (k == 3), ny = y + (k == 2) - (k == 0);
                if (nx < 0 || nx >= W || ny < 0 || ny >= Hh) continue;
                const unsigned long long bit = 1ull << (nx + kPad);
                if (!(open_rows[ny] & bit) || (seen[ny] & bit)) continue;
                seen[ny] |= bit; q[tail++] = (uint16_t)(ny * 64 + nx);
            }
        }
        for (int d = 1; d < nd; d++) if (clen[d] >= clen[best]) best = d;
    }
    const int ex = cex[best], ey = cey[best];

    // ---- dir-to-exit field: tree walk from the exit
    for (int y = 0; y < smax; y++) seen[y] = 0;
    {
        int head = 0, tail = 0;
        q[tail++] = (uint16_t)(ey * 64 + ex); seen[ey] |= 1ull << (ex + kPad);
        while (head < tail) {
            const int c = q[head++], x = c & 63, y = c >> 6;
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
                q[tail++] = (uint16_t)(ny * 64 + nx);
            }
        }
    }
    // ---- start->exit path: follow the field; `seen` is reused as the on-path bitmap for set_key
    for (int y = 0; y < smax; y++) seen[y] = 0;
    int spl = 1, p1x = sx, p1y = sy;
    {
        int x = sx, y = sy;
        seen[y] |= 1ull << (x + kPad);
        while ((x != ex || y != ey) && spl < 4096) {
            const int k = (int)(((dhi[y] >> (x + kPad)) & 1ull) << 1 | ((dlo[y] >> (x + kPad)) & 1ull));
            x += (k == 1) - (k == 3); y += (k == 2) - (k == 0);
            if (x < 0 || x >= W || y < 0 || y >= Hh) { x = sx; y = sy; break; }  // unreachable on a connected maze; never walk out of the arrays
            seen[y] |= 1ull << (x + kPad);
            if (spl == 1) { p1x = x; p1y = y; }
            spl++;
        }
    }
    // ---- key (set_key, maze.py:252-259)
    int kx = sx, ky = sy;
    bool kfound = false;
    for (int tries = 0; tries < 65536 && !kfound; tries++) {
        const int tx = rng.randint(0, W - 1), ty = rng.randint(0, Hh - 1);
        if (!is_open(tx, ty) || (tx == ex && ty == ey) || (tx == sx && ty == sy) || ((seen[ty] >> (tx + kPad)) & 1ull)) continue;
        kx = tx; ky = ty; kfound = true;
    }
    for (int pass = 0; pass < 2 && !kfound; pass++)  // reference: infinite loop; same way out as the oracle
        for (int yy = 0; yy < Hh && !kfound; yy++)
            for (int xx = 0; xx < W && !kfound; xx++)
                if (is_open(xx, yy) && !(xx == sx && yy == sy) && !(xx == ex && yy == ey) && (pass == 1 || !((seen[yy] >> (xx + kPad)) & 1ull))) { kx = xx; ky = yy; kfound = true; }

    // ---- pool entry in its final HBM form
    ulonglong2* g = pool_grid + (size_t)p * rows;
    for (int r = 0; r < rows; r++) {
        const int y = r - kPad;
        const unsigned long long o = (y >= 0 && y < smax) ? open_rows[y] : 0ull;
        g[r] = make_ulonglong2(~o, 0ull);
    }
    ulonglong2* dd = pool_d2e + (size_t)p * smax;
    for (int y = 0; y < smax; y++) dd[y] = make_ulonglong2(dlo[y], dhi[y]);
    pool_hdr[p] = make_uint4((uint32_t)W | ((uint32_t)Hh << 8) | ((uint32_t)sx << 16) | ((uint32_t)sy << 24),
                             (uint32_t)p1x | ((uint32_t)p1y << 8) | ((uint32_t)ex << 16) | ((uint32_t)ey << 24),
                             (uint32_t)kx | ((uint32_t)ky << 8) | ((uint32_t)spl << 16), maze_id);
    }
}

cudaError_t launch_generate(const mm_state* st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty, uint64_t seed,
                            uint32_t id_base, int id_mod, int id_mul, void* scratch, int max_blocks, const uint8_t* only, int height_cells, cudaStream_t stream) {
    int blocks = (n + 63) / 64;
    if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
    k_generate<<<blocks, 64, 0, stream>>>((ulonglong2*)st->pool_grid, (ulonglong2*)st->pool_d2e, (uint4*)st->pool_hdr, first, n,
                                                 st->smax + 2 * MM_PAD, st->smax, side_lo, side_hi, rand_start, difficulty, seed, id_base, id_mod, id_mul, (uint16_t*)scratch, only, height_cells);
    return cudaGetLastError();
}

}  // namespace mm
