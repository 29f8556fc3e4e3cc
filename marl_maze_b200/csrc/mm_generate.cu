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
//
// Second generation (the first -- grid planes in local memory, DFS stack / BFS queue in a global scratch, 229 ms per 1 Mi mazes of side 49: every
// step of the serial carve waited on an L2 round trip -- is kept under -DMM_K1_V1 for A/B runs).  The carve is inherently serial per maze, so
// the lever is the latency of each dependent step: here ALL of a maze's working state lives in shared memory, [row][lane] interleaved (a lane's rows are
// 256 bytes apart: conflict-free whatever row each lane is on):
//   open plane (+2 guard rows / 2 guard bits each side that read "visited", so the carve needs no bounds tests), direction lo / hi planes, on-path
//   plane.  There is no DFS stack: every cell the carve opens records the direction back to where it was opened from, and popping a room is walking
//   two cells along that pointer.
// The two breadth-first passes of the first generation (candidate-exit path lengths; dir-to-exit field) are gone: a perfect maze is a tree, the carve
// records every new cell's direction back to its parent (the tree rooted at the start) as it opens it, a candidate's path length is then a climb to
// the start, and the dir-to-exit field IS that parent field except on the start -> exit path, whose pointers one climb from the exit turns around.
// (A stackless tree walk from the exit was tried first: 55 % of the kernel's instructions at 11 of 32 lanes active.)  One warp per block (no block-level
// synchronisation anywhere); in an incremental refill the consumed slots are compacted warp-wide first, so every lane of a pass carries a maze.
#include "mm_env.cuh"

namespace mm {

constexpr int kMaxSide = MM_MAX_SIDE;
static_assert(kMaxSide + 2 * kPad <= 64, "a maze row with its wall border is one 64-bit word");

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

// The same stream with an 8-word ring in shared memory ([word][lane]) instead of a 4-word register buffer: a lane may hold two Philox blocks, so the
// lanes of a warp can refill TOGETHER (k1 carve loop) although they consume at different rates -- the ten rounds then run once per ~4 carve steps of
// the warp instead of in nearly every step for whichever lane happened to run dry.  Word order and values are those of PhiloxStream.
struct PhiloxRing {
    uint32_t k0, k1, ctr, rd, wr;
    uint32_t* ring;   // this lane's column: ring[32 * j], j = 0..7
    __device__ void seed(uint64_t s, uint32_t id, uint32_t* r) { k0 = (uint32_t)s ^ id; k1 = (uint32_t)(s >> 32) + 0x632BE5ABu; ctr = 0; rd = wr = 0; ring = r; }
    __device__ int have() const { return (int)(wr - rd); }
    __device__ void refill() {   // have() <= 4
        uint32_t b[4];
        philox4x32_10(ctr++, 0, 0, 0, k0, k1, b);
#pragma unroll
        for (int j = 0; j < 4; j++) ring[32 * ((wr + j) & 7u)] = b[j];
        wr += 4;
    }
    __device__ uint32_t next() {
        if (rd == wr) refill();
        return ring[32 * (rd++ & 7u)];
    }
    __device__ uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
    __device__ int randint(int a, int b) { return a + (int)below((uint32_t)(b - a + 1)); }
};

#ifdef MM_K1_V1
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
    int best = 0, best_len = 0;
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

#else
constexpr int K1_T = 32;   // one warp per block

struct K1Args {
    ulonglong2* pool_grid; ulonglong2* pool_d2e; uint4* pool_hdr;
    int first, n, rows, smax, side_lo, side_hi, rand_start, difficulty;
    uint64_t seed; uint32_t id_base; int id_mod, id_mul;
    const uint8_t* only; int height_cells, slots_per_warp;
};

__device__ __forceinline__ int k1_dx(int k) { return (int)((0x19u >> (2 * k)) & 3u) - 1; }   // N, E, S, W -> 0, 1, 0, -1
__device__ __forceinline__ int k1_dy(int k) { return (int)((0x64u >> (2 * k)) & 3u) - 1; }   //             -> -1, 0, 1, 0

// One maze, one thread.  s = this block's planes, element (row r of the plane at row offset off) = s[(off + r) * K1_T + lane].
// Called by the WHOLE warp (the carve loop votes); lanes without a maze pass i < 0.
__device__ __forceinline__ void k1_generate_one(const K1Args& A, unsigned long long* __restrict__ s, uint32_t* __restrict__ ring, const int lane, const int i_in) {
    const bool live = i_in >= 0;
    const int i = live ? i_in : 0;
    const int smax = A.smax;
    const int oO = 2, oL = smax + 4, oH = 2 * smax + 4, oP = 3 * smax + 4;   // open (rows -2 .. smax+1), dir lo, dir hi, on-path
#define K1_AT(off, r) s[((off) + (r)) * K1_T + lane]
    const int p = A.first + i;
    const uint32_t maze_id = A.id_mod ? A.id_base + (uint32_t)(i % A.id_mod) * (uint32_t)A.id_mul + (uint32_t)(i / A.id_mod) : A.id_base + (uint32_t)i;
    PhiloxRing rng; rng.seed(A.seed, maze_id, ring + lane);

    // rand_sizes (maze.py:171-174): one draw, square.  height_cells > 0: Maze(default_size=[w, h]) with rand_sizes False (maze.py:26-27) -- every maze
    // is 2w-1 wide and 2h-1 high and no size is drawn
    const int W = A.height_cells > 0 ? A.side_lo * 2 - 1 : rng.randint(A.side_lo, A.side_hi) * 2 - 1;  // maze.py:172
    const int Hh = A.height_cells > 0 ? A.height_cells * 2 - 1 : W;
    const int S = max(W, Hh);
    int sx, sy;
    if (A.rand_start) { sx = rng.randint(0, (W - 1) / 2) * 2; sy = rng.randint(0, (Hh - 1) / 2) * 2; }  // maze.py:231-234
    else { sx = ((W / 2) % 2 == 0) ? W / 2 : W / 2 - 1; sy = 0; }

    // guards: two bits left and right of the W columns in every maze row, two rows above and below read as "already visited" during the carve
    const unsigned long long guard = (3ull << (kPad - 2)) | (3ull << (W + kPad));
    for (int r = -2; r < smax + 2; r++) K1_AT(oO, r) = (r >= 0 && r < Hh) ? guard : (r >= -2 && r < Hh + 2) ? ~0ull : 0ull;
    for (int r = 0; r < smax; r++) { K1_AT(oL, r) = 0ull; K1_AT(oH, r) = 0ull; K1_AT(oP, r) = 0ull; }

    // ---- carve (maze.py:180-201)
    {
        int x = sx, y = sy;
        float cc = 0.f;
        const float inc = __fdiv_rn(1.0f, (float)(10 * S));
        K1_AT(oO, y) |= 1ull << (x + kPad);
        bool done = !live;
        for (int it = 0; it < 2 * 64 * 64; it++) {   // every room is entered once and left once; bounded so that a bug cannot hang the GPU
            // warp-synchronous refill: a step draws at most twice; when any lane could run dry, every lane with room for a block takes one
            if (__any_sync(kFull, !done && rng.have() < 2)) { if (!done && rng.have() <= 4) rng.refill(); }
            if (__all_sync(kFull, done)) break;
            if (done) continue;
            const unsigned long long rN = K1_AT(oO, y - 2), rC = K1_AT(oO, y), rS = K1_AT(oO, y + 2);
            const uint32_t av = (uint32_t)((~rN >> (x + kPad)) & 1ull) | (uint32_t)((~rC >> (x + 2 + kPad)) & 1ull) << 1 |
                                (uint32_t)((~rS >> (x + kPad)) & 1ull) << 2 | (uint32_t)((~rC >> (x - 2 + kPad)) & 1ull) << 3;   // unvisited rooms N, E, S, W
            const int nn = __popc(av);
            bool go = false;
            // (0,1]: with a 24-bit draw an exact 0 would (once in 2^24) lose against corridor_const == 0 and stop the carve at a
            // single cell; Python's 53-bit random() never does in practice
            if (nn) go = __fmul_rn((float)((rng.next() >> 8) + 1u), 1.0f / 16777216.0f) > cc;
            if (go) {
                const uint32_t pick = rng.below((uint32_t)nn);
                uint32_t m = av;
                for (uint32_t j = 0; j < pick; j++) m &= m - 1;
                const int k = __ffs(m) - 1, dx = k1_dx(k), dy = k1_dy(k);
                // both new cells (the opened wall, the room) point BACK along the move: the parent pointers of the tree rooted at the start.  They are
                // also the DFS stack: popping a room is walking two cells along its pointer (the reference's list of coordinates, maze.py:181-201)
                const int back = (k + 2) & 3;
                const unsigned long long bw = 1ull << (x + dx + kPad), br = 1ull << (x + 2 * dx + kPad);
                K1_AT(oO, y + dy) |= bw;
                if (back & 1) K1_AT(oL, y + dy) |= bw;
                if (back & 2) K1_AT(oH, y + dy) |= bw;
                x += 2 * dx; y += 2 * dy;
                K1_AT(oO, y) |= br;
                if (back & 1) K1_AT(oL, y) |= br;
                if (back & 2) K1_AT(oH, y) |= br;
                cc = __fadd_rn(cc, inc);
            } else {
                if (x == sx && y == sy) { done = true; continue; }   // the start popped: the stack is empty
                const int k = (int)(((K1_AT(oH, y) >> (x + kPad)) & 1ull) << 1 | ((K1_AT(oL, y) >> (x + kPad)) & 1ull));
                x += 2 * k1_dx(k); y += 2 * k1_dy(k);
                cc = 0.f;
            }
        }
    }
    if (!live) return;
    // the guards have done their work: from here on a bit outside the maze reads "wall"
    for (int r = -2; r < smax + 2; r++) { const unsigned long long v = K1_AT(oO, r); K1_AT(oO, r) = (r >= 0 && r < Hh) ? (v & ~guard) : 0ull; }
    auto is_open = [&](int x, int y) { return (K1_AT(oO, y) >> (x + kPad)) & 1ull; };   // any (x, y) within one cell of the maze

    // ---- candidate exits (set_end, maze.py:239-250), all draws first: get_shortest_path consumes no randomness
    int cex[8], cey[8], clen[8];
    const int nd = A.difficulty < 8 ? A.difficulty : 8;
#pragma unroll
    for (int d = 0; d < 8; d++) {
        cex[d] = sx; cey[d] = sy; clen[d] = 0;
        if (d < nd) {
            const int coin = rng.randint(0, 1);
            const int x = coin == 0 ? 0 : W - 1;
            bool found = false;
            for (int tries = 0; tries < 4096 && !found; tries++) {
                const int y = rng.randint(0, Hh - 1);
                if (x == sx && y == sy) continue;
                if (is_open(x, y)) { cex[d] = x; cey[d] = y; found = true; }
            }
            if (!found) {  // the reference spins forever here (an edge without an eligible cell: its generator CAN leave a partial maze).
                // Way out shared with the oracle: first open cell != start in row-major order.
                for (int yy = 0; yy < Hh && !found; yy++)
                    for (int xx = 0; xx < W && !found; xx++)
                        if (is_open(xx, yy) && !(xx == sx && yy == sy)) { cex[d] = xx; cey[d] = yy; found = true; }
            }
        }
    }
    // The carve left, in the lo / hi planes, every open cell's direction to its PARENT in the tree rooted at the start.  A perfect maze is a tree, so
    //   - a candidate exit's path length is its depth: climb the parent pointers to the start (no search);
    //   - the dir-to-exit field differs from the parent pointers only ON the start -> exit path (a cell off the path starts its way to the exit by
    //     going up to its parent; a cell on it goes down the path): climb once from the exit and turn the pointers of that path around.
    auto parent_dir = [&](int x, int y) { return (int)(((K1_AT(oH, y) >> (x + kPad)) & 1ull) << 1 | ((K1_AT(oL, y) >> (x + kPad)) & 1ull)); };
    int best = 0, best_len = 0;
    if (nd > 1) {
#pragma unroll
        for (int d = 0; d < 8; d++) {
            if (d < nd) {
                int x = cex[d], y = cey[d], len = 1;
                while ((x != sx || y != sy) && len < 4096) { const int k = parent_dir(x, y); x += k1_dx(k); y += k1_dy(k); len++; }
                clen[d] = len;
            }
        }
        best_len = clen[0];
#pragma unroll
        for (int d = 1; d < 8; d++) if (d < nd && clen[d] >= best_len) { best = d; best_len = clen[d]; }
    }
    int ex = cex[0], ey = cey[0];
#pragma unroll
    for (int d = 1; d < 8; d++) if (d == best) { ex = cex[d]; ey = cey[d]; }

    // ---- start -> exit path (climbed from the exit): on-path plane for set_key, its length, its second cell; pointers of the path turned towards the exit
    int spl = 1, p1x = sx, p1y = sy;
    {
        int x = ex, y = ey;
        K1_AT(oP, y) |= 1ull << (x + kPad);
        int k = parent_dir(x, y);                     // read before the cell's entry is rewritten
        { const unsigned long long bit = 1ull << (x + kPad); K1_AT(oL, y) &= ~bit; K1_AT(oH, y) &= ~bit; }   // the exit's own entry is 0
        while ((x != sx || y != sy) && spl < 4096) {
            p1x = x; p1y = y;                         // ends as the cell the path enters right after the start
            x += k1_dx(k); y += k1_dy(k);             // the parent
            const int down = (k + 2) & 3;             // from the parent back to the cell we came from: its direction to the exit
            const unsigned long long bit = 1ull << (x + kPad);
            k = parent_dir(x, y);
            K1_AT(oL, y) = (K1_AT(oL, y) & ~bit) | ((down & 1) ? bit : 0ull);
            K1_AT(oH, y) = (K1_AT(oH, y) & ~bit) | ((down & 2) ? bit : 0ull);
            K1_AT(oP, y) |= bit;
            spl++;
        }
    }
    // ---- key (set_key, maze.py:252-259)
    int kx = sx, ky = sy;
    bool kfound = false;
    for (int tries = 0; tries < 65536 && !kfound; tries++) {
        const int tx = rng.randint(0, W - 1), ty = rng.randint(0, Hh - 1);
        if (!is_open(tx, ty) || (tx == ex && ty == ey) || (tx == sx && ty == sy) || ((K1_AT(oP, ty) >> (tx + kPad)) & 1ull)) continue;
        kx = tx; ky = ty; kfound = true;
    }
    for (int pass = 0; pass < 2 && !kfound; pass++)  // reference: infinite loop; same way out as the oracle
        for (int yy = 0; yy < Hh && !kfound; yy++)
            for (int xx = 0; xx < W && !kfound; xx++)
                if (is_open(xx, yy) && !(xx == sx && yy == sy) && !(xx == ex && yy == ey) && (pass == 1 || !((K1_AT(oP, yy) >> (xx + kPad)) & 1ull))) { kx = xx; ky = yy; kfound = true; }

    // ---- pool entry in its final HBM form
    ulonglong2* g = A.pool_grid + (size_t)p * A.rows;
    for (int r = 0; r < A.rows; r++) {
        const int y = r - kPad;
        const unsigned long long o = (y >= 0 && y < smax) ? K1_AT(oO, y) : 0ull;
        g[r] = make_ulonglong2(~o, 0ull);
    }
    ulonglong2* dd = A.pool_d2e + (size_t)p * smax;
    for (int y = 0; y < smax; y++) dd[y] = make_ulonglong2(K1_AT(oL, y), K1_AT(oH, y));
    A.pool_hdr[p] = make_uint4((uint32_t)W | ((uint32_t)Hh << 8) | ((uint32_t)sx << 16) | ((uint32_t)sy << 24),
                               (uint32_t)p1x | ((uint32_t)p1y << 8) | ((uint32_t)ex << 16) | ((uint32_t)ey << 24),
                               (uint32_t)kx | ((uint32_t)ky << 8) | ((uint32_t)spl << 16), maze_id);
#undef K1_AT
}

__host__ __device__ inline int k1_words(int smax) { return 4 * smax + 4; }   // 64-bit words per maze: four planes, two guard rows above and below the open plane

__global__ void __launch_bounds__(K1_T) k_generate(const K1Args A) {
    extern __shared__ __align__(16) unsigned long long k1_smem[];
    __shared__ int pend[64];
    __shared__ uint32_t ring[8 * K1_T];   // PhiloxRing: [word][lane]
    const int lane = threadIdx.x;
    const int lo = blockIdx.x * A.slots_per_warp, hi = min(A.n, lo + A.slots_per_warp);
    int base = lo, npend = 0;
    // this warp's slots, the consumed ones compacted (incremental refill: typically a sixth of them) so that every lane of a pass carries a maze
    while (base < hi || npend > 0) {
        while (npend < 32 && base < hi) {
            const int i = base + lane;
            const bool take = i < hi && (A.only == nullptr || A.only[i] != 0);
            const uint32_t m = __ballot_sync(kFull, take);
            if (take) pend[npend + __popc(m & ((1u << lane) - 1u))] = i;
            npend += __popc(m); base += 32;
        }
        __syncwarp();
        const int cnt = min(npend, 32), rem = npend - cnt;
        const int mine = lane < cnt ? pend[lane] : -1;
        const int carry = lane < rem ? pend[cnt + lane] : 0;
        __syncwarp();
        if (lane < rem) pend[lane] = carry;
        npend = rem;
        __syncwarp();
        k1_generate_one(A, k1_smem, ring, lane, mine);
        __syncwarp();
    }
}

cudaError_t launch_generate(const mm_state* st, int first, int n, int side_lo, int side_hi, int rand_start, int difficulty, uint64_t seed,
                            uint32_t id_base, int id_mod, int id_mul, void* scratch, int max_blocks, const uint8_t* only, int height_cells, cudaStream_t stream) {
    (void)scratch;   // the first generation kept its DFS stack / BFS queue there
    if (n <= 0) return cudaSuccess;
    const int smem = k1_words(st->smax) * 8 * K1_T;
    static PerDeviceFlag configured;
    static int smem_set[kMaxDevices] = {};
    const int dslot = current_device_slot();
    if (configured.first_time() || smem_set[dslot] < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_generate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem > 48 * 1024 ? smem : 48 * 1024);
        // every resident warp brings 28 (side 25) .. 62 KB (side 54) of planes and the kernel has no use for L1: ask for the whole array as shared memory,
        // otherwise the driver's default carve-out decides how many warps an SM carves with
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_generate, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { configured.retract(); return e; }
        smem_set[dslot] = smem;
    }
    // a masked (incremental) build gives every warp a long run of slots to compact; a full build one pass per warp
    int spw = only ? 512 : 32;
    int blocks = (n + spw - 1) / spw;
    const int cap = max_blocks > 0 ? 2 * max_blocks : 0;   // max_blocks counts 64-maze blocks (the first generation's block size)
    if (cap > 0 && blocks > cap) { spw = ((n + cap - 1) / cap + 31) / 32 * 32; blocks = (n + spw - 1) / spw; }
    K1Args A{(ulonglong2*)st->pool_grid, (ulonglong2*)st->pool_d2e, (uint4*)st->pool_hdr, first, n, st->smax + 2 * MM_PAD, st->smax, side_lo, side_hi, rand_start,
             difficulty, seed, id_base, id_mod, id_mul, only, height_cells, spw};
    k_generate<<<blocks, K1_T, smem, stream>>>(A);
    return cudaGetLastError();
}
#endif

}  // namespace mm
