/*
 * maze_oracle.c -- CPU restatement of the MARL-Maze environment hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity ORACLE for the CUDA kernels in marl_maze_b200/csrc.  It is a deliberately literal,
 * scalar, one-environment-at-a-time restatement of the reference's Python (rhuangr/MARL-Maze):
 * explicit exit_route stacks, a byte-per-cell layout, the agent_positions map, Python's Mersenne
 * Twister -- none of the reformulations the kernels use (bit planes, dir-to-exit field, lane pairs).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path never does.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this oracle is
 * pinned against the reference ITSELF, imported unmodified in the build container by
 * tools/make_golden.py; the recorded traces live in tests/golden/ and tests/test_oracle_golden.py
 * replays them (observations bit-exact, state, rewards, dones, mazes generated from Python seeds).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OBS_DIM 65
#define N_AGENTS 2
#define VISION_DEFAULT 4 /* maze_agent.py:16 vision_range=4; per agent (Agent.__init__), see omaze_set_vision */

static const int DELTAS[4][2] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}}; /* maze.py:19, maze_agent.py:7 */

/* ------------------------------------------------------------------------------------------------
 * Python's random module (CPython Modules/_randommodule.c + Lib/random.py, 3.12): MT19937.
 * The reference draws every maze from the module-level `random` (maze.py:172,188,189,232,233,242,245,255).
 * ------------------------------------------------------------------------------------------------ */
typedef struct { uint32_t mt[624]; int idx; } PyMT;

static void mt_init_genrand(PyMT *r, uint32_t s) {
    r->mt[0] = s;
    for (int i = 1; i < 624; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->idx = 624;
}
static void mt_init_by_array(PyMT *r, const uint32_t *key, int klen) {
    mt_init_genrand(r, 19650218u);
    int i = 1, j = 0, k = (624 > klen ? 624 : klen);
    for (; k; k--) {
        r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (k = 623; k; k--) {
        r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
    }
    r->mt[0] = 0x80000000u;
}
static uint32_t mt_u32(PyMT *r) {
    if (r->idx >= 624) {
        uint32_t *mt = r->mt;
        for (int k = 0; k < 624; k++) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->idx = 0;
    }
    uint32_t y = r->mt[r->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* random.seed(int): key = 32-bit limbs of abs(seed), least significant first */
static void mt_seed_u64(PyMT *r, uint64_t seed) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    mt_init_by_array(r, key, key[1] ? 2 : 1);
}
static double mt_random(PyMT *r) { /* random.random() */
    uint32_t a = mt_u32(r) >> 5, b = mt_u32(r) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}
static uint32_t mt_randbelow(PyMT *r, uint32_t n) { /* Random._randbelow_with_getrandbits */
    int k = 0; for (uint32_t t = n; t; t >>= 1) k++;
    uint32_t v = mt_u32(r) >> (32 - k);
    while (v >= n) v = mt_u32(r) >> (32 - k);
    return v;
}

/* ------------------------------------------------------------------------------------------------
 * Counter-based generator stream (Philox4x32-10).  The reference's generator is stream-bound to MT;
 * the CUDA generator (K1) draws from Philox keyed by (seed, maze id) instead.  The oracle runs the SAME
 * carving algorithm (maze.py:170-273) under either stream: MT pins the algorithm against the reference,
 * Philox pins K1 bit-exactly against the oracle.   Draw mapping for the Philox stream (shared with
 * csrc/mm_generate.cuh): word i of the stream = lane (i&3) of Philox(counter=(i>>2,0,0,0), key=(seed_lo ^ id, seed_hi + C)).
 *   random()      -> (float)((w >> 8) + 1) * 2^-24 in (0,1], compared in fp32 against an fp32 corridor accumulator
 *   randbelow(n)  -> (uint64)w * n >> 32
 * ------------------------------------------------------------------------------------------------ */
typedef struct { uint32_t key0, key1; uint32_t ctr; uint32_t buf[4]; int have; } Philox;

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static void philox_seed(Philox *p, uint64_t seed, uint32_t maze_id) {
    p->key0 = (uint32_t)seed ^ maze_id; p->key1 = (uint32_t)(seed >> 32) + 0x632BE5ABu; p->ctr = 0; p->have = 0;
}
static uint32_t philox_u32(Philox *p) {
    if (!p->have) { philox4x32_10(p->ctr++, 0, 0, 0, p->key0, p->key1, p->buf); p->have = 4; }
    return p->buf[4 - p->have--];
}

typedef struct { int kind; /* 0 = Python MT, 1 = Philox */ PyMT mt; Philox px; } Rng;

static uint32_t rng_below(Rng *r, uint32_t n) {
    if (r->kind == 0) return mt_randbelow(&r->mt, n);
    return (uint32_t)(((uint64_t)philox_u32(&r->px) * n) >> 32);
}
static int rng_randint(Rng *r, int a, int b) { return a + (int)rng_below(r, (uint32_t)(b - a + 1)); } /* random.randint */

/* ------------------------------------------------------------------------------------------------
 * State: maze_agent.py:16-57 (Agent fields), maze.py:22-53 (Maze fields)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int x, y, direction, tag;
    int vision;                                 /* vision_range, maze_agent.py:16,19 */
    int has_last_mark, lm_x, lm_y;              /* last_mark_pos (None | tuple) */
    int knows_end, sees_end, other_knows_end;
    int next_move_to_exit[4];
    int exit_len;
    int *route; int route_len, route_cap; int route_none; /* exit_route: None | list used as a stack */
    int has_key, sees_key, team_has_key;
    int ols_x, ols_y;                           /* other_last_seen */
    long time_from_last_seen;
    int current_t;
    int memory[4];                              /* deque(maxlen=4), oldest first */
    int min_x, max_x, min_y, max_y, width_est, height_est;
} OAgent;

typedef struct {
    int width, height;
    uint8_t *layout; int layout_cap;            /* layout[y*width+x]: 0 path, 1 wall, 2/3 marks */
    int start_x, start_y, end_x, end_y;
    int key_present, key_x, key_y;              /* maze.key: tuple | 0 (maze.py:158) */
    int *path; int path_len, path_cap;          /* shortest_path as x0,y0,x1,y1,... */
    int shortest_path_len;
    int current_t, max_timestep;
    OAgent agents[N_AGENTS];
    int in_map[N_AGENTS];                       /* agent_positions: which agents are registered */
    /* generation parameters, maze.py:48-53 */
    int rand_sizes, rr0, rr1, rand_start, difficulty, ds0, ds1;
    Rng rng;
    int error;
} OMaze;

static void route_reserve(OAgent *a, int n) {
    if (a->route_cap < n) { a->route_cap = n * 2 + 16; a->route = (int *)realloc(a->route, sizeof(int) * a->route_cap); }
}
static void route_push(OAgent *a, int d) { route_reserve(a, a->route_len + 1); a->route[a->route_len++] = d; }
static void route_copy(OAgent *dst, const OAgent *src) { /* [dir for dir in self.exit_route] */
    route_reserve(dst, src->route_len + 1);
    memcpy(dst->route, src->route, sizeof(int) * src->route_len);
    dst->route_len = src->route_len; dst->route_none = 0;
}

static void agent_init(OAgent *a, int tag) { /* maze_agent.py:16-57 */
    memset(a, 0, sizeof(*a));
    a->tag = tag; a->vision = VISION_DEFAULT; a->direction = 2; a->exit_len = -1; a->route_none = 1;
    for (int i = 0; i < 4; i++) a->memory[i] = -1;
    a->width_est = a->height_est = 1;
    /* other_last_seen = None until the first reset; never read before that */
}

OMaze *omaze_new(int max_timestep, int difficulty, int rand_start, int rand_sizes, int rr0, int rr1, int ds0, int ds1) {
    OMaze *m = (OMaze *)calloc(1, sizeof(OMaze)); /* maze.py:22-53 */
    m->width = ds0 * 2 - 1; m->height = ds1 * 2 - 1;
    m->max_timestep = max_timestep; m->difficulty = difficulty; m->rand_start = rand_start;
    m->rand_sizes = rand_sizes; m->rr0 = rr0; m->rr1 = rr1; m->ds0 = ds0; m->ds1 = ds1;
    agent_init(&m->agents[0], 2); agent_init(&m->agents[1], 3); /* main.py:18-19 tags */
    m->rng.kind = 0; mt_seed_u64(&m->rng.mt, 0);
    return m;
}
/* Agent(..., vision_range=r) per agent (maze_agent.py:16): how far the rays of get_visibility_features / get_dead_ends reach */
void omaze_set_vision(OMaze *m, int r0, int r1) { m->agents[0].vision = r0; m->agents[1].vision = r1; }
void omaze_free(OMaze *m) {
    if (!m) return;
    for (int i = 0; i < N_AGENTS; i++) free(m->agents[i].route);
    free(m->layout); free(m->path); free(m);
}
void omaze_seed(OMaze *m, uint64_t seed) { m->rng.kind = 0; mt_seed_u64(&m->rng.mt, seed); } /* random.seed(int) */
void omaze_seed_philox(OMaze *m, uint64_t seed, uint32_t maze_id) { m->rng.kind = 1; philox_seed(&m->rng.px, seed, maze_id); }

static int is_valid_cell(const OMaze *m, int x, int y) { return 0 <= x && x < m->width && 0 <= y && y < m->height; } /* maze.py:166 */
#define LAY(m, x, y) ((m)->layout[(y) * (m)->width + (x)])

/* ------------------------------------------------------------------------------------------------
 * Generation: maze.py:170-273
 * ------------------------------------------------------------------------------------------------ */
static void set_start(OMaze *m) { /* maze.py:229-237 */
    if (m->rand_start) {
        m->start_x = rng_randint(&m->rng, 0, (m->width - 1) / 2) * 2;
        m->start_y = rng_randint(&m->rng, 0, (m->height - 1) / 2) * 2;
    } else {
        m->start_x = ((m->width / 2) % 2 == 0) ? m->width / 2 : m->width / 2 - 1;
        m->start_y = 0;
    }
}
static void set_end(OMaze *m) { /* maze.py:239-250 */
    int coin = rng_randint(&m->rng, 0, 1);
    int x = coin == 0 ? 0 : m->width - 1;
    for (long tries = 0;; tries++) {
        /* The reference loops forever when the chosen edge has no eligible cell -- which its own generator can produce (a carve that
         * is popped early at every frontier leaves a partial maze; seen once in 262 144 small mazes).  Oracle and K1 share a defined
         * way out: after 4096 draws take the first open cell != start in row-major order, and raise the error flag. */
        if (tries >= 4096) {
            m->error |= 2; m->end_x = m->start_x; m->end_y = m->start_y;
            for (int yy = 0, found = 0; yy < m->height && !found; yy++)
                for (int xx = 0; xx < m->width && !found; xx++)
                    if (LAY(m, xx, yy) == 0 && !(xx == m->start_x && yy == m->start_y)) { m->end_x = xx; m->end_y = yy; found = 1; }
            break;
        }
        int y = rng_randint(&m->rng, 0, m->height - 1);
        if (x == m->start_x && y == m->start_y) continue;
        if (LAY(m, x, y) == 0) { m->end_x = x; m->end_y = y; break; }
    }
}
/* maze.py:261-273: stack DFS that carries the path; parent links recorded at push time give the same path */
static int get_shortest_path(OMaze *m, int sx, int sy, int ex, int ey, int **out, int *out_cap) {
    int n = m->width * m->height;
    int *parent = (int *)malloc(sizeof(int) * n), *stack = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) parent[i] = -2;
    int sp = 0; stack[sp++] = sy * m->width + sx; parent[sy * m->width + sx] = -1;
    int found = -1;
    while (sp) {
        int c = stack[--sp], x = c % m->width, y = c / m->width;
        if (x == ex && y == ey) { found = c; break; }
        for (int k = 0; k < 4; k++) {
            int nx = x + DELTAS[k][0], ny = y + DELTAS[k][1];
            if (is_valid_cell(m, nx, ny) && LAY(m, nx, ny) == 0 && parent[ny * m->width + nx] == -2) {
                parent[ny * m->width + nx] = c; stack[sp++] = ny * m->width + nx;
            }
        }
    }
    int len = 0;
    if (found >= 0) {
        for (int c = found; c != -1; c = parent[c]) len++;
        if (*out_cap < 2 * len) { *out_cap = 2 * len; *out = (int *)realloc(*out, sizeof(int) * *out_cap); }
        int i = len - 1;
        for (int c = found; c != -1; c = parent[c], i--) { (*out)[2 * i] = c % m->width; (*out)[2 * i + 1] = c / m->width; }
    }
    free(parent); free(stack);
    return len;
}
static int in_path(const OMaze *m, int x, int y) {
    for (int i = 0; i < m->path_len; i++) if (m->path[2 * i] == x && m->path[2 * i + 1] == y) return 1;
    return 0;
}
static void set_key(OMaze *m) { /* maze.py:252-259 */
    for (long tries = 0;; tries++) {
        if (tries >= 65536) { /* reference: infinite loop (no open cell off the path).  Shared way out: first open cell that is not
                                 start / exit / on the path, else first open cell that is not start / exit, else the start. */
            m->error |= 4; m->key_x = m->start_x; m->key_y = m->start_y; m->key_present = 1;
            for (int pass = 0, found = 0; pass < 2 && !found; pass++)
                for (int yy = 0; yy < m->height && !found; yy++)
                    for (int xx = 0; xx < m->width && !found; xx++)
                        if (LAY(m, xx, yy) == 0 && !(xx == m->start_x && yy == m->start_y) && !(xx == m->end_x && yy == m->end_y) && (pass == 1 || !in_path(m, xx, yy))) {
                            m->key_x = xx; m->key_y = yy; found = 1;
                        }
            break;
        }
        int x = rng_randint(&m->rng, 0, m->width - 1), y = rng_randint(&m->rng, 0, m->height - 1);
        if (LAY(m, x, y) == 1 || (x == m->end_x && y == m->end_y) || (x == m->start_x && y == m->start_y) || in_path(m, x, y)) continue;
        m->key_x = x; m->key_y = y; m->key_present = 1; break;
    }
}
void omaze_build(OMaze *m) { /* maze.py:170-218 */
    if (m->rand_sizes) { int size = rng_randint(&m->rng, m->rr0, m->rr1) * 2 - 1; m->height = size; m->width = size; }
    int n = m->width * m->height;
    if (m->layout_cap < n) { m->layout_cap = n; m->layout = (uint8_t *)realloc(m->layout, n); }
    memset(m->layout, 1, n);
    set_start(m);
    int *stack = (int *)malloc(sizeof(int) * n); int sp = 0;
    stack[sp++] = m->start_y * m->width + m->start_x;
    const int mx = m->width > m->height ? m->width : m->height;
    double cc = 0.0; float ccf = 0.0f; const double inc = 1.0 / (10 * mx); const float incf = 1.0f / (float)(10 * mx);
    while (sp) {
        int c = stack[sp - 1], x = c % m->width, y = c / m->width;
        LAY(m, x, y) = 0;
        int nb[4], nn = 0; /* maze.py:220-227: 2-away, in bounds, still wall; order N,E,S,W */
        for (int k = 0; k < 4; k++) {
            int nx = x + DELTAS[k][0] * 2, ny = y + DELTAS[k][1] * 2;
            if (is_valid_cell(m, nx, ny) && LAY(m, nx, ny) == 1) nb[nn++] = ny * m->width + nx;
        }
        int go = 0;
        if (nn) { /* `neighbors and random.random() > corridor_const` short-circuits, maze.py:188 */
            if (m->rng.kind == 0) go = mt_random(&m->rng.mt) > cc;
            else go = ((float)((philox_u32(&m->rng.px) >> 8) + 1u) * (1.0f / 16777216.0f)) > ccf; /* (0,1]: like random() it never
                                                                                                   loses against corridor_const == 0 */
        }
        if (go) {
            int nc = nb[rng_below(&m->rng, (uint32_t)nn)]; /* random.choice */
            int x2 = nc % m->width, y2 = nc / m->width;
            LAY(m, (x + x2) / 2, (y + y2) / 2) = 0;
            stack[sp++] = nc;
            cc += inc; ccf += incf;
        } else { sp--; cc = 0.0; ccf = 0.0f; }
    }
    free(stack);
    /* maze.py:203-217: `difficulty` candidate exits; dict keyed by length => the LAST candidate of maximal length wins */
    int best_len = 0, best_ex = 0, best_ey = 0; int *cand = NULL, cand_cap = 0;
    for (int d = 0; d < m->difficulty; d++) {
        set_end(m);
        int len = get_shortest_path(m, m->start_x, m->start_y, m->end_x, m->end_y, &cand, &cand_cap);
        if (len >= best_len) {
            best_len = len; best_ex = m->end_x; best_ey = m->end_y;
            if (m->path_cap < 2 * len) { m->path_cap = 2 * len; m->path = (int *)realloc(m->path, sizeof(int) * m->path_cap); }
            memcpy(m->path, cand, sizeof(int) * 2 * len); m->path_len = len;
        }
    }
    free(cand);
    m->end_x = best_ex; m->end_y = best_ey; m->shortest_path_len = best_len;
    set_key(m);
}

/* Inject a maze instead of generating one (the parity route: layouts recorded from the reference). */
void omaze_inject(OMaze *m, int width, int height, const uint8_t *layout, int p0x, int p0y, int p1x, int p1y,
                  int ex, int ey, int kx, int ky, int spl) {
    m->width = width; m->height = height;
    int n = width * height;
    if (m->layout_cap < n) { m->layout_cap = n; m->layout = (uint8_t *)realloc(m->layout, n); }
    memcpy(m->layout, layout, n);
    m->start_x = p0x; m->start_y = p0y; m->end_x = ex; m->end_y = ey;
    m->key_x = kx; m->key_y = ky; m->key_present = 1;
    if (m->path_cap < 4) { m->path_cap = 4; m->path = (int *)realloc(m->path, sizeof(int) * 4); }
    m->path[0] = p0x; m->path[1] = p0y; m->path[2] = p1x; m->path[3] = p1y; m->path_len = 2;
    m->shortest_path_len = spl;
}

/* ------------------------------------------------------------------------------------------------
 * Observation: maze_agent.py:89-358
 * ------------------------------------------------------------------------------------------------ */
static void agent_neighbors(const OMaze *m, const OAgent *a, int x, int y, int nb[4]) { /* maze_agent.py:347-358 */
    for (int i = 0; i < 4; i++) {
        const int *d = DELTAS[(i + a->direction) % 4];
        int nx = x + d[0], ny = y + d[1];
        nb[i] = is_valid_cell(m, nx, ny) && LAY(m, nx, ny) != 1;
    }
}
static void update_maze_minmax(OAgent *a, int direction, int nx, int ny) { /* maze_agent.py:313-328 */
    if (direction == 0 && ny < a->min_y) a->min_y = ny;
    else if (direction == 1 && nx > a->max_x) a->max_x = nx;
    else if (direction == 2 && ny > a->max_y) a->max_y = ny;
    else if (direction == 3 && nx < a->min_x) a->min_x = nx;
}
static int pymod4(int v) { return ((v % 4) + 4) % 4; } /* Python % is non-negative for a positive modulus */

typedef struct { double own_mark[4], others_mark[4]; int agents[4], key[4], other_dir[4]; double other_last_pos[2]; } Vis;

static void get_visibility_features(OMaze *m, int self_idx, Vis *v) { /* maze_agent.py:188-277 */
    OAgent *s = &m->agents[self_idx];
    memset(v, 0, sizeof(*v));
    int n_vis = 0; /* len(visible_agents) so far */
    s->time_from_last_seen += 1;
    s->sees_end = (s->x == m->end_x && s->y == m->end_y);
    s->sees_key = 0;
    for (int o = 0; o < N_AGENTS; o++) { /* :199-213 same-cell test against the other agent's LIVE x,y */
        if (o == self_idx) continue;
        OAgent *ag = &m->agents[o];
        if (ag->x == s->x && ag->y == s->y) {
            s->time_from_last_seen = 0;
            for (int i = 0; i < 4; i++) v->agents[i] = 1;
            n_vis += 4;
            s->ols_x = ag->x; s->ols_y = ag->y;
            s->team_has_key = s->team_has_key || ag->has_key;
            s->other_knows_end = s->other_knows_end || ag->knows_end;
            v->other_dir[ag->direction] = 1;
            if (s->knows_end && !ag->knows_end) {
                route_copy(ag, s);
                s->other_knows_end = 1; ag->knows_end = 1; ag->other_knows_end = 1;
            }
        }
    }
    for (int dir = 0; dir < 4; dir++) { /* :215-269 */
        int nx = s->x, ny = s->y;
        const int ad = (dir + s->direction) % 4;
        for (int j = 1; j <= s->vision; j++) {
            nx += DELTAS[ad][0]; ny += DELTAS[ad][1];
            if (nx < 0 || nx >= m->width || ny < 0 || ny >= m->height || LAY(m, nx, ny) == 1) break;
            if (nx == m->end_x && ny == m->end_y) { /* :227-233 */
                s->knows_end = 1; s->sees_end = 1;
                if (s->exit_len == -1) {
                    s->route_len = 0; s->route_none = 0;
                    for (int q = 0; q < j; q++) route_push(s, ad);
                    s->exit_len = j;
                }
            }
            if (m->key_present && nx == m->key_x && ny == m->key_y) { s->sees_key = 1; v->key[dir] = 1; } /* :235-237 */
            for (int o = 0; o < N_AGENTS; o++) { /* :239-260 agent_positions lookup */
                if (!m->in_map[o] || o == self_idx) continue;
                OAgent *ag = &m->agents[o];
                if (ag->x != nx || ag->y != ny) continue;
                s->time_from_last_seen = 0;
                s->ols_x = ag->x; s->ols_y = ag->y;
                s->other_knows_end = s->other_knows_end || ag->knows_end;
                s->team_has_key = s->team_has_key || ag->has_key;
                v->other_dir[ag->direction] = 1;
                if (n_vis < 4) v->agents[dir] = 1; /* extend([one-hot]); with 2 agents this is the only entry */
                n_vis += 4;
                if (j == 1 && s->knows_end && !ag->knows_end) {
                    route_copy(ag, s);
                    if (s->route_len > 0 && ad == s->route[s->route_len - 1]) ag->route_len--;
                    else route_push(ag, (ad + 2) % 4);
                    s->other_knows_end = 1; ag->knows_end = 1; ag->other_knows_end = 1;
                }
            }
            int cell = LAY(m, nx, ny);
            if (cell == s->tag) v->own_mark[dir] += 1.0 / s->vision;       /* :263-264 */
            else if (cell > 1) v->others_mark[dir] += 1.0 / s->vision;    /* :266-267 */
            update_maze_minmax(s, ad, nx, ny);                          /* :269 */
        }
    }
    /* update_maze_dims, maze_agent.py:330-335 */
    int we = s->max_x - s->min_x, he = s->max_y - s->min_y;
    s->width_est = we != 0 ? we : 1; s->height_est = he != 0 ? he : 1;
    v->other_last_pos[0] = (double)(s->ols_x - s->min_x) / (double)s->width_est;   /* :272 */
    v->other_last_pos[1] = (double)(s->max_y - s->ols_y) / (double)s->height_est;  /* :273 */
}

static void get_dead_ends(OMaze *m, OAgent *s, double dead[4], int move_mask[4]) { /* maze_agent.py:143-185 */
    int nb0[4], nb[4];
    agent_neighbors(m, s, s->x, s->y, nb0);
    for (int d = 0; d < 4; d++) dead[d] = nb0[d] ? 0.0 : 1.0;
    const double distance = 1.0 / s->vision;
    for (int d = 0; d < 4; d++) {
        if (dead[d] == 1.0) continue;
        int nx = s->x, ny = s->y;
        const int *dl = DELTAS[(d + s->direction) % 4];
        for (int j = 1; j <= s->vision; j++) {
            nx += dl[0]; ny += dl[1];
            agent_neighbors(m, s, nx, ny, nb);
            if (nb[(d + 1) % 4] || nb[pymod4(d - 1)]) break;
            int cnt = nb[0] + nb[1] + nb[2] + nb[3];
            if (cnt == 1) { dead[d] = 1.0 - j * distance; break; }
            else if (!nb[d]) break;
        }
    }
    if (!s->sees_end && !s->sees_key) for (int d = 0; d < 4; d++) move_mask[d] = (dead[d] == 0.0);
    else for (int d = 0; d < 4; d++) move_mask[d] = nb0[d];
}

static void get_direction_from(const OAgent *s, int ox, int oy, int out[4]) { /* maze_agent.py:297-311 */
    if (ox == s->x && oy == s->y) { out[0] = out[1] = out[2] = out[3] = 1; return; }
    out[0] = out[1] = out[2] = out[3] = 0;
    if (oy > s->y) out[pymod4(2 - s->direction)] = 1; else if (oy < s->y) out[pymod4(0 - s->direction)] = 1;
    if (ox > s->x) out[pymod4(1 - s->direction)] = 1; else if (ox < s->x) out[pymod4(3 - s->direction)] = 1;
}

/* maze_agent.py:89-140.  obs values are produced in double (python int/int true division) and cast to f32
 * exactly where the reference does (torch.as_tensor(..., dtype=float32), PPO.py:144 / networks.py:32). */
static void get_observations(OMaze *m, int idx, float *obs, uint8_t *mask) {
    OAgent *s = &m->agents[idx];
    double o[OBS_DIM]; int k = 0;
    Vis v; double dead[4]; int mm[4];
    for (int i = 0; i < 4; i++) o[k++] = (i == s->direction);
    get_visibility_features(m, idx, &v);
    get_dead_ends(m, s, dead, mm);
    for (int i = 0; i < 4; i++) o[k++] = dead[i];
    for (int i = 0; i < 4; i++) o[k++] = v.own_mark[i];
    for (int i = 0; i < 4; i++) o[k++] = v.others_mark[i];
    for (int i = 0; i < 4; i++) o[k++] = v.agents[i];
    for (int i = 0; i < 4; i++) o[k++] = v.other_dir[i];
    for (int i = 0; i < 4; i++) o[k++] = v.key[i];
    for (int i = 0; i < 4; i++) for (int q = 0; q < 4; q++) o[k++] = (s->memory[i] == q); /* get_memory :289-294 */
    int lm[4] = {0, 0, 0, 0};
    if (s->has_last_mark) get_direction_from(s, s->lm_x, s->lm_y, lm); /* :105 */
    for (int i = 0; i < 4; i++) o[k++] = lm[i];
    o[k++] = (double)(s->x - s->min_x) / (double)s->width_est;   /* :107 */
    o[k++] = (double)(s->max_y - s->y) / (double)s->height_est;  /* :108 */
    o[k++] = v.other_last_pos[0]; o[k++] = v.other_last_pos[1];
    o[k++] = s->sees_end;
    int nm[4] = {0, 0, 0, 0};
    if (!s->route_none && s->route_len > 0) nm[pymod4(s->route[s->route_len - 1] - s->direction)] = 1; /* :114-115 */
    else nm[0] = nm[1] = nm[2] = nm[3] = 1;
    for (int i = 0; i < 4; i++) { s->next_move_to_exit[i] = nm[i]; o[k++] = nm[i]; }
    o[k++] = s->exit_len < 40 ? (double)s->exit_len / 40.0 : 1.0;  /* :120 */
    o[k++] = s->other_knows_end; o[k++] = s->has_key; o[k++] = s->team_has_key;
    o[k++] = s->time_from_last_seen < 40 ? (double)s->time_from_last_seen / 40.0 : 1.0; /* :125 */
    o[k++] = (double)s->current_t / (double)m->max_timestep;       /* :127 */
    o[k++] = (2 - s->tag == 0); o[k++] = (2 - s->tag == -1);       /* id[2-tag]=1 with python negative index :128-130 */
    for (int i = 0; i < OBS_DIM; i++) obs[i] = (float)o[i];
    int anykey = v.key[0] | v.key[1] | v.key[2] | v.key[3];
    if (anykey) { /* :132-134 argmax = first ray showing the key */
        int f = v.key[0] ? 0 : v.key[1] ? 1 : v.key[2] ? 2 : 3;
        for (int i = 0; i < 4; i++) mm[i] = (i == f);
    }
    int anyagent = v.agents[0] | v.agents[1] | v.agents[2] | v.agents[3];
    for (int i = 0; i < 4; i++) mask[i] = (uint8_t)mm[i];
    mask[4] = (uint8_t)(anyagent && s->x == m->end_x && s->x == m->end_y); /* (self.x, self.x) == maze.end -- sic, :136 */
    mask[5] = (uint8_t)(LAY(m, s->x, s->y) != s->tag);                   /* :135 */
}

/* ------------------------------------------------------------------------------------------------
 * reset / step: maze.py:55-163, maze_agent.py:59-79
 * ------------------------------------------------------------------------------------------------ */
static void agent_reset(OAgent *a, int x, int y) { /* maze_agent.py:59-79 (time_from_last_seen is NOT reset) */
    a->current_t = 0; a->x = x; a->y = y; a->ols_x = x; a->ols_y = y;
    a->min_x = a->max_x = x; a->min_y = a->max_y = y; a->width_est = a->height_est = 1;
    a->direction = 2; a->has_last_mark = 0;
    for (int i = 0; i < 4; i++) { a->memory[i] = -1; a->next_move_to_exit[i] = 0; }
    a->knows_end = a->other_knows_end = a->sees_end = 0;
    a->exit_len = -1; a->route_none = 1; a->route_len = 0;
    a->has_key = a->team_has_key = a->sees_key = 0;
}
/* maze.py:55-72 after build_maze(): obs of agent i is computed INSIDE the placement loop */
static void reset_agents(OMaze *m, float *obs, uint8_t *masks) {
    m->current_t = 0;
    m->in_map[0] = m->in_map[1] = 0;
    for (int i = 0; i < N_AGENTS; i++) {
        agent_reset(&m->agents[i], m->path[2 * i], m->path[2 * i + 1]);
        m->in_map[i] = 1;
        get_observations(m, i, obs + i * OBS_DIM, masks + i * 6);
    }
}
void omaze_reset(OMaze *m, float *obs, uint8_t *masks) { omaze_build(m); reset_agents(m, obs, masks); }
/* the reference's two public agent setters, for tests of the host surface: Agent.reset(x, y) maze_agent.py:59-79, Agent.move(x, y, direction) :85-87 */
void omaze_agent_reset(OMaze *m, int a, int x, int y) { agent_reset(&m->agents[a], x, y); }
void omaze_agent_move(OMaze *m, int a, int x, int y, int direction) { m->agents[a].x = x; m->agents[a].y = y; m->agents[a].direction = direction; }
void omaze_reset_injected(OMaze *m, int width, int height, const uint8_t *layout, int p0x, int p0y, int p1x, int p1y,
                          int ex, int ey, int kx, int ky, int spl, float *obs, uint8_t *masks) {
    omaze_inject(m, width, height, layout, p0x, p0y, p1x, p1y, ex, ey, kx, ky, spl);
    reset_agents(m, obs, masks);
}

static int single_agent_step(OMaze *m, OAgent *a, int move, int mark) { /* maze.py:124-163; returns got_key */
    int got_key = 0;
    a->current_t = m->current_t;
    if (mark == 1) { LAY(m, a->x, a->y) = (uint8_t)a->tag; a->has_last_mark = 1; a->lm_x = a->x; a->lm_y = a->y; }
    /* moves outside 0..4 are outside the reference's domain (it would walk like move % 4 and then raise IndexError in get_memory,
     * maze_agent.py:289-294): treated as `stop` + error flag, the same convention as the out-of-bounds move below */
    if (move < 0 || move > 4) { m->error |= 1; return 0; }
    if (move != 4) {
        int direction = (move + a->direction) % 4;
        int nx = a->x + DELTAS[direction][0], ny = a->y + DELTAS[direction][1];
        if (!is_valid_cell(m, nx, ny)) { m->error |= 1; return 0; } /* reference prints then indexes out of range (maze.py:141-145); outside the parity domain */
        if (a->knows_end) { /* :148-154 */
            if (a->route_len > 0 && direction == a->route[a->route_len - 1]) { a->route_len--; a->exit_len -= 1; }
            else { route_push(a, (direction + 2) % 4); a->exit_len += 1; }
        }
        a->x = nx; a->y = ny; a->direction = direction; /* Agent.move :85-87 */
        if (m->key_present && nx == m->key_x && ny == m->key_y) { m->key_present = 0; a->has_key = 1; a->team_has_key = 1; got_key = 1; }
        a->memory[0] = a->memory[1]; a->memory[1] = a->memory[2]; a->memory[2] = a->memory[3]; a->memory[3] = move;
    }
    return got_key;
}

void omaze_step(OMaze *m, const int *action /* move0,mark0,move1,mark1 */, float *obs, uint8_t *masks, float *reward, uint8_t *done) {
    m->current_t += 1; /* maze.py:75 */
    int agents_have_key = 0, first_key_find = 0;
    for (int i = 0; i < N_AGENTS; i++) {
        first_key_find += single_agent_step(m, &m->agents[i], action[2 * i], action[2 * i + 1]);
        agents_have_key += m->agents[i].has_key;
    }
    m->in_map[0] = m->in_map[1] = 1; /* :92-97 */
    int exit_ready = 1;
    for (int i = 0; i < N_AGENTS; i++) { /* :102-106 */
        get_observations(m, i, obs + i * OBS_DIM, masks + i * 6);
        exit_ready = exit_ready && m->agents[i].team_has_key && m->agents[i].knows_end;
    }
    if (exit_ready) { /* :107-113 */
        for (int i = 0; i < N_AGENTS; i++) {
            OAgent *a = &m->agents[i]; uint8_t *mk = masks + i * 6;
            if (a->x != m->end_x || a->y != m->end_y) {
                mk[0] = mk[1] = mk[2] = mk[3] = 0;
                int f = 0; for (int q = 3; q >= 0; q--) if (a->next_move_to_exit[q]) f = q; /* np.argmax: first max */
                if (!(a->next_move_to_exit[0] | a->next_move_to_exit[1] | a->next_move_to_exit[2] | a->next_move_to_exit[3])) f = 0;
                mk[f] = 1;
            } else { mk[0] = mk[1] = mk[2] = mk[3] = 0; mk[4] = 1; }
        }
    }
    float r = first_key_find * 0.5f; uint8_t d = 0; /* :115-121 */
    int same = (m->agents[0].x == m->agents[1].x && m->agents[0].y == m->agents[1].y);
    if (agents_have_key && same && m->agents[0].x == m->end_x && m->agents[0].y == m->end_y) { r = 1.0f; d = 1; }
    else if (m->current_t >= m->max_timestep) d = 1;
    *reward = r; *done = d;
}

/* ------------------------------------------------------------------------------------------------ getters */
void omaze_get_maze(const OMaze *m, int *hdr /*[12]*/, uint8_t *layout /* width*height, may be NULL */) {
    hdr[0] = m->width; hdr[1] = m->height; hdr[2] = m->start_x; hdr[3] = m->start_y; hdr[4] = m->end_x; hdr[5] = m->end_y;
    hdr[6] = m->key_present ? m->key_x : -1; hdr[7] = m->key_present ? m->key_y : -1; hdr[8] = m->shortest_path_len;
    hdr[9] = m->path_len > 1 ? m->path[2] : -1; hdr[10] = m->path_len > 1 ? m->path[3] : -1; hdr[11] = m->current_t;
    if (layout) memcpy(layout, m->layout, (size_t)m->width * m->height);
}
int omaze_get_path(const OMaze *m, int *out, int cap) {
    int n = m->path_len < cap ? m->path_len : cap; memcpy(out, m->path, sizeof(int) * 2 * n); return m->path_len;
}
#define N_AGENT_FIELDS 18
void omaze_get_agents(const OMaze *m, int *out /*[2][18]*/) { /* same order as tools/ref_harness.py AGENT_FIELDS */
    for (int i = 0; i < N_AGENTS; i++) {
        const OAgent *a = &m->agents[i]; int *o = out + i * N_AGENT_FIELDS;
        o[0] = a->x; o[1] = a->y; o[2] = a->direction; o[3] = a->knows_end; o[4] = a->other_knows_end; o[5] = a->has_key;
        o[6] = a->team_has_key; o[7] = a->exit_len; o[8] = (int)a->time_from_last_seen; o[9] = a->ols_x; o[10] = a->ols_y;
        o[11] = a->has_last_mark ? a->lm_x : -1; o[12] = a->has_last_mark ? a->lm_y : -1;
        o[13] = a->min_x; o[14] = a->max_x; o[15] = a->min_y; o[16] = a->max_y; o[17] = a->route_none ? -1 : a->route_len;
    }
}
int omaze_error(const OMaze *m) { return m->error; }

/* ------------------------------------------------------------------------------------------------
 * Batch driver: E independent literal environments stepped in a loop (OpenMP over envs when asked).
 * Mirrors the host-visible contract of the CUDA path (auto-reset from a maze pool, env e takes pool maze
 * (e + episode*E) mod P) so that tests and the CPU baseline drive both sides with the same calls.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int width, height, p0x, p0y, p1x, p1y, ex, ey, kx, ky, spl;
    uint8_t *layout;
} OPoolMaze;
typedef struct {
    int E, P; OMaze **envs; OPoolMaze *pool; int *episode; int *maze_idx;
    int8_t **dkey, **dexit; int *dir_episode; /* guided-action driver cache (test helper, not reference code) */
} OBatch;

OBatch *obatch_new(int E, int P, int max_timestep) {
    OBatch *b = (OBatch *)calloc(1, sizeof(OBatch));
    b->E = E; b->P = P;
    b->envs = (OMaze **)calloc(E, sizeof(OMaze *));
    for (int e = 0; e < E; e++) b->envs[e] = omaze_new(max_timestep, 1, 1, 0, 0, 0, 4, 4);
    b->pool = (OPoolMaze *)calloc(P, sizeof(OPoolMaze));
    b->episode = (int *)calloc(E, sizeof(int)); b->maze_idx = (int *)calloc(E, sizeof(int));
    b->dkey = (int8_t **)calloc(E, sizeof(int8_t *)); b->dexit = (int8_t **)calloc(E, sizeof(int8_t *));
    b->dir_episode = (int *)calloc(E, sizeof(int));
    return b;
}
void obatch_set_vision(OBatch *b, int r0, int r1) { for (int e = 0; e < b->E; e++) omaze_set_vision(b->envs[e], r0, r1); }
void obatch_free(OBatch *b) {
    if (!b) return;
    for (int e = 0; e < b->E; e++) omaze_free(b->envs[e]);
    for (int p = 0; p < b->P; p++) free(b->pool[p].layout);
    for (int e = 0; e < b->E; e++) { free(b->dkey[e]); free(b->dexit[e]); }
    free(b->dkey); free(b->dexit); free(b->dir_episode);
    free(b->envs); free(b->pool); free(b->episode); free(b->maze_idx); free(b);
}
void obatch_set_pool_maze(OBatch *b, int p, int width, int height, const uint8_t *layout, int p0x, int p0y, int p1x, int p1y,
                          int ex, int ey, int kx, int ky, int spl) {
    OPoolMaze *q = &b->pool[p];
    q->width = width; q->height = height; q->p0x = p0x; q->p0y = p0y; q->p1x = p1x; q->p1y = p1y;
    q->ex = ex; q->ey = ey; q->kx = kx; q->ky = ky; q->spl = spl;
    q->layout = (uint8_t *)realloc(q->layout, (size_t)width * height);
    memcpy(q->layout, layout, (size_t)width * height);
}
/* generate pool maze p with the oracle's own generator under the Philox stream (pins K1) */
void obatch_generate_pool_maze(OBatch *b, int p, int S, int rand_start, int difficulty, uint64_t seed, uint32_t maze_id) {
    OMaze *g = omaze_new(1, difficulty, rand_start, 0, 0, 0, (S + 1) / 2, (S + 1) / 2);
    omaze_seed_philox(g, seed, maze_id);
    omaze_build(g);
    obatch_set_pool_maze(b, p, g->width, g->height, g->layout, g->path[0], g->path[1], g->path[2], g->path[3],
                         g->end_x, g->end_y, g->key_x, g->key_y, g->shortest_path_len);
    omaze_free(g);
}
static void obatch_reset_env(OBatch *b, int e, float *obs, uint8_t *masks) {
    int p = (int)(((long)e + (long)b->episode[e] * b->E) % b->P);
    const OPoolMaze *q = &b->pool[p];
    b->maze_idx[e] = p;
    omaze_reset_injected(b->envs[e], q->width, q->height, q->layout, q->p0x, q->p0y, q->p1x, q->p1y, q->ex, q->ey, q->kx, q->ky, q->spl,
                         obs + (size_t)e * 2 * OBS_DIM, masks + (size_t)e * 12);
    b->episode[e]++;
}
void obatch_reset_all(OBatch *b, float *obs, uint8_t *masks, int threads) {
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int e = 0; e < b->E; e++) { b->episode[e] = 0; obatch_reset_env(b, e, obs, masks); }
}
void obatch_reset_masked(OBatch *b, const uint8_t *which, float *obs, uint8_t *masks) {
    for (int e = 0; e < b->E; e++) if (which[e]) obatch_reset_env(b, e, obs, masks);
}
/* actions: [E][2][2] u8 (move, mark).  auto_reset: on done the emitted obs/masks are the reset ones (PPO.py:127-130). */
void obatch_step(OBatch *b, const uint8_t *actions, float *obs, uint8_t *masks, float *reward, uint8_t *done, int auto_reset, int threads) {
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int e = 0; e < b->E; e++) {
        int act[4] = {actions[4 * e], actions[4 * e + 1], actions[4 * e + 2], actions[4 * e + 3]};
        omaze_step(b->envs[e], act, obs + (size_t)e * 2 * OBS_DIM, masks + (size_t)e * 12, reward + e, done + e);
        if (auto_reset && done[e]) obatch_reset_env(b, e, obs, masks);
    }
}
void obatch_get_agents(const OBatch *b, int *out /*[E][2][18]*/) { for (int e = 0; e < b->E; e++) omaze_get_agents(b->envs[e], out + (size_t)e * 2 * N_AGENT_FIELDS); }
void obatch_get_env(const OBatch *b, int *out /*[E][4]: t, key_x, key_y, maze_idx*/) {
    for (int e = 0; e < b->E; e++) {
        const OMaze *m = b->envs[e];
        out[4 * e] = m->current_t; out[4 * e + 1] = m->key_present ? m->key_x : -1; out[4 * e + 2] = m->key_present ? m->key_y : -1; out[4 * e + 3] = b->maze_idx[e];
    }
}
void obatch_get_layout(const OBatch *b, int e, uint8_t *out) { const OMaze *m = b->envs[e]; memcpy(out, m->layout, (size_t)m->width * m->height); }
OMaze *obatch_env(OBatch *b, int e) { return b->envs[e]; }
int obatch_errors(const OBatch *b) { int r = 0; for (int e = 0; e < b->E; e++) r |= b->envs[e]->error; return r; }
/* uniform mask-legal actions with a per-env xorshift stream; CPU-baseline driver only (keeps python out of the timed loop) */
void obatch_random_actions(const OBatch *b, const uint8_t *masks, uint8_t *actions, uint64_t *rng_state /*[E]*/) {
    for (int e = 0; e < b->E; e++) {
        uint64_t s = rng_state[e];
        for (int a = 0; a < 2; a++) {
            const uint8_t *mk = masks + (size_t)e * 12 + a * 6;
            int legal[5], n = 0; for (int k = 0; k < 5; k++) if (mk[k]) legal[n++] = k;
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            actions[4 * e + 2 * a] = (uint8_t)(n ? legal[(s >> 33) % (uint64_t)n] : 4);
            actions[4 * e + 2 * a + 1] = (uint8_t)(mk[5] ? ((s >> 11) & 1) : 0);
        }
        rng_state[e] = s;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Test-driver helper, NOT part of the restated reference: mask-legal actions biased towards the key and then
 * the exit (privileged BFS over the maze), so that batched parity runs exercise key pickup, route sharing, the
 * exit_ready mask override and reward-1 terminations.  Mirrors tools/ref_harness.py guided_action in spirit.
 * ------------------------------------------------------------------------------------------------ */
static void bfs_dirs(const OMaze *m, int tx, int ty, int8_t *out) {
    int n = m->width * m->height; int *q = (int *)malloc(sizeof(int) * n); int head = 0, tail = 0;
    for (int i = 0; i < n; i++) out[i] = -2;
    out[ty * m->width + tx] = -1; q[tail++] = ty * m->width + tx;
    while (head < tail) {
        int c = q[head++], x = c % m->width, y = c / m->width;
        for (int k = 0; k < 4; k++) {
            int nx = x + DELTAS[k][0], ny = y + DELTAS[k][1];
            if (is_valid_cell(m, nx, ny) && LAY(m, nx, ny) != 1 && out[ny * m->width + nx] == -2) { out[ny * m->width + nx] = (int8_t)((k + 2) % 4); q[tail++] = ny * m->width + nx; }
        }
    }
    free(q);
}
void obatch_guided_actions(OBatch *b, const uint8_t *masks, uint8_t *actions, uint64_t *rng_state, int p_follow_1024, int p_mark_1024, int threads) {
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int e = 0; e < b->E; e++) {
        OMaze *m = b->envs[e];
        if (!b->dkey[e] || b->dir_episode[e] != b->episode[e]) {
            b->dkey[e] = (int8_t *)realloc(b->dkey[e], 64 * 64); b->dexit[e] = (int8_t *)realloc(b->dexit[e], 64 * 64);
            bfs_dirs(m, m->key_x, m->key_y, b->dkey[e]); bfs_dirs(m, m->end_x, m->end_y, b->dexit[e]);
            b->dir_episode[e] = b->episode[e];
        }
        uint64_t s = rng_state[e];
        for (int a = 0; a < 2; a++) {
            const uint8_t *mk = masks + (size_t)e * 12 + a * 6; const OAgent *ag = &m->agents[a];
            int legal[5], n = 0; for (int k = 0; k < 5; k++) if (mk[k]) legal[n++] = k;
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            int move = -1;
            if ((int)((s >> 40) & 1023) < p_follow_1024) {
                const int8_t *d = m->key_present ? b->dkey[e] : b->dexit[e];
                int ad = d[ag->y * m->width + ag->x];
                if (ad >= 0) { int rel = ((ad - ag->direction) % 4 + 4) % 4; if (mk[rel]) move = rel; }
            }
            if (move < 0) move = n ? legal[(s >> 20) % (uint64_t)n] : 4;
            actions[4 * e + 2 * a] = (uint8_t)move;
            actions[4 * e + 2 * a + 1] = (uint8_t)(mk[5] ? ((int)((s >> 5) & 1023) < p_mark_1024) : 0);
        }
        rng_state[e] = s;
    }
}

/* ------------------------------------------------------------------------------------------------
 * CPU-baseline driver (bench.py cpu_baseline / --impl reference): every env runs `steps` steps of uniform mask-legal
 * random actions with auto-reset, env-major (each thread keeps one env hot in cache for all its steps -- the best
 * case for the CPU), OpenMP over envs.  Observations are produced every step exactly as the reference does and
 * left in the per-env slot of `obs`.  Returns the number of env-steps executed.
 * ------------------------------------------------------------------------------------------------ */
/* stagger > 0: env e runs (e * 2654435761 mod 2^32) mod stagger steps instead of `steps` -- the benchmark's phase spreader: after it the
 * envs sit at uniformly spread episode times, so a timed window sees the steady-state rate of truncations and resets (bench.py does the
 * same on the GPU with masked resets). */
long obatch_run_random_ex(OBatch *b, int steps, int stagger, int threads, uint64_t seed, float *obs, uint8_t *masks, double *reward_sum) {
    double rs = 0.0;
    long total = 0;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 8) reduction(+ : rs, total)
    for (int e = 0; e < b->E; e++) {
        uint64_t s = (seed + (uint64_t)e + 1) * 0x9E3779B97F4A7C15ull;
        float *o = obs + (size_t)e * 2 * OBS_DIM; uint8_t *mk = masks + (size_t)e * 12;
        const int n_steps = stagger > 0 ? (int)(((uint32_t)e * 2654435761u) % (uint32_t)stagger) : steps;
        total += n_steps;
        for (int t = 0; t < n_steps; t++) {
            int act[4];
            for (int a = 0; a < 2; a++) {
                int legal[5], n = 0; for (int k = 0; k < 5; k++) if (mk[a * 6 + k]) legal[n++] = k;
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                act[2 * a] = n ? legal[(s >> 33) % (uint64_t)n] : 4;
                act[2 * a + 1] = mk[a * 6 + 5] ? (int)((s >> 11) & 1) : 0;
            }
            float r; uint8_t d;
            omaze_step(b->envs[e], act, o, mk, &r, &d);
            rs += r;
            if (d) obatch_reset_env(b, e, obs, masks);
        }
    }
    if (reward_sum) *reward_sum = rs;
    return total;
}
long obatch_run_random(OBatch *b, int steps, int threads, uint64_t seed, float *obs, uint8_t *masks, double *reward_sum) {
    return obatch_run_random_ex(b, steps, 0, threads, seed, obs, masks, reward_sum);
}
