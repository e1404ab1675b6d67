// minecraft_3D_maze search machinery (bit-mask level, FIFO label-correcting player search, far tile), shared by the
// plain problem (step_maze3d.cu) and the holey problems on the bordered map (step_maze3d_holey.cu).
#pragma once
#include "step_search.cuh"

namespace pcgrl {

constexpr int MAZE_QCAP = 256;    // FIFO ring capacity (entries); measured maximum with push filtering: 50
#ifndef PCGRL_UF_RUNS32
#define PCGRL_UF_RUNS32 1   // A/B on B200 (14^3): u16 per-cell parents / u32 run parents -> 2.12 / 2.26e7 env-steps/s
#endif
#ifndef PCGRL_MAZE_BATCH
#define PCGRL_MAZE_BATCH 0   // 1: the player search pops up to 32 queue entries per round (search_batch), 0: one per round.
                             // A/B on B200 (14^3, 65 536 envs, bench workload): 1.45e7 against 2.25e7 env-steps/s -- the
                             // queue rarely holds more than 1-3 entries (a player's frontier in a 14^3 level is narrow:
                             // 1.4 entries per round on 50 % AIR maps), and a round costs about four one-pop iterations.
                             // Parity is green either way (fixtures, traces, CPU-restatement rollouts, holey problems).
#endif
#if PCGRL_MAZE_BATCH
#define PCGRL_MAZE_SEARCH search_batch
#else
#define PCGRL_MAZE_SEARCH search
#endif
#ifndef PCGRL_MAZE_WARPS
#define PCGRL_MAZE_WARPS 6   // A/B on B200 (14^3, 65 536 envs): 4 / 6 / 8 / 16 warps per CTA -> 1.81 / 2.13 / 2.05 / 1.41e7 env-steps/s
#endif
constexpr int MAZE_WARPS = PCGRL_MAZE_WARPS;     // small CTAs, several per SM: finer-grained tile barriers
constexpr int MAZE_CTAS_PER_SM = 32 / MAZE_WARPS;
constexpr int MAZE_MAX_CTAS = 160 * MAZE_CTAS_PER_SM;   // sizes the global scratch (>= 148 SMs x 8 CTAs)
constexpr int MAZE_NJ_SLICE = 16 * 16 * 16 * 2;         // bytes of nj per warp (maps up to 16^3)
constexpr int MAZE_ORDER_SLICE = (16 / 2 + 1) * 16 * 16 * 2;   // bytes of the first-recording order list per warp
constexpr int MAZE_PAR_SLICE = 16 * 16 * 16 * 2;        // bytes of the parent pointers per warp (holey maze only)
constexpr int MAZE_SLICE = MAZE_NJ_SLICE + MAZE_ORDER_SLICE + MAZE_PAR_SLICE;

struct MazeLayout {
    int best, q_cl, nj, order, q_nj, q_par, col, row, total, order_cap, best_bytes;
};
__host__ __device__ inline MazeLayout maze_layout(int Z, int Y, int X, int row_stride) {
    MazeLayout L;
    const int cells = Z * Y * X;
    int bb = 2 * cells;
    if (bb < row_stride) bb = row_stride;      // the raw grid is staged here first
#if PCGRL_UF_RUNS32
    if (bb < Z * Y * 32) bb = Z * Y * 32;      // ... and the region count's 32-bit run parents live here last
#endif
    bb = (bb + 15) / 16 * 16;
    L.best_bytes = bb;
    L.order_cap = (Z / 2 + 1) * Y * X;
    int o = 0;
    L.best = o;  o += bb;
    L.q_cl = o;  o += 4 * MAZE_QCAP;
    L.nj = -1;   // global scratch, see make_ctx
    L.order = -1;   // global scratch too: appended by lane 0, re-read a handful of times per grid
    L.q_nj = o;  o += 2 * MAZE_QCAP;
    L.q_par = o; o += 2 * MAZE_QCAP;   // parent cell of each queued entry (only filled by search<true>)
    L.col = o;   o += 2 * Y * X;
    L.row = o;   o += 2 * Z * Y;
    L.total = (o + 15) / 16 * 16;
    return L;
}

struct Maze3DProb {
    static constexpr int K = 3;   // regions, path-length, n_jump

    struct Ctx {
        int Z, Y, X, cells, R;
        uint32_t magic_xy, magic_x;
        uint16_t *best, *nj, *order, *par, *q_nj, *q_par, *col, *row;
        uint32_t* q_cl;
        int best_bytes, order_cap;
        int32_t* status;
    };

    __device__ static Ctx make_ctx(const KParams& p, uint8_t* ws, int global_warp) {
        return make_ctx_dims(p, ws, global_warp, p.d0, p.d1, p.d2);
    }
    // (Z, Y, X): the dimensions the search runs on -- the map's own, or the bordered map's for the holey problems
    __device__ static Ctx make_ctx_dims(const KParams& p, uint8_t* ws, int global_warp, int Z, int Y, int X) {
        Ctx c;
        c.Z = Z; c.Y = Y; c.X = X;
        c.cells = Z * Y * X;
        c.R = c.Z * c.Y;
        c.magic_xy = div_magic(c.X * c.Y);
        c.magic_x = div_magic(c.X);
        const MazeLayout L = maze_layout(c.Z, c.Y, c.X, p.row_stride);
        c.best = (uint16_t*)(ws + L.best);
        c.q_cl = (uint32_t*)(ws + L.q_cl);
        c.nj = (uint16_t*)((uint8_t*)p.scratch + (size_t)global_warp * MAZE_SLICE);
        c.order = (uint16_t*)((uint8_t*)p.scratch + (size_t)global_warp * MAZE_SLICE + MAZE_NJ_SLICE);
        c.par = (uint16_t*)((uint8_t*)p.scratch + (size_t)global_warp * MAZE_SLICE + MAZE_NJ_SLICE + MAZE_ORDER_SLICE);
        c.q_nj = (uint16_t*)(ws + L.q_nj);
        c.q_par = (uint16_t*)(ws + L.q_par);
        c.col = (uint16_t*)(ws + L.col);
        c.row = (uint16_t*)(ws + L.row);
        c.best_bytes = L.best_bytes;
        c.order_cap = L.order_cap;
        c.status = p.status;
        return c;
    }

    // ---- helper_3D._passable for one direction: foothold reached from (x, y, z) going (dx, dy) ----------
    // Returns false if the direction offers no move; else the foothold cell index, the number of path
    // entries the move appends (1 walk, 2 stairs / level jump, 3 jump up / down) and whether it is a jump.
    // Branch-free: with w = the neighbour column's AIR bits z-2 .. z+3 (bits below 0 and from Z up read as "not
    // AIR", which is what the reference's explicit range tests amount to) the four kinds of move are mutually
    // exclusive bit patterns --
    //   walk       !w1  w2  w3                      (:229-237)      step down  !w0  w1  w2  w3   (:243-252)
    //   step up    !w2  w3  w4  and head-room c0[z+2] (:259-266)    jump        w0..w4 all AIR, c0[z+2], landing column in range (:279-288)
    // and so are the three landings of a jump on the column two cells away (u = its bits z-2 .. z+3):
    //   level !u1 u2 u3 u4 (:289-296)   up !u2 u3 u4 u5 (:297-304)   down !u0 u1 u2 u3 (:305-312)
    // so the four lanes of a pop run one straight instruction stream instead of a cascade of divergent tests.
    __device__ static __forceinline__ bool move(const Ctx& c, int x, int y, int z, int dx, int dy, int& ncell,
                                                int& cost, int& jump) {
        const int nx = x + dx, ny = y + dy;
        if (nx < 0 || ny < 0 || nx >= c.X || ny >= c.Y) return false;                    // :225
        const uint32_t cn = c.col[ny * c.X + nx], c0 = c.col[y * c.X + x];
        const uint32_t w = ((cn << 2) >> z) & 0x3Fu;
        const uint32_t h2 = (c0 >> (z + 2)) & 1u;
        const uint32_t w0 = w & 1u, w1 = (w >> 1) & 1u, w2 = (w >> 2) & 1u, w3 = (w >> 3) & 1u, w4 = (w >> 4) & 1u;
        const uint32_t walk = (w1 ^ 1u) & w2 & w3;
        const uint32_t down = (w0 ^ 1u) & w1 & w2 & w3;
        const uint32_t up = (w2 ^ 1u) & w3 & w4 & h2;
        const int jx = nx + dx, jy = ny + dy;
        const bool j_in = jx >= 0 && jy >= 0 && jx < c.X && jy < c.Y;
        const uint32_t jp = ((w & 0x1Fu) == 0x1Fu) & h2 & (uint32_t)j_in;
        const uint32_t cj = j_in ? c.col[jy * c.X + jx] : 0u;
        const uint32_t u = ((cj << 2) >> z) & 0x3Fu;
        const uint32_t u0 = u & 1u, u1 = (u >> 1) & 1u, u2 = (u >> 2) & 1u, u3 = (u >> 3) & 1u, u4 = (u >> 4) & 1u,
                       u5 = (u >> 5) & 1u;
        const uint32_t j_level = jp & (u1 ^ 1u) & u2 & u3 & u4;
        const uint32_t j_up = jp & (u2 ^ 1u) & u3 & u4 & u5;
        const uint32_t j_down = jp & (u0 ^ 1u) & u1 & u2 & u3;
        const uint32_t landed = j_level | j_up | j_down;
        jump = (int)jp;
        const int nz = z + (int)(up | j_up) - (int)(down | j_down);
        cost = (int)(walk + 2u * (down | up | j_level) + 3u * (j_up | j_down));
        ncell = landed ? (nz * c.Y + jy) * c.X + jx : (nz * c.Y + ny) * c.X + nx;
        return (walk | down | up | landed) != 0u;
    }

    // ---- helper_3D.run_dijkstra from `start` (warp-uniform FIFO; lanes 0..3 evaluate the 4 directions) -----
    // Leaves order[0..n) = cells in first-recording order, best[c] & 0x7FFF = len(paths[c]), nj[c] = jumps[c].
    // PARENTS: also leave par[c] = the cell the recorded path reaches c from (par[start] = start).  The start tile
    // needs head-room like every recorded tile (helper_3D.py:443-445), else nothing is recorded.
    template <bool PARENTS = false>
    __device__ static int search(Ctx& c, int start, int lane, bool& overflow) {
        const int XY = c.X * c.Y;
        unsigned head = 0, tail = 1;
        int n = 0;
        {
            const int sz = div_by(start, c.magic_xy), srem = start - sz * XY;
            if (sz + 1 >= c.Z || !((c.col[srem] >> (sz + 1)) & 1u)) return 0;
        }
        if (lane == 0) {
            c.q_cl[0] = (uint32_t)start | (1u << 12);
            c.q_nj[0] = 0;
            if (PARENTS) c.q_par[0] = (uint16_t)start;
            c.best[start] = 1;
        }
        __syncwarp();
        const int dx = lane == 0 ? 1 : (lane == 2 ? -1 : 0);       // helper_3D.py:220 direction order
        const int dy = lane == 1 ? 1 : (lane == 3 ? -1 : 0);
        while (head != tail) {
            const uint32_t cl = c.q_cl[head & (MAZE_QCAP - 1)];
            const int nj_e = c.q_nj[head & (MAZE_QCAP - 1)];
            ++head;
            const int cell = cl & 0xFFF, ln = cl >> 12;
            const int z = div_by(cell, c.magic_xy), rem = cell - z * XY, y = div_by(rem, c.magic_x), x = rem - y * c.X;
            const uint16_t b = c.best[cell];
            const bool first = !(b & 0x8000u);
            if (lane == 0) {
                if (first) {
                    if (n < c.order_cap) c.order[n] = (uint16_t)cell;
                    c.best[cell] = b | 0x8000u;
                }
                c.nj[cell] = (uint16_t)nj_e;
                if (PARENTS) c.par[cell] = c.q_par[(head - 1) & (MAZE_QCAP - 1)];
            }
            n += first;
            bool valid = false;
            int ncell = 0, cost = 0, jump = 0;
            if (lane < 4) {
                valid = move(c, x, y, z, dx, dy, ncell, cost, jump);
                if (valid) {
                    const int nb = c.best[ncell] & 0x7FFF;
                    if (nb != 0 && nb <= ln + cost) valid = false;   // would be skipped at pop (:437-440)
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const unsigned pos = (tail + __popc(m & ((1u << lane) - 1u))) & (MAZE_QCAP - 1);
                c.q_cl[pos] = (uint32_t)ncell | ((uint32_t)(ln + cost) << 12);
                c.q_nj[pos] = (uint16_t)(nj_e + jump);
                if (PARENTS) c.q_par[pos] = (uint16_t)cell;
                c.best[ncell] = (uint16_t)((c.best[ncell] & 0x8000u) | (ln + cost));
            }
            tail += __popc(m);
            if (tail - head > (unsigned)MAZE_QCAP || n > c.order_cap) {
                overflow = true;
                break;
            }
            __syncwarp();
        }
        __syncwarp();
        return n < c.order_cap ? n : c.order_cap;
    }

    // ---- the same search, up to 32 queue entries per round ------------------------------------------------------------
    // The reference pushes EVERY move and decides at pop time (`0 < len(old) <= len(path)`: skip).  Here a round takes
    // the next B <= 32 entries; entry i is effective iff it is shorter than what is recorded for its cell and than
    // every earlier entry of the round for the same cell (__match_any on the cell, then a loop over the few lanes that
    // share one) -- exactly the sequential pop test.  The first effective entry of an unrecorded cell appends it to the
    // recording order, the LAST effective entry of a cell leaves length / jumps / parent.  Then every effective lane
    // evaluates its four moves and appends the candidates lane by lane, direction by direction (the sequential push
    // order), dropping only those that are certain to be skipped when popped: the target already holds a recording
    // that is not longer (recordings only get shorter; this round's are all in place before any of its candidates
    // is popped).  No "best pushed" filter as in search(): it depends on the order of the pushes inside a round.
    // On a queue overflow the cells are cleared and the search falls back to search().
    template <bool PARENTS = false>
    __device__ static int search_batch(Ctx& c, int start, int lane, bool& overflow) {
        const int XY = c.X * c.Y;
        unsigned head = 0, tail = 1;
        int n = 0;
        {
            const int sz = div_by(start, c.magic_xy), srem = start - sz * XY;
            if (sz + 1 >= c.Z || !((c.col[srem] >> (sz + 1)) & 1u)) return 0;
        }
        if (lane == 0) {
            c.q_cl[0] = (uint32_t)start | (1u << 12);
            c.q_nj[0] = 0;
            if (PARENTS) c.q_par[0] = (uint16_t)start;
        }
        __syncwarp();
        const unsigned lt = (1u << lane) - 1u;
        bool spilled = false;
        for (;;) {
            const unsigned avail = tail - head;
            if (!avail) break;
            const int B = (int)min(32u, avail);
            const bool mine = lane < B;
            uint32_t cl = 0;
            int nj_e = 0, par_e = 0;
            if (mine) {
                const unsigned slot = (head + lane) & (MAZE_QCAP - 1);
                cl = c.q_cl[slot];
                nj_e = c.q_nj[slot];
                if (PARENTS) par_e = c.q_par[slot];
            }
            head += B;
            const int cell = cl & 0xFFF, ln = cl >> 12;
            const uint16_t rb = mine ? c.best[cell] : (uint16_t)0;
            const unsigned mm = __ballot_sync(0xffffffffu, mine);
            unsigned grp = 0;
            if (mine) grp = __match_any_sync(mm, cell);
            int minlow = 0x7FFFFFFF;
            const unsigned dups = __ballot_sync(0xffffffffu, mine && (grp & (grp - 1u)) != 0u);
            for (unsigned m = dups; m; m &= m - 1u) {       // rare: several entries of one cell in the same round
                const int j = __ffs(m) - 1;
                const int lj = __shfl_sync(0xffffffffu, ln, j), cj = __shfl_sync(0xffffffffu, cell, j);
                if (mine && j < lane && cj == cell) minlow = min(minlow, lj);
            }
            const int rec_len = (rb & 0x8000u) ? (int)(rb & 0x7FFF) : 0x7FFFFFFF;
            const bool eff = mine && ln < min(rec_len, minlow);                       // helper_3D.py:437-440
            const bool first = eff && !(rb & 0x8000u) && (grp & lt) == 0u;
            const unsigned fm = __ballot_sync(0xffffffffu, first);
            const unsigned em = __ballot_sync(0xffffffffu, eff);
            if (first) {
                const int at = n + __popc(fm & lt);
                if (at < c.order_cap) c.order[at] = (uint16_t)cell;
            }
            n += __popc(fm);
            if (eff && (grp & em & ~lt & ~(1u << lane)) == 0u) {     // the last effective entry of this cell in the round
                c.best[cell] = (uint16_t)(0x8000u | ln);
                c.nj[cell] = (uint16_t)nj_e;
                if (PARENTS) c.par[cell] = (uint16_t)par_e;
            }
            __syncwarp();
            // moves of the effective entries (helper_3D.py:455-484), direction order of :220
            uint32_t cand[4];
            uint16_t cnj[4];
            unsigned ok = 0;
            if (eff) {
                const int z = div_by(cell, c.magic_xy), rem = cell - z * XY, y = div_by(rem, c.magic_x), x = rem - y * c.X;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const int dx = d == 0 ? 1 : (d == 2 ? -1 : 0), dy = d == 1 ? 1 : (d == 3 ? -1 : 0);
                    int ncell = 0, cost = 0, jump = 0;
                    if (move(c, x, y, z, dx, dy, ncell, cost, jump)) {
                        const uint16_t r = c.best[ncell];
                        if (!(r & 0x8000u) || (int)(r & 0x7FFF) > ln + cost) {
                            ok |= 1u << d;
                            cand[d] = (uint32_t)ncell | ((uint32_t)(ln + cost) << 12);
                            cnj[d] = (uint16_t)(nj_e + jump);
                        }
                    }
                }
            }
            const int cnt = __popc(ok);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (tail + total - head > (unsigned)MAZE_QCAP || n > c.order_cap) {
                spilled = true;
                break;
            }
            unsigned at = tail + incl - cnt;
#pragma unroll
            for (int d = 0; d < 4; ++d)
                if (ok & (1u << d)) {
                    const unsigned slot = at & (MAZE_QCAP - 1);
                    c.q_cl[slot] = cand[d];
                    c.q_nj[slot] = cnj[d];
                    if (PARENTS) c.q_par[slot] = (uint16_t)cell;
                    ++at;
                }
            tail += total;
            __syncwarp();
        }
        __syncwarp();
        if (spilled) {       // more candidates in flight than the ring holds: start over with the filtered, one-pop loop
            clear_search(c, n < c.order_cap ? n : c.order_cap, lane);
            return search<PARENTS>(c, start, lane, overflow);
        }
        return n;
    }

    // first maximum of len(paths[.]) in insertion order (np.argmax, helper_3D.py:538-541)
    __device__ static __forceinline__ void far_tile(const Ctx& c, int n, int lane, int& cell, int& dist) {
        uint32_t key = 0;
        for (int i = lane; i < n; i += 32) {
            const uint32_t k = ((uint32_t)(c.best[c.order[i]] & 0x7FFF) << 16) | (uint32_t)(0xFFFF - i);
            key = max(key, k);
        }
        key = __reduce_max_sync(0xffffffffu, key);
        cell = c.order[0xFFFF - (key & 0xFFFF)];
        dist = key >> 16;
    }

    __device__ static __forceinline__ void clear_search(Ctx& c, int n, int lane) {
        for (int i = lane; i < n; i += 32) c.best[c.order[i]] = 0;
        __syncwarp();
    }

    __device__ static void stats(const KParams& p, Ctx& c, const int8_t* grid, int lane, int32_t* out) {
        const int Z = c.Z, Y = c.Y, X = c.X, R = c.R;
        // ---- stage the grid (coalesced 128-bit loads), build the row / column AIR masks --------------------
        {
            uint4* stage = (uint4*)c.best;
            const uint4* src = (const uint4*)grid;
            for (int i = lane; i < p.row_stride / 16; i += 32) stage[i] = src[i];
            __syncwarp();
            const uint8_t* g = (const uint8_t*)c.best;
            for (int r = lane; r < R; r += 32) {
                uint32_t m = 0;
                for (int x = 0; x < X; ++x) m |= (uint32_t)(g[r * X + x] == 0) << x;
                c.row[r] = (uint16_t)m;
            }
            for (int q = lane; q < Y * X; q += 32) {
                uint32_t m = 0;
                for (int z = 0; z < Z; ++z) m |= (uint32_t)(g[z * Y * X + q] == 0) << z;
                c.col[q] = (uint16_t)m;
            }
            __syncwarp();
            uint4* bz = (uint4*)c.best;
            for (int i = lane; i < c.best_bytes / 16; i += 32) bz[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
        }

        // ---- calc_longest_path (helper_3D.py:503-563) ---------------------------------------------------------
        int final_value = 0, last_far = -1;
        uint32_t planes = 0;
        bool overflow = false, index_error = false;
        for (int z = 1; z + 1 < Z; ++z) {
            if ((planes >> z) & 1u) continue;                                            // :515
            // start tiles of this plane: AIR with head-room (:520) standing on a solid tile (:525)
            uint32_t cand = 0;
            if (lane < Y) cand = c.row[z * Y + lane] & c.row[(z + 1) * Y + lane] & ~(uint32_t)c.row[(z - 1) * Y + lane];
            const unsigned rows_with = __ballot_sync(0xffffffffu, cand != 0);
            if (!rows_with) continue;
            const int y = __ffs(rows_with) - 1;
            const int x = __ffs(__shfl_sync(0xffffffffu, cand, y)) - 1;
            const int start = (z * Y + y) * X + x;

            int n = PCGRL_MAZE_SEARCH(c, start, lane, overflow);                         // :529
            uint32_t mark = 0;
            for (int i = lane; i < n; i += 32) {                                          // :531
                const int cell = c.order[i];
                const int cz = div_by(cell, c.magic_xy), rem = cell - cz * X * Y, cy = div_by(rem, c.magic_x),
                          cx = rem - cy * X;
                mark |= (1u << cx) | (1u << cy) | (1u << cz);
            }
            mark = __reduce_or_sync(0xffffffffu, mark);
            if (mark >> Z) index_error = true;
            planes |= mark;
            int far, dist;
            far_tile(c, n, lane, far, dist);                                             // :538-541
            clear_search(c, n, lane);
            n = PCGRL_MAZE_SEARCH(c, far, lane, overflow);                               // :548
            far_tile(c, n, lane, far, dist);                                             // :549-552
            last_far = far;                                                              // :553 last component wins
            if (dist > final_value) final_value = dist;                                  // :558
            clear_search(c, n, lane);
            if (overflow) break;
        }

        // ---- calc_num_regions (helper_3D.py:396-406): 6-neighbour AIR components by union-find over runs ---
#if PCGRL_UF_RUNS32
        const int regions = count_regions_runs32(c.row, Z, Y, X, (uint32_t*)c.best, lane);
#else
        const int regions = count_regions_rows(c.row, Z, Y, X, c.best, lane);
#endif

        if (lane == 0) {
            // jumps[far] of the last processed component: nj[] still holds that search's recordings (lane 0
            // wrote them, so its own load sees them)
            out[0] = regions;
            out[1] = final_value;
            out[2] = last_far >= 0 ? c.nj[last_far] : 0;
            if (p.status && (index_error || overflow)) atomicOr(p.status, (index_error ? 2 : 0) | (overflow ? 4 : 0));
        }
    }
};

}  // namespace pcgrl
