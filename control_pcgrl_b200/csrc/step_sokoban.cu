// sokoban: fused env step with warp-per-grid get_stats, including the reference's solver.
//
// Reference path replaced (paths relative to /root/reference/control_pcgrl/envs/probs/sokoban/):
//   sokoban_prob.py:160-180 + sokoban_ctrl_prob.py:58-65   get_stats (counts, regions, dist-win, sol-length, ratio)
//   sokoban_prob.py:99-148                                  _run_game: BFS -> A*(1) -> A*(0.5) -> A*(0), 10 000 iterations each
//   sokoban/engine.py:56-74 BFSAgent, :96-119 AStarAgent, :121-363 State (update, deadlocks, heuristic, key)
//
// State = player position + crate positions in list order (the reference's visited key lists the crates in
// order, engine.py:329-335), one byte each (position = y * BW + x in the '#'-framed level), packed in a uint4.
// A node = state + depth + heuristic; parents are not kept (only len(solution) = depth is observable).
// Node pools, the A* heap and the visited hash table live in a per-warp slice of the global scratch; the
// level itself (solid / deadlock bit masks, target list) lives in shared memory.
//
// Exactness notes:
//   * iterations count every pop, also of already visited keys; win is tested on pop, before the visited test;
//   * a state wins iff its heuristic is 0 (#crates == #targets > 0 is a precondition of running the solver and
//     the greedy crate->target matching is then a perfect matching at distance 0), so the stored heuristic
//     doubles as the win flag;
//   * queue.PriorityQueue == heapq on Node.__lt__ = h + balance * depth: the binary heap below performs CPython's
//     heappush / heappop moves (_siftdown / _siftup) with `<` only; keys are the exact integers 2h + (2 balance) d;
//   * if the BFS empties its queue without a win, each A* would pop exactly the same multiset of nodes
//     (1 + sum of children over all reachable states) and end with best heuristic = min over reachable states,
//     so dist-win is taken from the BFS directly; the A* passes only run when the BFS hit the iteration cap.
#include "step_search.cuh"

namespace pcgrl {

constexpr int SOK_POWER = 10000;                 // sokoban_prob.py:40 _solver_power
constexpr int SOK_NODES = 4 * SOK_POWER + 8;     // root + <= 4 children per iteration
constexpr int SOK_TABLE = 32768;                 // visited hash slots (<= 10 000 insertions per search)
constexpr int SOK_MAX_CRATES = 15;
#ifndef PCGRL_SOK_CTAS_PER_SM
#define PCGRL_SOK_CTAS_PER_SM 2   // A/B on B200: 2 / 3 / 4 CTAs per SM change nothing (bounded by the longest solve)
#endif
constexpr int SOK_MAX_CTAS_PER_SM = PCGRL_SOK_CTAS_PER_SM;
constexpr int SOK_MAX_CTAS = 160 * SOK_MAX_CTAS_PER_SM;   // sizes the global scratch (>= 148 SMs x CTAs per SM)

struct SokScratch {
    // byte offsets inside one warp's slice
    static constexpr size_t header = 0;                                    // u32 generation counter (+pad)
    static constexpr size_t state = 16;                                    // uint4[SOK_NODES]
    static constexpr size_t meta = state + sizeof(uint4) * SOK_NODES;      // u32[SOK_NODES]: depth | h << 16
    static constexpr size_t heap = meta + 4 * (size_t)SOK_NODES;           // u32[SOK_NODES]: idx | key << 16
    static constexpr size_t table = heap + 4 * (size_t)SOK_NODES;          // u32[SOK_TABLE]: idx | gen << 16
    static constexpr size_t total = (table + 4 * (size_t)SOK_TABLE + 255) / 256 * 256;
};

int64_t sokoban_scratch_bytes() { return (int64_t)SokScratch::total * SOK_MAX_CTAS * SEARCH_WARPS; }

struct SokobanProb {
    static constexpr int K = 7;   // player crate target regions dist-win sol-length ratio

    struct Ctx {
        int H, W, BW, BH;
        // shared memory
        uint8_t* stage;      // raw grid bytes
        uint32_t* solid;     // [8] bit p = '#'
        uint32_t* dead;      // [8] bit p = static deadlock cell
        uint8_t* tpos;       // [16] target positions in scan order
        uint8_t* cpos;       // [256] scratch list (crates, then corners)
        uint16_t* rows;      // [16] passable rows, then [256] union-find parents
        // global scratch slice
        uint32_t* hdr;
        uint4* state;
        uint32_t *meta, *heap, *table;
        int nt;              // #targets == #crates while solving
        uint32_t magic_bw;
    };

    __host__ __device__ static int smem_bytes(int row_stride) {
        return (row_stride + 15) / 16 * 16 + 32 + 32 + 16 + 256 + 2 * (16 + 256);
    }

    __device__ static Ctx make_ctx(const KParams& p, uint8_t* ws, int global_warp) {
        Ctx c;
        c.H = p.d0; c.W = p.d1;
        c.BW = c.W + 2; c.BH = c.H + 2;
        int o = 0;
        c.stage = ws + o;             o += (p.row_stride + 15) / 16 * 16;
        c.solid = (uint32_t*)(ws + o); o += 32;
        c.dead = (uint32_t*)(ws + o);  o += 32;
        c.tpos = ws + o;              o += 16;
        c.cpos = ws + o;              o += 256;
        c.rows = (uint16_t*)(ws + o);
        uint8_t* g = (uint8_t*)p.scratch + (size_t)global_warp * SokScratch::total;
        c.hdr = (uint32_t*)(g + SokScratch::header);
        c.state = (uint4*)(g + SokScratch::state);
        c.meta = (uint32_t*)(g + SokScratch::meta);
        c.heap = (uint32_t*)(g + SokScratch::heap);
        c.table = (uint32_t*)(g + SokScratch::table);
        c.nt = 0;
        c.magic_bw = div_magic(c.BW);
        return c;
    }

    // ---- state helpers (registers) -------------------------------------------------------------------------
    __device__ static __forceinline__ uint32_t get_byte(const uint4& s, int i) {
        const uint32_t w = i < 4 ? s.x : (i < 8 ? s.y : (i < 12 ? s.z : s.w));
        return (w >> ((i & 3) * 8)) & 0xFFu;
    }
    __device__ static __forceinline__ void set_byte(uint4& s, int i, uint32_t v) {
        const uint32_t sh = (i & 3) * 8, m = ~(0xFFu << sh), b = v << sh;
        if (i < 4) s.x = (s.x & m) | b;
        else if (i < 8) s.y = (s.y & m) | b;
        else if (i < 12) s.z = (s.z & m) | b;
        else s.w = (s.w & m) | b;
    }
    // index (1..15) of the state byte holding crate position `pos`, or 0 (engine.py:257-261 checkCrateLocation)
    __device__ static __forceinline__ int crate_byte(const uint4& s, uint32_t pos) {
        const uint32_t pat = pos * 0x01010101u;
        uint32_t m = __vcmpeq4(s.x, pat) & 0xFFFFFF00u;    // byte 0 is the player
        if (m) return (__ffs(m) - 1) >> 3;
        m = __vcmpeq4(s.y, pat);
        if (m) return 4 + ((__ffs(m) - 1) >> 3);
        m = __vcmpeq4(s.z, pat);
        if (m) return 8 + ((__ffs(m) - 1) >> 3);
        m = __vcmpeq4(s.w, pat);
        if (m) return 12 + ((__ffs(m) - 1) >> 3);
        return 0;
    }
    __device__ static __forceinline__ bool bit(const uint32_t* m, uint32_t p) { return (m[p >> 5] >> (p & 31)) & 1u; }

    // engine.py:279-293 getHeuristic: greedy crate -> nearest remaining target (first minimum, strict '>')
    __device__ static int heuristic(const Ctx& c, const uint4& s) {
        uint32_t remaining = (1u << c.nt) - 1u;
        int total = 0;
        for (int i = 0; i < c.nt; ++i) {
            const int cp = get_byte(s, 1 + i), cy = div_by(cp, c.magic_bw), cx = cp - cy * c.BW;
            int best = c.BW + c.BH, bi = -1, bd = 0;
            for (uint32_t r = remaining; r; r &= r - 1) {
                const int j = __ffs(r) - 1;
                const int tp = c.tpos[j], ty = div_by(tp, c.magic_bw);
                const int d = abs(cx - (tp - ty * c.BW)) + abs(cy - ty);
                if (bi < 0) {          // bestMatch defaults to the first remaining target
                    bi = j;
                    bd = d;
                }
                if (best > d) {
                    best = d;
                    bi = j;
                    bd = d;
                }
            }
            total += bd;
            remaining &= ~(1u << bi);
        }
        return total;
    }

    __device__ static __forceinline__ uint32_t hash_state(const uint4& s) {
        uint32_t h = s.x * 0x9E3779B1u;
        h = (h ^ (h >> 15)) + s.y * 0x85EBCA77u;
        h = (h ^ (h >> 13)) + s.z * 0xC2B2AE3Du;
        h = (h ^ (h >> 16)) + s.w * 0x27D4EB2Fu;
        h ^= h >> 15;
        h *= 0x2C1B3C6Du;
        h ^= h >> 12;
        return h;
    }

    // visited set (lane 0 only): true if `s` was already there, else inserts node `idx`
    __device__ static bool visited_test_and_set(const Ctx& c, const uint4& s, int idx, uint32_t gen) {
        uint32_t i = hash_state(s) & (SOK_TABLE - 1);
        for (;;) {
            const uint32_t slot = c.table[i];
            if ((slot >> 16) != gen) {
                c.table[i] = (uint32_t)idx | (gen << 16);
                return false;
            }
            const uint4 o = c.state[slot & 0xFFFFu];
            if (o.x == s.x && o.y == s.y && o.z == s.z && o.w == s.w) return true;
            i = (i + 1) & (SOK_TABLE - 1);
        }
    }

    // ---- CPython heapq on (key, idx) entries, lane 0 only --------------------------------------------------
    __device__ static __forceinline__ void sift_down(uint32_t* heap, int start, int pos) {
        const uint32_t item = heap[pos];
        while (pos > start) {
            const int parent = (pos - 1) >> 1;
            const uint32_t pe = heap[parent];
            if ((item >> 16) < (pe >> 16)) {
                heap[pos] = pe;
                pos = parent;
                continue;
            }
            break;
        }
        heap[pos] = item;
    }
    __device__ static __forceinline__ void heap_push(uint32_t* heap, int& n, uint32_t entry) {
        heap[n] = entry;
        sift_down(heap, 0, n);
        ++n;
    }
    __device__ static __forceinline__ uint32_t heap_pop(uint32_t* heap, int& n) {
        const uint32_t last = heap[--n];
        if (n == 0) return last;
        const uint32_t ret = heap[0];
        int pos = 0, child = 1;
        while (child < n) {
            const int right = child + 1;
            uint32_t ce = heap[child];
            if (right < n) {
                const uint32_t re = heap[right];
                if (!((ce >> 16) < (re >> 16))) {
                    child = right;
                    ce = re;
                }
            }
            heap[pos] = ce;
            pos = child;
            child = 2 * pos + 1;
        }
        heap[pos] = last;
        sift_down(heap, 0, pos);
        return ret;
    }

    // ---- one search: mode 0 = BFSAgent, 1..3 = AStarAgent with 2*balance = 2, 1, 0 ---------------------------
    // Returns true on win (depth in `out_depth`); else best heuristic in `out_h`; `exhausted` = queue ran empty.
    __device__ static bool search(Ctx& c, const uint4 root, int root_h, int mode, int lane, int& out_depth, int& out_h,
                                  bool& exhausted) {
        const int bal2 = mode == 1 ? 2 : (mode == 2 ? 1 : 0);
        // new generation of the visited table (cleared only when the 16-bit tag wraps)
        uint32_t gen = 0;
        if (lane == 0) gen = *c.hdr + 1;
        gen = __shfl_sync(0xffffffffu, gen, 0);
        if (gen > 0xFFFFu) {
            for (int i = lane; i < SOK_TABLE; i += 32) c.table[i] = 0;
            gen = 1;
        }
        if (lane == 0) {
            *c.hdr = gen;
            c.state[0] = root;
            c.meta[0] = (uint32_t)root_h << 16;
            c.heap[0] = (uint32_t)(2 * root_h) << 16;
        }
        __syncwarp();
        int n_nodes = 1, head = 0, hn = 1;
        int best_h = 0x7FFFFFFF, best_d = 0, iters = 0;
        const int delta = lane == 0 ? -1 : (lane == 1 ? 1 : (lane == 2 ? -c.BW : c.BW));   // engine.py:3 L R U D
        for (;;) {
            const bool empty = mode == 0 ? head >= n_nodes : hn == 0;
            if (iters >= SOK_POWER || empty) {
                exhausted = empty;
                break;
            }
            ++iters;
            // pop + visited test on lane 0, broadcast
            int cur = 0, seen = 0;
            if (lane == 0) {
                cur = mode == 0 ? head : (int)(heap_pop(c.heap, hn) & 0xFFFFu);
            }
            cur = __shfl_sync(0xffffffffu, cur, 0);
            if (mode == 0) ++head;
            const uint4 s = c.state[cur];
            const uint32_t m = c.meta[cur];
            const int h = m >> 16, d = m & 0xFFFFu;
            if (h == 0) {                                     // checkWin (engine.py:64 / :108)
                out_depth = d;
                return true;
            }
            if (lane == 0) seen = visited_test_and_set(c, s, cur, gen);
            seen = __shfl_sync(0xffffffffu, seen, 0);
            if (mode != 0) hn = __shfl_sync(0xffffffffu, hn, 0);
            if (seen) continue;
            if (h < best_h || (h == best_h && d < best_d)) {  // engine.py:66-70
                best_h = h;
                best_d = d;
            }
            // children (Node.getChildren, engine.py:14-24): lane = direction
            bool valid = false;
            uint4 cs = s;
            int ch = 0;
            if (lane < 4) {
                const uint32_t pp = s.x & 0xFFu, np = pp + delta;
                if (!bit(c.solid, np)) {
                    const int cb = crate_byte(s, np);
                    if (cb == 0) {                            // plain move
                        cs.x = (s.x & 0xFFFFFF00u) | np;
                        valid = true;
                        ch = h;
                    } else {                                  // push crate `cb - 1`
                        const uint32_t cp = np + delta;
                        if (!bit(c.solid, cp) && crate_byte(s, cp) == 0) {
                            set_byte(cs, cb, cp);
                            cs.x = (cs.x & 0xFFFFFF00u) | np;
                            bool deadlocked = false;          // checkDeadlock: any crate on a deadlock cell
                            for (int i = 0; i < c.nt; ++i) deadlocked |= bit(c.dead, get_byte(cs, 1 + i));
                            if (!deadlocked) {
                                valid = true;
                                ch = heuristic(c, cs);
                            }
                        }
                    }
                }
            }
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const int idx = n_nodes + __popc(vm & ((1u << lane) - 1u));
                c.state[idx] = cs;
                c.meta[idx] = ((uint32_t)ch << 16) | (uint32_t)(d + 1);
            }
            const int nc = __popc(vm);
            if (mode != 0) {
                __syncwarp();
                if (lane == 0) {
                    for (int k = 0; k < nc; ++k) {
                        const int idx = n_nodes + k;
                        const uint32_t mm = c.meta[idx];
                        const uint32_t key = 2u * (mm >> 16) + (uint32_t)bal2 * (mm & 0xFFFFu);
                        heap_push(c.heap, hn, (uint32_t)idx | (key << 16));
                    }
                }
                hn = __shfl_sync(0xffffffffu, hn, 0);
            }
            n_nodes += nc;
            __syncwarp();
        }
        out_h = best_h;
        return false;
    }

    __device__ static void stats(const KParams& p, Ctx& c, const int8_t* grid, int lane, int32_t* out) {
        const int H = c.H, W = c.W, BW = c.BW, BH = c.BH;
        // ---- stage the grid; counts; passable rows ------------------------------------------------------------
        {
            uint4* st = (uint4*)c.stage;
            const uint4* src = (const uint4*)grid;
            for (int i = lane; i < p.row_stride / 16; i += 32) st[i] = src[i];
            __syncwarp();
        }
        int n_player = 0, n_crate = 0, n_target = 0;
        for (int i = lane; i < p.cells; i += 32) {
            const int t = c.stage[i];
            n_player += t == 2;
            n_crate += t == 3;
            n_target += t == 4;
        }
        n_player = __reduce_add_sync(0xffffffffu, n_player);
        n_crate = __reduce_add_sync(0xffffffffu, n_crate);
        n_target = __reduce_add_sync(0xffffffffu, n_target);
        uint16_t* rows = c.rows;
        for (int y = lane; y < H; y += 32) {
            uint32_t m = 0;
            for (int x = 0; x < W; ++x) m |= (uint32_t)(c.stage[y * W + x] != 1) << x;
            rows[y] = (uint16_t)m;
        }
        __syncwarp();
        const int regions = count_regions_rows(rows, 1, H, W, rows + 16, lane);

        int dist_win = W * H * (W + H), sol_len = 0;                    // sokoban_prob.py:171
        const bool run = n_player == 1 && n_crate == n_target && n_crate > 0 && regions == 1;   // :174-179
        if (run && n_crate > SOK_MAX_CRATES) {
            if (lane == 0 && p.status) atomicOr(p.status, 8);           // more crates than a packed state holds
        } else if (run) {
            // ---- the '#'-framed level (sokoban_prob.py:99-123, engine.py:137-186) -----------------------------
            uint4 root = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            int ncr = 0, ntg = 0;
            uint32_t player = 0;
            for (int w = 0; w < 8; ++w) {
                const int pidx = w * 32 + lane;
                int t = 1;
                if (pidx < BW * BH) {
                    const int y = div_by(pidx, c.magic_bw), x = pidx - y * BW;
                    if (x > 0 && y > 0 && x < BW - 1 && y < BH - 1) t = c.stage[(y - 1) * W + (x - 1)];
                }
                const unsigned sm = __ballot_sync(0xffffffffu, t == 1);
                const unsigned cm = __ballot_sync(0xffffffffu, t == 3);
                const unsigned tm = __ballot_sync(0xffffffffu, t == 4);
                const unsigned pm = __ballot_sync(0xffffffffu, t == 2);
                if (lane == 0) {
                    c.solid[w] = sm;
                    c.dead[w] = 0;
                }
                if (t == 3) c.cpos[ncr + __popc(cm & ((1u << lane) - 1u))] = (uint8_t)pidx;
                if (t == 4) c.tpos[ntg + __popc(tm & ((1u << lane) - 1u))] = (uint8_t)pidx;
                if (pm) player = w * 32 + (__ffs(pm) - 1);
                ncr += __popc(cm);
                ntg += __popc(tm);
            }
            __syncwarp();
            c.nt = ntg;
            root.x = (root.x & 0xFFFFFF00u) | player;
            for (int i = 0; i < ncr; ++i) set_byte(root, 1 + i, c.cpos[i]);
            __syncwarp();
            // ---- static deadlocks (engine.py:203-246): non-target corners, and wall-hugging runs between two
            //      corners of the same row / column --------------------------------------------------------------
            uint32_t tmask[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) tmask[w] = 0;
            for (int j = 0; j < ntg; ++j) tmask[c.tpos[j] >> 5] |= 1u << (c.tpos[j] & 31);
            for (int w = 0; w < 8; ++w) {
                const int pidx = w * 32 + lane;
                bool corner = false;
                if (pidx < BW * BH && !bit(c.solid, pidx)) {
                    const bool up = bit(c.solid, pidx - BW), dn = bit(c.solid, pidx + BW), lf = bit(c.solid, pidx - 1),
                               rt = bit(c.solid, pidx + 1);
                    corner = ((up && lf) || (up && rt) || (dn && lf) || (dn && rt)) && !((tmask[w] >> lane) & 1u);
                }
                const unsigned cm = __ballot_sync(0xffffffffu, corner);
                if (lane == 0) c.dead[w] = cm;
            }
            __syncwarp();
            uint32_t cornerm[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) cornerm[w] = c.dead[w];
            __syncwarp();
            for (int pidx = lane; pidx < BW * BH; pidx += 32) {
                if (!((cornerm[pidx >> 5] >> (pidx & 31)) & 1u)) continue;
                // walk right, then down, from this corner; every further corner reached closes a run
                for (int dir = 0; dir < 2; ++dir) {
                    const int step = dir == 0 ? 1 : BW;
                    const int sa = dir == 0 ? BW : 1;     // the two sides that must hold a wall
                    int q = pidx + step, last_corner = -1;
                    while (!bit(c.solid, q) && !((tmask[q >> 5] >> (q & 31)) & 1u) &&
                           (bit(c.solid, q - sa) || bit(c.solid, q + sa))) {
                        if ((cornerm[q >> 5] >> (q & 31)) & 1u) last_corner = q;
                        q += step;
                    }
                    for (int r = pidx + step; r < last_corner; r += step) atomicOr(&c.dead[r >> 5], 1u << (r & 31));
                }
            }
            __syncwarp();
            // ---- _run_game (sokoban_prob.py:124-148) -------------------------------------------------------------
            const int root_h = heuristic(c, root);
            int depth = 0, hbest = 0;
            bool exhausted = false;
            bool won = search(c, root, root_h, 0, lane, depth, hbest, exhausted);
            if (!won && !exhausted) {
                for (int mode = 1; mode <= 3 && !won; ++mode) won = search(c, root, root_h, mode, lane, depth, hbest, exhausted);
            }
            if (won) {
                dist_win = 0;
                sol_len = depth;
            } else {
                dist_win = hbest;
            }
        }
        if (lane == 0) {
            out[0] = n_player;
            out[1] = n_crate;
            out[2] = n_target;
            out[3] = regions;
            out[4] = dist_win;
            out[5] = sol_len;
            out[6] = abs(n_crate - n_target);
        }
    }
};

cudaError_t launch_sokoban(const KParams& p, cudaStream_t s, bool& supported) {
    supported = p.ndim == 2 && p.d0 <= 14 && p.d1 <= 14 && p.scratch != nullptr;
    if (!supported) return cudaSuccess;
    return launch_search<SokobanProb>(p, s, SokobanProb::smem_bytes(p.row_stride), SOK_MAX_CTAS_PER_SM, SOK_MAX_CTAS);
}

}  // namespace pcgrl
