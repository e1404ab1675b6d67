// smb (Super Mario Bros): fused env step with warp-per-grid get_stats, including the A* playthrough.
//
// Reference path replaced (paths relative to /root/reference/control_pcgrl/envs/):
//   probs/smb/smb_prob.py:132-154   get_stats (dist-floor, disjoint-tubes, enemies, empty, noise + play stats)
//   probs/smb/smb_prob.py:96-130    _run_game: level framing, A*(balance 1) then A*(balance 0), 10 000 iterations each
//   probs/smb/smb/engine.py:105-129 AStarAgent, :131-262 State (update physics, key, heuristic, win / lose)
//   helper.py:40-72 get_floor_dist, :74-121 get_type_grouping, :123-145 get_changes
//
// The level is the map with three extra columns on each side (smb_prob.py:99-113): rows above `H-3` are open
// and carry the exit marker in column W+4, row H-3 holds the player at x = 1 and a block at x = W+4, the last
// two rows are solid on both sides.  A node packs (x, y, airTime, last jump x, max jump gap) in one word and
// (jumps, depth) in another; jumps-dist (smb_prob.py:147-152) only needs the last jump x and the largest gap
// between consecutive jump x's, so the reference's jump_locs list is folded into those two numbers.
// Node pool and heap live in a per-warp slice of the global scratch; the solid bit map and the visited-key
// bit map (x, y, airTime) live in shared memory.
//
// Exactness: every pop counts as an iteration (also lose / already visited nodes); all
// four children are always pushed; the heap performs CPython's heappush / heappop moves with `<` on
// h + balance * depth only, because which of several equal-priority nodes is expanded first decides the
// jumps / jump positions that are reported.
#include "step_search.cuh"

namespace pcgrl {

constexpr int SMB_POWER = 10000;                 // smb_prob.py:21 _solver_power
constexpr int SMB_NODES = 4 * SMB_POWER + 8;
#ifndef PCGRL_SMB_WARPS
#define PCGRL_SMB_WARPS 8
#endif
constexpr int SMB_WARPS = PCGRL_SMB_WARPS;   // warps per CTA
#ifndef PCGRL_SMB_CTAS_PER_SM
#define PCGRL_SMB_CTAS_PER_SM 4   // A/B on B200 (116x16, 65 536 envs): 2 / 3 / 4 CTAs per SM -> 7.0 / 8.5 / 9.7e6 env-steps/s
#endif
constexpr int SMB_MAX_CTAS_PER_SM = PCGRL_SMB_CTAS_PER_SM;
constexpr int SMB_MAX_CTAS = 160 * SMB_MAX_CTAS_PER_SM;   // sizes the global scratch (>= 148 SMs x CTAs per SM)
constexpr int SMB_YOFF = 8;                      // y >= -5 (four rows per jump, re-jump possible from y = -1)

struct SmbScratch {
    static constexpr size_t nodes = 0;                                        // uint2[SMB_NODES]
    static constexpr size_t heap = nodes + sizeof(uint2) * (size_t)SMB_NODES; // u32[SMB_NODES]: idx | key << 16
    static constexpr size_t total = (heap + 4 * (size_t)SMB_NODES + 255) / 256 * 256;
};

int64_t smb_scratch_bytes() { return (int64_t)SmbScratch::total * SMB_MAX_CTAS * SMB_WARPS; }

struct SmbLayout {
    int stage, solid, visited, total, visited_bytes;
};
__host__ __device__ inline SmbLayout smb_layout(int H, int W, int row_stride) {
    SmbLayout L;
    int o = 0;
    L.stage = o;   o += (row_stride + 15) / 16 * 16;
    L.solid = o;   o += H * 16;                                   // 4 words per row (level width <= 128)
    L.visited_bytes = ((H + SMB_YOFF) * (W + 6) + 15) / 16 * 16;   // 8 airTime slots per (x, y) = one byte
    L.visited = o; o += L.visited_bytes;
    L.total = (o + 15) / 16 * 16;
    return L;
}

struct SmbProb {
    static constexpr int K = 9;   // dist-floor disjoint-tubes enemies empty noise jumps jumps-dist dist-win sol-length

    struct Ctx {
        int H, W, LW, exit_x, visited_bytes;
        uint8_t* stage;
        uint32_t* solid;     // [H][4]
        uint8_t* visited;    // [(y + YOFF) * LW + x] bit airTime
        uint2* nodes;
        uint32_t* heap;
    };

    __device__ static Ctx make_ctx(const KParams& p, uint8_t* ws, int global_warp) {
        Ctx c;
        c.H = p.d0; c.W = p.d1;
        c.LW = c.W + 6;
        c.exit_x = c.W + 4;
        const SmbLayout L = smb_layout(c.H, c.W, p.row_stride);
        c.stage = ws + L.stage;
        c.solid = (uint32_t*)(ws + L.solid);
        c.visited = ws + L.visited;
        c.visited_bytes = L.visited_bytes;
        uint8_t* g = (uint8_t*)p.scratch + (size_t)global_warp * SmbScratch::total;
        c.nodes = (uint2*)(g + SmbScratch::nodes);
        c.heap = (uint32_t*)(g + SmbScratch::heap);
        return c;
    }

    __device__ static __forceinline__ bool solid_at(const Ctx& c, int x, int y) {
        return (c.solid[y * 4 + (x >> 5)] >> (x & 31)) & 1u;
    }
    // engine.py:190-193 checkMovableLocation
    __device__ static __forceinline__ bool movable(const Ctx& c, int x, int y) {
        if (y < 0) return true;
        return !(x < 0 || x >= c.LW || y >= c.H || solid_at(c, x, y));
    }

    struct Node {
        int x, y, air, last_jx, gap, jumps, depth;
    };
    __device__ static __forceinline__ uint2 pack(const Node& n) {
        return make_uint2((uint32_t)n.x | ((uint32_t)(n.y + SMB_YOFF) << 7) | ((uint32_t)n.air << 15) |
                              ((uint32_t)n.last_jx << 18) | ((uint32_t)n.gap << 25),
                          (uint32_t)n.jumps | ((uint32_t)n.depth << 16));
    }
    __device__ static __forceinline__ Node unpack(const uint2 v) {
        Node n;
        n.x = v.x & 0x7F;
        n.y = (int)((v.x >> 7) & 0xFF) - SMB_YOFF;
        n.air = (v.x >> 15) & 0x7;
        n.last_jx = (v.x >> 18) & 0x7F;
        n.gap = (v.x >> 25) & 0x7F;
        n.jumps = v.y & 0xFFFF;
        n.depth = v.y >> 16;
        return n;
    }

    // engine.py:195-237 State.update for a node that is neither won nor lost
    __device__ static __forceinline__ Node update(const Ctx& c, Node n, int dx, int dy) {
        bool ground = false;
        if (n.y < c.H - 1 && n.y >= -1) ground = solid_at(c, n.x, n.y + 1);
        int nx = n.x, ny = n.y;
        if (dx != 0 && movable(c, nx + dx, ny)) nx += dx;
        if (dy == -1) {
            if (ground && movable(c, nx, ny - 1)) {
                n.air = 5;
                n.jumps += 1;
                n.gap = max(n.gap, n.x - n.last_jx);   // jump_locs.append((x, y)) with the pre-move x
                n.last_jx = n.x;
            }
        } else if (n.air > 0) {
            n.air = 1;
        }
        if (n.air > 1) {
            n.air -= 1;
            if (movable(c, nx, ny - 1)) ny -= 1;
            else n.air = 1;
        } else if (n.air == 1) {
            n.air = 0;
        } else if (movable(c, nx, ny + 1)) {
            ny += 1;
        }
        n.x = nx;
        n.y = ny;
        n.depth += 1;
        return n;
    }

    // CPython heapq on (key << 16 | idx) entries (lane 0 only); see step_sokoban.cu
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

    // engine.py:105-129 AStarAgent.getSolution.  Returns true on win; `res` = winning node or the best node.
    __device__ static bool search(Ctx& c, int balance, int lane, Node& res) {
        for (int i = lane; i < c.visited_bytes / 16; i += 32) ((uint4*)c.visited)[i] = make_uint4(0, 0, 0, 0);
        Node root;
        root.x = 1; root.y = c.H - 3; root.air = 0; root.last_jx = 0; root.gap = 0; root.jumps = 0; root.depth = 0;
        if (lane == 0) {
            c.nodes[0] = pack(root);
            c.heap[0] = (uint32_t)(c.exit_x - root.x) << 16;
        }
        __syncwarp();
        int n_nodes = 1, hn = 1, iters = 0;
        int best_h = 0x7FFFFFFF, best_d = 0;
        Node best = root;
        const int dx = lane & 1, dy = (lane & 2) ? -1 : 0;       // engine.py:3: (0,0) (1,0) (0,-1) (1,-1)
        while (iters < SMB_POWER && hn > 0) {
            ++iters;
            int cur = 0;
            if (lane == 0) cur = (int)(heap_pop(c.heap, hn) & 0xFFFFu);
            cur = __shfl_sync(0xffffffffu, cur, 0);
            hn = __shfl_sync(0xffffffffu, hn, 0);
            const Node n = unpack(c.nodes[cur]);
            if (n.y >= c.H) continue;                             // checkLose: skipped, still an iteration
            if (n.x >= c.exit_x) {                                // checkWin
                res = n;
                return true;
            }
            const int vidx = (n.y + SMB_YOFF) * c.LW + n.x;
            const uint32_t vb = c.visited[vidx];
            if ((vb >> n.air) & 1u) continue;
            const int h = c.exit_x - n.x;
            if (h < best_h || (h == best_h && n.depth < best_d)) {
                best_h = h;
                best_d = n.depth;
                best = n;
            }
            uint32_t key = 0;
            if (lane == 0) c.visited[vidx] = (uint8_t)(vb | (1u << n.air));
            if (lane < 4) {
                const Node ch = update(c, n, dx, dy);
                c.nodes[n_nodes + lane] = pack(ch);
                key = (uint32_t)((c.exit_x - ch.x) + balance * ch.depth);
            }
            // heap pushes in direction order (lane 0), keys fetched by shuffle
            for (int k = 0; k < 4; ++k) {
                const uint32_t kk = __shfl_sync(0xffffffffu, key, k);
                if (lane == 0) {
                    c.heap[hn] = (uint32_t)(n_nodes + k) | (kk << 16);
                    sift_down(c.heap, 0, hn);
                    ++hn;
                }
            }
            hn = __shfl_sync(0xffffffffu, hn, 0);
            n_nodes += 4;
            __syncwarp();
        }
        res = best;
        return false;
    }

    __device__ static void stats(const KParams& p, Ctx& c, const int8_t* grid, int lane, int32_t* out) {
        const int H = c.H, W = c.W;
        {
            uint4* st = (uint4*)c.stage;
            const uint4* src = (const uint4*)grid;
            for (int i = lane; i < p.row_stride / 16; i += 32) st[i] = src[i];
            __syncwarp();
        }
        const uint8_t* g = c.stage;
        // ---- map statistics (smb_prob.py:134-139) --------------------------------------------------------------
        int dist_floor = 0, tubes = 0, enemies = 0, empty = 0, noise = 0;
        const uint32_t magic_w = div_magic(W);
        for (int i = lane; i < p.cells; i += 32) {
            const int y = div_by(i, magic_w), x = i - y * W, t = g[i];
            empty += t == 0;
            if (t == 2) {                                        // helper.py:40-46 distance of an enemy to the floor
                ++enemies;
                int d = H - 1;
                for (int dy = 1; y + dy < H; ++dy) {
                    const int f = g[i + dy * W];
                    if (f == 1 || f == 3 || f == 4) {
                        d = dy - 1;
                        break;
                    }
                }
                dist_floor += d;
            }
            if (t == 6) {                                        // helper.py:103 tubes with exactly one tube beside
                const int nb = (x > 0 && g[i - 1] == 6) + (x < W - 1 && g[i + 1] == 6);
                tubes += nb == 1;
            }
            noise += (x > 0 && g[i - 1] != t) + (y > 0 && g[i - W] != t);   // helper.py:123 get_changes h + v
        }
        dist_floor = __reduce_add_sync(0xffffffffu, dist_floor);
        tubes = __reduce_add_sync(0xffffffffu, tubes);
        enemies = __reduce_add_sync(0xffffffffu, enemies);
        empty = __reduce_add_sync(0xffffffffu, empty);
        noise = __reduce_add_sync(0xffffffffu, noise);

        // ---- the framed level's solid bit map (smb_prob.py:96-113; " # ## #": codes 1, 3, 4, 6 are solid) --------
        for (int y = lane; y < H; y += 32) {
            uint32_t w[4] = {0, 0, 0, 0};
            for (int x = 0; x < W; ++x) {
                const int t = g[y * W + x];
                if (t == 1 || t == 3 || t == 4 || t == 6) w[(x + 3) >> 5] |= 1u << ((x + 3) & 31);
            }
            if (y == H - 3) {
                w[(W + 4) >> 5] |= 1u << ((W + 4) & 31);
            } else if (y > H - 3) {
                w[0] |= 7u;
                for (int x = W + 3; x < W + 6; ++x) w[x >> 5] |= 1u << (x & 31);
            }
            for (int k = 0; k < 4; ++k) c.solid[y * 4 + k] = w[k];
        }
        __syncwarp();

        // ---- _run_game (smb_prob.py:115-130) -----------------------------------------------------------------------
        Node res;
        bool won = search(c, 1, lane, res);
        if (!won) {
            __syncwarp();
            won = search(c, 0, lane, res);
        }
        if (lane == 0) {
            out[0] = dist_floor;
            out[1] = tubes;
            out[2] = enemies;
            out[3] = empty;
            out[4] = noise;
            out[5] = res.jumps;
            out[6] = max(res.gap, W - res.last_jx);              // smb_prob.py:147-152 with _width = map_shape[1]
            out[7] = won ? 0 : c.exit_x - res.x;
            out[8] = won ? res.depth : 0;
        }
        __syncwarp();
    }
};

cudaError_t launch_smb(const KParams& p, cudaStream_t s, bool& supported) {
    // level width W + 6 <= 128 (7-bit x), H + 8 <= 255, at least one open row above the player's
    supported = p.ndim == 2 && p.d1 <= 122 && p.d0 >= 4 && p.d0 <= 240 && p.scratch != nullptr;
    if (!supported) return cudaSuccess;
    const SmbLayout L = smb_layout(p.d0, p.d1, p.row_stride);
    if (L.total * SMB_WARPS > 200 * 1024) {
        supported = false;
        return cudaSuccess;
    }
    return launch_search<SmbProb, SMB_WARPS>(p, s, L.total, SMB_MAX_CTAS_PER_SM, SMB_MAX_CTAS);
}

}  // namespace pcgrl
