// binary / zelda on maps larger than 32x32 (up to 64x64): warp-per-grid bit-boards.
//
// The thread-per-grid kernels (step_bitboard.cu, step_split.cu) keep a whole board in one thread's registers,
// which stops at 32 rows of 32 cells.  The reference's `binary_bigger` / `zelda_bigger` tasks are 64x64
// (configs/task/binary_bigger.yaml:5, zelda_bigger.yaml:5): here a board is spread over a warp -- lane l holds
// rows 2l and 2l+1 as two 64-bit words -- and the searches run level-synchronously with warp-uniform control
// flow: x+-1 are shifts inside a word, y+-1 one word from the same lane and one from a neighbouring lane
// (__shfl_up / __shfl_down), "frontier empty" is a ballot, "lowest cell" is the first lane of that ballot and a
// find-first-set.  Same restatements as the register machines (bitboard_machines.cuh):
//   binary (envs/probs/binary/binary_prob.py:152-158, envs/helper.py:200-276): isolated cells by popcount, first
//     sweeps one component at a time (start = lowest remaining cell, far = lowest cell of the last level), second
//     sweeps of all components as one multi-source BFS;
//   zelda (envs/probs/zelda/zelda_ctrl_prob.py:90-168): tile counts, regions by flood fill, BFS from the player
//     (nearest-enemy, d(player -> key)), BFS from the key over walkable + door (d(key -> door)), -1 when unreached.
// The step around it (representation update, counters, reward) is the warp-per-grid skeleton of step_search.cuh.
#include "step_search.cuh"

namespace pcgrl {

struct Rows2 {
    uint64_t a, b;   // rows 2 * lane and 2 * lane + 1
};
__device__ __forceinline__ Rows2 r2(uint64_t a, uint64_t b) {
    Rows2 r;
    r.a = a;
    r.b = b;
    return r;
}
__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int lane) {
    const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, 1), hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), 1);
    return lane == 0 ? 0ull : ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_down64(uint64_t v, int lane) {
    const uint32_t lo = __shfl_down_sync(0xffffffffu, (uint32_t)v, 1), hi = __shfl_down_sync(0xffffffffu, (uint32_t)(v >> 32), 1);
    return lane == 31 ? 0ull : ((uint64_t)hi << 32) | lo;
}
// 4-neighbour dilation of f (bits beyond the map width / height are never set in `av`, so nothing leaks)
__device__ __forceinline__ Rows2 dilate(const Rows2 f, int lane) {
    const uint64_t up = shfl_up64(f.b, lane), dn = shfl_down64(f.a, lane);
    return r2((f.a << 1) | (f.a >> 1) | up | f.b, (f.b << 1) | (f.b >> 1) | f.a | dn);
}
__device__ __forceinline__ bool any2(const Rows2 x) { return __any_sync(0xffffffffu, (x.a | x.b) != 0); }
__device__ __forceinline__ int popc2(const Rows2 x) { return __reduce_add_sync(0xffffffffu, __popcll(x.a) + __popcll(x.b)); }
// one-hot board of the lowest cell of x in row-major order (all zero when x is empty)
__device__ __forceinline__ Rows2 lowest2(const Rows2 x, int lane) {
    const unsigned has = __ballot_sync(0xffffffffu, (x.a | x.b) != 0);
    Rows2 r = r2(0, 0);
    if (has && lane == __ffs(has) - 1) {
        if (x.a) r.a = x.a & (0 - x.a);
        else r.b = x.b & (0 - x.b);
    }
    return r;
}
// level-synchronous BFS from `front` over `avail` (front excluded from avail by the caller); returns the number of
// levels; `last` = the last non-empty level.  hit / hit_level: the first level (>= 1) that touches `probe`.
__device__ __forceinline__ int sweep(Rows2 front, Rows2& avail, int lane, Rows2* last = nullptr, const Rows2* probe = nullptr,
                                     int* hit_level = nullptr, const Rows2* probe2 = nullptr, int* hit_level2 = nullptr) {
    int level = 0;
    for (;;) {
        const Rows2 d = dilate(front, lane);
        const Rows2 n = r2(d.a & avail.a, d.b & avail.b);
        if (!any2(n)) break;
        avail.a &= ~n.a;
        avail.b &= ~n.b;
        front = n;
        ++level;
        if (probe && *hit_level == 0 && __any_sync(0xffffffffu, ((n.a & probe->a) | (n.b & probe->b)) != 0)) *hit_level = level;
        if (probe2 && __any_sync(0xffffffffu, ((n.a & probe2->a) | (n.b & probe2->b)) != 0)) *hit_level2 = level;
    }
    if (last) *last = front;
    return level;
}

// membership masks of the lane's two rows: bit x of plane q = tile at (row, x) is in tile set `sets[q]`
template <int P>
__device__ __forceinline__ void load_rows(const KParams& p, const int8_t* grid, int lane, const uint32_t (&sets)[P],
                                          Rows2 (&plane)[P]) {
    const int H = p.d0, W = p.d1;
#pragma unroll
    for (int q = 0; q < P; ++q) plane[q] = r2(0, 0);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int y = 2 * lane + half;
        if (y >= H) continue;
        const int8_t* row = grid + y * W;
        for (int x = 0; x < W; ++x) {
            const uint32_t t = (uint32_t)row[x] & 31u;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                const uint64_t bit = (uint64_t)((sets[q] >> t) & 1u) << x;
                if (half == 0) plane[q].a |= bit;
                else plane[q].b |= bit;
            }
        }
    }
}

struct BigCtx {
    int unused;
};

struct BigBinaryProb {
    static constexpr int K = 2;   // regions, path-length
    using Ctx = BigCtx;
    __device__ static Ctx make_ctx(const KParams&, uint8_t*, int) { return Ctx{0}; }
    __device__ static void stats(const KParams& p, Ctx&, const int8_t* grid, int lane, int32_t* out) {
        const uint32_t sets[1] = {0x1u};   // {empty}
        Rows2 pl[1];
        load_rows<1>(p, grid, lane, sets, pl);
        const Rows2 pass = pl[0];
        const Rows2 nb = dilate(pass, lane);
        const Rows2 iso = r2(pass.a & ~nb.a, pass.b & ~nb.b);
        int regions = popc2(iso);
        const Rows2 base = r2(pass.a & ~iso.a, pass.b & ~iso.b);
        Rows2 avail = base, fars = r2(0, 0);
        for (;;) {   // first sweeps: one component at a time, in row-major order of its first tile
            const Rows2 start = lowest2(avail, lane);
            if (!any2(start)) break;
            ++regions;
            avail.a &= ~start.a;
            avail.b &= ~start.b;
            Rows2 last;
            sweep(start, avail, lane, &last);
            const Rows2 far = lowest2(last, lane);   // np.argmax: row-major-first cell of the last level
            fars.a |= far.a;
            fars.b |= far.b;
        }
        Rows2 av2 = r2(base.a & ~fars.a, base.b & ~fars.b);
        const int path = any2(fars) ? sweep(fars, av2, lane) : 0;   // joint second sweep from every far tile
        if (lane == 0) {
            out[0] = regions;
            out[1] = path;
        }
    }
};

struct BigZeldaProb {
    static constexpr int K = 7;   // player key door enemies regions nearest-enemy path-length
    using Ctx = BigCtx;
    __device__ static Ctx make_ctx(const KParams&, uint8_t*, int) { return Ctx{0}; }
    __device__ static void stats(const KParams& p, Ctx&, const int8_t* grid, int lane, int32_t* out) {
        // walkable {0,2,3,5,6,7} (zelda_ctrl_prob.py:101-104), player, key, door, enemies {5,6,7}
        const uint32_t sets[5] = {0xEDu, 0x04u, 0x08u, 0x10u, 0xE0u};
        Rows2 pl[5];
        load_rows<5>(p, grid, lane, sets, pl);
        const Rows2 walk = pl[0];
        const int n_player = popc2(pl[1]), n_key = popc2(pl[2]), n_door = popc2(pl[3]), n_enemy = popc2(pl[4]);
        int regions = 0;
        {
            Rows2 avail = walk;
            for (;;) {
                const Rows2 start = lowest2(avail, lane);
                if (!any2(start)) break;
                ++regions;
                avail.a &= ~start.a;
                avail.b &= ~start.b;
                sweep(start, avail, lane);
            }
        }
        int near = 0, dkey = -1, ddoor = -1;
        if (n_player == 1 && (n_enemy > 0 || (n_key == 1 && n_door == 1))) {
            Rows2 avail = r2(walk.a & ~pl[1].a, walk.b & ~pl[1].b);
            sweep(pl[1], avail, lane, nullptr, &pl[4], &near, &pl[2], &dkey);
            if (n_key == 1 && n_door == 1) {
                Rows2 av2 = r2((walk.a | pl[3].a) & ~pl[2].a, (walk.b | pl[3].b) & ~pl[2].b);
                int unused = 1;   // hit_level of the first probe is only written while it is 0
                sweep(pl[2], av2, lane, nullptr, &pl[3], &unused, &pl[3], &ddoor);
            }
        }
        if (lane == 0) {
            out[0] = n_player;
            out[1] = n_key;
            out[2] = n_door;
            out[3] = n_enemy;
            out[4] = regions;
            out[5] = near;
            out[6] = (n_player == 1 && n_key == 1 && n_door == 1) ? dkey + ddoor : 0;
        }
    }
};

// maps with a side above 32 (up to 64x64); smaller maps belong to the register machines
cudaError_t launch_bigboard(const KParams& p, int problem, cudaStream_t s, bool& supported) {
    supported = p.ndim == 2 && p.d0 <= 64 && p.d1 <= 64 && (problem == PCGRL_PROB_BINARY || problem == PCGRL_PROB_ZELDA);
    if (!supported) return cudaSuccess;
    if (problem == PCGRL_PROB_BINARY) return launch_search<BigBinaryProb, SEARCH_WARPS>(p, s, 16);
    return launch_search<BigZeldaProb, SEARCH_WARPS>(p, s, 16);
}

}  // namespace pcgrl
