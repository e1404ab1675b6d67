// Fused batched env-step kernel for the bit-board problems (maps up to 32x32):
//   representation update -> get_stats -> ControlWrapper reward, one launch for the whole env shard.
//
// Reference path replaced (paths relative to /root/reference/control_pcgrl/):
//   envs/pcgrl_env.py:267-342 (step), envs/reps/{narrow,turtle,wide,ca}_rep.py (update),
//   envs/probs/binary/binary_prob.py:152-158, envs/probs/zelda/zelda_ctrl_prob.py:90-168 (get_stats),
//   envs/helper.py:200-276 (calc_num_regions, run_dijkstra, calc_longest_path),
//   control_wrappers.py:216-244,318-345 (reward = loss - last_loss).
//
// Work decomposition per CTA (TILE consecutive envs):
//   A  thread-per-env: apply the action to the int8 grid in HBM (single-byte scatter, or whole-map rewrite
//      for the cellular rep), advance counters, decide done, flag envs whose map changed;
//   B  all threads: 128-bit coalesced loads of the changed grids, packed into per-plane bit-boards in
//      shared memory (one bit per cell; a "plane" is a set of tile codes);
//   C  thread-per-grid: every thread pulls changed envs from a shared-memory queue and runs the stat
//      searches as level-synchronous bit-board BFS held entirely in registers (NW 32-bit words): shifts
//      inside a word for x+-1, funnel shifts between neighbouring words for y+-1, no cross-lane traffic;
//   D  thread-per-env: fp64 loss(new stats) - loss(old stats), write stats / reward / done / counters.
#include "pcgrl_device.cuh"
#include "step_common.cuh"

namespace pcgrl {

#ifndef PCGRL_THREADS
#define PCGRL_THREADS 128
#endif
#ifndef PCGRL_MIN_CTAS
#define PCGRL_MIN_CTAS 10  // boards of <= 8 words: 48 registers, 10 CTAs (40 warps) per SM; A/B 1 / 10 -> 2.88 / 2.95e9
#endif
#ifndef PCGRL_CARVEOUT
#define PCGRL_CARVEOUT 0
#endif
#ifndef PCGRL_SKIP_CLAIM
#define PCGRL_SKIP_CLAIM 0
#endif
#ifndef PCGRL_STATIC_ITEMS
#define PCGRL_STATIC_ITEMS 0
#endif
#ifndef PCGRL_EXPAND_R
#define PCGRL_EXPAND_R 3
#endif
#ifndef PCGRL_WARP_BATCH
#define PCGRL_WARP_BATCH 0  // 1: warps claim batches of 32 changed envs; 0: per-thread claims
#endif
#ifndef PCGRL_BATCH_K
#define PCGRL_BATCH_K 0     // lanes of a warp that must be waiting for a transition before the warp takes it
#endif
#ifndef PCGRL_TILE8
#define PCGRL_TILE8 256   // envs per CTA when a board is <= 8 words (binary 16x16); A/B: 2 envs per thread is best
#endif
constexpr int THREADS = PCGRL_THREADS;
// envs per CTA, bounded so the shared-memory bit-boards stay under the 48 KB static limit
__host__ __device__ constexpr int tile_for(int bbw) { return bbw == 8 ? PCGRL_TILE8 : bbw < 8 ? 256 : (bbw <= 32 ? 256 : (bbw <= 80 ? 128 : 64)); }

// ------------------------------------------------------------------------------------------------
// Problem policies: planes (tile-code sets packed to bit-boards) + the per-thread stats state machine.
// Each machine is a flat loop: one board expansion per trip, with rare transitions -- so the 32 grids
// that share a warp execute the same instruction stream whatever phase each of them is in.
// ------------------------------------------------------------------------------------------------
struct BinaryProb {
    static constexpr int P = 1;
    static constexpr int K = 2;  // regions, path-length
    __host__ __device__ static constexpr uint32_t plane_mask(int p) { return 0x1u; }  // {empty}
};

struct ZeldaProb {
    static constexpr int P = 5;
    static constexpr int K = 7;  // player key door enemies regions nearest-enemy path-length
    // tiles: empty 0, solid 1, player 2, key 3, door 4, bat 5, scorpion 6, spider 7 (zelda_prob.py:20)
    __host__ __device__ static constexpr uint32_t plane_mask(int p) {
        return p == 0 ? 0xEDu   /* walkable {0,2,3,5,6,7}  zelda_ctrl_prob.py:101-104 */
             : p == 1 ? 0x04u   /* player */
             : p == 2 ? 0x08u   /* key */
             : p == 3 ? 0x10u   /* door */
             :          0xE0u;  /* enemies {5,6,7} */
    }
};

struct BinaryHoleyProb {
    static constexpr int P = 2;  // plane 0 {empty}; plane 1 holds no tile: the machine keeps the exit cell there
    static constexpr int K = 3;  // regions, path-length, connected-path-length
    __host__ __device__ static constexpr uint32_t plane_mask(int p) { return p == 0 ? 0x1u : 0x0u; }
};

// binary: regions + double-sweep longest path over the {empty} plane.
//   helper.calc_longest_path runs, per component, BFS(first tile) -> far tile -> BFS(far) and keeps the max.
//   Exact restatement used here: (1) isolated cells are components with eccentricity 0: count them with a
//   popcount; (2) for the other components run the first sweep one component at a time (start = lowest
//   remaining cell, far = lowest cell of the last non-empty level == np.argmax); (3) the second sweeps of
//   all components run at once as a single multi-source BFS from the set of far tiles -- components are
//   disconnected, so the number of levels until the joint frontier dies is max_c ecc(far_c).
template <int NW, bool TWO>
struct BinaryMachine {
    using Prob = BinaryProb;
    using B = Board<NW, TWO>;
    uint32_t avail[NW], front[NW], fars[NW];
    uint32_t* base;  // shared-memory copy of the non-isolated passable cells (re-read for the joint sweep)
    int phase, level, ncomp;

    __device__ __forceinline__ void init(uint32_t* bb /* [P][NW] in shared memory */, const KParams&, int64_t) {
        uint32_t pass[NW], ones[NW], nb[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            pass[i] = bb[i];
            ones[i] = 0xFFFFFFFFu;
        }
        B::expand_and(pass, ones, nb);
        ncomp = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            const uint32_t iso = pass[i] & ~nb[i];
            ncomp += __popc(iso);
            avail[i] = pass[i] & ~iso;
            bb[i] = avail[i];
            front[i] = 0;
            fars[i] = 0;
        }
        base = bb;
        phase = 0;
        level = 0;
    }
    // one board expansion; returns false (and changes nothing) when the frontier has died
    __device__ __forceinline__ bool expand() {
        uint32_t n[NW];
        if (!B::expand_and(front, avail, n)) return false;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            avail[i] = B::minus_subset(avail[i], n[i]);
            front[i] = n[i];
        }
        ++level;
        return true;
    }
    // the frontier died: next component / next phase; returns true when out[] holds the K stats
    __device__ __forceinline__ bool transition(int* out) {
        if (phase == 0) {
#if PCGRL_OPT_BORROW
            uint32_t t[NW];
            B::minus_one(front, t);           // far tile of the component just swept (nothing on entry)
#pragma unroll
            for (int i = 0; i < NW; ++i) fars[i] |= front[i] & ~t[i];
            if (B::minus_one(avail, t)) {     // first tile of the next component
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    front[i] = avail[i] & ~t[i];
                    avail[i] &= t[i];
                }
                ++ncomp;
                return false;
            }
#else
            uint32_t lo[NW];
            B::lowest(front, lo);             // far tile of the component just swept (nothing on entry)
#pragma unroll
            for (int i = 0; i < NW; ++i) fars[i] |= lo[i];
            B::lowest(avail, lo);             // first tile of the next component
            if (B::any(lo)) {
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    front[i] = lo[i];
                    avail[i] ^= lo[i];
                }
                ++ncomp;
                return false;
            }
#endif
            phase = 1;  // joint second sweep from every far tile
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                front[i] = fars[i];
                avail[i] = base[i] & ~fars[i];
                any |= fars[i];
            }
            level = 0;
            if (any) return false;
        }
        out[0] = ncomp;
        out[1] = level;
        return true;
    }
};

// binary_holey (envs/probs/binary/binary_holey_prob.py:59-93): the stats are taken on the BORDERED map
// (pcgrl_holey_env.py:52-53) whose border is solid except for the entrance and the exit.  The board built by
// phase B holds the level (one row per word); init() moves it one cell down-right into the border frame and
// digs the two holes.  regions = flood fill over the bordered board; then ONE BFS from the entrance:
// path-length = its last level (np.max of the dijkstra map), connected-path-length = the level that reaches
// the exit (0 when it never does: the reference maps -1 to 0, :69-77).
template <int NW, bool TWO>
struct BinaryHoleyMachine {
    static_assert(!TWO, "the bordered board keeps one row per word");
    using Prob = BinaryHoleyProb;
    using B = Board<NW, TWO>;
    uint32_t avail[NW], front[NW];
    uint32_t* bb;   // plane 0: bordered passable board, plane 1: the exit cell
    int phase, level, regions, connected, ey, ex;

    __device__ __forceinline__ void init(uint32_t* planes, const KParams& p, int64_t env) {
        bb = planes;
        const int32_t* h = p.holes + env * 4;
        ey = h[0];
        ex = h[1];
        const int xy = h[2], xx = h[3];
        uint32_t pass[NW], ones[NW], nb[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) pass[i] = i > 0 ? planes[i - 1] << 1 : 0u;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            uint32_t xm = 0;
            if (i == xy && (unsigned)xx < 32u) xm = 1u << xx;
            if (i == ey && (unsigned)ex < 32u) pass[i] |= 1u << ex;
            pass[i] |= xm;
            planes[i] = pass[i];
            planes[NW + i] = xm;
            ones[i] = 0xFFFFFFFFu;
        }
        B::expand_and(pass, ones, nb);
        regions = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            const uint32_t iso = pass[i] & ~nb[i];
            regions += __popc(iso);
            avail[i] = pass[i] & ~iso;
            front[i] = 0;
        }
        phase = 0;
        level = 0;
        connected = 0;
    }
    __device__ __forceinline__ bool expand() {
        uint32_t n[NW];
        if (!B::expand_and(front, avail, n)) return false;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            avail[i] = B::minus_subset(avail[i], n[i]);
            front[i] = n[i];
        }
        ++level;
        if (phase == 1) {
            uint32_t hit = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) hit |= n[i] & bb[NW + i];
            if (hit) connected = level;
        }
        return true;
    }
    __device__ __forceinline__ bool transition(int* out) {
        if (phase == 0) {
            uint32_t t[NW];
            if (B::minus_one(avail, t)) {
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    front[i] = avail[i] & ~t[i];
                    avail[i] &= t[i];
                }
                ++regions;
                return false;
            }
            phase = 1;   // BFS from the entrance over the whole bordered board
            level = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                front[i] = (i == ey && (unsigned)ex < 32u) ? 1u << ex : 0u;
                avail[i] = bb[i] & ~front[i];
            }
            return false;
        }
        out[0] = regions;
        out[1] = level;
        out[2] = connected;
        return true;
    }
};

// zelda: tile counts, regions over the walkable plane, then (player == 1) BFS from the player:
// nearest-enemy = first level >= 1 that touches an enemy, d(player->key) = level that touches the key;
// then (key == 1 && door == 1) BFS from the key over walkable+door: d(key->door).  Unreached = -1 each
// (run_dijkstra's fill value), added raw (zelda_ctrl_prob.py:134-150).
template <int NW, bool TWO>
struct ZeldaMachine {
    using Prob = ZeldaProb;
    using B = Board<NW, TWO>;
    uint32_t avail[NW], front[NW];
    const uint32_t* bb;  // planes in shared memory: walk, player, key, door, enemy
    int phase, level, regions, near, dkey, ddoor;
    int n_player, n_key, n_door, n_enemy;

    __device__ __forceinline__ void load(int plane, uint32_t (&x)[NW]) const {
#pragma unroll
        for (int i = 0; i < NW; ++i) x[i] = bb[plane * NW + i];
    }
    __device__ __forceinline__ void init(uint32_t* planes, const KParams&, int64_t) {
        bb = planes;
        uint32_t walk[NW], t[NW], ones[NW], nb[NW];
        load(0, walk);
        load(1, t);
        n_player = B::popcount(t);
        load(2, t);
        n_key = B::popcount(t);
        load(3, t);
        n_door = B::popcount(t);
        load(4, t);
        n_enemy = B::popcount(t);
#pragma unroll
        for (int i = 0; i < NW; ++i) ones[i] = 0xFFFFFFFFu;
        B::expand_and(walk, ones, nb);
        regions = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            const uint32_t iso = walk[i] & ~nb[i];
            regions += __popc(iso);
            avail[i] = walk[i] & ~iso;
            front[i] = 0;
        }
        phase = 0;
        level = 0;
        near = 0;
        dkey = -1;
        ddoor = -1;
    }
    __device__ __forceinline__ bool expand() {
        uint32_t n[NW];
        if (!B::expand_and(front, avail, n)) return false;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            avail[i] ^= n[i];
            front[i] = n[i];
        }
        ++level;
        if (phase == 1) {
            uint32_t t[NW];
            load(4, t);
            if (near == 0 && B::any_and(n, t)) near = level;
            load(2, t);
            if (B::any_and(n, t)) dkey = level;
        } else if (phase == 2) {
            uint32_t t[NW];
            load(3, t);
            if (B::any_and(n, t)) ddoor = level;
        }
        return true;
    }
    __device__ __forceinline__ bool transition(int* out) {
        if (phase == 0) {  // region flood fill, one component at a time
            uint32_t t[NW];
            if (B::minus_one(avail, t)) {
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    front[i] = avail[i] & ~t[i];
                    avail[i] &= t[i];
                }
                ++regions;
                return false;
            }
            if (n_player == 1 && (n_enemy > 0 || (n_key == 1 && n_door == 1))) {
                phase = 1;  // BFS from the player over the walkable plane
                uint32_t walk[NW];
                load(0, walk);
                load(1, front);
#pragma unroll
                for (int i = 0; i < NW; ++i) avail[i] = walk[i] & ~front[i];
                level = 0;
                return false;
            }
            phase = 3;
        } else if (phase == 1) {
            if (n_key == 1 && n_door == 1) {
                phase = 2;  // BFS from the key over walkable + door
                uint32_t walk[NW], door[NW];
                load(0, walk);
                load(3, door);
                load(2, front);
#pragma unroll
                for (int i = 0; i < NW; ++i) avail[i] = (walk[i] | door[i]) & ~front[i];
                level = 0;
                return false;
            }
            phase = 3;
        }
        out[0] = n_player;
        out[1] = n_key;
        out[2] = n_door;
        out[3] = n_enemy;
        out[4] = regions;
        out[5] = near;
        out[6] = (n_player == 1 && n_key == 1 && n_door == 1) ? dkey + ddoor : 0;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// 16 tile codes (one 128-bit load) -> one 16-bit membership mask per plane (bit i = cell i in the plane).
// Tile codes are < 8 for every bit-board problem, so a plane is an 8-entry 0/1 table that PRMT looks up
// for 4 cells at once; the multiply then gathers the four 0/1 bytes into bits 24..27.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t lut_bytes(uint32_t mask, int first) {
    return ((mask >> first) & 1u) | (((mask >> (first + 1)) & 1u) << 8) | (((mask >> (first + 2)) & 1u) << 16) |
           (((mask >> (first + 3)) & 1u) << 24);
}
template <class Prob>
__device__ __forceinline__ void pack16(const uint4 v, uint32_t (&out)[Prob::P]) {
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < Prob::P; ++q) out[q] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t t = w4[k] | (w4[k] >> 4);
        const uint32_t sel = ((t & 0xFFu) | ((t >> 8) & 0xFF00u)) & 0x7777u;  // 4 nibbles = 4 tile codes
#pragma unroll
        for (int q = 0; q < Prob::P; ++q) {
            const uint32_t b = __byte_perm(lut_bytes(Prob::plane_mask(q), 0), lut_bytes(Prob::plane_mask(q), 4), sel);
            out[q] |= ((b * 0x01020408u) >> 24) << (4 * k);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <class Machine, int NW, bool TWO>
__global__ void __launch_bounds__(THREADS, NW <= 8 ? PCGRL_MIN_CTAS : 1) k_step_bitboard(const KParams p) {
    using Prob = typename Machine::Prob;
    constexpr int P = Prob::P;
    constexpr int K = Prob::K;
    constexpr int BBW = P * NW;  // board words per env
    constexpr int TILE = tile_for(BBW);
    static_assert(TILE % 32 == 0 && THREADS % 32 == 0, "phase_a's warp ballots need every lane of a warp in the same trip");

    __shared__ uint32_t s_bb[TILE * BBW];
    __shared__ int32_t s_stats[TILE * K];
    __shared__ int16_t s_list[TILE];   // compact list of tile-local env indices that need stats
    __shared__ int16_t s_slot[TILE];   // env -> position in s_list, or -1
    __shared__ uint8_t s_flag[TILE];   // cellular: map changed
    __shared__ int s_count, s_next;

    // One CTA per tile.  (A persistent grid with a device-side tile counter, as in step_search.cuh, was A/B-tested
    // here and ran 2 % slower: the extra barriers cost more than the partly filled last wave.)
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * TILE;
    const int64_t n_total = p.n_envs;
    const int tile_n = (int)min((int64_t)TILE, n_total - base);
    const int8_t* grids_in = (p.mode == MODE_STATS) ? p.stats_grids : p.grids;

    if (tid == 0) {
        s_count = 0;
        s_next = 0;
    }
    for (int e = tid; e < TILE; e += THREADS) {
        s_slot[e] = -1;
        s_flag[e] = 0;
    }
    __syncthreads();

    // ---------------- phase A: representation update / reset, counters, change flag ----------------
    phase_a<THREADS, TILE>(p, base, tile_n, s_list, s_slot, s_flag, &s_count);
    const int M = s_count;

    // ---------------- phase B: coalesced 128-bit grid loads -> per-plane bit-boards in smem --------
    // 16 cells (one uint4) are turned into a 16-bit membership mask per plane with SIMD-in-register ops:
    // a PRMT table lookup maps 4 tile codes to 4 0/1 bytes, a multiply gathers them into a nibble.
    {
        const int chunks = p.row_stride / 16;
        const int W = (p.ndim == 2) ? p.d1 : p.d2;
        const bool word_is_32_bytes = TWO ? (W == 16) : (W == 32);
        if (word_is_32_bytes) {
            // a board word is exactly 32 consecutive cells: one thread builds one word, no atomics
            for (int i = tid; i < M * NW; i += THREADS) {
                const int slot = i / NW, j = i - slot * NW;
                const int8_t* g = grids_in + (base + s_list[slot]) * p.row_stride + j * 32;
                uint32_t lo[P], hi[P];
#pragma unroll
                for (int q = 0; q < P; ++q) lo[q] = hi[q] = 0;
                if (2 * j < chunks) pack16<Prob>(*(const uint4*)g, lo);
                if (2 * j + 1 < chunks) pack16<Prob>(*(const uint4*)(g + 16), hi);
#pragma unroll
                for (int q = 0; q < P; ++q) s_bb[slot * BBW + q * NW + j] = lo[q] | (hi[q] << 16);
            }
        } else {
            for (int i = tid; i < M * BBW; i += THREADS) s_bb[i] = 0;
            __syncthreads();
            for (int i = tid; i < M * chunks; i += THREADS) {
                const int slot = i / chunks, c = i - slot * chunks;
                uint32_t lin[P];
                pack16<Prob>(*(const uint4*)(grids_in + (base + s_list[slot]) * p.row_stride + c * 16), lin);
                uint32_t* bb = s_bb + slot * BBW;
                int cell = c * 16;
                const int end = min(cell + 16, p.cells);
                int y = cell / W, x = cell - y * W;
                while (cell < end) {   // one row segment at a time (at most ceil(16/W)+1 of them)
                    const int len = min(W - x, end - cell);
                    const int word = TWO ? (y >> 1) : y;
                    const int sh = TWO ? ((y & 1) * 16 + x) : x;
                    const uint32_t m = (1u << len) - 1u;
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const uint32_t seg = (lin[q] >> (cell - c * 16)) & m;
                        if (seg) atomicOr(&bb[q * NW + word], seg << sh);
                    }
                    cell += len;
                    x = 0;
                    ++y;
                }
            }
        }
    }
    __syncthreads();

    // ---------------- phase C: thread-per-grid stat searches, dynamic queue --------------------------
    // Each thread pulls the next changed env from the shared-memory queue as soon as it finishes one, so
    // lanes with short searches do not idle behind long ones; the search itself is a flat loop (one board
    // expansion per trip) so lanes in different phases of different grids execute the same instructions.
    // Only as many threads as finish in whole rounds take part (R = ceil(M/THREADS) items each), so a warp's
    // lanes run out of work together instead of a few lanes dragging a mostly idle warp through a last round.
    {
        Machine m;
#if PCGRL_WARP_BATCH
        // Warp-synchronous batches: a warp claims 32 consecutive changed envs at a time and its lanes start their
        // searches together, so at most ONE warp of the CTA runs a partly filled batch (with per-thread claims
        // the few envs beyond THREADS were picked up by lanes scattered over every warp, each of which then ran
        // a whole extra search at 1/32..7/32 lane occupancy).
        for (;;) {
            int b0 = 0;
            if ((tid & 31) == 0) b0 = atomicAdd(&s_next, 32);
            b0 = __shfl_sync(0xffffffffu, b0, 0);
            if (b0 >= M) break;
            const int item = b0 + (tid & 31);
            if (item < M) {
                m.init(s_bb + item * BBW, p, base + s_list[item]);
                int out[K];
                for (;;) {
                    if (m.expand()) continue;
                    if (m.transition(out)) break;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) s_stats[item * K + k] = out[k];
            }
            __syncwarp();
        }
#else
#if PCGRL_STATIC_ITEMS
        // static round-robin claims: the changed envs beyond THREADS all land in the lowest warp(s), so only those
        // run a second, partly filled round (dynamic claims scatter them over lanes of every warp, and each such
        // warp then runs a whole extra search at 1/32 lane occupancy)
#define PCGRL_NEXT_ITEM(cur) ((cur) + THREADS)
        int item = tid;
#else
#define PCGRL_NEXT_ITEM(cur) ((PCGRL_SKIP_CLAIM && M <= THREADS) ? M : atomicAdd(&s_next, 1))
        int item = atomicAdd(&s_next, 1);
#endif
        bool active = item < M;
        if (active) m.init(s_bb + item * BBW, p, base + s_list[item]);
#if PCGRL_BATCH_K > 0
        // Expansions and transitions are different instruction streams; with 32 independent grids per warp some
        // lane needs a transition on almost every trip, so an unbatched loop pays for both streams every trip
        // with the transition stream running at 1/8 lane occupancy.  Lanes whose frontier died therefore WAIT
        // (masked off during the expansions) until PCGRL_BATCH_K lanes of the warp are waiting, or nobody can
        // expand; then all of them take their transition together.
        bool pend = active;   // a fresh machine has an empty frontier: its first act is a transition
        while (__any_sync(0xffffffffu, active)) {
            if (active && !pend) pend = !m.expand();
            const unsigned pb = __ballot_sync(0xffffffffu, pend);
            const unsigned cb = __ballot_sync(0xffffffffu, active && !pend);
            if (__popc(pb) >= PCGRL_BATCH_K || cb == 0) {
                if (pend) {
                    int out[K];
                    while (m.transition(out)) {
#pragma unroll
                        for (int k = 0; k < K; ++k) s_stats[item * K + k] = out[k];
                        item = atomicAdd(&s_next, 1);
                        active = item < M;
                        if (!active) break;
                        m.init(s_bb + item * BBW, p, base + s_list[item]);
                    }
                    pend = false;
                }
            }
        }
#else
        while (active) {
            int out[K];
            // PCGRL_EXPAND_R expansions per trip: the transition stream (a few lanes) is then paid once per R
            // expansions of the warp instead of once per expansion; a lane whose frontier dies early idles for
            // the rest of the trip
            bool alive = m.expand();
#pragma unroll
            for (int r = 1; r < PCGRL_EXPAND_R; ++r)
                if (alive) alive = m.expand();
            if (alive) continue;
            if (m.transition(out)) {
#pragma unroll
                for (int k = 0; k < K; ++k) s_stats[item * K + k] = out[k];
                item = PCGRL_NEXT_ITEM(item);
                active = item < M;
                if (active) m.init(s_bb + item * BBW, p, base + s_list[item]);
            }
        }
#endif
#endif
    }
    __syncthreads();

    // ---------------- phase D: reward (fp64), stats / outputs --------------------------------------
    phase_d<THREADS, K>(p, base, tile_n, s_slot, s_stats);
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
template <class Machine, int NW, bool TWO>
static cudaError_t launch(const KParams& p, cudaStream_t s) {
    constexpr int TILE = tile_for(Machine::Prob::P * NW);
    const int64_t ctas = (p.n_envs + TILE - 1) / TILE;
    if (ctas == 0) return cudaSuccess;
#if PCGRL_CARVEOUT > 0
    static bool carved = false;   // ask for enough shared memory per SM that PCGRL_MIN_CTAS CTAs fit
    if (!carved) {
        cudaFuncSetAttribute(k_step_bitboard<Machine, NW, TWO>, cudaFuncAttributePreferredSharedMemoryCarveout, PCGRL_CARVEOUT);
        carved = true;
    }
#endif
    k_step_bitboard<Machine, NW, TWO><<<(unsigned)ctas, THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

template <template <int, bool> class Machine>
static cudaError_t dispatch_shape(const KParams& p, cudaStream_t s, bool& supported) {
    const int H = (p.ndim == 2) ? p.d0 : -1, W = p.d1;
    supported = true;
    if (p.ndim != 2) {
        supported = false;
        return cudaSuccess;
    }
    if (W <= 16 && H <= 16) {
        const int nw = (H + 1) / 2;
        if (nw <= 1) return launch<Machine<1, true>, 1, true>(p, s);
        if (nw <= 2) return launch<Machine<2, true>, 2, true>(p, s);
        if (nw <= 4) return launch<Machine<4, true>, 4, true>(p, s);
        return launch<Machine<8, true>, 8, true>(p, s);
    }
    if (W <= 32 && H <= 32) {
        if (H <= 16) return launch<Machine<16, false>, 16, false>(p, s);
        return launch<Machine<32, false>, 32, false>(p, s);
    }
    supported = false;
    return cudaSuccess;
}

cudaError_t launch_bitboard(const KParams& p, int problem, cudaStream_t s, bool& supported) {
    if (problem == PCGRL_PROB_BINARY) return dispatch_shape<BinaryMachine>(p, s, supported);
    if (problem == PCGRL_PROB_BINARY_HOLEY) {
        // the bordered board: H + 2 rows of W + 2 cells, one row per word
        supported = p.ndim == 2 && p.d1 + 2 <= 32 && p.d0 + 2 <= 32;
        if (!supported) return cudaSuccess;
        if (p.d0 + 2 <= 18) return launch<BinaryHoleyMachine<18, false>, 18, false>(p, s);
        return launch<BinaryHoleyMachine<32, false>, 32, false>(p, s);
    }
    if (problem == PCGRL_PROB_ZELDA) return dispatch_shape<ZeldaMachine>(p, s, supported);
    supported = false;
    return cudaSuccess;
}

}  // namespace pcgrl
