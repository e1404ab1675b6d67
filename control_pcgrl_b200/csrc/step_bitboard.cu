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
#include "bitboard_machines.cuh"

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
#ifndef PCGRL_THETA_F
#define PCGRL_THETA_F 0     // expand while >= THETA_F/32 of a warp's active lanes are alive (0: fixed windows of PCGRL_EXPAND_R);
                            // A/B 8 / 12 / 16 / 20: 0.60 / 0.53 / 0.47 / 0.45 ms per step against 0.385 with fixed windows
#endif
#ifndef PCGRL_CONVERGE
#define PCGRL_CONVERGE 0    // 1: every lane of a warp starts every search trip together (vote at the loop head); A/B in the
                            // fused kernel 0.393 against 0.385 ms without (it does pay in the split kernel: 0.381 / 0.432)
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
#if PCGRL_THETA_F > 0
        // Rounds: the warp keeps expanding while at least PCGRL_THETA_F/32 of its active lanes still have a live
        // frontier (the others wait), then every waiting lane takes its transition (and claims its next grid) in
        // ONE pass of the transition stream, so both streams run with many lanes.
        bool alive = false;   // a fresh machine has an empty frontier: its first act is a transition
        for (;;) {
            const unsigned act = __ballot_sync(0xffffffffu, active);
            if (!act) break;
            const int need = max(1, (__popc(act) * PCGRL_THETA_F) >> 5);
            for (;;) {
                if (__popc(__ballot_sync(0xffffffffu, active && alive)) < need) break;
                if (active && alive) alive = m.expand();
            }
            if (active && !alive) {
                int out[K];
                if (m.transition(out)) {
#pragma unroll
                    for (int k = 0; k < K; ++k) s_stats[item * K + k] = out[k];
                    if constexpr (Machine::HAS_CACHE)
                        if (p.cache && p.mode != MODE_STATS)
                            m.store_cache((uint32_t*)(p.cache + (base + s_list[item]) * p.cache_stride));
                    item = PCGRL_NEXT_ITEM(item);
                    active = item < M;
                    if (active) m.init(s_bb + item * BBW, p, base + s_list[item]);
                } else {
                    alive = true;
                }
            }
        }
#elif PCGRL_CONVERGE
        // One trip = PCGRL_EXPAND_R board expansions, then the transition stream for the lanes whose frontier died.
        // The warp-wide vote at the top makes every lane start every trip together (finished lanes stay in the
        // loop until the whole warp is done): without it the lanes that loop back early run ahead of the ones in a
        // transition and the warp falls apart into groups that issue both streams separately.
        for (;;) {
            if (!__any_sync(0xffffffffu, active)) break;
            bool dead = false;
            if (active) {
                bool alive = m.expand();
#pragma unroll
                for (int r = 1; r < PCGRL_EXPAND_R; ++r)
                    if (alive) alive = m.expand();
                dead = !alive;
            }
            __syncwarp();
            int out[K];
            if (dead && m.transition(out)) {
#pragma unroll
                for (int k = 0; k < K; ++k) s_stats[item * K + k] = out[k];
                if constexpr (Machine::HAS_CACHE)
                    if (p.cache && p.mode != MODE_STATS)
                        m.store_cache((uint32_t*)(p.cache + (base + s_list[item]) * p.cache_stride));
                item = PCGRL_NEXT_ITEM(item);
                active = item < M;
                if (active) m.init(s_bb + item * BBW, p, base + s_list[item]);
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
                if constexpr (Machine::HAS_CACHE)
                    if (p.cache && p.mode != MODE_STATS)
                        m.store_cache((uint32_t*)(p.cache + (base + s_list[item]) * p.cache_stride));
                item = PCGRL_NEXT_ITEM(item);
                active = item < M;
                if (active) m.init(s_bb + item * BBW, p, base + s_list[item]);
            }
        }
#endif
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
