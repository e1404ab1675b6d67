// Split env-step path for the bit-board problems: three launches per step instead of the fused kernel's one,
// so that every lane of every search warp has a grid to work on.
//
//   k_split_act    thread-per-env: representation update (single-byte scatter into the int8 grid), counters, done,
//                  zero reward for unchanged envs; envs whose map changed are appended (CTA-aggregated atomic) to a
//                  GLOBAL work list as (env, cell).
//   k_split_stats  the stat searches over the work list.  Two variants:
//       static  (any bit-board machine): a warp takes 32 consecutive items, stages their grids as bit-boards in
//               shared memory cooperatively, then every lane runs one from-scratch search -- full warps whatever
//               the per-tile change count was (the fused kernel's tiles hold 128 +- 8 changed envs for its 128
//               threads: the overflow items ran as a second round at 1-3 lanes per warp).
//       incremental (binary): the one changed cell only touches the components next to it.  From the per-env
//               cache (passable board, one far tile per non-isolated component, a cell of the component that
//               attains path-length) the machine floods the affected region U, redoes calc_longest_path's two
//               sweeps inside U only, and re-sweeps the untouched components only when the one that held the
//               maximum was touched: 34 instead of 93 board expansions and 2.3 instead of 13.7 component
//               transitions per changed grid on random 16x16 maps, bit-identical results (helper.py:255-276 is a
//               per-component computation; components the edit does not touch keep their first tile, far tile
//               and eccentricity).  Work per item is uneven (0..135 expansions), so lanes claim items one by one
//               from their warp's slice of the list instead of running in batches.
//   k_split_out    thread-per-changed-env: fp64 loss(new) - loss(old), stats / reward / packed record; the list is
//                  handed from step to step through ping-pong counters (see the header layout below).
//   k_step_inc     all three in ONE launch for mid-size shards; k_split_stats_inc_multi / k_wait_list: the progressive
//                  host pipeline (measured, off by default).  step_lanegroup.cu holds the small-shard kernel.
//
// Reference path replaced: the same as step_bitboard.cu (envs/pcgrl_env.py:267-342, envs/reps/*_rep.py,
// envs/probs/binary/binary_prob.py:152-158, envs/probs/zelda/zelda_ctrl_prob.py:90-168, envs/helper.py:200-276,
// control_wrappers.py:216-244, 318-345).
#include <cstdlib>
#include "pcgrl_device.cuh"
#include "step_common.cuh"
#include "bitboard_machines.cuh"

namespace pcgrl {

#ifndef PCGRL_SPLIT_MIN_CTAS
#define PCGRL_SPLIT_MIN_CTAS 10
#endif
#ifndef PCGRL_INC_CTAS_PER_SM
#define PCGRL_INC_CTAS_PER_SM 8      // 128-thread CTAs of the incremental search resident per SM (A/B 4 / 6 / 8 / 12 at
                                     // R=3: 0.412 / 0.393 / 0.384 / 0.378 ms per step)
#endif
#ifndef PCGRL_INC_CTAS_PER_SM_HOST
#define PCGRL_INC_CTAS_PER_SM_HOST 4 // the same inside the host pipeline: most of the SM's registers stay free, so the
                                     // update / output kernels of the neighbouring chunks run next to the search
                                     // (e2e at 4 chunks, 2 / 3 / 4 / 5 / 8 CTAs per SM: 2.48 / 2.60 / 2.51 / 2.36 / 2.11e9 with the
                                     // copies on the compute streams; with separate copy streams 3 / 4 / 5 / 8:
                                     // 2.56 / 2.75 / 2.71 / 2.42e9, profiles/r02_e2e_cps_chunks_split.txt)
#endif
#ifndef PCGRL_PROG_CTAS_PER_SM
#define PCGRL_PROG_CTAS_PER_SM 6
#endif
#ifndef PCGRL_INC_MIN_CLAIM
#define PCGRL_INC_MIN_CLAIM 12       // lanes that must be waiting before a warp hands out new items: the claim / init
                                     // stream then runs with that many lanes (A/B 1 / 4 / 8 at R=3: 0.412 / 0.390 / 0.387; at
                                     // R=12, 8 / 12 / 16 / 24: 0.2978 / 0.2996 / 0.3022 / 0.3263 ms, profiles/r02_ab_inc_claim_window.txt)
#endif
#ifndef PCGRL_INC_DYNAMIC
#define PCGRL_INC_DYNAMIC 0          // > 0: items per fetch from the global list counter (see k_split_stats_inc); 0: static slices.
                                     // A/B static / 16 / 32 / 64 (400 steps, 1 Mi envs, profiles/r02_ab_inc_dynamic.txt): 0.300-0.304 /
                                     // 0.304 / 0.295 / 0.313 ms per step; as the default (32) the 800-step bench line did not move
                                     // (0.2909 against 0.2913 ms) and e2e fell 2.60 -> 2.51e9: the tail is not where the time goes.  Off.
#endif
#ifndef PCGRL_INC_EXPAND_R
#define PCGRL_INC_EXPAND_R 12        // board expansions per trip.  Every trip also issues each transition stream that
                                     // some lane needs (flood done / next component / second sweep / re-sweep / finish,
                                     // ~580 instructions at 2-3 lanes), so few long trips beat many short ones although
                                     // lanes idle inside the window: A/B 2 / 3 / 4 / 6 / 8 / 12 -> 0.466 / 0.412 / 0.356 /
                                     // 0.325 / 0.308 / 0.303 ms per step (1 Mi envs, 8 CTAs per SM); 16 / 20 -> 0.3018 / 0.3140
                                     // against 0.2996 for 12 in a later run
#endif
#ifndef PCGRL_CONVERGE
#define PCGRL_CONVERGE 1             // 1: every lane of a warp starts every search trip together (vote at the loop head)
#endif
#ifndef PCGRL_THETA
#define PCGRL_THETA 0                // static search: expand while >= THETA/32 of the running lanes are alive (0: fixed
                                     // windows).  A/B 8 / 12 / 16 / 20 / 24: 0.59 / 0.50 / 0.45 / 0.43 / 0.42 ms per step
                                     // against 0.38 with fixed windows of 3 -- waiting lanes cost more than the
                                     // fuller transition stream saves
#endif
#ifndef PCGRL_INC_THETA
#define PCGRL_INC_THETA 0            // the same for the incremental search (16: 0.334 ms against 0.303 with R=12)
#endif
#ifndef PCGRL_EXPAND_R_SPLIT
#define PCGRL_EXPAND_R_SPLIT 3       // board expansions per trip of the static search (as in the fused kernel)
#endif

constexpr int ACT_THREADS = 256;
constexpr int STAT_THREADS = 128;
// threads per CTA of the static search: as many warps as keep the staged bit-boards under 48 KB of shared memory
__host__ __device__ constexpr int stat_threads(int bbw) { return bbw <= 96 ? 128 : 64; }
constexpr int OUT_THREADS = 256;
// Work list of one launch: p.wl_hdr = 16 header ints, p.worklist = the body: (env, cell) pairs for up to n_envs items, then n_stats new
// stats per item.  The headers of all pipeline chunks live together at the front of pcgrl_state.worklist, where no
// body ever lands, so they stay zero whatever chunking the previous step used.
// Header (all zero before the first step): [0], [1] item counters used alternately, [3] the step count, whose parity
// picks the counter, [4] the step's final item count, [5] search warps that reported the list (progressive host
// pipeline), [6], [7] fetch counters of the dynamic distribution, alternating like [0], [1].  k_split_act adds to counter [par]; the search kernel clears
// counter [par ^ 1] for the next step and publishes the count in [4]; k_split_out reads [4] and bumps [3] -- every
// word is written by exactly one thread of a kernel none of whose CTAs reads it, so the three launches need no
// "last CTA out" atomics or fences to hand the list over.
__device__ __forceinline__ int wl_parity(const KParams& p) { return *(volatile const int*)(p.wl_hdr + 3) & 1; }
__device__ __forceinline__ int wl_count_search(const KParams& p) {
    const int par = wl_parity(p);
    const int count = *(volatile const int*)(p.wl_hdr + par);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.wl_hdr[par ^ 1] = 0;
        p.wl_hdr[6 + (par ^ 1)] = 0;    // the next step's fetch counter (k_split_stats_inc, dynamic distribution)
        p.wl_hdr[4] = count;
    }
    return count;
}
__device__ __forceinline__ int2* wl_items(const KParams& p) { return (int2*)p.worklist; }
__device__ __forceinline__ int32_t* wl_stats(const KParams& p) { return p.worklist + 2 * p.n_envs; }

// ------------------------------------------------------------------------------------------------
// k_split_act
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ACT_THREADS) k_split_act(const KParams p) {
    __shared__ int s_warp[ACT_THREADS / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t gid = (int64_t)blockIdx.x * ACT_THREADS + tid;
    bool need = false;
    int cell = -1;
    if (gid < p.n_envs) need = step_counters(p, gid, apply_action(p, gid, &cell));
    const unsigned bal = __ballot_sync(0xffffffffu, need);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < ACT_THREADS / 32; ++w) {
            const int c = s_warp[w];
            s_warp[w] = tot;
            tot += c;
        }
        s_base = tot ? atomicAdd(p.wl_hdr + wl_parity(p), tot) : 0;
    }
    __syncthreads();
    if (need) wl_items(p)[s_base + s_warp[warp] + __popc(bal & ((1u << lane) - 1u))] = make_int2((int)gid, cell);
}

// ------------------------------------------------------------------------------------------------
// k_split_stats, static variant: one from-scratch search per lane, 32 consecutive list items per warp
// ------------------------------------------------------------------------------------------------
template <class Machine, int NW, bool TWO>
__global__ void __launch_bounds__(stat_threads(Machine::Prob::P * NW), NW <= 8 ? PCGRL_SPLIT_MIN_CTAS : 1)
k_split_stats(const KParams p) {
    using Prob = typename Machine::Prob;
    constexpr int P = Prob::P;
    constexpr int K = Prob::K;
    constexpr int BBW = P * NW;
    constexpr int THREADS = stat_threads(BBW);
    __shared__ uint32_t s_bb[THREADS * BBW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int count = wl_count_search(p);
    const int first = (blockIdx.x * (THREADS / 32) + warp) * 32;
    if (first >= count) return;   // warps are independent: no CTA-wide barrier anywhere in this kernel
    const int n_here = min(32, count - first);
    const int2* items = wl_items(p) + first;
    uint32_t* wbb = s_bb + warp * 32 * BBW;

    // stage the warp's grids as per-plane bit-boards (same packing as phase B of the fused kernel)
    {
        const int chunks = p.row_stride / 16;
        const int W = p.d1;
        const bool word_is_32_bytes = TWO ? (W == 16) : (W == 32);
        if (word_is_32_bytes) {
            for (int i = lane; i < n_here * NW; i += 32) {
                const int slot = i / NW, j = i - slot * NW;
                const int8_t* g = p.grids + (int64_t)items[slot].x * p.row_stride + j * 32;
                uint32_t lo[P], hi[P];
#pragma unroll
                for (int q = 0; q < P; ++q) lo[q] = hi[q] = 0;
                if (2 * j < chunks) pack16<Prob>(*(const uint4*)g, lo);
                if (2 * j + 1 < chunks) pack16<Prob>(*(const uint4*)(g + 16), hi);
#pragma unroll
                for (int q = 0; q < P; ++q) wbb[slot * BBW + q * NW + j] = lo[q] | (hi[q] << 16);
            }
        } else {
            for (int i = lane; i < n_here * BBW; i += 32) wbb[i] = 0;
            __syncwarp();
            for (int i = lane; i < n_here * chunks; i += 32) {
                const int slot = i / chunks, c = i - slot * chunks;
                uint32_t lin[P];
                pack16<Prob>(*(const uint4*)(p.grids + (int64_t)items[slot].x * p.row_stride + c * 16), lin);
                uint32_t* bb = wbb + slot * BBW;
                int cell = c * 16;
                const int end = min(cell + 16, p.cells);
                int y = cell / W, x = cell - y * W;
                while (cell < end) {
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
    __syncwarp();
#if !PCGRL_CONVERGE
    if (lane >= n_here) return;
#endif
    const bool mine = lane < n_here;
    const int64_t env = mine ? items[lane].x : 0;
    Machine m;
    if (mine) m.init(wbb + lane * BBW, p, env);
    int out[K];
#if PCGRL_CONVERGE
    // One trip = PCGRL_EXPAND_R_SPLIT board expansions, then the transition stream for the lanes whose frontier
    // died.  The warp-wide vote at the top makes every lane start every trip together: without it the lanes that
    // `continue` run ahead of the ones in a transition, the warp falls apart into groups that never re-merge, and
    // both streams are issued once per group (ncu: transitions at 3.7 lanes, expansions at 19).
    bool running = mine;
#if PCGRL_THETA > 0
    // Rounds instead of fixed windows: the warp keeps expanding while at least PCGRL_THETA/32 of its running lanes
    // still have a live frontier (the others wait), then every waiting lane takes its transition in ONE pass of
    // the transition stream -- both instruction streams run with many lanes instead of one of them with a few.
    bool alive = false;   // a fresh machine has an empty frontier: its first act is a transition
    for (;;) {
        const unsigned run = __ballot_sync(0xffffffffu, running);
        if (!run) break;
        const int need = max(1, (__popc(run) * PCGRL_THETA) >> 5);
        for (;;) {
            if (__popc(__ballot_sync(0xffffffffu, running && alive)) < need) break;
            if (running && alive) alive = m.expand();
        }
        if (running && !alive) {
            if (m.transition(out)) running = false;
            else alive = true;
        }
    }
#else
    for (;;) {
        if (!__any_sync(0xffffffffu, running)) break;
        bool dead = false;
        if (running) {
            bool alive = m.expand();
#pragma unroll
            for (int r = 1; r < PCGRL_EXPAND_R_SPLIT; ++r)
                if (alive) alive = m.expand();
            dead = !alive;
        }
        __syncwarp();
        if (dead && m.transition(out)) running = false;
    }
#endif
    if (!mine) return;
#else
    for (;;) {
        bool alive = m.expand();
#pragma unroll
        for (int r = 1; r < PCGRL_EXPAND_R_SPLIT; ++r)
            if (alive) alive = m.expand();
        if (alive) continue;
        if (m.transition(out)) break;
    }
#endif
    int32_t* o = wl_stats(p) + (int64_t)(first + lane) * K;
#pragma unroll
    for (int k = 0; k < K; ++k) o[k] = out[k];
    if constexpr (Machine::HAS_CACHE)
        if (p.cache) m.store_cache((uint32_t*)(p.cache + env * p.cache_stride));
}

// ------------------------------------------------------------------------------------------------
// BinaryIncMachine: regions / path-length of a binary map after a one-cell edit, from the per-env cache.
//
// calc_num_regions (helper.py:200-210) counts components; calc_longest_path (:255-276) takes, per component in
// row-major order of its first tile, BFS(first tile) -> far tile (np.argmax: row-major-first maximum) -> BFS(far)
// and keeps the largest second-sweep maximum.  Both are sums / maxima of per-component terms, and a component the
// edit does not touch keeps its first tile, far tile and eccentricity.  With D+ the cells that became passable, D-
// the cells that became solid, P' the new passable board:
//   U  = flood fill in P' from D+ and from the passable neighbours of D-   (every new component that is affected)
//   A  = U | D-                                          (covers every OLD component that is affected, entirely)
//   regions' = regions - #old components in A + #components of U
//              (#old in A = far tiles in A + old isolated cells in A: each non-isolated component owns one far tile)
//   path'    = max(two-sweep maximum inside U, maximum over the untouched components), where the latter is the
//              old path-length if the cached cell of a maximal component lies outside A, else the joint second
//              sweep from the remaining far tiles.
// Validated bit-for-bit against the from-scratch restatement on random rollouts before it was written for the GPU
// (and on the GPU by tests/test_gpu_split.py: every step of every env equals pcgrl_stats recomputed from scratch).
// ------------------------------------------------------------------------------------------------
template <int NW, bool TWO>
struct BinaryIncMachine {
    using B = Board<NW, TWO>;
    static constexpr int SMEM_WORDS = 3 * NW;   // per thread: P' | U (non-isolated part) | far tiles of the rest
    uint32_t avail[NW], front[NW], fars[NW];
    uint32_t* sm;
    const uint32_t* cache;   // this env's cache row (still the OLD state until finish())
    int phase, level, regions, lu, mcu, mold, lold;
    bool hit;

    static __device__ __forceinline__ int lowest_pos(const uint32_t (&x)[NW]) {
        int pos = -1;   // branch-free: this runs in the few-lane transition stream
#pragma unroll
        for (int i = NW - 1; i >= 0; --i) {
            const int f = __ffs(x[i]);
            pos = f ? i * 32 + f - 1 : pos;
        }
        return pos;
    }

    // bitpos: bit position of the edited cell in the board (word * 32 + bit)
    __device__ __forceinline__ void init(uint32_t* smem, const uint32_t* c, int bitpos, int regions_old, int path_old) {
        sm = smem;
        cache = c;
        regions = regions_old;
        lold = path_old;
        const uint4* c4 = (const uint4*)c;
        uint32_t pn[NW], d[NW], seed[NW];
        static_assert(NW % 4 == 0 || NW < 4, "cache rows are read as 128-bit vectors");
        if constexpr (NW >= 4) {
#pragma unroll
            for (int i = 0; i < NW / 4; ++i) {
                const uint4 v = c4[i];
                pn[4 * i] = v.x, pn[4 * i + 1] = v.y, pn[4 * i + 2] = v.z, pn[4 * i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NW; ++i) pn[i] = c[i];
        }
        mold = (int)c[2 * NW];
        const int w = bitpos >> 5;
        const uint32_t bit = 1u << (bitpos & 31);
        uint32_t now_pass = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            d[i] = i == w ? bit : 0u;
            pn[i] ^= d[i];
            now_pass |= pn[i] & d[i];
            sm[i] = pn[i];
        }
        // seed of the flood: the cell itself if it became passable, else its passable neighbours
        B::expand_and(d, pn, seed);
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            front[i] = now_pass ? d[i] : seed[i];
            avail[i] = pn[i] & ~front[i];
            fars[i] = 0;
        }
        phase = 0;
        level = 0;
        lu = 0;
        mcu = -1;
        hit = false;
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
        return true;
    }
    // the frontier died; returns true when the search is over (call finish())
    __device__ __forceinline__ bool transition() {
        if (phase == 0) {
            // flood done: U = P' \\ avail.  Account for the old components it swallowed, set up the sweeps inside U.
            uint32_t pn[NW], po[NW], ones[NW], nb[NW], nbo[NW];
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                pn[i] = sm[i];
                po[i] = cache[i];
                ones[i] = 0xFFFFFFFFu;
            }
            B::expand_and(pn, ones, nb);
            B::expand_and(po, ones, nbo);
            int k_old = 0, k_iso = 0;
            uint32_t hitm = 0;
            const int mw = mold >> 5;
            const uint32_t mb = mold >= 0 ? 1u << (mold & 31) : 0u;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const uint32_t u = pn[i] & ~avail[i];
                const uint32_t a = u | (po[i] & ~pn[i]);           // U | D-
                const uint32_t fo = cache[NW + i];
                k_old += __popc(fo & a) + __popc(po[i] & ~nbo[i] & a);
                sm[2 * NW + i] = fo & ~a;                          // far tiles of the untouched components
                if (i == mw) hitm |= a & mb;
                const uint32_t iso = u & ~nb[i];
                k_iso += __popc(iso);
                avail[i] = u & ~iso;
                sm[NW + i] = avail[i];
                front[i] = 0;
            }
            hit = hitm != 0;
            regions += k_iso - k_old;
            phase = 1;
        }
        if (phase == 1) {   // first sweeps inside U, one component at a time (BinaryMachine phase 0)
            uint32_t t[NW];
            B::minus_one(front, t);
#pragma unroll
            for (int i = 0; i < NW; ++i) fars[i] |= front[i] & ~t[i];
            if (B::minus_one(avail, t)) {
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    front[i] = avail[i] & ~t[i];
                    avail[i] &= t[i];
                }
                ++regions;
                return false;
            }
            phase = 2;      // joint second sweep inside U
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                front[i] = fars[i];
                avail[i] = sm[NW + i] & ~fars[i];
                any |= fars[i];
            }
            level = 0;
            if (any) return false;
        }
        if (phase == 2) {
            lu = level;
            mcu = lowest_pos(front);
            if (hit) {      // the component that held the maximum was touched: re-sweep the untouched ones
                uint32_t any = 0;
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    front[i] = sm[2 * NW + i];
                    avail[i] = sm[i] & ~sm[NW + i] & ~front[i];   // isolated cells may stay: never reached
                    any |= front[i];
                }
                phase = 3;
                level = 0;
                lold = 0;
                mold = -1;
                if (any) return false;
                return true;
            }
            return true;
        }
        // phase 3 died
        lold = level;
        mold = lowest_pos(front);
        return true;
    }
    // results + the new cache row
    __device__ __forceinline__ void finish(int* out, uint32_t* c) const {
        out[0] = regions;
        out[1] = max(lu, lold);
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            c[i] = sm[i];
            c[NW + i] = sm[2 * NW + i] | fars[i];
        }
        c[2 * NW] = (uint32_t)(lu > lold ? mcu : mold);
    }
};

// ------------------------------------------------------------------------------------------------
// k_split_stats, incremental variant: persistent-size grid, every warp owns a contiguous slice of the work list
// and its lanes claim items from it one by one (ballot-ranked at a convergent point, no atomics).
// ------------------------------------------------------------------------------------------------
template <int NW, bool TWO>
__global__ void __launch_bounds__(STAT_THREADS) k_split_stats_inc(const KParams p) {
    using M = BinaryIncMachine<NW, TWO>;
    __shared__ uint32_t s_work[STAT_THREADS * M::SMEM_WORDS];
    const int tid = threadIdx.x, lane = tid & 31;
    const int count = wl_count_search(p);
#if PCGRL_INC_DYNAMIC > 0
    // Dynamic distribution: a warp fetches the next PCGRL_INC_DYNAMIC items of the list from a global counter (header
    // words 6 / 7, alternating with the step parity like the item counters; block 0 clears the one the NEXT step uses)
    // whenever its lanes need items.  With static slices of ~110 items whose cost varies 0..135 expansions each, the
    // slowest of the 4 736 warps carries ~25 % more work than the average one and the kernel ends with a thin tail.
    int* ctr = p.wl_hdr + 6 + wl_parity(p);     // (wl_count_search cleared it one step ago, whichever split kernel ran)
    int lo = 0, hi = 0;
    bool drained = count == 0, last_batch = false;
#else
    const int n_warps = gridDim.x * (STAT_THREADS / 32);
    const int gw = blockIdx.x * (STAT_THREADS / 32) + (tid >> 5);
    const int per = (count + n_warps - 1) / n_warps;
    int lo = min(count, gw * per);
    const int hi = min(count, lo + per);
    if (lo >= hi) return;
#endif
    const int2* items = wl_items(p);
    int32_t* stats_out = wl_stats(p);
    const int W = p.d1;
    M m;
    bool active = false, alive = false;
    int item = 0;
    int64_t env = 0;
    for (;;) {
        unsigned need = __ballot_sync(0xffffffffu, !active);
#if PCGRL_INC_DYNAMIC > 0
        // up to two rounds per trip: lanes left over when a batch runs out take theirs from the next one at once
        for (int round = 0; round < 2 && need; ++round) {
            if (!(round || __popc(need) >= PCGRL_INC_MIN_CLAIM || need == 0xffffffffu || (last_batch && lo < hi))) break;
            if (lo >= hi) {
                if (drained) break;
                int b = 0;
                if (lane == 0) b = atomicAdd(ctr, PCGRL_INC_DYNAMIC);
                b = __shfl_sync(0xffffffffu, b, 0);
                lo = min(count, b);
                hi = min(count, b + PCGRL_INC_DYNAMIC);
                last_batch = b + PCGRL_INC_DYNAMIC >= count;      // the counter is past the end: nothing after this batch
                if (lo >= hi) {
                    drained = true;
                    break;
                }
            }
#else
        if (need && lo < hi && (__popc(need) >= PCGRL_INC_MIN_CLAIM || need == 0xffffffffu || hi - lo < PCGRL_INC_MIN_CLAIM)) {
#endif
            const int mine = lo + __popc(need & ((1u << lane) - 1u));
            if (!active && mine < hi) {
                item = mine;
                const int2 it = items[item];
                env = it.x;
                const int y = it.y / W, x = it.y - y * W;
                const int bitpos = TWO ? y * 16 + x : y * 32 + x;
                const int32_t* st = p.stats + env * 2;
                m.init(s_work + tid * M::SMEM_WORDS, (const uint32_t*)(p.cache + env * p.cache_stride), bitpos, st[0], st[1]);
                active = true;
                alive = true;
            }
            lo = min(hi, lo + __popc(need));
#if PCGRL_INC_DYNAMIC > 0
            if (lo < hi) break;                                    // the batch served every idle lane
            need = __ballot_sync(0xffffffffu, !active);
#endif
        }
        const unsigned act = __ballot_sync(0xffffffffu, active);
#if PCGRL_INC_DYNAMIC > 0
        if (!act) {
            if (drained) break;
            continue;          // every lane is idle: the next trip fetches a batch
        }
#else
        if (!act) break;
#endif
#if PCGRL_INC_THETA > 0
        const int want = max(1, (__popc(act) * PCGRL_INC_THETA) >> 5);
        for (;;) {
            if (__popc(__ballot_sync(0xffffffffu, active && alive)) < want) break;
            if (active && alive) alive = m.expand();
        }
        if (active && !alive) {
            if (m.transition()) {
                int out[2];
                m.finish(out, (uint32_t*)(p.cache + env * p.cache_stride));
                *(int2*)(stats_out + (int64_t)item * 2) = make_int2(out[0], out[1]);
                active = false;
            } else {
                alive = true;
            }
        }
#else
        if (active) {
            bool alive = m.expand();
#pragma unroll
            for (int r = 1; r < PCGRL_INC_EXPAND_R; ++r)
                if (alive) alive = m.expand();
            if (!alive && m.transition()) {
                int out[2];
                m.finish(out, (uint32_t*)(p.cache + env * p.cache_stride));
                *(int2*)(stats_out + (int64_t)item * 2) = make_int2(out[0], out[1]);
                active = false;
            }
        }
#endif
    }
}

// ------------------------------------------------------------------------------------------------
// k_split_stats_inc_multi: ONE incremental search over the work lists of every chunk of the host pipeline.
//
// The chunked pipeline used to run one search launch per chunk on rotating streams; every launch has a tail in which
// its lanes run dry, and concurrent launches share the SMs so that the first chunk's result was ready only after 2/3
// of the whole step (profiles/r02_host_trace.txt: kernels done at 190 / 270 / 274 / 361 us).  Here every warp walks
// the lists in chunk order -- its slice of list 0, then of list 1, ... -- and its lanes keep claiming across the list
// boundaries, so the SIMT fill is that of the whole-shard search while chunk c is finished after about (c + 1) / n of
// the kernel.  A warp reports a list (atomicAdd on header word 5, after a fence) once it has claimed its whole
// slice of it and none of its lanes still works on one of its items; when all warps have reported, the chunk's
// k_wait_list lets the output kernel and the download go (k_wait_list runs on another stream while this kernel is
// still busy with the later lists).
// ------------------------------------------------------------------------------------------------
template <int NW, bool TWO>
__global__ void __launch_bounds__(STAT_THREADS, NW >= 8 ? PCGRL_PROG_CTAS_PER_SM : 1) k_split_stats_inc_multi(const KParams p) {
    using M = BinaryIncMachine<NW, TWO>;
    __shared__ uint32_t s_work[STAT_THREADS * M::SMEM_WORDS];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int n_lists = p.ml_lists;
    const int n_warps = gridDim.x * (STAT_THREADS / 32);
    const int gw = blockIdx.x * (STAT_THREADS / 32) + (tid >> 5);
    if (blockIdx.x == 0 && tid < n_lists) {   // publish every list's count, clear the other counter (as wl_count_search)
        int32_t* hdr = p.wl_hdr + 16 * tid;
        const int par = *(volatile const int*)(hdr + 3) & 1;
        hdr[4] = *(volatile const int*)(hdr + par);
        hdr[par ^ 1] = 0;
        hdr[6 + (par ^ 1)] = 0;
    }
    const int W = p.d1;
    int cl = -1, lo = 0, hi = 0, sig = 0;                  // warp-uniform: list being claimed from, slice, next list to report
    M m;
    bool active = false;
    int my_list = 0, my_item = 0;                          // per lane; pointers are rebuilt from them (registers)
    for (;;) {
        unsigned need = __ballot_sync(0xffffffffu, !active);
        if (need && (__popc(need) >= PCGRL_INC_MIN_CLAIM || need == 0xffffffffu || (cl == n_lists - 1 && hi - lo < PCGRL_INC_MIN_CLAIM))) {
            for (;;) {
                while (lo >= hi && cl + 1 < n_lists) {     // next list: this warp's slice of it
                    ++cl;
                    const int32_t* hdr = p.wl_hdr + 16 * cl;
                    const int par = *(volatile const int*)(hdr + 3) & 1;
                    const int count = *(volatile const int*)(hdr + par);
                    const int per_w = (count + n_warps - 1) / n_warps;
                    lo = min(count, gw * per_w);
                    hi = min(count, lo + per_w);
                }
                if (lo >= hi) break;                       // every list is claimed
                const int mine = lo + __popc(need & lt);
                if (!active && mine < hi) {
                    const int64_t first = (int64_t)cl * p.ml_per;
                    const int2 it = ((const int2*)(p.worklist + 4 * first))[mine];
                    const int y = it.y / W, x = it.y - y * W;
                    const int bitpos = TWO ? y * 16 + x : y * 32 + x;
                    const int32_t* st = p.stats + 2 * (first + it.x);
                    m.init(s_work + tid * M::SMEM_WORDS, (const uint32_t*)(p.cache + (first + it.x) * p.cache_stride), bitpos, st[0], st[1]);
                    my_item = mine;
                    my_list = cl;
                    active = true;
                }
                lo = min(hi, lo + __popc(need));
                need = __ballot_sync(0xffffffffu, !active);
                if (!need || lo < hi) break;               // lanes left over at the end of a slice go on to the next list
            }
        }
        const unsigned act = __ballot_sync(0xffffffffu, active);
        // lists below the oldest one that is still in flight or not claimed to the end are finished for this warp
        const int oldest = __reduce_min_sync(0xffffffffu, active ? my_list : (lo < hi ? cl : cl + 1));
        while (sig < oldest) {
            __threadfence();                               // every lane's stats / cache rows of that list
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                atomicAdd(p.wl_hdr + 16 * sig + 5, 1);
            }
            ++sig;
        }
        if (!act) {
            if (lo >= hi && cl + 1 >= n_lists) break;
            continue;
        }
        if (active) {
            bool alive = m.expand();
#pragma unroll
            for (int r = 1; r < PCGRL_INC_EXPAND_R; ++r)
                if (alive) alive = m.expand();
            if (!alive && m.transition()) {
                int out[2];
                m.finish(out, const_cast<uint32_t*>(m.cache));
                const int64_t first = (int64_t)my_list * p.ml_per;
                const int64_t cap = min(p.ml_per, p.n_envs - first);
                *(int2*)(p.worklist + 4 * first + 2 * cap + 2 * (int64_t)my_item) = make_int2(out[0], out[1]);
                active = false;
            }
        }
    }
}

// One thread waits until every search warp has reported the chunk's list, then re-arms the counter for the next
// step.  Launched AFTER the search kernel (so it can never sit in front of it in a hardware queue) on the stream
// that carries the chunk's output kernel and download.  A wait of more than two seconds sets status bit 5.
__global__ void k_wait_list(int32_t* hdr, int target, int32_t* status) {
    if (threadIdx.x != 0) return;
    volatile int* f = hdr + 5;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*f < target) {
        __nanosleep(400);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) {
            if (status) atomicOr(status, 32);
            break;
        }
    }
    __threadfence();
    *f = 0;
}

// ------------------------------------------------------------------------------------------------
// k_step_inc: the whole binary env step in ONE launch around the incremental search (no global work list).
// Every warp owns a contiguous slice of the shard's envs and loops over three kinds of work, each issued with
// (nearly) all 32 lanes:
//   update  the next 32 envs of the slice: representation update, counters, done, zero reward for unchanged
//           maps (thread-per-env, as k_split_act); the changed ones go into the warp's item ring in shared memory;
//   search  lanes pick items from the ring as they become free (at least PCGRL_INC_MIN_CLAIM at a time) and run
//           BinaryIncMachine trips; a finished lane writes its env's cache row and parks (env, regions,
//           path-length) in the warp's result ring;
//   output  once 32 results wait (or nothing else is left): one lane per result does the fp64 reward and writes
//           stats / reward / packed record (as k_split_out).
// The update and output passes are memory-latency work; in the split path they were two separate kernels (42 +
// 30 us of the 303 us step at 1 Mi envs) during which the ALU pipes idled, here other warps' searches cover them.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ bool inc_update_env(const KParams& p, int64_t gid, int* cell) {
    return step_counters(p, gid, apply_action(p, gid, cell));
}
__device__ __noinline__ void inc_output_env(const KParams& p, int64_t env, int regions, int path) {
    const int32_t nw[2] = {regions, path};
    finish_env<2>(p, env, nw);
}

#ifndef PCGRL_STEP_INC_MIN_CTAS
#define PCGRL_STEP_INC_MIN_CTAS 6
#endif
#ifndef PCGRL_STEP_INC_ROUNDS
#define PCGRL_STEP_INC_ROUNDS 1      // update rounds (of 32 envs) per warp the grid is sized for, see launch_split
#endif
template <int NW, bool TWO>
__global__ void __launch_bounds__(STAT_THREADS, PCGRL_STEP_INC_MIN_CTAS) k_step_inc(const __grid_constant__ KParams p) {
    using M = BinaryIncMachine<NW, TWO>;
    constexpr int WARPS = STAT_THREADS / 32;
    constexpr int RING = 64;
    __shared__ uint32_t s_work[STAT_THREADS * M::SMEM_WORDS];
    __shared__ int2 s_items[WARPS][RING];     // (env, cell) waiting for a lane
    __shared__ int4 s_res[WARPS][RING];       // (env, regions, path-length, -) waiting for the output pass
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t n_warps = (int64_t)gridDim.x * WARPS;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t per = ((p.n_envs + n_warps - 1) / n_warps + 31) / 32 * 32;
    int64_t e_next = min(p.n_envs, gw * per);
    const int64_t e_hi = min(p.n_envs, e_next + per);
    if (e_next >= e_hi) return;
    int2* items = s_items[warp];
    int4* res = s_res[warp];
    const int W = p.d1;
    int it_head = 0, it_cnt = 0, res_head = 0, res_cnt = 0;     // warp-uniform
    M m;
    bool active = false;
    int64_t env = 0;
    for (;;) {
        // ---- update: refill the item ring from the warp's env slice
        if (it_cnt < 32 && e_next < e_hi) {
            const int64_t gid = e_next + lane;
            int cell = -1;
            const bool need = gid < e_hi && inc_update_env(p, gid, &cell);
            const unsigned bal = __ballot_sync(0xffffffffu, need);
            if (need) items[(it_head + it_cnt + __popc(bal & lt)) & (RING - 1)] = make_int2((int)gid, cell);
            it_cnt += __popc(bal);
            e_next += 32;
            __syncwarp();
        }
        const bool drained = e_next >= e_hi;
        // ---- hand items to free lanes
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle && it_cnt > 0 && (__popc(idle) >= PCGRL_INC_MIN_CLAIM || idle == 0xffffffffu || drained)) {
            const int rank = __popc(idle & lt);
            if (!active && rank < it_cnt) {
                const int2 it = items[(it_head + rank) & (RING - 1)];
                env = it.x;
                const int y = it.y / W, x = it.y - y * W;
                const int bitpos = TWO ? y * 16 + x : y * 32 + x;
                const int32_t* st = p.stats + env * 2;
                m.init(s_work + tid * M::SMEM_WORDS, (const uint32_t*)(p.cache + env * p.cache_stride), bitpos, st[0], st[1]);
                active = true;
            }
            const int taken = min(__popc(idle), it_cnt);
            it_head += taken;
            it_cnt -= taken;
        }
        const unsigned act = __ballot_sync(0xffffffffu, active);
        // ---- output: 32 results at a time, or whatever is left at the end
        if (res_cnt >= 32 || (res_cnt > 0 && !act && it_cnt == 0 && drained)) {
            const int n_out = min(32, res_cnt);
            if (lane < n_out) {
                const int4 r = res[(res_head + lane) & (RING - 1)];
                inc_output_env(p, r.x, r.y, r.z);
            }
            res_head += n_out;
            res_cnt -= n_out;
            __syncwarp();
        }
        if (!act) {
            if (it_cnt == 0 && drained && res_cnt == 0) break;
            continue;
        }
        // ---- search: one trip of every active lane
        bool finished = false;
        int out[2] = {0, 0};
        if (active) {
            bool alive = m.expand();
#pragma unroll
            for (int r = 1; r < PCGRL_INC_EXPAND_R; ++r)
                if (alive) alive = m.expand();
            if (!alive && m.transition()) {
                m.finish(out, (uint32_t*)(p.cache + env * p.cache_stride));
                finished = true;
                active = false;
            }
        }
        const unsigned fin = __ballot_sync(0xffffffffu, finished);
        if (finished) res[(res_head + res_cnt + __popc(fin & lt)) & (RING - 1)] = make_int4((int)env, out[0], out[1], 0);
        res_cnt += __popc(fin);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k_split_out
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(OUT_THREADS) k_split_out(const KParams p) {
    const int count = *(volatile const int*)(p.wl_hdr + 4);
    const int i = blockIdx.x * OUT_THREADS + threadIdx.x;
    if (i == 0) p.wl_hdr[3] = p.wl_hdr[3] + 1;   // the next step uses the other counter (no CTA of this kernel reads [3])
    if (i < count) {
        const int64_t env = wl_items(p)[i].x;
        int32_t nw[K];
        const int32_t* o = wl_stats(p) + (int64_t)i * K;
#pragma unroll
        for (int k = 0; k < K; ++k) nw[k] = o[k];
        finish_env<K>(p, env, nw);
    }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
// resident CTAs per SM of the persistent incremental kernels; PCGRL_INC_CPS overrides the compiled default (A/B)
static int inc_ctas_per_sm(bool host_chunk) {
    static int env_v = -1;
    if (env_v < 0) {
        const char* e = getenv("PCGRL_INC_CPS");
        env_v = e ? atoi(e) : 0;
    }
    if (env_v > 0) return env_v;
    return host_chunk ? PCGRL_INC_CTAS_PER_SM_HOST : PCGRL_INC_CTAS_PER_SM;
}
static int g_n_sm = 0;
static cudaError_t sm_count(int& n) {
    if (!g_n_sm) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&g_n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    }
    n = g_n_sm;
    return cudaSuccess;
}

// CTAs of the progressive pipeline's search: one short of filling the SMs' register files (8 CTAs of 64 registers x
// 128 threads), so that the chunks' wait / output kernels become resident beside it.  PCGRL_PROG_CPS overrides (A/B).
static int64_t prog_search_ctas(int64_t n, int n_sm) {
    static int env_v = -1;
    if (env_v < 0) {
        const char* e = getenv("PCGRL_PROG_CPS");
        env_v = e ? atoi(e) : 0;
    }
    const int64_t want = (n + STAT_THREADS - 1) / STAT_THREADS;
    const int64_t cap = (int64_t)n_sm * (env_v > 0 ? env_v : PCGRL_PROG_CTAS_PER_SM);
    return want < cap ? want : cap;
}
// search warps of that launch (= reports every chunk's k_wait_list waits for); < 0 on error
int split_prog_search_warps(int64_t n) {
    int n_sm = 0;
    if (sm_count(n_sm) != cudaSuccess) return -1;
    return (int)prog_search_ctas(n, n_sm) * (STAT_THREADS / 32);
}
// can this config run the progressive host pipeline (binary, one-cell edits, maps <= 16x16, cache + work list)?
bool split_prog_supported(const KParams& p, int problem) {
    return problem == PCGRL_PROB_BINARY && p.ndim == 2 && p.d0 <= 16 && p.d1 <= 16 && p.cache && p.worklist &&
           p.rep != PCGRL_REP_CELLULAR && p.action_kind != PCGRL_ACT_PATCH;
}

template <class Machine, int NW, bool TWO>
static cudaError_t launch_split(const KParams& p, cudaStream_t s, int incremental) {
    constexpr int K = Machine::Prob::K;
    const int64_t n = p.n_envs;
    if (n == 0) return cudaSuccess;
    cudaError_t e;
    if constexpr (Machine::HAS_CACHE && TWO) {
        if (incremental == 2) {     // the whole step in one launch (k_step_inc)
            int n_sm = 0;
            if ((e = sm_count(n_sm)) != cudaSuccess) return e;
            // envs per warp = 32 * m.  A warp's first 32 envs yield ~16 changed ones, i.e. half-empty warps for the whole
            // search; with two update rounds per warp the ring refills and every lane gets an item.  And the grid is
            // sized so that EVERY CTA gets work: with the plain cap (6 CTAs per SM) a 128 Ki-env shard ran 64 envs per warp
            // on 512 of its 888 CTAs while the others exited at once.  A/B of m (profiles/r02_step_inc_rounds.txt, kernel
            // ms, m = 1 / 2 / 3 / 4): 32 Ki envs 0.042 / 0.048 / 0.062 / 0.075, 64 Ki 0.060 / 0.056 / 0.070 / 0.077,
            // 128 Ki 0.095 / 0.075 / 0.083 / 0.089.  PCGRL_STEP_INC_ROUNDS overrides.
            static int rounds = -1;
            if (rounds < 0) {
                const char* ev = getenv("PCGRL_STEP_INC_ROUNDS");
                rounds = ev ? atoi(ev) : 0;
            }
            const int64_t cap = (int64_t)n_sm * inc_ctas_per_sm(p.host_chunk != 0);
            int64_t m = rounds > 0 ? rounds : (n >= (48 << 10) ? 2 : PCGRL_STEP_INC_ROUNDS);
            const int64_t need_m = (n + STAT_THREADS * cap - 1) / (STAT_THREADS * cap);
            if (m < need_m) m = need_m;
            const int64_t want = (n + STAT_THREADS * m - 1) / (STAT_THREADS * m);
            k_step_inc<NW, TWO><<<(unsigned)(want < cap ? want : cap), STAT_THREADS, 0, s>>>(p);
            return cudaGetLastError();
        }
    }
    if constexpr (Machine::HAS_CACHE && TWO) {
        if (p.split_phase == 2) {   // progressive host pipeline: one search over every chunk's list
            int n_sm = 0;
            if ((e = sm_count(n_sm)) != cudaSuccess) return e;
            k_split_stats_inc_multi<NW, TWO><<<(unsigned)prog_search_ctas(n, n_sm), STAT_THREADS, 0, s>>>(p);
            return cudaGetLastError();
        }
    }
    if (p.split_phase == 3) {       // ... and per chunk: wait for its list, then the output kernel
        k_wait_list<<<1, 32, 0, s>>>(p.wl_hdr, p.ml_target, p.status);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        k_split_out<K><<<(unsigned)((n + OUT_THREADS - 1) / OUT_THREADS), OUT_THREADS, 0, s>>>(p);
        return cudaGetLastError();
    }
    k_split_act<<<(unsigned)((n + ACT_THREADS - 1) / ACT_THREADS), ACT_THREADS, 0, s>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (p.split_phase == 1) return cudaSuccess;
    bool ran_inc = false;
    if constexpr (Machine::HAS_CACHE && TWO) {
        if (incremental == 1) {
            int n_sm = 0;
            if ((e = sm_count(n_sm)) != cudaSuccess) return e;
            const int64_t want = (n + STAT_THREADS - 1) / STAT_THREADS;
            const int64_t cap = (int64_t)n_sm * inc_ctas_per_sm(p.host_chunk != 0);
            k_split_stats_inc<NW, TWO><<<(unsigned)(want < cap ? want : cap), STAT_THREADS, 0, s>>>(p);
            ran_inc = true;
        }
    }
    if (!ran_inc) {
        constexpr int T = stat_threads(Machine::Prob::P * NW);
        k_split_stats<Machine, NW, TWO><<<(unsigned)((n + T - 1) / T), T, 0, s>>>(p);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    k_split_out<K><<<(unsigned)((n + OUT_THREADS - 1) / OUT_THREADS), OUT_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

template <template <int, bool> class Machine>
static cudaError_t dispatch_split(const KParams& p, cudaStream_t s, int incremental, bool& supported) {
    const int H = p.d0, W = p.d1;
    supported = p.ndim == 2;
    if (!supported) return cudaSuccess;
    if (W <= 16 && H <= 16) {
        const int nw = (H + 1) / 2;
        if (nw <= 1) return launch_split<Machine<1, true>, 1, true>(p, s, incremental);
        if (nw <= 2) return launch_split<Machine<2, true>, 2, true>(p, s, incremental);
        if (nw <= 4) return launch_split<Machine<4, true>, 4, true>(p, s, incremental);
        return launch_split<Machine<8, true>, 8, true>(p, s, incremental);
    }
    if (W <= 32 && H <= 32) {
        if (H <= 16) return launch_split<Machine<16, false>, 16, false>(p, s, 0);
        return launch_split<Machine<32, false>, 32, false>(p, s, 0);
    }
    supported = false;
    return cudaSuccess;
}

// MODE_STEP of a bit-board problem through the split path.  `incremental` asks for the cached incremental search
// (binary only; needs p.cache).  Three kernels are launched: *launches is incremented by the caller accordingly.
// incremental: 0 = from-scratch searches, 1 = incremental search as the middle kernel of the split path, 2 = the
// fused incremental step (k_step_inc, one launch).  *n_launches = kernels launched.
cudaError_t launch_bitboard_split(const KParams& p, int problem, cudaStream_t s, int incremental, bool& supported,
                                  int& n_launches) {
    supported = false;
    n_launches = p.split_phase == 0 ? 3 : p.split_phase == 3 ? 2 : 1;
    if (p.mode != MODE_STEP || !p.worklist || p.rep == PCGRL_REP_CELLULAR) return cudaSuccess;
    if (p.split_phase != 0 && (!split_prog_supported(p, problem) || incremental != 1)) return cudaSuccess;
    if (problem == PCGRL_PROB_BINARY) {
        const int inc = p.cache != nullptr && p.d0 <= 16 && p.d1 <= 16 ? incremental : 0;
        if (inc == 2) n_launches = 1;
        return dispatch_split<BinaryMachine>(p, s, inc, supported);
    }
    if (problem == PCGRL_PROB_ZELDA) return dispatch_split<ZeldaMachine>(p, s, 0, supported);
    if (problem == PCGRL_PROB_BINARY_HOLEY) {
        supported = p.ndim == 2 && p.d1 + 2 <= 32 && p.d0 + 2 <= 32;
        if (!supported) return cudaSuccess;
        if (p.d0 + 2 <= 18) return launch_split<BinaryHoleyMachine<18, false>, 18, false>(p, s, 0);
        return launch_split<BinaryHoleyMachine<32, false>, 32, false>(p, s, 0);
    }
    return cudaSuccess;
}

// bytes per env of the incremental search cache (0: this config has none)
int bitboard_cache_stride(int problem, int ndim, int d0, int d1, int rep, int action_kind) {
    if (problem != PCGRL_PROB_BINARY || ndim != 2 || d0 > 16 || d1 > 16) return 0;
    if (rep == PCGRL_REP_CELLULAR || action_kind == PCGRL_ACT_PATCH) return 0;   // whole-map / multi-cell edits
    const int nw = (d0 + 1) / 2;
    const int NW = nw <= 1 ? 1 : nw <= 2 ? 2 : nw <= 4 ? 4 : 8;
    return (2 * NW + 4) * 4;
}

}  // namespace pcgrl
