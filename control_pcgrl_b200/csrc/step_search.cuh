// Skeleton of the warp-per-grid step kernels (problems whose get_stats is an order-dependent sequential
// search: minecraft_3D_maze, sokoban, smb).  Same phases as the bit-board kernel (step_bitboard.cu), but
//   - the grid of CTAs is persistent (one or more CTAs per SM looping over tiles of TILE envs), so that the
//     per-warp search workspaces (shared memory, plus global scratch for the big node pools) are bounded by
//     the number of resident warps, not by the number of envs;
//   - phase C hands each changed env to a whole warp: lane-parallel where the work is data-parallel (grid
//     load, bit-board construction, flood fill, child generation), warp-uniform where the reference's
//     queue order has to be reproduced step by step.
#pragma once
#include "pcgrl_device.cuh"
#include "step_common.cuh"

namespace pcgrl {

constexpr int SEARCH_WARPS = 8;
constexpr int SEARCH_THREADS = SEARCH_WARPS * 32;
constexpr int SEARCH_TILE = 64;   // envs per CTA iteration
constexpr int SEARCH_MAX_CTAS = 320;   // persistent-grid cap that sizes the global scratch (>= 2 CTAs x 148 SMs)

// Number of connected components of the set bits of a Z x Y x 16 bit-board (rows[z * Y + y] = mask over x),
// 6-neighbour in 3D, 4-neighbour when Z == 1 (helper.py:200-210 / helper_3D.py:396-406 calc_num_regions).
// Whole warp; avail / f0 / f1 are Z*Y-entry scratch boards.  Single-cell components are counted with one
// popcount pass; the others are flood-filled one at a time, a level per pass over the rows.
__device__ inline int count_regions_rows(const uint16_t* row, int Z, int Y, uint16_t* avail, uint16_t* f0,
                                         uint16_t* f1, int lane) {
    const int R = Z * Y;
    int regions = 0;
    {
        int iso_cnt = 0;
        for (int r = lane; r < R; r += 32) {
            const int z = r / Y, y = r - z * Y;
            const uint32_t a = row[r];
            const uint32_t nb = (a << 1) | (a >> 1) | (y > 0 ? row[r - 1] : 0u) | (y < Y - 1 ? row[r + 1] : 0u) |
                                (z > 0 ? row[r - Y] : 0u) | (z < Z - 1 ? row[r + Y] : 0u);
            const uint32_t iso = a & ~nb;
            iso_cnt += __popc(iso);
            avail[r] = (uint16_t)(a & ~iso);
            f0[r] = 0;
        }
        regions = __reduce_add_sync(0xffffffffu, iso_cnt);
        __syncwarp();
    }
    for (;;) {
        int first = 0xFFFF;
        for (int r = lane; r < R; r += 32)
            if (avail[r]) {
                first = r;
                break;
            }
        first = __reduce_min_sync(0xffffffffu, first);
        if (first == 0xFFFF) break;
        ++regions;
        if (lane == 0) {
            const uint32_t a = avail[first], bit = a & (0u - a);
            f0[first] = (uint16_t)bit;
            avail[first] = (uint16_t)(a ^ bit);
        }
        __syncwarp();
        uint16_t *cur = f0, *nxt = f1;
        for (;;) {
            uint32_t any = 0;
            for (int r = lane; r < R; r += 32) {
                const int z = r / Y, y = r - z * Y;
                const uint32_t f = cur[r];
                const uint32_t nb = (f << 1) | (f >> 1) | (y > 0 ? cur[r - 1] : 0u) | (y < Y - 1 ? cur[r + 1] : 0u) |
                                    (z > 0 ? cur[r - Y] : 0u) | (z < Z - 1 ? cur[r + Y] : 0u);
                const uint32_t a = avail[r], nf = nb & a;
                nxt[r] = (uint16_t)nf;
                avail[r] = (uint16_t)(a ^ nf);
                any |= nf;
            }
            __syncwarp();
            uint16_t* t = cur;
            cur = nxt;
            nxt = t;
            if (!__any_sync(0xffffffffu, any != 0)) break;
        }
        for (int r = lane; r < R; r += 32) f0[r] = 0;   // the seed board must be empty for the next component
        __syncwarp();
    }
    return regions;
}

// Prob must provide:
//   static constexpr int K;
//   struct Ctx;                                         per-warp context (pointers into its workspaces)
//   static __device__ Ctx make_ctx(const KParams&, uint8_t* warp_smem, int global_warp);
//   static __device__ void stats(const KParams&, Ctx&, const int8_t* grid, int lane, int32_t* out /*smem [K]*/);
template <class Prob>
__global__ void __launch_bounds__(SEARCH_THREADS) k_step_search(const KParams p, const int smem_per_warp) {
    constexpr int K = Prob::K;
    constexpr int TILE = SEARCH_TILE;
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    __shared__ int32_t s_stats[TILE * K];
    __shared__ int16_t s_list[TILE];
    __shared__ int16_t s_slot[TILE];
    __shared__ uint8_t s_flag[TILE];
    __shared__ int s_count, s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int8_t* grids_in = (p.mode == MODE_STATS) ? p.stats_grids : p.grids;
    typename Prob::Ctx ctx = Prob::make_ctx(p, dyn_smem + (size_t)warp * smem_per_warp, blockIdx.x * SEARCH_WARPS + warp);
    const int64_t n_tiles = (p.n_envs + TILE - 1) / TILE;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        const int tile_n = (int)min((int64_t)TILE, p.n_envs - base);
        __syncthreads();   // previous tile's phase D is done with the shared lists
        if (tid == 0) {
            s_count = 0;
            s_next = 0;
        }
        for (int e = tid; e < TILE; e += SEARCH_THREADS) {
            s_slot[e] = -1;
            s_flag[e] = 0;
        }
        __syncthreads();
        phase_a<SEARCH_THREADS, TILE>(p, base, tile_n, s_list, s_slot, s_flag, &s_count);
        const int M = s_count;

        // phase C: one warp per changed env, pulled from the shared queue
        for (;;) {
            int item = 0;
            if (lane == 0) item = atomicAdd(&s_next, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= M) break;
            const int8_t* grid = grids_in + (base + s_list[item]) * p.row_stride;
            Prob::stats(p, ctx, grid, lane, s_stats + item * K);
            __syncwarp();
        }
        __syncthreads();
        phase_d<SEARCH_THREADS, K>(p, base, tile_n, s_slot, s_stats);
    }
}

// host side: persistent launch sized from the occupancy the dynamic shared memory allows
template <class Prob>
static cudaError_t launch_search(const KParams& p, cudaStream_t s, int smem_per_warp, int max_ctas_per_sm = 1 << 20) {
    static int n_sm = 0;
    cudaError_t e;
    if (!n_sm) {
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    }
    const int dyn = smem_per_warp * SEARCH_WARPS;
    if ((e = cudaFuncSetAttribute(k_step_search<Prob>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)) != cudaSuccess)
        return e;
    int per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_step_search<Prob>, SEARCH_THREADS, dyn)) != cudaSuccess)
        return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const int64_t tiles = (p.n_envs + SEARCH_TILE - 1) / SEARCH_TILE;
    if (tiles == 0) return cudaSuccess;
    if (per_sm > max_ctas_per_sm) per_sm = max_ctas_per_sm;
    int64_t cap = (int64_t)n_sm * per_sm;
    if (max_ctas_per_sm < (1 << 20) && cap > SEARCH_MAX_CTAS) cap = SEARCH_MAX_CTAS;   // scratch is sized for this
    const int ctas = (int)(tiles < cap ? tiles : cap);
    k_step_search<Prob><<<ctas, SEARCH_THREADS, dyn, s>>>(p, smem_per_warp);
    return cudaGetLastError();
}

}  // namespace pcgrl
