// Shared device-side helpers for the PCGRL step kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pcgrl_b200.h"

namespace pcgrl {

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_STATS = 2 };

// Everything a kernel needs, passed by value (fits the 4 KB parameter space comfortably).
struct KParams {
    int32_t mode;
    int32_t rep, action_kind, ndim;
    int32_t d0, d1, d2;           // numpy axis order; 2D: (H, W, 1)
    int32_t cells, row_stride;
    int32_t n_tiles, n_stats;
    int32_t max_iterations, max_changes;
    int32_t act_h, act_w;
    int32_t targets_per_env;
    int32_t init_random_probs;
    float   init_cdf[PCGRL_MAX_TILES];
    double  weights[PCGRL_MAX_STATS];
    int64_t n_envs, env_offset;
    int8_t* grids;
    int32_t *pos, *n_step, *iteration, *changes, *stats;
    const double* targets;
    float* reward;
    uint8_t *done, *changed;
    int32_t* status;
    void* scratch;
    const void* actions;
    // reset
    const uint8_t* mask;
    const int8_t* src_grids;
    const int32_t* src_pos;
    uint64_t seed, epoch;
    // stats-only mode
    const int8_t* stats_grids;
    int32_t* stats_out;
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: no RNG state lives in HBM.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ------------------------------------------------------------------------------------------------
// ControlWrapper.get_loss (control_wrappers.py:318-345) for one env, in fp64 from integer stats.
//   scalar target:  -|trg - val| * w
//   range (lo,hi):  -min_j |lo + j - val| * w   over j = 0 .. ceil(hi-lo)-1   (np.arange(lo,hi), hi excluded)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double control_loss(const int32_t* st, const double* trg, const double* w, int K) {
    double loss = 0.0;
    for (int k = 0; k < K; ++k) {
        const double wk = w[k];
        if (wk == 0.0) continue;
        const double lo = trg[2 * k], hi = trg[2 * k + 1], v = (double)st[k];
        double d;
        if (isnan(hi)) {
            d = fabs(lo - v);
        } else {
            const double n = ceil(hi - lo);
            double j = rint(v - lo);
            j = fmin(fmax(j, 0.0), n - 1.0);
            d = fabs(lo + j - v);
        }
        loss += -d * wk;
    }
    return loss;
}

// ------------------------------------------------------------------------------------------------
// Sub-warp groups: G consecutive lanes own one level grid as a bit-board, one 32-bit word per lane.
//   TWO = true : a word holds two rows of up to 16 cells (row 2l in bits 0..15, row 2l+1 in 16..31)
//   TWO = false: a word holds one row of up to 32 cells
// Bit order (lane, bit) is monotonic in the row-major cell index, so "lowest set bit of the group"
// == "first cell in the reference's y-outer/x-inner scan" (envs/helper.py:23-25).
// ------------------------------------------------------------------------------------------------
template <int G, bool TWO>
struct Group {
    static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group size");
    int lane;      // lane in warp
    int lig;       // lane in group
    int gbase;     // first lane of the group
    uint32_t gmask;  // warp mask of the group's lanes
    uint32_t mask_l, mask_r;

    __device__ __forceinline__ Group(int width) {
        lane = threadIdx.x & 31;
        lig = lane & (G - 1);
        gbase = lane & ~(G - 1);
        gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << gbase);
        // horizontal shifts must not leak between the two 16-bit rows of a word (only possible when the
        // rows are exactly 16 wide; narrower rows have never-passable padding bits in between)
        mask_l = (TWO && width == 16) ? 0xFFFEFFFEu : 0xFFFFFFFFu;
        mask_r = (TWO && width == 16) ? 0x7FFF7FFFu : 0xFFFFFFFFu;
    }

    // 4-neighbourhood dilation of the board (one bit per cell); caller ANDs with the passable set.
    // Must be executed by all lanes of the group (uses group-masked shuffles).
    __device__ __forceinline__ uint32_t expand(uint32_t f) const {
        uint32_t up = __shfl_up_sync(gmask, f, 1, G);      // word of the rows above
        uint32_t dn = __shfl_down_sync(gmask, f, 1, G);    // word of the rows below
        if (G == 1 || lig == 0) up = 0;
        if (G == 1 || lig == G - 1) dn = 0;
        uint32_t u, d;
        if (TWO) {
            u = __funnelshift_l(up, f, 16);   // cell (y,x) <- (y-1,x): low row from prev word's high row
            d = __funnelshift_r(f, dn, 16);   // cell (y,x) <- (y+1,x)
        } else {
            u = up;
            d = dn;
        }
        return ((f << 1) & mask_l) | ((f >> 1) & mask_r) | u | d;
    }

    __device__ __forceinline__ bool any(uint32_t x) const { return (__ballot_sync(gmask, x != 0) & gmask) != 0; }

    // One-hot board holding only the lowest set cell of x (all-zero if x is empty).
    __device__ __forceinline__ uint32_t lowest(uint32_t x, bool& found) const {
        const uint32_t b = __ballot_sync(gmask, x != 0) & gmask;
        found = b != 0;
        const int first = found ? (__ffs(b) - 1) : gbase;
        const uint32_t w = __shfl_sync(gmask, x, first);
        return (found && lane == first) ? (w & (0u - w)) : 0u;
    }

    // sum over the group's lanes
    __device__ __forceinline__ int sum(int v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        return v;
    }
    __device__ __forceinline__ int bcast(int v, int src_lig) const { return __shfl_sync(gmask, v, gbase + src_lig); }
};

}  // namespace pcgrl
