// Observation writer: Cropped + OneHotEncoding + ToImage (+ ControlWrapper target channels).
//
// Reference (relative to /root/reference/control_pcgrl/):
//   wrappers.py:407-437  Cropped._transform    out[o] = map[pos + o - obs//2] + 1, 0 when out of bounds
//   wrappers.py:232-257  OneHotEncoding        np.eye(dim)[.]  (dim = C+1 behind a crop, C otherwise)
//   wrappers.py:140-150  ToImage               channels last
//   control_wrappers.py:189-214 observe_metric_trgs   2*n_ctrl constant planes PREPENDED:
//                                              (trg / range, metric / range), tuple trg -> midpoint
// One thread per output pixel writes all of its channels (channels-last => contiguous per thread, and
// consecutive threads write consecutive pixels => coalesced stores).  The int8 grid rows are re-read
// through L1/L2; nothing else is loaded.
#include "pcgrl_device.cuh"

namespace pcgrl {

struct ObsParams {
    int32_t ndim, d0, d1, d2, row_stride, n_tiles, n_stats;
    int32_t crop, o0, o1, o2;
    int32_t n_ctrl;
    int32_t ctrl_idx[PCGRL_MAX_STATS];
    double ctrl_range[PCGRL_MAX_STATS];
    int32_t targets_per_env;
    int64_t n_envs;
    const int8_t* grids;
    const int32_t* pos;
    const int32_t* stats;
    const double* targets;
    const uint8_t* static_mask;   // NULL: no static_builds plane
    void* out;
};

template <typename T>
__global__ void __launch_bounds__(256) k_observe(const ObsParams p) {
    const int64_t pix_per_env = (int64_t)p.o0 * p.o1 * p.o2;
    const int64_t total = p.n_envs * pix_per_env;
    const int n_map_ch = p.crop ? p.n_tiles + 1 : p.n_tiles;
    const int n_ch = 2 * p.n_ctrl + n_map_ch + (p.static_mask ? 1 : 0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t env = i / pix_per_env;
        int r = (int)(i - env * pix_per_env);
        const int q2 = r % p.o2;
        r /= p.o2;
        const int q1 = r % p.o1;
        const int q0 = r / p.o1;
        int hot, frozen = 0;
        if (p.crop) {
            const int32_t* pos = p.pos + env * 3;
            const int s0 = pos[0] + q0 - p.o0 / 2, s1 = pos[1] + q1 - p.o1 / 2;
            const int s2 = (p.ndim == 3) ? pos[2] + q2 - p.o2 / 2 : 0;
            if ((unsigned)s0 < (unsigned)p.d0 && (unsigned)s1 < (unsigned)p.d1 && (unsigned)s2 < (unsigned)p.d2)
                hot = p.grids[env * p.row_stride + (s0 * p.d1 + s1) * p.d2 + s2] + 1;
            else
                hot = 0;
            if (p.static_mask) {
                // 'static_builds' (wrappers.py:451-459): the (dims+2) BORDERED mask goes through the same pad and
                // crop as the map, so window cell o shows bordered cell pos + o - obs//2 = map cell (that - 1);
                // border cells are always frozen (envs/reps/wrappers.py:310-312), beyond the border is padding (0)
                const int b0 = s0 - 1, b1 = s1 - 1, b2 = (p.ndim == 3) ? s2 - 1 : 0;
                const bool in_bordered = b0 >= -1 && b0 <= p.d0 && b1 >= -1 && b1 <= p.d1 &&
                                         (p.ndim != 3 || (b2 >= -1 && b2 <= p.d2));
                if (in_bordered) {
                    const bool inner = (unsigned)b0 < (unsigned)p.d0 && (unsigned)b1 < (unsigned)p.d1 &&
                                       (unsigned)b2 < (unsigned)p.d2;
                    frozen = inner ? p.static_mask[env * p.row_stride + (b0 * p.d1 + b1) * p.d2 + b2] != 0 : 1;
                }
            }
        } else {
            hot = p.grids[env * p.row_stride + (q0 * p.d1 + q1) * p.d2 + q2];
        }
        T* o = (T*)p.out + i * n_ch;
        for (int c = 0; c < p.n_ctrl; ++c) {
            const int k = p.ctrl_idx[c];
            const double* trg = p.targets + ((p.targets_per_env ? env * p.n_stats : 0) + k) * 2;
            double t = trg[0];
            if (!isnan(trg[1])) t = (trg[0] + trg[1]) / 2;
            o[2 * c] = (T)(t / p.ctrl_range[c]);
            o[2 * c + 1] = (T)((double)p.stats[env * p.n_stats + k] / p.ctrl_range[c]);
        }
        o += 2 * p.n_ctrl;
        for (int c = 0; c < n_map_ch; ++c) o[c] = (T)(c == hot ? 1 : 0);
        if (p.static_mask) o[n_map_ch] = (T)frozen;
    }
}

cudaError_t launch_observe(const pcgrl_config& cfg, const pcgrl_state& st, const pcgrl_obs_args& a, cudaStream_t s) {
    ObsParams p;
    p.ndim = cfg.ndim;
    p.d0 = cfg.dims[0];
    p.d1 = cfg.dims[1];
    p.d2 = cfg.ndim == 3 ? cfg.dims[2] : 1;
    p.row_stride = cfg.row_stride;
    p.n_tiles = cfg.n_tiles;
    p.n_stats = cfg.n_stats;
    p.crop = a.crop;
    p.o0 = a.obs_dims[0];
    p.o1 = a.obs_dims[1];
    p.o2 = cfg.ndim == 3 ? a.obs_dims[2] : 1;
    p.n_ctrl = a.n_ctrl;
    for (int i = 0; i < PCGRL_MAX_STATS; ++i) {
        p.ctrl_idx[i] = a.ctrl_idx[i];
        p.ctrl_range[i] = a.ctrl_range[i];
    }
    p.targets_per_env = cfg.targets_per_env;
    p.n_envs = st.n_envs;
    p.grids = st.grids;
    p.pos = st.pos;
    p.stats = st.stats;
    p.targets = st.targets;
    p.static_mask = nullptr;
    if (a.static_channel) {
        if (!a.crop || !st.static_mask) return cudaErrorInvalidValue;
        p.static_mask = st.static_mask;
    }
    p.out = a.out;
    if (!a.crop && (p.o0 != p.d0 || p.o1 != p.d1 || p.o2 != p.d2)) return cudaErrorInvalidValue;
    if (a.out_kind == 0 && a.n_ctrl > 0) return cudaErrorInvalidValue;  // target planes are fractional
    const int64_t total = st.n_envs * (int64_t)p.o0 * p.o1 * p.o2;
    if (total == 0) return cudaSuccess;
    const int64_t want = (total + 255) / 256;
    const unsigned blocks = (unsigned)(want < 148 * 32 ? want : 148 * 32);
    if (a.out_kind == 0)
        k_observe<uint8_t><<<blocks, 256, 0, s>>>(p);
    else if (a.out_kind == 1)
        k_observe<float><<<blocks, 256, 0, s>>>(p);
    else if (a.out_kind == 2)
        k_observe<double><<<blocks, 256, 0, s>>>(p);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace pcgrl
