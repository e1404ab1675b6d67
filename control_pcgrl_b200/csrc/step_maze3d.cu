// minecraft_3D_maze: fused env step with warp-per-grid get_stats.
//
// Reference path replaced (paths relative to /root/reference/control_pcgrl/envs/):
//   probs/minecraft/minecraft_3D_maze_prob.py:143-181 (get_stats: regions, path-length, n_jump)
//   helper_3D.py:396-406 + 354-383   calc_num_regions / _flood_fill (6-neighbour AIR components)
//   helper_3D.py:503-563             calc_longest_path (outer scan, whole-plane visited marks, last n_jump wins)
//   helper_3D.py:422-490             run_dijkstra (FIFO label-correcting search over player footholds)
//   helper_3D.py:214-319             _passable (walk / step down / step up / 1-gap jumps of a 2-tall player)
//
// Data layout per warp (shared memory): the level as bit masks -- col[y][x] = 16-bit mask over z of AIR
// cells (every _passable test is a few bit tests on 3 column words), row[z][y] = mask over x (flood fill,
// start-tile scan) -- plus per-cell u16 `best` (shortest length pushed so far | recorded flag), the
// first-recording order list and a ring-buffer FIFO.  The per-cell jump counts `nj` (written at every pop,
// read ONCE per grid: only the far tile of the last component matters) live in a per-warp slice of the global
// scratch instead (fire-and-forget stores), and so does the first-recording order list (appended by lane 0,
// re-read a handful of times per grid: plane marks, far tile, clearing): 7.8 KB instead of 16.4 KB of shared
// memory per warp, i.e. 28 instead of 12 resident warps per SM for this latency-bound search.
//
// Exactness notes (the same restatement is checked on the CPU against the reference):
//   * the reference pops (cell, path, n_jump) entries FIFO and skips an entry when a path of length <= its
//     own is already recorded.  An entry that is no shorter than an entry queued (or recorded) earlier for the
//     same cell is therefore always skipped at pop time, and a skipped pop has no side effect -- so it is
//     filtered at push time instead (`best`), which leaves the order of all effective pops unchanged and
//     makes every pop a recording.
//   * far tile = first maximum in dict-insertion order of `paths` = first-recording order (`order`).
//   * visited_map[np.array(list(paths.keys()))] = 1 marks the z-planes whose index equals ANY coordinate of
//     any recorded cell (helper_3D.py:531); a coordinate >= depth raises IndexError there, reported here
//     as status bit 1.
#include "maze3d_search.cuh"

namespace pcgrl {

int64_t maze3d_scratch_bytes() { return (int64_t)MAZE_SLICE * MAZE_MAX_CTAS * MAZE_WARPS; }

cudaError_t launch_maze3d(const KParams& p, cudaStream_t s, bool& supported) {
    supported = p.ndim == 3 && p.d0 <= 16 && p.d1 <= 16 && p.d2 <= 16 && p.d0 >= 1 && p.scratch != nullptr;
    if (!supported) return cudaSuccess;
    const MazeLayout L = maze_layout(p.d0, p.d1, p.d2, p.row_stride);
    return launch_search<Maze3DProb, MAZE_WARPS>(p, s, L.total, MAZE_CTAS_PER_SM, MAZE_MAX_CTAS);
}

}  // namespace pcgrl
