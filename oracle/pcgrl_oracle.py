"""TEST INFRASTRUCTURE ONLY -- CPU restatement of control-pcgrl's env-step hot path.

This is the oracle for the CUDA path in control_pcgrl_b200/.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs may import it; the product never does.

Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py) against
fixtures under tests/golden/ that oracle/gen_golden.py produced by running the *real* reference
(smearle/control-pcgrl @ 8bde536, through oracle/refshim.py) in the build container, and -- when
/root/reference is present -- against the live reference (tests/test_oracle_vs_reference.py).

All `file:line` citations are relative to /root/reference/control_pcgrl/.  Grids are integer numpy
arrays indexed [y, x] (2D) or [z, y, x] (3D) holding tile codes (index into the problem's tile
list, envs/probs/problem.py:59-64); the reference's string maps (envs/helper.py:506-515) are an
implementation detail that does not affect any result.
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np

# --------------------------------------------------------------------------------------------
# problem tables (tile codes, stat order) -- SURVEY.md section E
# --------------------------------------------------------------------------------------------
TILES = {
    "binary": ["empty", "solid"],                                                   # binary_prob.py:17
    "zelda": ["empty", "solid", "player", "key", "door", "bat", "scorpion", "spider"],  # zelda_prob.py:20
    "sokoban": ["empty", "solid", "player", "crate", "target"],                     # sokoban_prob.py:26
    "smb": ["empty", "solid", "enemy", "brick", "question", "coin", "tube"],        # smb_prob.py:12
    "minecraft_3D_maze": ["AIR", "DIRT"],                                           # minecraft_3D_maze_prob.py:26
    "binary_holey": ["empty", "solid"],                                             # binary_holey_prob.py:12-14
    "minecraft_2D_maze": ["AIR", "DIRT"],                                           # minecraft_2D_maze_prob.py:40-41
    "minecraft_3D_holey_maze": ["AIR", "DIRT"],
    "minecraft_3D_dungeon_holey": ["AIR", "DIRT", "CHEST", "SKULL", "PUMPKIN"],     # minecraft_3D_holey_dungeon_prob.py:18
}
STAT_NAMES = {
    "binary": ["regions", "path-length"],
    "zelda": ["player", "key", "door", "enemies", "regions", "nearest-enemy", "path-length"],
    "sokoban": ["player", "crate", "target", "regions", "dist-win", "sol-length", "ratio"],
    "smb": ["dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist",
            "dist-win", "sol-length"],
    "minecraft_3D_maze": ["regions", "path-length", "n_jump"],
    "binary_holey": ["regions", "path-length", "connected-path-length"],
    "minecraft_2D_maze": ["regions", "path-length"],
    # minecraft_3D_holey_maze_prob.py:124-130 (+ the carried length of the path found by THIS call, which the next
    # call reports as path-length, :92-93)
    "minecraft_3D_holey_maze": ["regions", "path-length", "connected-path-length", "n_jump", "_next-path-length"],
    # minecraft_3D_holey_dungeon_prob.py:99-106
    "minecraft_3D_dungeon_holey": ["regions", "path-length", "chests", "enemies", "nearest-enemy", "n_jump"],
}
# tile init probabilities used by reset when no grid is supplied
INIT_PROBS = {
    "binary": [0.5, 0.5],                                               # binary_prob.py:25
    "zelda": [0.58, 0.3, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02],           # zelda_prob.py:26
    "sokoban": [0.45, 0.4, 0.05, 0.05, 0.05],                           # sokoban_prob.py:32-38
    "smb": [0.75, 0.1, 0.01, 0.04, 0.01, 0.02, 0.02],                   # smb_prob.py:18
    "minecraft_3D_maze": [1.0, 0.0],                                    # minecraft_3D_maze_prob.py:37
    "minecraft_2D_maze": [0.5, 0.5],                                    # minecraft_2D_maze_prob.py:19
}

_ZELDA_WALK = [0, 2, 3, 5, 7, 6]        # empty, player, key, bat, spider, scorpion (zelda_ctrl_prob.py:101-104)
_ZELDA_WALK_DOOR = [0, 2, 3, 4, 5, 7, 6]


# --------------------------------------------------------------------------------------------
# 2D stat helpers (envs/helper.py)
# --------------------------------------------------------------------------------------------
def tile_positions(grid):
    """helper.py:19-26 get_tile_locations: code -> [(x, y), ...] in row-major scan order."""
    out = {}
    h, w = grid.shape
    rows = grid.tolist()
    for y in range(h):
        row = rows[y]
        for x in range(w):
            out.setdefault(row[x], []).append((x, y))
    return out


def _ordered_tiles(positions, codes):
    """helper.py:153-157 _get_certain_tiles: concatenation per tile type, in list order."""
    seq = []
    for c in codes:
        seq.extend(positions.get(c, ()))
    return seq


def bfs_distances(grid, passable, sx, sy):
    """helper.py:225-240 run_dijkstra: unit-weight 4-neighbour distances, -1 = not reached.

    The reference pops a FIFO list and relabels only when the stored label is larger, which for
    unit weights is plain breadth-first search; a source that is not passable leaves all -1.
    """
    h, w = grid.shape
    dist = np.full((h, w), -1, dtype=np.int64)
    ok = np.isin(grid, list(passable))
    if not ok[sy, sx]:
        return dist
    okl = ok.tolist()
    d = [[-1] * w for _ in range(h)]
    d[sy][sx] = 0
    q = deque([(sx, sy)])
    while q:
        cx, cy = q.popleft()
        nd = d[cy][cx] + 1
        for nx, ny in ((cx - 1, cy), (cx + 1, cy), (cx, cy - 1), (cx, cy + 1)):
            if 0 <= nx < w and 0 <= ny < h and okl[ny][nx] and d[ny][nx] < 0:
                d[ny][nx] = nd
                q.append((nx, ny))
    return np.array(d, dtype=np.int64)


def count_regions(grid, passable, positions=None):
    """helper.py:200-210 calc_num_regions (+ _flood_fill :173-187): # 4-connected components."""
    h, w = grid.shape
    ok = np.isin(grid, list(passable)).tolist()
    seen = [[False] * w for _ in range(h)]
    n = 0
    for y in range(h):
        for x in range(w):
            if ok[y][x] and not seen[y][x]:
                n += 1
                seen[y][x] = True
                st = [(x, y)]
                while st:
                    cx, cy = st.pop()
                    for nx, ny in ((cx - 1, cy), (cx + 1, cy), (cx, cy - 1), (cx, cy + 1)):
                        if 0 <= nx < w and 0 <= ny < h and ok[ny][nx] and not seen[ny][nx]:
                            seen[ny][nx] = True
                            st.append((nx, ny))
    return n


def longest_shortest_path(grid, passable, positions=None):
    """helper.py:255-276 calc_longest_path: the double-sweep heuristic, per component.

    Component start tiles are taken in (passable-type list order, then row-major) order
    (:257, :153-157); far tile = row-major-first maximum of the first sweep (np.argmax, :265);
    result = max over components of the second sweep's maximum (strict '>', :268).
    """
    if positions is None:
        positions = tile_positions(grid)
    h, w = grid.shape
    visited = np.zeros((h, w), dtype=bool)
    best = 0
    for (x, y) in _ordered_tiles(positions, passable):
        if visited[y, x]:
            continue
        d1 = bfs_distances(grid, passable, x, y)
        visited |= d1 >= 0
        my, mx = np.unravel_index(int(np.argmax(d1)), d1.shape)
        d2 = bfs_distances(grid, passable, int(mx), int(my))
        best = max(best, int(d2.max()))
    return best


def range_reward(new, old, low, high):
    """helper.py:550-560 get_range_reward (legacy reward; not called by step at this commit)."""
    if low <= new <= high and low <= old <= high:
        return 0
    if old <= high and new <= high:
        return min(new, low) - min(old, low)
    if old >= low and new >= low:
        return max(old, high) - max(new, high)
    if new > high and old < low:
        return high - new + old - low
    if new < low and old > high:
        return high - old + new - low
    return None


_INF = float("inf")
# Legacy Problem.get_reward tables: stat -> (low, high) band of get_range_reward and the weight the
# (non-ctrl) Problem class hard-codes.  binary_prob.py:170-178 (:37-40); zelda_prob.py:135-153 (:33-41,
# _max_enemies 5, _target_enemy_dist 4); sokoban_prob.py:185-229 (:44-52, _max_crates 3);
# smb_prob.py:156-176 (:25-35).  minecraft_3D_maze's get_reward is commented out upstream.
RANGE_BANDS = {
    "binary": {"regions": (1, 1), "path-length": (125, 125)},
    "zelda": {"player": (1, 1), "key": (1, 1), "door": (1, 10), "enemies": (2, 5), "regions": (1, 1),
              "nearest-enemy": (4, _INF), "path-length": (_INF, _INF)},
    "sokoban": {"player": (1, 1), "crate": (1, 3), "target": (1, 3), "regions": (1, 1), "ratio": (-_INF, -_INF),
                "dist-win": (-_INF, -_INF), "sol-length": (_INF, _INF)},
    "smb": {"dist-floor": (0, 0), "disjoint-tubes": (0, 0), "enemies": (10, 30), "empty": (900, _INF),
            "noise": (0, 0), "jumps": (20, _INF), "jumps-dist": (0, 0), "dist-win": (0, 0)},
    "minecraft_2D_maze": {"regions": (1, 1), "path-length": (_INF, _INF)},   # minecraft_2D_maze_prob.py:106-115
}
RANGE_WEIGHTS = {
    "binary": {"regions": 100, "path-length": 100},
    "zelda": {"player": 3, "key": 3, "door": 3, "regions": 5, "enemies": 1, "nearest-enemy": 2, "path-length": 1},
    "sokoban": {"player": 3, "crate": 2, "target": 2, "regions": 5, "ratio": 2, "dist-win": 0.0, "sol-length": 1},
    "smb": {"dist-floor": 2, "disjoint-tubes": 1, "enemies": 1, "empty": 1, "noise": 4, "jumps": 2, "jumps-dist": 2,
            "dist-win": 5},
    "minecraft_2D_maze": {"regions": 5, "path-length": 1},                   # minecraft_2D_maze_prob.py:24-27
}


def legacy_reward(problem, new, old):
    """Problem.get_reward of the non-ctrl problem classes: sum of weighted get_range_reward terms."""
    return sum(range_reward(new[k], old[k], *band) * RANGE_WEIGHTS[problem][k]
               for k, band in RANGE_BANDS[problem].items())


# --------------------------------------------------------------------------------------------
# per-problem get_stats
# --------------------------------------------------------------------------------------------
def binary_stats(grid):
    """probs/binary/binary_prob.py:152-158."""
    pos = tile_positions(grid)
    return {"regions": count_regions(grid, [0]),
            "path-length": longest_shortest_path(grid, [0], pos)}


def zelda_stats(grid):
    """probs/zelda/zelda_ctrl_prob.py:90-168."""
    pos = tile_positions(grid)
    n = lambda codes: sum(len(pos.get(c, ())) for c in codes)
    st = {"player": n([2]), "key": n([3]), "door": n([4]), "enemies": n([5, 7, 6]),
          "regions": count_regions(grid, _ZELDA_WALK), "nearest-enemy": 0, "path-length": 0}
    if st["player"] == 1:                                             # :107
        px, py = pos[2][0]
        enemies = _ordered_tiles(pos, [7, 5, 6])                      # spider, bat, scorpion (:110-113)
        d_player = None
        if enemies:                                                   # :117-132
            d_player = bfs_distances(grid, _ZELDA_WALK, px, py)
            reach = [int(d_player[ey, ex]) for ex, ey in enemies if d_player[ey, ex] > 0]
            st["nearest-enemy"] = min(reach) if reach else 0
        if st["key"] == 1 and st["door"] == 1:                        # :134-150
            kx, ky = pos[3][0]
            dx, dy = pos[4][0]
            if d_player is None:
                d_player = bfs_distances(grid, _ZELDA_WALK, px, py)
            d_key = bfs_distances(grid, _ZELDA_WALK_DOOR, kx, ky)
            st["path-length"] = int(d_player[ky, kx]) + int(d_key[dy, dx])   # -1 each when unreachable
    return st


def bordered_with_holes(grid, holes, border_tile=1, empty_tile=0):
    """The map Problem.get_stats sees in the holey envs (pcgrl_holey_env.py:52-53 `_get_rep_map` ->
    `_bordered_map`): the level inside a one-tile border (reps/representation.py:162-164) into which the
    entrance and the exit were dug as empty tiles (reps/wrappers.py:139-142).  holes = (entrance_y, entrance_x,
    exit_y, exit_x) in bordered coordinates (holey_prob.py:41-42: coords[0] is y)."""
    h, w = grid.shape
    b = np.full((h + 2, w + 2), border_tile, dtype=np.int64)
    b[1:-1, 1:-1] = grid
    ey, ex, xy, xx = (int(v) for v in holes)
    b[ey, ex] = empty_tile
    b[xy, xx] = empty_tile
    return b


def binary_holey_stats(grid, holes):
    """probs/binary/binary_holey_prob.py:59-93 on the bordered map: one BFS from the entrance;
    path-length = its largest distance, connected-path-length = distance of the exit (0 when unreachable);
    regions over the bordered map (the holes are empty tiles)."""
    b = bordered_with_holes(grid, holes)
    ey, ex, xy, xx = (int(v) for v in holes)
    d = bfs_distances(b, [0], ex, ey)                                 # :62 run_dijkstra(x, y, ...)
    connected = int(d[xy, xx])                                        # :63
    return {"regions": count_regions(b, [0]), "path-length": int(d.max()),          # :65-66, :89
            "connected-path-length": 0 if connected == -1 else connected}            # :69-77


def holey_border_idxs(h, w):
    """holey_prob.py:20-30 get_border_idxs: non-corner border cells of the (h+2, w+2) map, row-major."""
    m = np.zeros((h + 2, w + 2), dtype=np.uint8)
    m[1:-1, 0] = m[1:-1, -1] = 1
    m[0, 1:-1] = m[-1, 1:-1] = 1
    return np.argwhere(m == 1)


def valid_holes(entrance, exit_, h, w):
    """holey_prob.py:74-90 _valid_holes, restated literally (its x/y naming and the `_width - 1` tests are
    the reference's own): each hole is pulled one cell inwards on at most one axis, then the two must differ
    by more than one cell on some axis."""
    pts = []
    for (x, y) in (entrance, exit_):
        x, y = int(x), int(y)
        if x == 0:
            x = 1
        elif x == w - 1:
            x = w - 2
        elif y == 0:
            y = 1
        elif y == h - 1:
            y = h - 2
        pts.append((x, y))
    return max(abs(pts[0][0] - pts[1][0]), abs(pts[0][1] - pts[1][1])) > 1


def bordered_with_holes_3d(grid, holes, border_tile=1, empty_tile=0):
    """The bordered 3D map of the holey minecraft problems with entrance and exit dug (foot + head tile each).
    holes = (ez, ey, ex, xz, xy, xx), the foot tiles; the head is the tile above -- except the reference's default
    exit (1, 1, 1) when no candidate was valid, whose head is the same tile (holey_prob_3D.py:86)."""
    g = np.asarray(grid)
    b = np.full(tuple(d + 2 for d in g.shape), border_tile, dtype=np.int64)
    b[1:-1, 1:-1, 1:-1] = g
    ez, ey, ex, xz, xy, xx = (int(v) for v in holes)
    ent = ((ez, ey, ex), (ez + 1, ey, ex))
    ext = ((xz, xy, xx), (xz, xy, xx) if (xz, xy, xx) == (1, 1, 1) else (xz + 1, xy, xx))
    for c in ent + ext:
        b[c] = empty_tile
    return b, ent, ext


def get_stats(problem, grid, holes=None, prev_path_length=0):
    if problem == "binary_holey":
        return binary_holey_stats(grid, holes)
    if problem == "minecraft_3D_holey_maze":
        from . import maze3d_oracle
        b, ent, ext = bordered_with_holes_3d(grid, holes)
        st, new_len = maze3d_oracle.maze3d_holey_stats(b, ent, ext, prev_path_length)
        st["_next-path-length"] = new_len
        return st
    if problem == "minecraft_3D_dungeon_holey":
        from . import maze3d_oracle
        b, ent, ext = bordered_with_holes_3d(grid, holes)
        return maze3d_oracle.maze3d_holey_dungeon_stats(b, ent, ext)
    if problem in ("binary", "minecraft_2D_maze"):
        # minecraft_2D_maze_prob.py:87-93: the same two helpers over ["AIR"], tile code 0 like binary's "empty"
        return binary_stats(grid)
    if problem == "zelda":
        return zelda_stats(grid)
    if problem == "sokoban":
        from . import search_oracle
        return search_oracle.sokoban_stats(grid)
    if problem == "smb":
        from . import search_oracle
        return search_oracle.smb_stats(grid)
    if problem == "minecraft_3D_maze":
        from . import maze3d_oracle
        return maze3d_oracle.maze3d_stats(grid)
    raise KeyError(problem)


def stats_vector(problem, stats):
    return [int(stats[k]) for k in STAT_NAMES[problem]]


# --------------------------------------------------------------------------------------------
# problem constants: static targets / bounds
# --------------------------------------------------------------------------------------------
def problem_constants(problem, map_shape):
    """static_trgs / cond_bounds / default weights exactly as the Problem constructors derive them.

    binary_prob.py:50-84; zelda_ctrl_prob.py:19-64 (+ zelda_prob.py:29); sokoban_ctrl_prob.py:8-56;
    smb_ctrl_prob.py:8-35; minecraft_3D_maze_prob.py:33-81.  Problem.__init__ (problem.py:30-31)
    sets _height,_width = map_shape[0], map_shape[1] *before* the subclass body for binary/zelda;
    sokoban/smb/minecraft overwrite them with hard-coded sizes inside __init__ (SURVEY A-15).
    """
    if problem == "binary":
        h, w = map_shape
        mp = np.ceil(w / 2) * h + np.floor(h / 2)
        return dict(static_trgs={"regions": 1, "path-length": mp},
                    cond_bounds={"regions": (0, w * np.ceil(h / 2)), "path-length": (0, mp)},
                    default_weights={"regions": 100, "path-length": 100})
    if problem in ("minecraft_3D_holey_maze", "minecraft_3D_dungeon_holey"):
        # minecraft_3D_holey_maze_prob.py:33-61 / minecraft_3D_holey_dungeon_prob.py:43-82 on the hard-coded 15^3
        w = h = l = 15
        mp = 2 * (h // 3) * (np.ceil(w / 2) * l + np.floor(l / 2))
        if problem == "minecraft_3D_holey_maze":
            return dict(static_trgs={"regions": 1, "path-length": 10 * mp, "n_jump": 5, "connected-path-length": 10 * mp},
                        cond_bounds={"regions": (0, np.ceil(w * l / 2 * h)), "path-length": (0, mp + 2),
                                     "connected-path-length": (0, mp + 2), "n_jump": (0, mp // 2)},
                        default_weights={"regions": 0, "path-length": 100, "connected-path-length": 120, "n_jump": 150})
        ma = w * h * l // 4
        return dict(static_trgs={"enemies": (2, 5), "regions": 1, "path-length": 10 * mp, "nearest-enemy": (5, mp // 2),
                                 "chests": 1, "n_jump": (2, 5)},
                    cond_bounds={"regions": (0, np.ceil(w * l / 2 * h)), "path-length": (0, mp), "chests": (0, ma),
                                 "n_jump": (0, mp // 2), "nearest-enemy": (0, mp // 2), "enemies": (0, ma)},
                    default_weights={"regions": 0, "path-length": 100, "chests": 300, "n_jump": 100, "enemies": 100,
                                     "nearest-enemy": 200})
    if problem == "minecraft_2D_maze":
        # minecraft_2D_maze_prob.py:15-33: not a controllable problem upstream (no static_trgs / cond_bounds)
        return dict(static_trgs={}, cond_bounds={}, default_weights={"regions": 5, "path-length": 1})
    if problem == "binary_holey":
        # binary_holey_prob.py:19-42 on top of BinaryProblem's constructor
        h, w = map_shape
        mp = np.ceil(w / 2) * h + np.floor(h / 2)
        return dict(static_trgs={"regions": 1, "path-length": mp + 2, "connected-path-length": mp + 2},
                    cond_bounds={"regions": (0, w * np.ceil(h / 2)), "path-length": (0, mp + 2),
                                 "connected-path-length": (0, mp + 2)},
                    default_weights={"regions": 100, "path-length": 0, "connected-path-length": 100})
    if problem == "zelda":
        h, w = map_shape
        mne = np.ceil(w / 2 + 1) * h
        mp = (np.ceil(w / 2) * h + np.floor(h / 2)) * 2 - 1
        return dict(
            static_trgs={"enemies": (2, 5), "path-length": mp, "nearest-enemy": (5, mne), "regions": 1,
                         "player": 1, "key": 1, "door": 1},
            cond_bounds={"nearest-enemy": (0, mne), "enemies": (0, w * h - 2), "player": (0, w * h - 2),
                         "key": (0, w * h - 2), "door": (0, w * h - 2), "regions": (0, w * h / 2),
                         "path-length": (0, mp)},
            default_weights={"player": 3, "key": 3, "door": 3, "regions": 5, "enemies": 1,
                             "nearest-enemy": 1, "path-length": 1})
    if problem == "sokoban":
        # sokoban_prob.py:30-31 hard-codes 5x5 before sokoban_ctrl_prob.py:11-56 derives these (A-15)
        w = h = 5
        mp = np.ceil(w / 2 + 1) * h
        return dict(
            static_trgs={"player": 1, "crate": (2, 3), "regions": 1, "ratio": 0, "dist-win": 0, "sol-length": mp},
            cond_bounds={"player": (1, w * h), "crate": (1, w * h / 2 - max(w, h)), "target": (1, w * h),
                         "ratio": (0, w * h), "dist-win": (0, w * h * (w + h)), "sol-length": (0, 2 * mp),
                         "regions": (0, w * h / 2)},
            default_weights={"player": 3, "crate": 1, "regions": 5, "ratio": 2, "dist-win": 0.0, "sol-length": 1})
    if problem == "smb":
        # smb_prob.py:16-17 hard-codes 116x16 before smb_ctrl_prob.py:8-35 derives these (A-15)
        w, h = 116, 16
        msl = np.ceil(w) * 3
        return dict(
            static_trgs={"dist-floor": 0, "disjoint-tubes": 0, "enemies": (10, 30), "empty": (900, w * h),
                         "noise": 0, "jumps": (20, w * h), "jumps-dist": 0, "dist-win": 0, "sol-length": msl},
            cond_bounds={"dist-floor": (0, w * h), "disjoint-tubes": (0, w * h), "enemies": (0, w * h),
                         "empty": (0, w), "noise": (0, w * h), "jumps": (0, w), "jumps-dist": (0, w * h),
                         "dist-win": (0, w), "sol-length": (0, msl)},
            default_weights={"dist-floor": 2, "disjoint-tubes": 1, "enemies": 1, "empty": 1, "noise": 4, "jumps": 2,
                             "jumps-dist": 2, "dist-win": 5, "sol-length": 1})
    if problem == "minecraft_3D_maze":
        # minecraft_3D_maze_prob.py:33-35 hard-codes 15^3 before :43-69 derives these (A-15)
        w = h = l = 15
        mp = 2 * (h // 3) * (np.ceil(w / 2) * l + np.floor(l / 2))
        return dict(
            static_trgs={"regions": 1, "path-length": 10 * mp, "n_jump": 5},
            cond_bounds={"regions": (0, np.ceil(w * l / 2 * h)), "path-length": (0, mp), "n_jump": (0, mp // 2)},
            default_weights={"regions": 0, "path-length": 100, "n_jump": 100})
    raise KeyError(problem)


# --------------------------------------------------------------------------------------------
# ControlWrapper loss / reward (control_wrappers.py)
# --------------------------------------------------------------------------------------------
def control_loss(stats, targets, weights, metrics):
    """control_wrappers.py:318-345 get_loss: sum over metrics of -|trg - val| * w.

    A tuple target (lo, hi) costs the minimum distance to the *integers* lo..hi-1 (np.arange, :339).
    """
    loss = 0
    for m in metrics:
        trg = targets[m]
        val = stats[m]
        if isinstance(trg, tuple):
            lm = -abs(np.arange(*trg) - val).min()
        else:
            lm = -abs(trg - val)
        loss += lm * weights[m]
    return loss


def control_metrics(static_trgs, ctrl_metrics):
    """control_wrappers.py:47-84: all_metrics = ctrl metrics + the static-target metrics."""
    names = list(ctrl_metrics or [])
    for k in static_trgs:
        if k not in names:
            names.append(k)
    return names


def metric_weights(default_weights, cfg_weights):
    """control_wrappers.py:41-45: zero for every problem metric, then overridden by cfg.task.weights."""
    w = {k: 0 for k in default_weights}
    w.update(cfg_weights)
    return w


# --------------------------------------------------------------------------------------------
# representations (envs/reps/*.py)
# --------------------------------------------------------------------------------------------
def narrow_coords(map_shape):
    """representation.py:197-198: C-order scan of every cell (last axis fastest)."""
    return np.argwhere(np.ones(map_shape, dtype=bool))


def multiaction_coords(map_shape, act_window):
    """envs/reps/wrappers.py:445-463 MultiActionRepresentation.get_act_coords (+ :404-411 inner pads).

    Positions the patch can be centred on: per axis arange(floor((a-1)/2), dim - ceil((a-1)/2)), combined as
    np.meshgrid(*ranges).T.reshape(-1, ndim): row-major in 2D; in 3D meshgrid's default 'xy' indexing makes the
    LAST axis the slowest, then axis 0, then axis 1.
    """
    nd = len(map_shape)
    lo = [(int(a) - 1) // 2 for a in act_window]
    hi = [int(d) - (int(a) - 1 - (int(a) - 1) // 2) for d, a in zip(map_shape, act_window)]
    rng = [range(lo[i], hi[i]) for i in range(nd)]
    if nd == 2:
        return np.array([[i, j] for i in rng[0] for j in rng[1]], dtype=np.int64)
    return np.array([[i, j, k] for k in rng[2] for i in rng[0] for j in rng[1]], dtype=np.int64)


def rep_update(rep, grid, state, action):
    """One representation update, in place.  Returns `change` (0/1).

    Representation wrappers (envs/reps/wrappers.py), when present in `state`:
      act_window  MultiActionRepresentation.update :466-528 (narrow only): the patch action.reshape(act_window)
                  replaces map[pos - l_pad : pos + r_pad + 1]; change = any cell differs
      static      StaticTileRepresentation.update :358-376: np.where(static < 1, new, old) undoes edits of frozen
                  cells, while `change` stays what the inner update saw on the pre-undo array

    state: dict with 'pos' (list, [y,x] or [z,y,x]) and 'n_step' (narrow only).
    narrow  reps/narrow_rep.py:89-102  (position refreshed with n_step *before* the increment)
    turtle  reps/turtle_rep.py:73-107  (moves clamp, never wrap; moves are never a change)
    wide    reps/wide_rep.py:35-40     (action = [*coords, tile])
    cellular reps/ca_rep.py:31-44      (action[C,*dims] -> argmax over axis 0, lowest index wins)
    """
    static = state.get("static")
    if static is not None:
        old = grid.copy()
        inner = dict(state)
        inner["static"] = None
        change = rep_update(rep, grid, inner, action)
        for k in ("pos", "n_step"):
            if k in inner:
                state[k] = inner[k]
        if change > 0:
            grid[...] = np.where(static < 1, grid, old)
        return change
    if rep == "narrow" and state.get("act_window") is not None:
        aw = [int(a) for a in state["act_window"]]
        patch = np.asarray(action).reshape(aw)
        tl = [state["pos"][i] - (aw[i] - 1) // 2 for i in range(len(aw))]
        assert all(t >= 0 for t in tl) and all(tl[i] + aw[i] <= grid.shape[i] for i in range(len(aw)))
        sl = tuple(slice(tl[i], tl[i] + aw[i]) for i in range(len(aw)))
        change = int(np.any(grid[sl] != patch))
        grid[sl] = patch
        coords = state["coords"]
        state["pos"] = [int(v) for v in coords[state["n_step"] % len(coords)]]
        state["n_step"] += 1
        return change
    if rep == "narrow":
        p = tuple(state["pos"])
        change = int(grid[p] != action)
        grid[p] = action
        coords = state["coords"]
        state["pos"] = [int(v) for v in coords[state["n_step"] % len(coords)]]
        state["n_step"] += 1
        return change
    if rep == "turtle":
        pos = state["pos"]
        if action < 4:
            axis, delta = ((0, -1), (0, 1), (1, -1), (1, 1))[action]      # turtle_rep.py:14
            pos[axis] = min(max(pos[axis] + delta, 0), grid.shape[axis] - 1)
            return 0
        tile = action - 4
        change = int(grid[tuple(pos)] != tile)
        grid[tuple(pos)] = tile
        return change
    if rep == "wide":
        p = tuple(int(v) for v in action[:-1])
        change = int(grid[p] != action[-1])
        grid[p] = action[-1]
        state["pos"] = list(p)
        return change
    if rep == "cellular":
        a = np.asarray(action)
        nxt = a.argmax(axis=0) if a.ndim == grid.ndim + 1 else a
        change = int(np.any(nxt != grid))
        grid[...] = nxt
        return change
    raise KeyError(rep)


# --------------------------------------------------------------------------------------------
# observation wrappers (control_pcgrl/wrappers.py)
# --------------------------------------------------------------------------------------------
def cropped_onehot(grid, pos, obs_window, n_tiles):
    """wrappers.py:407-437 Cropped + :232-257 OneHotEncoding(dim = C+1, channel 0 = out of bounds)."""
    ow = tuple(int(v) for v in obs_window)
    idx = np.zeros(ow, dtype=np.int64)
    for o in np.ndindex(*ow):
        src = tuple(pos[i] + o[i] - ow[i] // 2 for i in range(len(ow)))
        if all(0 <= src[i] < grid.shape[i] for i in range(len(ow))):
            idx[o] = int(grid[src]) + 1
    return np.eye(n_tiles + 1)[idx]


def static_builds_crop(static, pos, obs_window):
    """wrappers.py:451-459 + 407-437: the 'static_builds' observation is the BORDERED mask (dims + 2, border = 1,
    envs/reps/wrappers.py:277,310-312) padded and cropped exactly like the (unbordered) map, so it is sampled one
    cell up-left of the map channels; beyond the border it is padding (0)."""
    ow = tuple(int(v) for v in obs_window)
    bordered = np.ones(tuple(d + 2 for d in static.shape), dtype=np.int64)
    bordered[tuple(slice(1, -1) for _ in static.shape)] = static
    out = np.zeros(ow, dtype=np.float64)
    for o in np.ndindex(*ow):
        src = tuple(pos[i] + o[i] - ow[i] // 2 for i in range(len(ow)))
        if all(0 <= src[i] < bordered.shape[i] for i in range(len(ow))):
            out[o] = bordered[src]
    return out


def agent_occupancy_crop(agent_pos, pos, obs_window, map_shape):
    """ShowAgentRepresentation (envs/reps/wrappers.py:201-232) through Cropped (wrappers.py:407-437, pad value 0 for
    every key but 'map'): 1 where any agent stands, seen from `pos`; out of the map = 0."""
    occ = np.zeros(tuple(map_shape), dtype=np.int64)
    for q in agent_pos:
        occ[tuple(int(v) for v in q)] = 1
    ow = tuple(int(v) for v in obs_window)
    out = np.zeros(ow, dtype=np.float64)
    for o in np.ndindex(*ow):
        src = tuple(pos[i] + o[i] - ow[i] // 2 for i in range(len(ow)))
        if all(0 <= src[i] < occ.shape[i] for i in range(len(ow))):
            out[o] = occ[src]
    return out


def full_onehot(grid, n_tiles):
    """ActionMapImagePCGRLWrapper (wrappers.py:502-526): one-hot of the whole map, no OOB channel."""
    return np.eye(n_tiles)[np.asarray(grid, dtype=np.int64)]


def target_channels(shape_hw, ctrl_metrics, targets, stats, cond_bounds):
    """control_wrappers.py:189-214: channels (2i, 2i+1) = (trg, metric) / |hi - lo|, constant planes."""
    out = np.zeros((*shape_hw, 2 * len(ctrl_metrics)))
    for i, k in enumerate(ctrl_metrics):
        trg = targets[k]
        if isinstance(trg, tuple):
            trg = (trg[0] + trg[1]) / 2
        rng = abs(cond_bounds[k][1] - cond_bounds[k][0])
        out[..., 2 * i] = trg / rng
        out[..., 2 * i + 1] = (stats[k] or 0) / rng
    return out


# --------------------------------------------------------------------------------------------
# the env step (envs/pcgrl_env.py + control_wrappers.py)
# --------------------------------------------------------------------------------------------
class OracleEnv:
    """One env: PcgrlEnv.reset/step (pcgrl_env.py:158-188, 267-342) under ControlWrapper
    (control_wrappers.py:174-187, 216-244).  Grids and start positions are inputs (the reference's
    RNG stream is not part of the contract)."""

    def __init__(self, problem, rep, map_shape, weights=None, controls=None, max_board_scans=3,
                 change_percentage=None, constants=None, reward_mode="control", act_window=None):
        self.problem, self.rep = problem, rep
        self.act_window = None if act_window is None else tuple(int(a) for a in act_window)
        self.reward_mode = reward_mode
        self.map_shape = tuple(map_shape)
        self.n_tiles = len(TILES[problem])
        c = constants or problem_constants(problem, self.map_shape)
        self.static_trgs = dict(c["static_trgs"])
        self.cond_bounds = dict(c["cond_bounds"])
        self.weights = metric_weights(c["default_weights"], weights or {})
        self.ctrl_metrics = list(controls or [])
        self.metrics_used = control_metrics(self.static_trgs, self.ctrl_metrics)
        self.targets = dict(self.static_trgs)
        cells = int(np.prod(self.map_shape))
        self.max_iterations = cells * max_board_scans + 1                       # pcgrl_env.py:241
        self.max_changes = None if change_percentage is None else max(int(change_percentage * cells), 1)
        self.coords = narrow_coords(self.map_shape) if self.act_window is None else \
            multiaction_coords(self.map_shape, self.act_window)

    def _get_stats(self):
        if self.holes is None:
            return get_stats(self.problem, self.grid)
        if self.problem == "minecraft_3D_holey_maze":
            # the problem object lives across resets: path-length reports the previous call's path (:92-93)
            st = get_stats(self.problem, self.grid, self.holes, getattr(self, "_prev_len", 0))
            self._prev_len = st["_next-path-length"]
            return st
        return get_stats(self.problem, self.grid, self.holes)

    def reset(self, grid, pos=None, targets=None, static=None, holes=None, agent_pos=None):
        # multi-agent turtle (envs/reps/wrappers.py:612-651): one position per agent, spawned by the wrapper
        self.agent_pos = None if agent_pos is None else [[int(v) for v in q] for q in agent_pos]
        self.holes = None if holes is None else [int(v) for v in holes]   # pcgrl_holey_env.py:44-45
        if targets:
            self.targets.update(targets)                                        # control_wrappers.py:170-178
        self.grid = np.array(grid, dtype=np.int64).reshape(self.map_shape)
        self.iteration = 0
        self.changes = 0
        self.state = {"coords": self.coords, "n_step": 0, "act_window": self.act_window,
                      "static": None if static is None else np.array(static, dtype=np.int64).reshape(self.map_shape),
                      "pos": [0] * len(self.map_shape) if pos is None else [int(v) for v in pos]}
        if self.rep == "narrow":
            self.state["pos"] = [int(v) for v in self.coords[0]]               # narrow_rep.py:43-50
        self.stats = self._get_stats()                                          # pcgrl_env.py:174-175
        self.last_loss = control_loss(self.stats, self.targets, self.weights, self.metrics_used)
        return self.stats

    def step(self, action, agent=None):
        """agent: MultiAgentWrapper.step (wrappers.py:724-731) is one full env step per agent, in dict order;
        MultiAgentTurtleRepresentation.update :631-647 runs the turtle update at that agent's position."""
        self.iteration += 1                                                     # pcgrl_env.py:279
        if agent is not None:
            self.state["pos"] = list(self.agent_pos[agent])
        change = rep_update(self.rep, self.grid, self.state, action)
        if agent is not None:
            self.agent_pos[agent] = [int(v) for v in self.state["pos"]]
        changed = change > 0
        if changed:
            self.changes += change
        done = self.iteration > self.max_iterations                             # :307
        if self.max_changes is not None:
            done = done or self.changes > self.max_changes                      # :308-309
        old_stats = self.stats
        if changed:
            self.stats = self._get_stats()                                      # :314-323
        if self.reward_mode == "range":                                         # legacy Problem.get_reward
            return legacy_reward(self.problem, self.stats, old_stats), bool(done), changed
        loss = control_loss(self.stats, self.targets, self.weights, self.metrics_used)
        reward = loss - self.last_loss                                          # control_wrappers.py:227-229
        self.last_loss = loss
        return reward, bool(done), changed

    @property
    def pos(self):
        return self.state["pos"]


def actionmap_unravel(action, h, w, n_tiles):
    """wrappers.py:304-323 ActionMap.step: flat -> (y, x, v) over (h, w, dim), then the env is
    stepped with [x, y, v] -- i.e. the wide rep writes _map[x, y] (transposed; SURVEY A-7)."""
    y, x, v = np.unravel_index(int(action), (h, w, n_tiles))
    return [int(x), int(y), int(v)]
