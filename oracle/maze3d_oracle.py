"""TEST INFRASTRUCTURE ONLY -- CPU restatement of minecraft_3D_maze get_stats (control-pcgrl @ 8bde536).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product never does.

Parity status: PINNED -- tests/test_oracle_golden.py checks it against tests/golden/stats_maze3d.npz
(written by oracle/gen_golden.py from the real helper_3D.py / Minecraft3DmazeProblem) and, in the build
container, tests/test_oracle_vs_reference.py compares it with the live reference on fresh random maps.

Citations are relative to /root/reference/control_pcgrl/envs/.  Grids are int arrays [z, y, x]
(0 = AIR, passable; 1 = DIRT).  The reference carries whole path *lists* through its FIFO queue
(helper_3D.py:430,484); only three scalars per cell are observable in the outputs, so this restatement
keeps, per cell: dist (= len(path), the start counts 1), njump, and the order of first recording
(dict insertion order of `paths`).  SURVEY.md section F.
"""
from __future__ import annotations

from collections import deque

import numpy as np

DIRS = ((1, 0), (0, 1), (-1, 0), (0, -1))          # helper_3D.py:220


def count_regions_3d(air):
    """helper_3D.py:396-406 calc_num_regions + :354-383 _flood_fill: 6-connected AIR components."""
    Z, Y, X = air.shape
    seen = np.zeros_like(air, dtype=bool)
    a = air.tolist()
    s = seen.tolist()
    n = 0
    for z in range(Z):
        for y in range(Y):
            for x in range(X):
                if a[z][y][x] and not s[z][y][x]:
                    n += 1
                    s[z][y][x] = True
                    st = [(x, y, z)]
                    while st:
                        cx, cy, cz = st.pop()
                        for dx, dy, dz in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
                            nx, ny, nz = cx + dx, cy + dy, cz + dz
                            if 0 <= nx < X and 0 <= ny < Y and 0 <= nz < Z and a[nz][ny][nx] and not s[nz][ny][nx]:
                                s[nz][ny][nx] = True
                                st.append((nx, ny, nz))
    return n


def moves(a, x, y, z, Z, Y, X):
    """helper_3D.py:214-319 _passable: for each direction at most one foothold, by branch priority
    walk (+1) -> step down (+2) -> step up (+2) -> jump level (+2) / up (+3) / down (+3).
    Yields (nx, ny, nz, cost, is_jump).  `a[z][y][x]` is True for passable (AIR) cells.
    The caller guarantees head-room at (x, y, z), i.e. z + 1 < Z (:443)."""
    out = []
    for dx, dy in DIRS:
        nx, ny, nz = x + dx, y + dy, z
        if nx < 0 or ny < 0 or nx >= X or ny >= Y:                     # :225
            continue
        if (nz == 0 or not a[nz - 1][ny][nx]) and a[nz][ny][nx] and a[nz + 1][ny][nx]:       # :229-237
            out.append((nx, ny, nz, 1, 0))
        elif ((nz - 1 == 0 or (nz - 1 > 0 and not a[nz - 2][ny][nx])) and a[nz - 1][ny][nx]
              and a[nz][ny][nx] and a[nz + 1][ny][nx]):                                      # :243-252
            out.append((nx, ny, nz - 1, 2, 0))
        elif (nz + 2 < Z and not a[nz][ny][nx] and a[nz + 1][ny][nx] and a[nz + 2][ny][nx]
              and a[nz + 2][y][x]):                                                           # :259-266
            out.append((nx, ny, nz + 1, 2, 0))
        else:
            jx, jy, jz = x + 2 * dx, y + 2 * dy, z
            if (nz - 2 >= 0 and nz + 2 < Z and a[nz + 2][ny][nx] and a[nz + 1][ny][nx] and a[nz][ny][nx]
                    and a[nz - 1][ny][nx] and a[nz - 2][ny][nx] and a[nz + 2][y][x]
                    and 0 <= jx < X and 0 <= jy < Y):                                        # :279-288
                if a[jz + 1][jy][jx] and a[jz + 2][jy][jx] and a[jz][jy][jx] and not a[jz - 1][jy][jx]:
                    out.append((jx, jy, jz, 2, 1))                                           # :289-296
                elif (jz + 3 < Z and a[jz + 3][jy][jx] and a[jz + 2][jy][jx] and a[jz + 1][jy][jx]
                      and not a[jz][jy][jx]):
                    out.append((jx, jy, jz + 1, 3, 1))                                       # :297-304
                elif a[jz][jy][jx] and a[jz + 1][jy][jx] and a[jz - 1][jy][jx] and not a[jz - 2][jy][jx]:
                    out.append((jx, jy, jz - 1, 3, 1))                                       # :305-312
    return out


def search(a, sx, sy, sz, Z, Y, X, counters=None):
    """helper_3D.py:422-490 run_dijkstra: FIFO label-correcting search.
    Returns (order, dist, njump): `order` = cells in first-recording order (dict insertion order of
    `paths`), dist[cell] = len(paths[cell]), njump[cell] = jumps[cell]."""
    dist, njump, order = {}, {}, []
    q = deque([(sx, sy, sz, 1, 0)])
    pops = 0
    maxq = 1
    while q:
        cx, cy, cz, ln, nj = q.popleft()
        pops += 1
        c = (cx, cy, cz)
        old = dist.get(c, 0)
        if 0 < old <= ln:                                              # :437-440
            continue
        if cz + 1 == Z or not a[cz + 1][cy][cx]:                       # :443-445 (never records)
            continue
        if old == 0:
            order.append(c)
        dist[c] = ln                                                   # :449
        njump[c] = nj                                                  # :450
        for nx, ny, nz, cost, j in moves(a, cx, cy, cz, Z, Y, X):      # :455-484
            q.append((nx, ny, nz, ln + cost, nj + j))
        if len(q) > maxq:
            maxq = len(q)
    if counters is not None:
        counters["pops"] = max(counters.get("pops", 0), pops)
        counters["maxq"] = max(counters.get("maxq", 0), maxq)
        counters["searches"] = counters.get("searches", 0) + 1
        counters["total_pops"] = counters.get("total_pops", 0) + pops
    return order, dist, njump


def _far(order, dist):
    """np.argmax over the insertion-ordered lengths (helper_3D.py:538-541): first maximum."""
    best, bd = None, -1
    for c in order:
        if dist[c] > bd:
            best, bd = c, dist[c]
    return best, bd


def longest_path_3d(air, counters=None):
    """helper_3D.py:503-563 calc_longest_path -> (path-length, n_jump).

    Reproduces: candidate scan in (z, y, x) order (:22-29, :513); the three skip rules (:515-526);
    `visited_map[np.array(list(paths.keys()))] = 1` (:531), which indexes axis 0 with every x, y and z
    value of every recorded cell, i.e. marks whole z-planes (and raises IndexError when a value is
    >= the depth, so only maps with X, Y <= Z work); n_jump overwritten by every processed component
    while the length keeps the maximum (:553-560)."""
    Z, Y, X = air.shape
    a = air.tolist()
    plane_visited = [False] * Z          # final_visited_map is only ever set by whole planes ...
    cell_visited = set()                 # ... or by the single no-head-room cells of :520-522
    final_value, n_jump = 0, 0
    for z in range(Z):
        for y in range(Y):
            for x in range(X):
                if not a[z][y][x]:
                    continue
                if plane_visited[z] or (x, y, z) in cell_visited:      # :515
                    continue
                if z + 1 == Z or not a[z + 1][y][x]:                   # :520-522
                    cell_visited.add((x, y, z))
                    continue
                if z - 1 < 0 or a[z - 1][y][x]:                        # :525-526
                    continue
                order, dist, _ = search(a, x, y, z, Z, Y, X, counters)
                for (cx, cy, cz) in order:                             # :531
                    for v in (cx, cy, cz):
                        if v >= Z:
                            raise IndexError(f"index {v} is out of bounds for axis 0 with size {Z}")
                        plane_visited[v] = True
                far, _ = _far(order, dist)
                order2, dist2, jumps2 = search(a, far[0], far[1], far[2], Z, Y, X, counters)
                far2, max_dist = _far(order2, dist2)
                n_jump = jumps2[far2]                                  # :553 (last component wins)
                if max_dist > final_value:                             # :558
                    final_value = max_dist
    return final_value, n_jump


def maze3d_stats(grid, counters=None):
    """probs/minecraft/minecraft_3D_maze_prob.py:143-181."""
    air = np.asarray(grid) == 0
    length, nj = longest_path_3d(air, counters)
    return {"regions": count_regions_3d(air), "path-length": int(length), "n_jump": int(nj)}


# --------------------------------------------------------------------------------------------
# minecraft_3D_holey_maze / minecraft_3D_holey_dungeon
# --------------------------------------------------------------------------------------------
def moves_with_traversed(a, x, y, z, Z, Y, X):
    """helper_3D.py:214-319 _passable with the tiles each move passes through (they are appended to the path
    before the foothold, helper_3D.py:484): walk [], step down [(nx,ny,z)], step up [(x,y,z+1)], jump level
    [(nx,ny,z)], jump up [(nx,ny,z),(nx,ny,z+1)], jump down [(nx,ny,z),(nx,ny,z-1)]."""
    out = []
    for (nx, ny, nz, cost, j), (dx, dy) in _moves_dirs(a, x, y, z, Z, Y, X):
        mx, my = x + dx, y + dy
        if j == 0:
            if cost == 1:
                trav = []
            elif nz == z - 1:
                trav = [(mx, my, z)]
            else:
                trav = [(x, y, z + 1)]
        elif nz == z:
            trav = [(mx, my, z)]
        elif nz == z + 1:
            trav = [(mx, my, z), (mx, my, z + 1)]
        else:
            trav = [(mx, my, z), (mx, my, z - 1)]
        out.append((nx, ny, nz, trav, j))
    return out


def _moves_dirs(a, x, y, z, Z, Y, X):
    """moves() paired with the direction that produced each foothold (same branch order)."""
    out = []
    for d in DIRS:
        saved = globals()["DIRS"]
        try:
            globals()["DIRS"] = (d,)
            for m in moves(a, x, y, z, Z, Y, X):
                out.append((m, d))
        finally:
            globals()["DIRS"] = saved
    return out


def search_paths(a, sx, sy, sz, Z, Y, X):
    """helper_3D.py:422-490 run_dijkstra keeping the path lists: -> (order, paths, njump)."""
    paths, njump, order = {}, {}, []
    q = deque([(sx, sy, sz, [(sx, sy, sz)], 0)])
    while q:
        cx, cy, cz, path, nj = q.popleft()
        c = (cx, cy, cz)
        old = paths.get(c, [])
        if 0 < len(old) <= len(path):                                  # :437-440
            continue
        if cz + 1 == Z or not a[cz + 1][cy][cx]:                       # :443-445
            continue
        if c not in paths:
            order.append(c)
        paths[c] = path
        njump[c] = nj
        for nx, ny, nz, trav, j in moves_with_traversed(a, cx, cy, cz, Z, Y, X):
            q.append((nx, ny, nz, path + trav + [(nx, ny, nz)], nj + j))
    return order, paths, njump


def remove_stacked(path):
    """helper_3D.py:657-675 remove_stacked_path_tiles: drop every tile that sits directly on top of another tile
    of the path (the iteration runs over a snapshot, so a tile goes iff the one below it was in the ORIGINAL set)."""
    s = set(tuple(e) for e in path)
    return [c for c in s if (c[0], c[1], c[2] - 1) not in s]


def maze3d_holey_stats(bordered, entrance, exit_, prev_path_len):
    """probs/minecraft/minecraft_3D_holey_maze_prob.py:71-124 on the bordered map (holes already dug).
    entrance / exit_: ((z, y, x) foot, (z, y, x) head).  Returns (stats, len(path_coords)) -- `path-length` is
    len(self.path_coords) of the PREVIOUS call (:92-93 assign in the wrong order), so the caller carries the
    second value to the next call (0 before the first)."""
    air = np.asarray(bordered) == 0
    Z, Y, X = air.shape
    a = air.tolist()
    ez, ey, ex = (int(v) for v in entrance[0])
    order, paths, jumps = search_paths(a, ex, ey, ez, Z, Y, X)        # :81
    xz, xy, xx = (int(v) for v in exit_[0])
    xc = (xx, xy, xz)                                                 # :83
    connected = len(paths[xc]) if xc in paths else -1                 # :84
    n_jump = jumps[xc] if xc in jumps else 0                          # :89
    best, bl = None, -1
    for c in order:                                                   # :90-92 np.argmax over insertion order
        if len(paths[c]) > bl:
            best, bl = c, len(paths[c])
    new_len = len(remove_stacked(paths[best]))                        # :94
    return ({"regions": count_regions_3d(air), "path-length": int(prev_path_len),
             "connected-path-length": int(connected), "n_jump": int(n_jump)}, new_len)


def maze3d_holey_dungeon_stats(bordered, entrance, exit_):
    """probs/minecraft/minecraft_3D_holey_dungeon_prob.py:95-146 on the bordered map (holes already dug).
    Tiles: AIR 0, DIRT 1, CHEST 2, SKULL 3, PUMPKIN 4; the player walks through everything but DIRT (:27), regions
    are counted over AIR only (:99).  entrance / exit_: ((z, y, x) foot, (z, y, x) head)."""
    g = np.asarray(bordered)
    Z, Y, X = g.shape
    a = (g != 1).tolist()
    chests = [(int(x), int(y), int(z)) for z, y, x in np.argwhere(g == 2)]          # get_tile_locations order z, y, x
    enemies = [(int(x), int(y), int(z)) for z, y, x in np.argwhere((g == 3) | (g == 4))]
    st = {"regions": count_regions_3d(g == 0), "path-length": 0, "chests": len(chests), "enemies": len(enemies),
          "nearest-enemy": 0, "n_jump": 0}
    ez, ey, ex = (int(v) for v in entrance[0])
    dist = nj = None
    if enemies:                                                                    # :113-124
        _, dist, nj = search(a, ex, ey, ez, Z, Y, X)
        best = 0
        for e in enemies:
            d = dist.get(e, 0)
            if d > 0 and (d < best or best == 0):
                best = d
        st["nearest-enemy"] = best
    if chests:                                                                     # :127-142
        c = chests[0]
        if dist is None:
            _, dist, nj = search(a, ex, ey, ez, Z, Y, X)
        st["path-length"] += dist.get(c, 0)
        st["n_jump"] += nj.get(c, 0)
        xz, xy, xx = (int(v) for v in exit_[0])
        _, dist2, nj2 = search(a, c[0], c[1], c[2], Z, Y, X)
        st["path-length"] += dist2.get((xx, xy, xz), 0)
        st["n_jump"] += nj2.get((xx, xy, xz), 0)
    return st
