"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the search-based get_stats of control-pcgrl @ 8bde536:
sokoban (BFS -> A*(1) -> A*(.5) -> A*(0) solver) and smb (A*(1) -> A*(0) playthrough).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product never does.

Parity status: PINNED -- tests/test_oracle_golden.py checks it against tests/golden/stats_sokoban.npz and
stats_smb.npz (written by oracle/gen_golden.py from the real SokobanCtrlProblem / SMBCtrlProblem), and in
the build container tests/test_oracle_vs_reference.py compares it with the live reference engines.

Citations are relative to /root/reference/control_pcgrl/envs/probs/.  The reference's A* uses
queue.PriorityQueue, i.e. the stdlib heapq on Node objects whose __lt__ compares
heuristic + balance * depth (sokoban/sokoban/engine.py:49-50, smb/smb/engine.py:50-51); results depend on
how heapq breaks ties, so the binary heap below restates CPython's heappush/heappop move for move
(Lib/heapq.py: _siftdown / _siftup) using only `<` -- this is also the spec the CUDA kernels follow.
"""
from __future__ import annotations

from collections import deque

import numpy as np

from .pcgrl_oracle import count_regions

SOLVER_POWER = 10000                                  # sokoban_prob.py:40, smb_prob.py:21


# --------------------------------------------------------------------------------------------
# CPython heapq, restated (only `<` on priorities; items are opaque)
# --------------------------------------------------------------------------------------------
class Heap:
    def __init__(self):
        self.pri = []
        self.item = []

    def __len__(self):
        return len(self.pri)

    def _siftdown(self, start, pos):
        pri, item = self.pri, self.item
        np_, ni = pri[pos], item[pos]
        while pos > start:
            parent = (pos - 1) >> 1
            if np_ < pri[parent]:
                pri[pos], item[pos] = pri[parent], item[parent]
                pos = parent
                continue
            break
        pri[pos], item[pos] = np_, ni

    def push(self, p, it):
        self.pri.append(p)
        self.item.append(it)
        self._siftdown(0, len(self.pri) - 1)

    def pop(self):
        pri, item = self.pri, self.item
        lp, li = pri.pop(), item.pop()
        if not pri:
            return lp, li
        rp, ri = pri[0], item[0]
        # _siftup(heap, 0) with the last element as the new item
        end, pos = len(pri), 0
        child = 1
        while child < end:
            right = child + 1
            if right < end and not pri[child] < pri[right]:
                child = right
            pri[pos], item[pos] = pri[child], item[child]
            pos = child
            child = 2 * pos + 1
        pri[pos], item[pos] = lp, li
        self._siftdown(0, pos)
        return rp, ri


# --------------------------------------------------------------------------------------------
# sokoban
# --------------------------------------------------------------------------------------------
SOK_DIRS = ((-1, 0), (1, 0), (0, -1), (0, 1))         # sokoban/engine.py:3


class SokobanLevel:
    """State.stringInitialize + intializeDeadlocks (sokoban/engine.py:137-246) for the bordered level that
    SokobanProblem._run_game builds (sokoban_prob.py:99-123): a '#' frame around the map."""

    def __init__(self, grid):
        g = np.asarray(grid)
        h, w = g.shape
        self.h, self.w = h + 2, w + 2
        self.solid = [[True] * self.w for _ in range(self.h)]
        self.targets, crates, self.player = [], [], None
        for y in range(h):
            for x in range(w):
                t = int(g[y, x])
                self.solid[y + 1][x + 1] = t == 1
                if t == 2:
                    self.player = (x + 1, y + 1)
                elif t == 3:
                    crates.append((x + 1, y + 1))
                elif t == 4:
                    self.targets.append((x + 1, y + 1))
        self.crates0 = tuple(crates)
        self.target_set = set(self.targets)
        self.deadlock = self._deadlocks()

    def _deadlocks(self):
        s, W, H = self.solid, self.w, self.h
        sign = lambda v: (v > 0) - (v < 0)
        dead = [[False] * W for _ in range(H)]
        corners = []
        for y in range(H):
            for x in range(W):
                if x == 0 or y == 0 or x == W - 1 or y == H - 1 or s[y][x]:
                    continue
                if ((s[y - 1][x] and s[y][x - 1]) or (s[y - 1][x] and s[y][x + 1]) or
                        (s[y + 1][x] and s[y][x - 1]) or (s[y + 1][x] and s[y][x + 1])):
                    if (x, y) not in self.target_set:
                        corners.append((x, y))
                        dead[y][x] = True
        for c1 in corners:                                            # engine.py:218-246
            for c2 in corners:
                dx, dy = sign(c1[0] - c2[0]), sign(c1[1] - c2[1])
                if (dx == 0 and dy == 0) or (dx != 0 and dy != 0):
                    continue
                walls = []
                x, y = c2
                if dx != 0:
                    x += dx
                    while x != c1[0]:
                        if (x, y) in self.target_set or s[y][x] or (not s[y - 1][x] and not s[y + 1][x]):
                            walls = []
                            break
                        walls.append((x, y))
                        x += dx
                if dy != 0:
                    y += dy
                    while y != c1[1]:
                        if (x, y) in self.target_set or s[y][x] or (not s[y][x - 1] and not s[y][x + 1]):
                            walls = []
                            break
                        walls.append((x, y))
                        y += dy
                for wx, wy in walls:
                    dead[wy][wx] = True
        return dead

    # state = (player, crates) with crates an index-ordered tuple (engine.py:329-335: the key lists
    # the crates in list order, so permuted crate sets are different states)
    def win(self, crates):                                            # engine.py:269-277
        if len(self.targets) != len(crates) or not crates:
            return False
        cs = set(crates)
        return all(t in cs for t in self.targets)

    def heuristic(self, crates):                                      # engine.py:279-293
        targets = list(self.targets)
        distance = 0
        for cx, cy in crates:
            best_dist, best = self.w + self.h, 0
            for i, (tx, ty) in enumerate(targets):
                d = abs(cx - tx) + abs(cy - ty)
                if best_dist > d:
                    best, best_dist = i, d
            tx, ty = targets[best]
            distance += abs(tx - cx) + abs(ty - cy)
            del targets[best]
        return distance

    def _movable(self, x, y, crates):                                 # engine.py:266-267
        return 0 <= x < self.w and 0 <= y < self.h and not self.solid[y][x] and (x, y) not in crates

    def children(self, player, crates):
        """Node.getChildren (engine.py:14-24) over State.update (:295-327)."""
        if self.win(crates):                                          # update() returns at once when won
            return []
        out = []
        px, py = player
        for dx, dy in SOK_DIRS:
            nx, ny = px + dx, py + dy
            if self._movable(nx, ny, crates):
                out.append(((nx, ny), crates))
                continue
            if (nx, ny) in crates:
                cx, cy = nx + dx, ny + dy
                if self._movable(cx, cy, crates):
                    i = crates.index((nx, ny))
                    nc = crates[:i] + ((cx, cy),) + crates[i + 1:]
                    if any(self.deadlock[y][x] for x, y in nc):       # crateMove and checkDeadlock()
                        continue
                    out.append(((nx, ny), nc))
        return out


def _better(h, depth, best):
    return best is None or h < best[0] or (h == best[0] and depth < best[1])


def sokoban_search(level, balance, max_iterations=SOLVER_POWER):
    """BFSAgent (balance is None, engine.py:56-74) / AStarAgent (engine.py:96-119).
    -> (won, depth of the returned node, heuristic of the returned node, iterations)."""
    root = (level.player, level.crates0)
    if balance is None:
        queue = deque([(root, 0)])
        pop = queue.popleft
        push = lambda st, d: queue.append((st, d))
    else:
        heap = Heap()
        queue = heap
        pop = lambda: heap.pop()[1]
        push = lambda st, d: heap.push(level.heuristic(st[1]) + balance * d, (st, d))
        push(root, 0)
        queue = heap
    visited = set()
    best = None
    iterations = 0
    while iterations < max_iterations and len(queue) > 0:
        iterations += 1
        st, depth = pop()
        if level.win(st[1]):
            return True, depth, 0, iterations
        if st not in visited:
            h = level.heuristic(st[1])
            if _better(h, depth, best):
                best = (h, depth)
            visited.add(st)
            for ch in level.children(*st):
                push(ch, depth + 1)
    return False, best[1], best[0], iterations


def sokoban_run_game(grid, counters=None):
    """SokobanProblem._run_game (sokoban_prob.py:99-148) -> (dist-win, sol-length)."""
    level = SokobanLevel(grid)
    h = 0
    for k, balance in enumerate((None, 1, 0.5, 0)):
        won, depth, h, iters = sokoban_search(level, balance)
        if counters is not None:
            counters["max_iters"] = max(counters.get("max_iters", 0), iters)
            counters["searches"] = counters.get("searches", 0) + 1
            if k > 0:
                counters["astar_runs"] = counters.get("astar_runs", 0) + 1
        if won:
            return 0, depth
    return h, 0


def sokoban_stats(grid, counters=None):
    """sokoban_prob.py:160-180 + sokoban_ctrl_prob.py:58-65."""
    g = np.asarray(grid)
    h, w = g.shape
    st = {"player": int((g == 2).sum()), "crate": int((g == 3).sum()), "target": int((g == 4).sum()),
          "regions": count_regions(g, [0, 2, 3, 4]), "dist-win": w * h * (w + h), "sol-length": 0}
    if st["player"] == 1 and st["crate"] == st["target"] and st["crate"] > 0 and st["regions"] == 1:
        st["dist-win"], st["sol-length"] = sokoban_run_game(g, counters)
    st["ratio"] = abs(st["crate"] - st["target"])
    return st


# --------------------------------------------------------------------------------------------
# smb
# --------------------------------------------------------------------------------------------
SMB_DIRS = ((0, 0), (1, 0), (0, -1), (1, -1))         # smb/engine.py:3
SMB_SOLID = (False, True, False, True, True, False, True)   # " # ## #" (smb_prob.py:97)


class SmbLevel:
    """The level SMBProblem._run_game builds (smb_prob.py:96-113): three extra columns on each side;
    rows above `height-3` are open with the exit marker '|' in the column after the map, row
    `height-3` holds the player '@' at x = 1 and a '#' at x = W+4, the rows below are '###' on both
    sides.  `height` is the problem's _height, i.e. map_shape[0] after adjust_param (SURVEY A-15)."""

    def __init__(self, grid, prob_height=None):
        g = np.asarray(grid)
        rows, cols = g.shape
        ph = rows if prob_height is None else prob_height
        self.h, self.w = rows, cols + 6
        self.solid = [[False] * self.w for _ in range(rows)]
        self.exit = -1
        self.player = None
        for i in range(rows):
            for j in range(cols):
                self.solid[i][j + 3] = SMB_SOLID[int(g[i, j])]
            if i < ph - 3:
                self.exit = cols + 4                                   # '|' of "   ...row... | "
            elif i == ph - 3:
                self.player = (1, i)
                self.solid[i][cols + 4] = True
            else:
                for x in (0, 1, 2, cols + 3, cols + 4, cols + 5):
                    self.solid[i][x] = True

    def movable(self, x, y):                                           # engine.py:190-193
        if y < 0:
            return True
        return not (x < 0 or x >= self.w or y >= self.h or self.solid[y][x])

    def update(self, s, dx, dy):
        """State.update (engine.py:195-237).  s = (x, y, airTime, jumps, last_jump_x, max_gap)."""
        x, y, air, jumps, last_jx, gap = s
        if x >= self.exit or y >= self.h:                              # checkOver
            return s
        dy = -1 if dy < 0 else 0
        ground = False
        if -1 <= y < self.h - 1:
            ground = self.solid[y + 1][x]
        nx, ny = x, y
        if dx != 0 and self.movable(nx + dx, ny):
            nx += dx
        if dy == -1:
            if ground and self.movable(nx, ny - 1):
                air = 5
                jumps += 1
                gap = max(gap, x - last_jx)                            # smb_prob.py:147-150, folded in
                last_jx = x
        elif air > 0:
            air = 1
        if air > 1:
            air -= 1
            if self.movable(nx, ny - 1):
                ny -= 1
            else:
                air = 1
        elif air == 1:
            air = 0
        elif self.movable(nx, ny + 1):
            ny += 1
        return (nx, ny, air, jumps, last_jx, gap)


def smb_search(level, balance, max_iterations=SOLVER_POWER):
    """smb/engine.py:105-129 AStarAgent.getSolution -> (won, node state, depth, iterations)."""
    heap = Heap()
    px, py = level.player
    root = (px, py, 0, 0, 0, 0)
    heap.push((level.exit - px) + balance * 0, (root, 0))
    visited = set()
    best = None
    iterations = 0
    while iterations < max_iterations and len(heap) > 0:
        iterations += 1
        _, (s, depth) = heap.pop()
        if s[1] >= level.h:                                            # checkLose: skipped, but counted
            continue
        if s[0] >= level.exit:
            return True, s, depth, iterations
        key = s[:3]
        if key not in visited:
            h = level.exit - s[0]
            if best is None or h < best[0] or (h == best[0] and depth < best[1]):
                best = (h, depth, s)
            visited.add(key)
            for dx, dy in SMB_DIRS:
                c = level.update(s, dx, dy)
                heap.push((level.exit - c[0]) + balance * (depth + 1), (c, depth + 1))
    return False, best[2], best[1], iterations


def smb_run_game(grid, prob_width=None, prob_height=None, counters=None):
    """SMBProblem._run_game + the jumps / jumps-dist post-processing of get_stats
    (smb_prob.py:96-130, 144-153) -> (sol-length, dist-win, jumps, jumps-dist)."""
    g = np.asarray(grid)
    level = SmbLevel(g, prob_height)
    pw = g.shape[1] if prob_width is None else prob_width
    for balance in (1, 0):
        won, s, depth, iters = smb_search(level, balance)
        if counters is not None:
            counters["max_iters"] = max(counters.get("max_iters", 0), iters)
            counters["total_iters"] = counters.get("total_iters", 0) + iters
            counters["searches"] = counters.get("searches", 0) + 1
        if won:
            break
    jumps_dist = max(s[5], pw - s[4])
    if won:
        return depth, 0, s[3], jumps_dist
    return 0, level.exit - s[0], s[3], jumps_dist


def smb_stats(grid, prob_width=None, prob_height=None, counters=None):
    """smb_prob.py:132-154 (helper.py: get_floor_dist :59, get_type_grouping :103, get_changes :123)."""
    g = np.asarray(grid)
    rows, cols = g.shape
    floor = np.isin(g, [1, 3, 4])              # solid, brick, question ("tube_left/right" never occur)
    dist_floor = 0
    for y, x in zip(*np.nonzero(g == 2)):      # helper.py:40-46 _calc_dist_floor
        d = rows - 1                           # no floor below: len(map) - 1
        for dy in range(1, rows):
            if y + dy >= rows:
                break
            if floor[y + dy, x]:
                d = dy - 1
                break
        dist_floor += d
    tube = g == 6
    left = np.zeros_like(tube)
    left[:, 1:] = tube[:, :-1]
    right = np.zeros_like(tube)
    right[:, :-1] = tube[:, 1:]
    disjoint = int((tube & (left.astype(int) + right.astype(int) == 1)).sum())
    noise = int((g[:, 1:] != g[:, :-1]).sum() + (g[1:, :] != g[:-1, :]).sum())
    sol, dist_win, jumps, jumps_dist = smb_run_game(g, prob_width, prob_height, counters)
    return {"dist-floor": int(dist_floor), "disjoint-tubes": disjoint, "enemies": int((g == 2).sum()),
            "empty": int((g == 0).sum()), "noise": noise, "jumps": int(jumps), "jumps-dist": int(jumps_dist),
            "dist-win": int(dist_win), "sol-length": int(sol)}
