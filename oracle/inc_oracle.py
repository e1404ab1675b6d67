"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the INCREMENTAL binary search the headline kernels run
(control_pcgrl_b200/csrc/step_split.cu `BinaryIncMachine`, step_lanegroup.cu `LaneGroupInc`).

The reference recomputes `regions` / `path-length` from scratch after every edit (probs/binary/binary_prob.py:152-158
-> helper.py:200-210 calc_num_regions, :255-276 calc_longest_path).  Both are sums / maxima of per-component terms,
and a component the edit does not touch keeps its first tile, its far tile (np.argmax of the first sweep, :265) and
its eccentricity.  The kernels therefore keep, per env, a cache of

    P     the passable cells                                   (bit y * W + x of a Python int here)
    F     the far tile of every non-isolated component         (one bit per such component)
    m     a cell of a component that attains path-length       (-1 if path-length is 0)

and after a one-cell edit only look at the part of the map the edit can affect.  This module states that update in
plain Python so that `tests/test_inc_oracle_cpu.py` can pin it, edit by edit, on the from-scratch oracle
(pcgrl_oracle.binary_stats, itself pinned on the reference's fixtures) -- without a GPU.  The product never imports it.

Parity status: pinned on pcgrl_oracle (tests/test_inc_oracle_cpu.py).
"""
from __future__ import annotations


class IncBinary:
    def __init__(self, grid):
        """grid: [H][W] tile codes, 0 = empty = passable (binary_prob.py:17)."""
        self.H, self.W = len(grid), len(grid[0])
        self.full = (1 << (self.H * self.W)) - 1
        self.not_col0 = sum(1 << (y * self.W + x) for y in range(self.H) for x in range(1, self.W))
        self.not_colL = sum(1 << (y * self.W + x) for y in range(self.H) for x in range(self.W - 1))
        self.P = sum(1 << (y * self.W + x) for y in range(self.H) for x in range(self.W) if grid[y][x] == 0)
        self._from_scratch()

    # ---- bit-board helpers -------------------------------------------------------------------------------------
    def nbr(self, s):
        """4-neighbourhood of the cells of s."""
        W = self.W
        return (((s << 1) & self.not_col0) | ((s >> 1) & self.not_colL) | (s << W) | (s >> W)) & self.full

    @staticmethod
    def lowest(s):
        return s & -s

    def bfs(self, front, avail):
        """Level-synchronous BFS from `front` inside `avail` (front not in avail).  -> (levels, last non-empty
        frontier, cells reached incl. the start)."""
        levels, reached = 0, front
        while True:
            n = self.nbr(front) & avail
            if not n:
                return levels, front, reached
            avail &= ~n
            reached |= n
            front = n
            levels += 1

    # ---- the from-scratch machine (BinaryMachine in bitboard_machines.cuh; what a reset stores) -------------------
    def _from_scratch(self):
        P = self.P
        iso = P & ~self.nbr(P)
        avail = P & ~iso
        self.regions = bin(iso).count("1")
        fars = 0
        while avail:                                     # first sweeps, one component at a time (row-major first tile)
            s = self.lowest(avail)
            _, last, reached = self.bfs(s, avail & ~s)
            fars |= self.lowest(last)                    # np.argmax: row-major-first maximum
            avail &= ~reached
            self.regions += 1
        self.F = fars
        levels, last, _ = self.bfs(fars, P & ~iso & ~fars) if fars else (0, 0, 0)   # joint second sweep
        self.path = levels
        self.m = self.lowest(last).bit_length() - 1 if fars else -1

    # ---- the incremental update ------------------------------------------------------------------------------------
    def flip(self, y, x):
        """The cell (y, x) changes passability.  Updates P, F, m, regions, path exactly as the kernels do."""
        c = 1 << (y * self.W + x)
        Po, Pn = self.P, self.P ^ c
        d_plus, d_minus = Pn & c, Po & c
        seed = d_plus if d_plus else self.nbr(c) & Pn
        _, _, U = self.bfs(seed, Pn & ~seed) if seed else (0, 0, 0)    # every NEW component the edit affects
        A = U | d_minus                                                  # covers every affected OLD component entirely
        iso_old = Po & ~self.nbr(Po)
        k_old = bin(self.F & A).count("1") + bin(iso_old & A).count("1")
        f_rest = self.F & ~A                                             # far tiles of the untouched components
        hit = self.m >= 0 and (A >> self.m) & 1
        iso = U & ~self.nbr(Pn)
        regions = self.regions - k_old + bin(iso).count("1")
        u_non = U & ~iso
        avail, fars = u_non, 0
        while avail:                                                     # first sweeps inside U
            s = self.lowest(avail)
            _, last, reached = self.bfs(s, avail & ~s)
            fars |= self.lowest(last)
            avail &= ~reached
            regions += 1
        lu, last, _ = self.bfs(fars, u_non & ~fars) if fars else (0, 0, 0)          # joint second sweep inside U
        mcu = self.lowest(last).bit_length() - 1 if fars else -1
        lold, mold = self.path, self.m
        if hit:                                                          # the maximal component was touched:
            if f_rest:                                                   # re-sweep the untouched ones
                lold, last, _ = self.bfs(f_rest, Pn & ~u_non & ~f_rest)
                mold = self.lowest(last).bit_length() - 1
            else:
                lold, mold = 0, -1
        self.P, self.F, self.regions = Pn, f_rest | fars, regions
        self.path = max(lu, lold)
        self.m = mcu if lu > lold else mold
        return self.regions, self.path
