"""CPU: the incremental regions / path-length update the headline kernels implement (oracle/inc_oracle.py, a plain
Python restatement of BinaryIncMachine / LaneGroupInc) against the from-scratch oracle (pcgrl_oracle.binary_stats, which
is pinned on the reference's fixtures): every edit of long random edit sequences, on every board width the kernels
instantiate, plus hand-built maps that merge / split / create / delete components and hit the component that holds the
longest path.  Also pins the cache invariants the kernels rely on: F holds exactly the far tile (np.argmax of the
first sweep) of every non-isolated component, and m lies in a component that attains path-length."""
import numpy as np
import pytest

from oracle import pcgrl_oracle as O
from oracle.inc_oracle import IncBinary


def _check(inc, grid, tag):
    want = O.binary_stats(grid)
    assert (inc.regions, inc.path) == (want["regions"], want["path-length"]), tag
    fresh = IncBinary(grid.tolist())
    assert inc.P == fresh.P and inc.F == fresh.F, tag             # far tiles are those of a from-scratch pass
    if inc.path > 0:                                               # m: a cell of a component whose two-sweep value is path
        assert inc.m >= 0 and (inc.P >> inc.m) & 1, tag
        comp = inc.bfs(1 << inc.m, inc.P & ~(1 << inc.m))[2]
        far = comp & inc.F
        assert bin(far).count("1") == 1, tag
        assert inc.bfs(far, comp & ~far)[0] == inc.path, tag
    else:
        assert inc.m == -1 or inc.bfs(1 << inc.m, inc.P & ~(1 << inc.m))[0] == 0, tag


@pytest.mark.parametrize("shape", [(16, 16), (10, 10), (7, 5), (3, 4), (2, 16), (1, 9), (16, 3)])
@pytest.mark.parametrize("density", [0.3, 0.5, 0.75])
def test_incremental_update_equals_from_scratch(shape, density):
    rng = np.random.default_rng(hash((shape, density)) & 0xFFFF)
    h, w = shape
    grid = (rng.random(shape) < density).astype(np.int64)         # 1 = solid
    inc = IncBinary(grid.tolist())
    _check(inc, grid, "reset")
    n = 0
    for t in range(160):
        if t < 60:                                                 # the narrow representation's scan order
            y, x = divmod(t % (h * w), w)
            new = int(rng.integers(0, 2))
            if new == grid[y, x]:
                continue
        else:
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        grid[y, x] ^= 1
        inc.flip(y, x)
        n += 1
        _check(inc, grid, (t, y, x))
    assert n > 60


def test_incremental_update_hand_built_maps():
    g = np.ones((16, 16), dtype=np.int64)
    g[0, :] = 0                                                    # a corridor: path 15, the maximal component
    g[5, 2:9] = 0                                                  # a second one: path 6
    g[10, 10] = 0                                                  # an isolated cell
    inc = IncBinary(g.tolist())
    _check(inc, g, "reset")
    assert (inc.regions, inc.path) == (3, 15)
    edits = [(0, 7),      # cut the maximal corridor in two (re-sweep of the untouched ones, new maximum 7 or 8)
             (0, 7),      # and mend it again
             (5, 5),      # split the second corridor
             (10, 11),    # the isolated cell gets a neighbour
             (10, 10),    # ... and disappears, leaving the neighbour isolated
             (1, 0), (2, 0), (3, 0), (4, 0), (4, 1), (4, 2),      # a bridge from the first corridor to the second
             (0, 0),      # remove an end cell of the maximal component
             (15, 15), (15, 15)]                                   # create and delete an isolated cell far away
    for y, x in edits:
        g[y, x] ^= 1
        inc.flip(y, x)
        _check(inc, g, (y, x))
    # empty the map cell by cell, then fill it again
    for y in range(16):
        for x in range(16):
            if g[y, x] == 0:
                g[y, x] = 1
                inc.flip(y, x)
    _check(inc, g, "all solid")
    assert (inc.regions, inc.path, inc.F, inc.m) == (0, 0, 0, -1)
    for y in range(0, 16, 3):
        for x in range(16):
            g[y, x] = 0
            inc.flip(y, x)
    _check(inc, g, "stripes")
