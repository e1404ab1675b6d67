"""Problem tables: tile codes, stat order, static targets / bounds / default weights.

Values are re-derived here from the reference's Problem constructors (paths relative to
/root/reference/control_pcgrl/envs/probs/); tests pin them against the real classes.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field


INF = float("inf")


def get_range_reward(new_value, old_value, low, high):
    """envs/helper.py:550-560 -- the legacy piecewise distance-to-band reward (public helper of the reference)."""
    if low <= new_value <= high and low <= old_value <= high:
        return 0
    if old_value <= high and new_value <= high:
        return min(new_value, low) - min(old_value, low)
    if old_value >= low and new_value >= low:
        return max(old_value, high) - max(new_value, high)
    if new_value > high and old_value < low:
        return high - new_value + old_value - low
    if new_value < low and old_value > high:
        return high - old_value + new_value - low
    return None


@dataclass
class ProblemSpec:
    name: str
    tiles: list
    stat_names: list
    init_probs: list
    border_tile: str
    ndim: int
    static_trgs: "OrderedDict[str, object]" = field(default_factory=OrderedDict)
    cond_bounds: dict = field(default_factory=dict)
    reward_weights: dict = field(default_factory=dict)   # the Problem's own _reward_weights (keys matter)
    # legacy Problem.get_reward (helper.get_range_reward): stat -> (low, high) band, and the weights the
    # non-ctrl Problem class hard-codes.  Empty = the reference defines no such reward for this problem.
    range_bands: dict = field(default_factory=dict)
    range_weights: dict = field(default_factory=dict)

    @property
    def n_tiles(self):
        return len(self.tiles)

    @property
    def n_stats(self):
        return len(self.stat_names)


def binary_spec(map_shape):
    """binary/binary_prob.py:17-84."""
    h, w = map_shape[0], map_shape[1]
    max_path = float(math.ceil(w / 2) * h + math.floor(h / 2))
    return ProblemSpec(
        name="binary", tiles=["empty", "solid"], stat_names=["regions", "path-length"],
        init_probs=[0.5, 0.5], border_tile="solid", ndim=2,
        static_trgs=OrderedDict([("regions", 1), ("path-length", max_path)]),
        cond_bounds={"regions": (0, float(w * math.ceil(h / 2))), "path-length": (0, max_path)},
        reward_weights={"regions": 100, "path-length": 100},
        # binary_prob.py:170-178 (weights :37-40)
        range_bands={"regions": (1, 1), "path-length": (125, 125)},
        range_weights={"regions": 100, "path-length": 100})


def binary_holey_spec(map_shape):
    """binary/binary_holey_prob.py:12-42 on top of BinaryProblem's constructor: stats are taken
    on the bordered map with an entrance and an exit dug into the border."""
    s = binary_spec(map_shape)
    mp = s.static_trgs["path-length"]
    s.name = "binary_holey"
    s.stat_names = ["regions", "path-length", "connected-path-length"]
    s.static_trgs = OrderedDict([("regions", 1), ("path-length", mp + 2), ("connected-path-length", mp + 2)])
    s.cond_bounds = {"regions": s.cond_bounds["regions"], "path-length": (0, mp + 2),
                     "connected-path-length": (0, mp + 2)}
    s.reward_weights = {"regions": 100, "path-length": 0, "connected-path-length": 100}
    # binary_holey_prob.py:106-117: the third term is get_range_reward over *path-length* again, weighted by
    # connected-path-length's weight -- not a per-stat sum, so the in-kernel range mode does not offer it
    s.range_bands, s.range_weights = {}, {}
    return s


def minecraft_2d_maze_spec(map_shape):
    """minecraft/minecraft_2D_maze_prob.py:15-33, 87-93, 106-115: the binary stats over the tiles
    AIR / DIRT.  The class is not a controllable problem upstream (no static_trgs / cond_bounds, and its
    constructor calls Problem.__init__ without the cfg it now requires), so there are no ControlWrapper targets to
    mirror: the reward it defines is the legacy range reward (regions -> 1, the longer the path the better)."""
    return ProblemSpec(
        name="minecraft_2D_maze", tiles=["AIR", "DIRT"], stat_names=["regions", "path-length"],
        init_probs=[0.5, 0.5], border_tile="DIRT", ndim=2,
        static_trgs=OrderedDict(), cond_bounds={}, reward_weights={"regions": 5, "path-length": 1},
        range_bands={"regions": (1, 1), "path-length": (INF, INF)},
        range_weights={"regions": 5, "path-length": 1})


def zelda_spec(map_shape):
    """zelda/zelda_prob.py:20-45 + zelda/zelda_ctrl_prob.py:17-75."""
    h, w = map_shape[0], map_shape[1]
    max_near = float(math.ceil(w / 2 + 1) * h)
    max_path = float((math.ceil(w / 2) * h + math.floor(h / 2)) * 2 - 1)
    return ProblemSpec(
        name="zelda", tiles=["empty", "solid", "player", "key", "door", "bat", "scorpion", "spider"],
        stat_names=["player", "key", "door", "enemies", "regions", "nearest-enemy", "path-length"],
        init_probs=[0.58, 0.3, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02], border_tile="solid", ndim=2,
        static_trgs=OrderedDict([("enemies", (2, 5)), ("path-length", max_path), ("nearest-enemy", (5, max_near)),
                                 ("regions", 1), ("player", 1), ("key", 1), ("door", 1)]),
        cond_bounds={"nearest-enemy": (0, max_near), "enemies": (0, w * h - 2), "player": (0, w * h - 2),
                     "key": (0, w * h - 2), "door": (0, w * h - 2), "regions": (0, w * h / 2),
                     "path-length": (0, max_path)},
        reward_weights={"player": 3, "key": 3, "door": 3, "regions": 5, "enemies": 1, "nearest-enemy": 1,
                        "path-length": 1},
        # zelda_prob.py:135-153 (ZeldaProblem; _max_enemies 5, _target_enemy_dist 4, weights :33-41)
        range_bands={"player": (1, 1), "key": (1, 1), "door": (1, 10), "enemies": (2, 5), "regions": (1, 1),
                     "nearest-enemy": (4, INF), "path-length": (INF, INF)},
        range_weights={"player": 3, "key": 3, "door": 3, "regions": 5, "enemies": 1, "nearest-enemy": 2,
                       "path-length": 1})


def sokoban_spec(map_shape):
    """sokoban/sokoban_prob.py:26-52 + sokoban/sokoban_ctrl_prob.py:11-56.  The constructor hard-codes 5x5
    before the targets / bounds are derived; adjust_param later overwrites _height/_width only (SURVEY A-15)."""
    w = h = 5
    max_path = float(math.ceil(w / 2 + 1) * h)
    return ProblemSpec(
        name="sokoban", tiles=["empty", "solid", "player", "crate", "target"],
        stat_names=["player", "crate", "target", "regions", "dist-win", "sol-length", "ratio"],
        init_probs=[0.45, 0.4, 0.05, 0.05, 0.05], border_tile="solid", ndim=2,
        static_trgs=OrderedDict([("player", 1), ("crate", (2, 3)), ("regions", 1), ("ratio", 0), ("dist-win", 0),
                                 ("sol-length", max_path)]),
        cond_bounds={"player": (1, w * h), "crate": (1, w * h / 2 - max(w, h)), "target": (1, w * h),
                     "ratio": (0, w * h), "dist-win": (0, w * h * (w + h)), "sol-length": (0, 2 * max_path),
                     "regions": (0, w * h / 2)},
        reward_weights={"player": 3, "crate": 1, "regions": 5, "ratio": 2, "dist-win": 0.0, "sol-length": 1},
        # sokoban_prob.py:185-229 (SokobanProblem; _max_crates 3, weights :44-52)
        range_bands={"player": (1, 1), "crate": (1, 3), "target": (1, 3), "regions": (1, 1), "ratio": (-INF, -INF),
                     "dist-win": (-INF, -INF), "sol-length": (INF, INF)},
        range_weights={"player": 3, "crate": 2, "target": 2, "regions": 5, "ratio": 2, "dist-win": 0.0,
                       "sol-length": 1})


def smb_spec(map_shape):
    """smb/smb_prob.py:12-37 + smb/smb_ctrl_prob.py:8-35 (targets / bounds from the hard-coded 116x16)."""
    w, h = 116, 16
    max_sol = float(math.ceil(w) * 3)
    return ProblemSpec(
        name="smb", tiles=["empty", "solid", "enemy", "brick", "question", "coin", "tube"],
        stat_names=["dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist", "dist-win",
                    "sol-length"],
        init_probs=[0.75, 0.1, 0.01, 0.04, 0.01, 0.02, 0.02], border_tile="empty", ndim=2,
        static_trgs=OrderedDict([("dist-floor", 0), ("disjoint-tubes", 0), ("enemies", (10, 30)),
                                 ("empty", (900, w * h)), ("noise", 0), ("jumps", (20, w * h)), ("jumps-dist", 0),
                                 ("dist-win", 0), ("sol-length", max_sol)]),
        cond_bounds={"dist-floor": (0, w * h), "disjoint-tubes": (0, w * h), "enemies": (0, w * h), "empty": (0, w),
                     "noise": (0, w * h), "jumps": (0, w), "jumps-dist": (0, w * h), "dist-win": (0, w),
                     "sol-length": (0, max_sol)},
        reward_weights={"dist-floor": 2, "disjoint-tubes": 1, "enemies": 1, "empty": 1, "noise": 4, "jumps": 2,
                        "jumps-dist": 2, "dist-win": 5, "sol-length": 1},
        # smb_prob.py:156-176 (sol-length is not part of the legacy sum)
        range_bands={"dist-floor": (0, 0), "disjoint-tubes": (0, 0), "enemies": (10, 30), "empty": (900, INF),
                     "noise": (0, 0), "jumps": (20, INF), "jumps-dist": (0, 0), "dist-win": (0, 0)},
        range_weights={"dist-floor": 2, "disjoint-tubes": 1, "enemies": 1, "empty": 1, "noise": 4, "jumps": 2,
                       "jumps-dist": 2, "dist-win": 5})


def minecraft_3d_maze_spec(map_shape):
    """minecraft/minecraft_3D_maze_prob.py:26-81 (targets / bounds from the hard-coded 15^3)."""
    w = h = l = 15
    max_path = float(2 * (h // 3) * (math.ceil(w / 2) * l + math.floor(l / 2)))
    return ProblemSpec(
        name="minecraft_3D_maze", tiles=["AIR", "DIRT"], stat_names=["regions", "path-length", "n_jump"],
        init_probs=[1.0, 0.0], border_tile="DIRT", ndim=3,
        static_trgs=OrderedDict([("regions", 1), ("path-length", 10 * max_path), ("n_jump", 5)]),
        cond_bounds={"regions": (0, float(math.ceil(w * l / 2 * h))), "path-length": (0, max_path),
                     "n_jump": (0, max_path // 2)},
        reward_weights={"regions": 0, "path-length": 100, "n_jump": 100})


def minecraft_3d_holey_maze_spec(map_shape):
    """minecraft/minecraft_3D_holey_maze_prob.py:26-61 on top of Minecraft3DmazeProblem's constructor.  Five stats:
    the four get_stats returns (:124-130) plus `_next-path-length`, the length found by this call, which the
    reference reports as path-length one call late (:92-93) -- carried in the stats so the kernel needs no other
    per-env state (weight 0, not a metric)."""
    s = minecraft_3d_maze_spec(map_shape)
    mp = s.cond_bounds["path-length"][1]
    s.name = "minecraft_3D_holey_maze"
    s.stat_names = ["regions", "path-length", "connected-path-length", "n_jump", "_next-path-length"]
    s.static_trgs = OrderedDict([("regions", 1), ("path-length", 10 * mp), ("n_jump", 5),
                                 ("connected-path-length", 10 * mp)])
    s.cond_bounds = {"regions": s.cond_bounds["regions"], "path-length": (0, mp + 2),
                     "connected-path-length": (0, mp + 2), "n_jump": (0, mp // 2)}
    s.reward_weights = {"regions": 0, "path-length": 100, "connected-path-length": 120, "n_jump": 150}
    return s


def minecraft_3d_dungeon_holey_spec(map_shape):
    """minecraft/minecraft_3D_holey_dungeon_prob.py:17-93 (sizes from Minecraft3DmazeProblem's hard-coded 15^3)."""
    w = h = l = 15
    mp = float(2 * (h // 3) * (math.ceil(w / 2) * l + math.floor(l / 2)))
    max_any = w * h * l // 4
    return ProblemSpec(
        name="minecraft_3D_dungeon_holey", tiles=["AIR", "DIRT", "CHEST", "SKULL", "PUMPKIN"],
        stat_names=["regions", "path-length", "chests", "enemies", "nearest-enemy", "n_jump"],
        init_probs=[1.0, 0.0, 0.0, 0.0, 0.0], border_tile="DIRT", ndim=3,
        static_trgs=OrderedDict([("enemies", (2, 5)), ("regions", 1), ("path-length", 10 * mp),
                                 ("nearest-enemy", (5, mp // 2)), ("chests", 1), ("n_jump", (2, 5))]),
        cond_bounds={"regions": (0, float(math.ceil(w * l / 2 * h))), "path-length": (0, mp), "chests": (0, max_any),
                     "n_jump": (0, mp // 2), "nearest-enemy": (0, mp // 2), "enemies": (0, max_any)},
        reward_weights={"regions": 0, "path-length": 100, "chests": 300, "n_jump": 100, "enemies": 100,
                        "nearest-enemy": 200})


_SPECS = {"binary": binary_spec, "zelda": zelda_spec, "sokoban": sokoban_spec, "smb": smb_spec,
          "minecraft_3D_maze": minecraft_3d_maze_spec, "binary_holey": binary_holey_spec,
          "minecraft_2D_maze": minecraft_2d_maze_spec, "minecraft_3D_holey_maze": minecraft_3d_holey_maze_spec,
          "minecraft_3D_dungeon_holey": minecraft_3d_dungeon_holey_spec}


def get_spec(problem: str, map_shape) -> ProblemSpec:
    if problem not in _SPECS:
        raise KeyError(f"problem {problem!r} is not supported by control_pcgrl_b200 yet "
                       f"(have: {sorted(_SPECS)})")
    return _SPECS[problem](tuple(int(v) for v in map_shape))


def register_spec(name, fn):
    _SPECS[name] = fn


PROBLEM_NAMES = ["binary", "zelda", "sokoban", "smb", "minecraft_3D_maze", "binary_holey", "minecraft_2D_maze",
                 "minecraft_3D_holey_maze", "minecraft_3D_dungeon_holey"]
HOLEY_PROBLEMS = ("binary_holey", "minecraft_3D_holey_maze", "minecraft_3D_dungeon_holey")
# envs/reps/__init__.py:11-23 (+ the stale 3D spellings)
REPRESENTATION_ALIASES = {"narrow": "narrow", "turtle": "turtle", "wide": "wide", "cellular": "cellular",
                          "narrow3D": "narrow", "turtle3D": "turtle", "wide3D": "wide", "cellular3D": "cellular"}
