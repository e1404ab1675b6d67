"""Problem tables: tile codes, stat order, static targets / bounds / default weights.

Values are re-derived here from the reference's Problem constructors (paths relative to
/root/reference/control_pcgrl/envs/probs/); tests pin them against the real classes.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field


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
        reward_weights={"regions": 100, "path-length": 100})


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
                        "path-length": 1})


_SPECS = {"binary": binary_spec, "zelda": zelda_spec}


def get_spec(problem: str, map_shape) -> ProblemSpec:
    if problem not in _SPECS:
        raise KeyError(f"problem {problem!r} is not supported by control_pcgrl_b200 yet "
                       f"(have: {sorted(_SPECS)})")
    return _SPECS[problem](tuple(int(v) for v in map_shape))


def register_spec(name, fn):
    _SPECS[name] = fn


PROBLEM_NAMES = ["binary", "zelda", "sokoban", "smb", "minecraft_3D_maze"]
# envs/reps/__init__.py:11-23 (+ the stale 3D spellings)
REPRESENTATION_ALIASES = {"narrow": "narrow", "turtle": "turtle", "wide": "wide", "cellular": "cellular",
                          "narrow3D": "narrow", "turtle3D": "turtle", "wide3D": "wide", "cellular3D": "cellular"}
