"""Plain-dataclass stand-in for the reference's hydra Config (control_pcgrl/configs/config.py:254-320).

Same field names as the reference, so a hydra/omegaconf cfg, an RLlib EnvContext dict or this dataclass can
all be handed to `make` / `make_env` / `BatchedPcgrlEnv`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Optional


@dataclass
class MultiagentConfig:
    n_agents: int = 0


@dataclass
class TaskConfig:
    problem: str = "binary"
    name: str = "binary"
    map_shape: tuple = (16, 16)
    obs_window: tuple = (32, 32)
    weights: dict = field(default_factory=dict)
    controls: Optional[list] = None


@dataclass
class Config:
    task: TaskConfig = field(default_factory=TaskConfig)
    representation: str = "narrow"
    max_board_scans: float = 3.0            # configs/config.py:290
    change_percentage: Optional[float] = None
    controls: Optional[list] = None
    render_mode: Optional[str] = None
    render: bool = False
    infer: bool = False
    evaluate: bool = False
    evaluation_env: bool = False
    act_window: Optional[list] = None
    static_tile_wrapper: bool = False
    static_prob: Optional[float] = None
    n_static_walls: Optional[int] = None
    show_agents: bool = False
    multiagent: MultiagentConfig = field(default_factory=MultiagentConfig)
    n_aux_tiles: int = 0
    train_reward_model: bool = False
    fixed_holes: bool = False               # holey problems (binary_holey_prob.py:47-49)
    env_name: str = ""

    def __post_init__(self):
        if not self.env_name:
            self.env_name = f"{self.task.problem}-{self.representation}-v0"


# configs/task/*.yaml defaults for the BASELINE.json problems
TASK_DEFAULTS = {
    "binary": dict(map_shape=(16, 16), obs_window=(32, 32), weights={"regions": 1, "path-length": 1}),
    "zelda": dict(map_shape=(7, 11), obs_window=(22, 22),
                  weights={"player": 3, "key": 3, "door": 3, "regions": 5, "enemies": 1, "nearest-enemy": 2,
                           "path-length": 1}),
    # no configs/task/*.yaml upstream for these: the structured configs of configs/config.py:89-165
    # (SokobanConfig, SMBConfig, MinecraftMazeConfig); sokoban's 5x5 is the Problem constructor's size
    "sokoban": dict(map_shape=(5, 5), obs_window=(10, 10),
                    weights={"player": 3, "crate": 2, "target": 2, "regions": 5, "ratio": 2, "dist-win": 0,
                             "sol-length": 1}),
    "smb": dict(map_shape=(116, 16), obs_window=(32, 32),
                weights={"dist-floor": 2, "disjoint-tubes": 1, "enemies": 1, "empty": 1, "noise": 4, "jumps": 2,
                         "jumps-dist": 2, "dist-win": 5, "sol-length": 1}),
    "binary_holey": dict(map_shape=(16, 16), obs_window=(32, 32),
                         weights={"regions": 1, "path-length": 0, "connected-path-length": 1}),
    # Minecraft2DmazeProblem's own size (minecraft_2D_maze_prob.py:17-18) and reward weights (:24-27)
    "minecraft_2D_maze": dict(map_shape=(14, 14), obs_window=(28, 28), weights={"regions": 5, "path-length": 1}),
    # the 3D holey problems (weights: minecraft_3D_holey_maze_prob.py:33-41, minecraft_3D_holey_dungeon_prob.py:75-82);
    # 7^3 is the size the reference's 3D experiments use and the bordered map must fit 16^3
    "minecraft_3D_holey_maze": dict(map_shape=(7, 7, 7), obs_window=(14, 14, 14),
                                    weights={"regions": 0, "path-length": 100, "connected-path-length": 120,
                                             "n_jump": 150}),
    "minecraft_3D_dungeon_holey": dict(map_shape=(7, 7, 7), obs_window=(14, 14, 14),
                                       weights={"regions": 0, "path-length": 100, "chests": 300, "n_jump": 100,
                                                "enemies": 100, "nearest-enemy": 200}),
    "minecraft_3D_maze": dict(map_shape=(15, 15, 15), obs_window=(30, 30, 30),
                              weights={"path-length": 100, "n_jump": 100, "regions": 0}),
}


def make_config(problem="binary", representation="narrow", map_shape=None, obs_window=None, weights=None,
                controls=None, max_board_scans=3.0, change_percentage=None, **extra) -> Config:
    d = TASK_DEFAULTS.get(problem, {})
    task = TaskConfig(problem=problem, name=problem,
                      map_shape=tuple(map_shape or d.get("map_shape", (16, 16))),
                      obs_window=tuple(obs_window or d.get("obs_window", map_shape or (16, 16))),
                      weights=dict(weights if weights is not None else d.get("weights", {})),
                      controls=controls)
    return Config(task=task, representation=representation, max_board_scans=max_board_scans,
                  change_percentage=change_percentage, controls=controls, **extra)


def _get(obj, name, default=None):
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


def normalise(cfg) -> SimpleNamespace:
    """Read the fields the step path needs from any cfg flavour (dataclass / namespace / dict)."""
    task = _get(cfg, "task")
    problem = _get(task, "problem")
    rep = _get(cfg, "representation")
    map_shape = tuple(int(v) for v in _get(task, "map_shape"))
    obs_window = _get(task, "obs_window")
    obs_window = tuple(int(v) for v in obs_window) if obs_window is not None else map_shape
    weights = dict(_get(task, "weights") or {})
    controls = _get(cfg, "controls")
    controls = list(controls) if controls else None
    return SimpleNamespace(problem=problem, representation=rep, map_shape=map_shape, obs_window=obs_window,
                           weights=weights, controls=controls,
                           max_board_scans=_get(cfg, "max_board_scans", 3.0),
                           change_percentage=_get(cfg, "change_percentage", None),
                           env_name=_get(cfg, "env_name", f"{problem}-{rep}-v0"),
                           # representation wrappers (envs/reps/wrappers.py wrap_rep :717-722)
                           act_window=(lambda a: None if a is None else tuple(int(v) for v in a))(
                               _get(cfg, "act_window", None)),
                           static_tile_wrapper=bool(_get(cfg, "static_tile_wrapper", False)),
                           static_prob=_get(cfg, "static_prob", None),
                           n_static_walls=_get(cfg, "n_static_walls", None),
                           # MultiAgentWrapper / MultiAgentTurtleRepresentation (cfg.multiagent.n_agents, 0 = off)
                           n_agents=int(_get(_get(cfg, "multiagent", None), "n_agents", 0) or 0),
                           show_agents=bool(_get(cfg, "show_agents", False)),
                           # HoleyProblem.adjust_param (binary_holey_prob.py:47-49)
                           fixed_holes=bool(_get(cfg, "fixed_holes", _get(task, "fixed_holes", False))))
