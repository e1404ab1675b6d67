"""TEST INFRASTRUCTURE ONLY -- loader that runs the *real* reference from /root/reference.

The reference tree is /root/reference in the build container; on the GPU box (no /root/reference) it is the
unmodified pip install under baseline/_ref that oracle/install_reference.sh makes, when present.
It is used by oracle/gen_golden.py to produce the committed fixtures under tests/golden/
and by the `needs_reference` tests that pin oracle/pcgrl_oracle.py against the reference.
Nothing under control_pcgrl_b200/ may import it.

The reference (smearle/control-pcgrl @ 8bde536) cannot be imported unmodified here:
gymnasium / ray / hydra / matplotlib are absent, `control_pcgrl/envs/probs/__init__.py:18-19`
writes a file at import time, and `reps/narrow_rep.py:93`, `reps/wide_rep.py:37` use an
idiom numpy 2 rejects.  The recipe below installs fake third-party
modules and stub *packages* (so the reference's own `__init__.py` files never run) and then
imports the reference's env, representation, problem and wrapper modules verbatim, in place.
No reference source is copied.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """First tree that holds the reference's env package: $PCGRL_REFERENCE_ROOT, the read-only mount of the build
    container, or the unmodified copy oracle/install_reference.sh pip-installs into baseline/_ref (git-ignored, but
    it travels to the GPU box with the repo snapshot, so the CPU arm there runs the reference's own code)."""
    cands = [os.environ.get("PCGRL_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "control_pcgrl", "envs")):
            return c
    return cands[1]


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "control_pcgrl", "envs"))


# --------------------------------------------------------------------------- fake gymnasium
class _Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(int(s) for s in shape)
        self.dtype = dtype


class _Discrete(_Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self):
        return int(np.random.randint(self.n))


class _MultiDiscrete(_Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        super().__init__(self.nvec.shape, np.int64)

    def sample(self):
        return np.array([np.random.randint(n) for n in self.nvec])


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        super().__init__(shape, dtype)
        # the reference's wrappers compute `high.max() - low.min() + 1` and feed it to np.eye
        # (control_pcgrl/wrappers.py:204-208,248), so low/high must carry the integer dtype.
        self.low = np.broadcast_to(np.asarray(low), self.shape).astype(dtype)
        self.high = np.broadcast_to(np.asarray(high), self.shape).astype(dtype)

    def sample(self):
        return (np.random.random(self.shape) * (self.high - self.low) + self.low).astype(self.dtype)


class _DictSpace(_Space):
    def __init__(self, spaces=None):
        super().__init__(None, None)
        self.spaces = dict(spaces or {})

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()


class _Env:
    metadata: dict = {}

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError


class _Wrapper(_Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, *, seed=None, options=None):
        return self.env.reset()

    def step(self, action, **kw):
        return self.env.step(action, **kw)

    def render(self, *a, **kw):
        return self.env.render(*a, **kw)


_INSTALLED = False


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, rel):
    m = _mod(name)
    m.__path__ = [os.path.join(REF_ROOT, rel)]
    return m


def install():
    """Put the fake third-party modules and the stub packages into sys.modules (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")

    # ---- gymnasium
    spaces = _mod("gymnasium.spaces", Discrete=_Discrete, MultiDiscrete=_MultiDiscrete, Box=_Box,
                  Dict=_DictSpace, Space=_Space)
    seeding = _mod("gymnasium.utils.seeding",
                   np_random=lambda seed=None: (np.random.default_rng(seed), seed))
    gutils = _mod("gymnasium.utils", seeding=seeding)
    env_checker = _mod("gymnasium.wrappers.env_checker", PassiveEnvChecker=type("PassiveEnvChecker", (_Wrapper,), {}))
    gwrappers = _mod("gymnasium.wrappers", env_checker=env_checker)
    registration = _mod("gymnasium.envs.registration", register=lambda **kw: None)
    genvs = _mod("gymnasium.envs", registration=registration)
    _mod("gymnasium", Env=_Env, Wrapper=_Wrapper, spaces=spaces, utils=gutils, wrappers=gwrappers,
         envs=genvs, make=make)

    # ---- ray
    class TaskSettableEnv(_Env):
        pass

    _mod("ray.rllib.env.apis.task_settable_env", TaskSettableEnv=TaskSettableEnv, TaskType=object)
    _mod("ray.rllib.env.apis")
    _mod("ray.rllib.env.env_context", EnvContext=dict)
    _mod("ray.rllib.env")
    _mod("ray.rllib.utils.annotations", override=lambda cls: (lambda f: f))
    _mod("ray.rllib.utils")
    _mod("ray.rllib", MultiAgentEnv=type("MultiAgentEnv", (_Env,), {}))
    _mod("ray")

    # ---- matplotlib (helper_3D.py:4)
    _mod("matplotlib.pyplot")
    _mod("matplotlib")

    # ---- stub packages: the reference's own __init__.py files are skipped
    _pkg("control_pcgrl", "control_pcgrl")
    _pkg("control_pcgrl.envs", "control_pcgrl/envs")
    probs = _pkg("control_pcgrl.envs.probs", "control_pcgrl/envs/probs")
    _pkg("control_pcgrl.configs", "control_pcgrl/configs")
    for game in ("binary", "zelda", "sokoban", "smb", "minecraft"):
        _pkg(f"control_pcgrl.envs.probs.{game}", f"control_pcgrl/envs/probs/{game}")
    _mod("control_pcgrl.configs.config", Config=object)
    noop = lambda *a, **k: None
    _mod("control_pcgrl.envs.probs.minecraft.mc_render",
         **{n: noop for n in ("erase_3D_path", "init_player_view", "spawn_3D_maze", "spawn_3D_border",
                              "spawn_3D_path", "get_3D_maze_blocks", "get_3D_path_blocks",
                              "get_erased_3D_path_blocks", "render_blocks", "spawn_base", "set_player_view",
                              "edit_3D_maze", "edit_bordered_3D_maze", "spawn_3D_bordered_map")})
    _mod("control_pcgrl.envs.probs.minecraft.utils", patch_grpc_evocraft_imports=noop)

    from control_pcgrl.envs.probs.binary.binary_prob import BinaryProblem
    from control_pcgrl.envs.probs.zelda.zelda_ctrl_prob import ZeldaCtrlProblem
    from control_pcgrl.envs.probs.sokoban.sokoban_ctrl_prob import SokobanCtrlProblem
    from control_pcgrl.envs.probs.smb.smb_ctrl_prob import SMBCtrlProblem
    from control_pcgrl.envs.probs.minecraft.minecraft_3D_maze_prob import Minecraft3DmazeProblem

    # control_pcgrl/envs/probs/__init__.py:31-58 -- the five problems BASELINE.json names
    probs.PROBLEMS = {
        "binary": BinaryProblem,
        "zelda": ZeldaCtrlProblem,
        "sokoban": SokobanCtrlProblem,
        "smb": SMBCtrlProblem,
        "minecraft_3D_maze": Minecraft3DmazeProblem,
    }

    # numpy-2 fix: `[0,1][np.bool_]` raises TypeError (narrow_rep.py:93, wide_rep.py:37).
    from control_pcgrl.envs.reps.narrow_rep import NarrowRepresentation
    from control_pcgrl.envs.reps.wide_rep import WideRepresentation
    from control_pcgrl.envs.reps.representation import Representation

    def narrow_update(self, action, **kwargs):
        change = int(self._map[tuple(self._pos)] != action)
        self._map[tuple(self._pos)] = action
        self._pos = self._act_coords[self.n_step % len(self._act_coords)]
        self._positions = [self._pos]
        self.n_step += 1
        Representation.update(self, action)
        return change, self._pos

    def wide_update(self, action):
        self._pos = action[:-1]
        change = int(self._map[tuple(action[:-1])] != action[-1])
        self._map[tuple(action[:-1])] = action[-1]
        Representation.update(self, action)
        return change, action[:-1]

    NarrowRepresentation.update = narrow_update
    WideRepresentation.update = wide_update
    _INSTALLED = True


def make(env_id, cfg=None, **kw):
    """control_pcgrl/__init__.py:8-37 entry-point choice, without the gym registry."""
    install()
    from control_pcgrl.envs.probs import PROBLEMS
    from control_pcgrl.envs.probs.problem import Problem3D
    prob, rep, _ = env_id.rsplit("-", 2)
    if issubclass(PROBLEMS[prob], Problem3D):
        from control_pcgrl.envs.pcgrl_env_3D import PcgrlEnv3D
        return PcgrlEnv3D(cfg=cfg, prob=prob, rep=rep)
    from control_pcgrl.envs.pcgrl_ctrl_env import PcgrlCtrlEnv
    return PcgrlCtrlEnv(cfg=cfg, prob=prob, rep=rep)


def make_cfg(problem, representation, map_shape, obs_window=None, weights=None, controls=None,
             max_board_scans=3, change_percentage=None):
    """Every cfg attribute the live path reads."""
    map_shape = tuple(int(s) for s in map_shape)
    obs_window = tuple(int(s) for s in (obs_window or map_shape))
    task = SimpleNamespace(name=problem, problem=problem, map_shape=map_shape, obs_window=obs_window,
                           weights=dict(weights or {}), controls=controls)
    return SimpleNamespace(
        task=task, representation=representation, render_mode=None, render=False, infer=False,
        evaluate=False, evaluation_env=False, change_percentage=change_percentage,
        max_board_scans=max_board_scans, act_window=None, static_tile_wrapper=False, static_prob=None,
        n_static_walls=None, show_agents=False, multiagent=SimpleNamespace(n_agents=0), n_aux_tiles=0,
        controls=controls, env_name=f"{problem}-{representation}-v0", train_reward_model=False)


def make_raw_env(cfg):
    return make(cfg.env_name, cfg=cfg)


def make_wrapped_env(cfg, raw_only=False):
    """rl/envs.py:28-66 wrapper stack (without the broken UniformNoiseyTargets, SURVEY A-25)."""
    install()
    from control_pcgrl import wrappers, control_wrappers
    rep = cfg.representation
    if raw_only:
        env = make(cfg.env_name, cfg=cfg)
    elif rep == "wide":
        env = wrappers.ActionMapImagePCGRLWrapper(cfg.env_name, cfg=cfg)
    elif rep in ("narrow", "turtle"):
        env = wrappers.CroppedImagePCGRLWrapper(game=cfg.env_name, cfg=cfg)
    else:
        raise ValueError(f"reference RL stack for {rep!r} is broken upstream (SURVEY A-9); use raw_only")
    return control_wrappers.ControlWrapper(env, ctrl_metrics=cfg.controls, cfg=cfg)


def inject_map(env, grid):
    """Make the next reset() start from `grid` (SURVEY A-3a: works for wrapped reps too)."""
    rep = env.unwrapped._rep.unwrapped
    rep._old_map = np.array(grid).copy()
    rep._random_start = False


def load_helpers():
    """The pure stat helpers / engines, imported in place."""
    install()
    from control_pcgrl.envs import helper, helper_3D
    from control_pcgrl.envs.probs.sokoban.sokoban import engine as sok
    from control_pcgrl.envs.probs.smb.smb import engine as smb
    return SimpleNamespace(h2=helper, h3=helper_3D, sokoban=sok, smb=smb)
