"""Single-env gym façade over the batched CUDA env (N = 1) -- the reference's API surface.

Mirrors (names, arguments, return tuples, attributes) of, relative to /root/reference/control_pcgrl/:
  envs/pcgrl_env.py        PcgrlEnv            reset/step/adjust_param/seed/get_border_tile/...
  envs/pcgrl_ctrl_env.py   PcgrlCtrlEnv        set_map, cond_bounds, static_trgs
  wrappers.py              CroppedImagePCGRLWrapper / ActionMapImagePCGRLWrapper / CAactionWrapper
  control_wrappers.py      ControlWrapper, UniformNoiseyTargets
  rl/envs.py               make_env
so that rl/train.py, evo/evolve.py and profile_env.py-style loops can switch to it.  Every number it
returns is computed by the CUDA kernels; this file only moves data and keeps the attribute surface
(`unwrapped._prob.*`, `unwrapped._rep.*`, `_rep_stats`, ...) consumers touch.

For throughput use BatchedPcgrlEnv / PcgrlVectorEnv directly: this façade pays a launch + sync per call.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import spaces
from .batched_env import BatchedPcgrlEnv
from .config import normalise
from .problems import REPRESENTATION_ALIASES


class _ProblemProxy:
    """Attribute surface of envs/probs/problem.py:Problem used by consumers."""

    def __init__(self, owner):
        self._o = owner
        spec = owner._b.spec
        self._tile_types = list(spec.tiles)
        self._height, self._width = owner._b.map_shape[0], owner._b.map_shape[1]
        self._length = owner._b.map_shape[2] if owner._b.ndim == 3 else None
        self.static_trgs = owner._b.metric_trgs          # aliased + mutated, like the reference (A-23)
        self.cond_bounds = owner._b.cond_bounds
        self._reward_weights = dict(spec.reward_weights)
        self._ctrl_reward_weights = dict(spec.reward_weights)
        self._border_tile = spec.border_tile
        self._empty_tile, self._wall_tile = spec.tiles[0], spec.tiles[1]
        self._border_size = (1, 1)
        self._tile_size = 16
        self._prob = dict(zip(spec.tiles, spec.init_probs))
        self.eval_maps = []
        self.path_coords, self.path_length = [], None
        self._hole_queue = []
        self.fixed_holes = getattr(owner._b, "fixed_holes", False)

    # HoleyProblem (envs/probs/holey_prob.py): the current entrance / exit, (y, x) in bordered coordinates
    @property
    def entrance_coords(self):
        return self._o._b.holes[0, :2].cpu().numpy()

    @property
    def exit_coords(self):
        return self._o._b.holes[0, 2:].cpu().numpy()

    def queue_holes(self, holes):
        """Holes to use on the next resets: a list of (entrance, exit) pairs (the reference's _hole_queue,
        holey_prob.py:43-44, without the ray actor that feeds it)."""
        self._hole_queue = list(holes)

    def get_tile_types(self):
        return self._tile_types

    def get_tile_int(self, tile):
        return self._tile_types.index(tile)

    def get_stats(self, level_map):
        """Problem.get_stats: accepts the reference's string map (list of lists of tile names) or an int
        array; runs the CUDA stats kernel on it."""
        arr = np.asarray(level_map)
        if arr.dtype.kind in "US":
            lut = {t: i for i, t in enumerate(self._tile_types)}
            arr = np.vectorize(lut.__getitem__, otypes=[np.int8])(arr)
        if self._o._b.holey:
            # the holey problems are handed the BORDERED map (pcgrl_holey_env.py:52-53) and read
            # self.entrance_coords / self.exit_coords (binary_holey_prob.py:62-63)
            inner = arr[tuple(slice(1, -1) for _ in arr.shape)]
            holes = np.concatenate([np.asarray(self.entrance_coords), np.asarray(self.exit_coords)])[None]
            st = self._o._b.compute_stats(inner[None].astype(np.int8), holes=holes)[0].tolist()
        else:
            st = self._o._b.compute_stats(arr[None].astype(np.int8))[0].tolist()
        out = OrderedDict(zip(self._o._b.stat_names, st))
        if "path-length" in out:
            self.path_length = out["path-length"]
        return out

    def get_debug_info(self, new_stats, old_stats):
        return dict(new_stats)

    def get_episode_over(self, new_stats, old_stats):
        return False

    def get_reward(self, new_stats, old_stats):
        """Legacy Problem.get_reward.  At this commit only BinaryProblem is registered with its own get_reward
        (binary_prob.py:170-178); the Ctrl problems return None (e.g. zelda_ctrl_prob.py:79) and the 3D maze's
        is commented out.  The batched kernels compute the same sum with reward_mode="range"."""
        spec = self._o._b.spec
        if spec.name != "binary":
            return None
        from .problems import get_range_reward
        return sum(get_range_reward(new_stats[k], old_stats[k], *spec.range_bands[k]) * spec.range_weights[k]
                   for k in spec.range_bands)


class _RepProxy:
    """Attribute surface of envs/reps/representation.py used by evolution (evo/evolve.py:1004-1116)."""

    def __init__(self, owner):
        self._o = owner
        self._random_start = True
        self._old_map = None
        self._bordered_map = None
        # multi-agent (envs/reps/wrappers.py:548-651): which agent the next observation is for (None: all of them)
        self._active_agent = None

    @property
    def n_agents(self):
        return self._o._b.n_agents

    @property
    def agent_positions(self):
        b = self._o._b
        return b.agent_pos[:, 0, :b.ndim].cpu().numpy().astype(np.int64)

    def set_active_agent(self, agent_name):
        self._active_agent = agent_name

    def _agent_index(self, agent_name):
        return int(str(agent_name).split("_")[-1])

    @property
    def unwrapped(self):
        return self

    @property
    def _map(self):
        return self._o._b.maps[0].cpu().numpy().astype(np.uint8)

    @_map.setter
    def _map(self, value):
        b = self._o._b
        b.maps[0].copy_(torch.as_tensor(np.asarray(value).astype(np.int8)).to(b.device))

    @property
    def _pos(self):
        return self._o._b.pos[0, :self._o._b.ndim].cpu().numpy()

    @property
    def _x(self):
        return int(self._o._b.pos[0, 1])

    @property
    def _y(self):
        return int(self._o._b.pos[0, 0])

    @property
    def n_step(self):
        return int(self._o._b.n_step[0])

    def _update_bordered_map(self):
        # representation.py:162-164 -- materialised lazily; nothing on the step path reads it
        m = self._map
        b = np.full(tuple(s + 2 for s in m.shape), self._o.get_border_tile(), dtype=int)
        b[tuple(slice(1, -1) for _ in m.shape)] = m
        self._bordered_map = b

    def get_observation(self):
        obs = {"map": self._map.copy()}
        if self._o._b.n_agents > 1:
            # MultiAgentRepresentationWrapper.get_observation (:588-604): one dict per agent, or the active agent's
            pos = self.agent_positions
            names = [f"agent_{i}" for i in range(len(pos))] if self._active_agent is None else [self._active_agent]
            return {k: {"map": obs["map"].copy(), "pos": pos[self._agent_index(k)]} for k in names}
        if self._o._b.representation in ("narrow", "turtle"):
            obs["pos"] = np.array(self._pos)
        if self._o._b.holey:
            # HoleyRepresentation.get_observation (envs/reps/wrappers.py:153-160): the bordered map with the
            # holes dug as empty tiles (:139-142), positions shifted by the border
            self._update_bordered_map()
            ey, ex, xy, xx = self._o._b.holes[0].tolist()
            self._bordered_map[ey, ex] = self._bordered_map[xy, xx] = 0
            obs["map"] = self._bordered_map.astype(np.uint8)
            if "pos" in obs:
                obs["pos"] = obs["pos"] + 1
        return obs

    def update(self, action, **kwargs):
        """Representation.update(action) -> (change, pos), without touching the env counters/stats
        semantics beyond what PcgrlEnv.step would do is not possible on the fused kernel, so this runs
        one fused step and rolls the episode counters back (evolution only reads the map)."""
        o = self._o
        b = o._b
        it, ch = b.iteration.clone(), b.changes.clone()
        o._launch_step(action)
        changed = int(b.changed[0])
        b.iteration.copy_(it)
        b.changes.copy_(ch)
        return changed, self._pos


class PcgrlEnv(spaces.GymEnv):
    """Drop-in for PcgrlEnv / PcgrlCtrlEnv / PcgrlEnv3D (envs/pcgrl_env.py:39-342)."""
    metadata = {"render.modes": []}

    def __init__(self, cfg, prob=None, rep=None, device="cuda:0"):
        c = normalise(cfg)
        # the reference builds the problem / representation from the `prob` / `rep` kwargs of the env id
        # (control_pcgrl/__init__.py:8-37, pcgrl_env.py:44-46); a cfg that names a different pair is a caller bug
        if prob is not None and prob != c.problem:
            raise ValueError(f"env id problem {prob!r} != cfg.task.problem {c.problem!r}")
        if rep is not None and REPRESENTATION_ALIASES[rep] != REPRESENTATION_ALIASES[c.representation]:
            raise ValueError(f"env id representation {rep!r} != cfg.representation {c.representation!r}")
        self.cfg = cfg
        self.render_mode = None
        self._repr_name = REPRESENTATION_ALIASES[rep or c.representation]
        self.map_shape = c.map_shape
        self.obs_window = c.obs_window
        # MultiActionRepresentation (cfg.act_window, envs/reps/wrappers.py:397-545): the batched env picks the
        # patch action kind itself
        action_kind = None if c.act_window is not None else \
            {"narrow": "int32", "turtle": "int32", "wide": "wide_coords", "cellular": "ca_logits"}[self._repr_name]
        self._b = BatchedPcgrlEnv(cfg, 1, device=device, action_kind=action_kind)
        self._prob = _ProblemProxy(self)
        self._rep = _RepProxy(self)
        self._rep_stats = None
        self.metrics = {k: None for k in self._b.static_trgs}
        self._get_stats_on_step = True
        self._has_been_assigned_map = False
        self.switch_env, self.cur_map_idx = False, 0
        self.cond_bounds = self._b.cond_bounds
        self.static_trgs = self._b.metric_trgs
        self.metric_trgs = self._b.metric_trgs
        sp = self._b.spec
        # pcgrl_env.py:82-88; problems without cond_bounds (minecraft_2D_maze) keep their raw weights
        self._reward_weights = {k: (v / (sp.cond_bounds[k][1] - sp.cond_bounds[k][0]) if k in sp.cond_bounds else v)
                                for k, v in sp.reward_weights.items()}
        self._ctrl_reward_weights = dict(self._reward_weights)
        self._np_random = np.random.default_rng()
        self.adjust_param(cfg)

    # -- properties that live on the device
    @property
    def _iteration(self):
        return int(self._b.iteration[0])

    @property
    def _changes(self):
        return int(self._b.changes[0])

    def adjust_param(self, cfg):
        b = self._b
        self._max_changes = b.max_changes
        self._max_iterations = b.max_iterations
        n_tiles, dims = b.n_tiles, b.map_shape
        rep = self._repr_name
        if b.act_window is not None:                                           # wrappers.py:438-443
            self.action_space = spaces.MultiDiscrete([n_tiles] * int(np.prod(b.act_window)))
        elif rep == "narrow":
            self.action_space = spaces.Discrete(n_tiles)                       # narrow_rep.py:65-68
        elif rep == "turtle":
            self.action_space = spaces.Discrete(4 + n_tiles)                   # turtle_rep.py:70-71
        elif rep == "wide":
            self.action_space = spaces.MultiDiscrete([*dims, n_tiles])         # wide_rep.py:23-24
        else:
            self.action_space = spaces.Dict({"map": spaces.Box(0, n_tiles - 1, shape=dims, dtype=np.uint8)})
        # HoleyRepresentation.get_observation_space (envs/reps/wrappers.py:162-174): map two cells larger, pos + 1
        grow = 2 if b.holey else 0
        obs = {"map": spaces.Box(low=0, high=n_tiles - 1, dtype=np.uint8, shape=tuple(d + grow for d in b.obs_window))}
        if rep in ("narrow", "turtle"):                                        # representation.py:223-228
            obs["pos"] = spaces.Box(low=np.zeros(len(dims)) + grow // 2,
                                    high=np.array([d - 1 for d in dims]) + grow // 2, dtype=np.uint8)
        self.observation_space = spaces.Dict(obs)

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        if seed is not None:
            self._b.seed = int(seed)
        return [seed]

    def get_map_dims(self):
        return tuple(self.map_shape) + (self.get_num_tiles(),)

    def get_num_tiles(self):
        return self._b.n_tiles

    def get_border_tile(self):
        return self._b.spec.tiles.index(self._b.spec.border_tile)

    def get_map(self):
        return self._rep._map

    def get_rep(self):
        return self._rep

    def set_map(self, init_map):                                               # pcgrl_ctrl_env.py:12-14
        self._rep._random_start = False
        self._rep._old_map = np.array(init_map).copy()

    def sample_tasks(self, n_maps):
        return [int(self._np_random.integers(max(len(self._prob.eval_maps), 1)))]

    def get_task(self):
        return self.cur_map_idx

    def set_task(self, map_idx):
        if map_idx is not None:
            self.cur_map_idx, self.switch_env = map_idx, True

    def _raw_obs(self):
        return self._rep.get_observation()

    def _stats_dict(self):
        return self._b.stats_dict(0)

    def reset(self, *, seed=None, options=None):
        grids = None
        if not self._rep._random_start and self._rep._old_map is not None:
            grids = self._rep._old_map[None]
        holes = None
        if self._b.holey and self._prob._hole_queue:                             # holey_prob.py:43-44
            (ent, ext), self._prob._hole_queue = self._prob._hole_queue[0], self._prob._hole_queue[1:]
            holes = np.concatenate([np.asarray(ent), np.asarray(ext)])[None]
        self._b.reset(grids=grids, holes=holes)
        self._rep_stats = self._stats_dict() if self._get_stats_on_step else None
        self.metrics = self._rep_stats
        return self._raw_obs(), {}

    def _launch_step(self, action):
        b = self._b
        rep = self._repr_name
        if b.n_agents > 1:
            # MultiAgentTurtleRepresentation.update (:631-647) takes {agent name: action}; MultiAgentWrapper hands it
            # one agent per env step (wrappers.py:724-731), which is what one kernel step does
            if not isinstance(action, dict) or len(action) != 1:
                raise ValueError("a multi-agent env steps one agent at a time: action = {'agent_i': a} "
                                 "(control_pcgrl_b200.MultiAgentWrapper does the round)")
            (name, act), = action.items()
            a = torch.tensor([int(np.asarray(act).reshape(-1)[0])], dtype=torch.int32, device=b.device)
            b.step(a, agent=self._rep._agent_index(name))
            return
        if b.act_window is not None:
            a = torch.tensor(np.asarray(action, dtype=np.int32).reshape(1, -1), device=b.device)
        elif rep in ("narrow", "turtle"):
            a = torch.tensor([int(np.asarray(action).reshape(-1)[0])], dtype=torch.int32, device=b.device)
        elif rep == "wide":
            a = torch.tensor(np.asarray(action, dtype=np.int32).reshape(1, -1), device=b.device)
        else:
            a = torch.tensor(np.asarray(action, dtype=np.float32).reshape(1, -1), device=b.device)
        b.step(a)

    def step(self, action):
        self._launch_step(action)
        b = self._b
        changed = bool(b.changed[0])
        done = bool(b.done[0])
        info = {}
        if changed:
            self._rep_stats = self._stats_dict()
            self.metrics = self._rep_stats
            info.update(self._rep_stats)
        info["iterations"] = self._iteration
        info["changes"] = self._changes
        info["max_iterations"] = self._max_iterations
        info["max_changes"] = self._max_changes
        self._last_reward = float(b.reward[0])
        return self._raw_obs(), None, done, done, info               # PcgrlEnv's own reward is None (:302)

    def render(self, *a, **k):
        raise NotImplementedError("rendering is out of scope for control_pcgrl_b200 (SURVEY.md row 19)")

    # PcgrlHoleyEnv (envs/pcgrl_holey_env.py:47-53)
    def get_empty_tile(self):
        return 0

    def _get_rep_map(self):
        if not self._b.holey:
            return self._rep._map
        return self._rep.get_observation()["map"]


PcgrlCtrlEnv = PcgrlEnv
PcgrlEnv3D = PcgrlEnv
PcgrlHoleyEnv = PcgrlEnv


class _ObsWrapper(spaces.GymWrapper):
    """Shared body of the three composite obs/action wrappers of control_pcgrl/wrappers.py."""

    def __init__(self, game, cfg, device="cuda:0"):
        from .registry import make
        env = make(game, cfg=cfg, device=device) if isinstance(game, str) else game
        super().__init__(env)
        b = env.unwrapped._b
        self._crop = b.representation in ("narrow", "turtle")
        # what _obs() returns: the batched observation (holey problems: window + 2 for the border frame; frozen
        # tiles: the extra static_builds plane) without ControlWrapper's target planes, which that wrapper adds
        shp = b.obs_shape()
        self.observation_space = spaces.Box(low=0, high=1, shape=(*shp[:-1], shp[-1] - 2 * len(b.ctrl_metrics)),
                                            dtype=np.float64)
        self.action_space = env.action_space

    def _obs(self):
        b = self.env.unwrapped._b
        if b.n_agents > 1:      # the same crop per agent, around its own position
            rp = self.env.unwrapped._rep
            names = [f"agent_{i}" for i in range(b.n_agents)] if rp._active_agent is None else [rp._active_agent]
            return {k: b.observe(dtype=torch.float64, agent=rp._agent_index(k))[0][..., 2 * len(b.ctrl_metrics):]
                    .cpu().numpy() for k in names}
        full = b.observe(dtype=torch.float64)[0]
        return full[..., 2 * len(b.ctrl_metrics):].cpu().numpy()

    def reset(self, *, seed=None, options=None):
        _, info = self.env.reset()
        return self._obs(), info

    def step(self, action, **kw):
        _, r, done, trunc, info = self.env.step(self._map_action(action))
        return self._obs(), r, done, trunc, info

    def _map_action(self, action):
        return action


class CroppedImagePCGRLWrapper(_ObsWrapper):
    """wrappers.py:443-476: Cropped -> OneHotEncoding -> ToImage (narrow / turtle)."""


class ActionMapImagePCGRLWrapper(_ObsWrapper):
    """wrappers.py:502-526: ActionMap -> OneHotEncoding -> ToImage (wide)."""

    def __init__(self, game, cfg, device="cuda:0"):
        super().__init__(game, cfg, device)
        b = self.env.unwrapped._b
        self.h, self.w = b.obs_window[0], b.obs_window[1]           # from the observation space (A-7)
        self.dim = b.n_tiles
        self.action_space = spaces.Discrete(self.h * self.w * self.dim)

    def _map_action(self, action):
        y, x, v = np.unravel_index(int(action), (self.h, self.w, self.dim))
        return [x, y, v]                                            # wrappers.py:320: writes _map[x, y]


class CAactionWrapper(_ObsWrapper):
    """wrappers.py:529-544, as intended (the upstream class is broken, SURVEY A-9): flat Box(0,1,(C*W*H,))
    reshaped channel-major to (C, W, H), argmax over channels."""

    def __init__(self, game, cfg, device="cuda:0"):
        super().__init__(game, cfg, device)
        b = self.env.unwrapped._b
        self.action_space = spaces.Box(0, 1, shape=(b.n_tiles * b.cells,), dtype=np.float32)


class ControlWrapper(spaces.GymWrapper):
    """control_wrappers.py:26-362.  Reward = loss_t - loss_{t-1}; optional target/metric channels."""

    def __init__(self, env, cfg, ctrl_metrics=None, rand_params=False):
        super().__init__(env)
        u = env.unwrapped
        b = u._b
        want = list(ctrl_metrics) if ctrl_metrics is not None else []
        if want != b.ctrl_metrics:
            raise ValueError(f"ctrl_metrics {want} must equal cfg.controls {b.ctrl_metrics} (the kernels read "
                             "per-env targets for exactly the controlled metrics)")
        self.controllable = ctrl_metrics is not None
        self.ctrl_metrics = want
        self.n_ctrl_metrics = len(want)
        self.ctrl_loss_metrics = want
        self.metric_weights = b.metric_weights
        self.static_metric_names = set(k for k in b.static_trgs if k not in want)
        self.cond_bounds = b.cond_bounds
        self.param_ranges = b.param_ranges
        self.metric_trgs = b.metric_trgs
        self.static_trgs = b.metric_trgs
        self.all_metrics = set(b.all_metrics)
        self.n_metrics = len(self.all_metrics)
        self.metrics = u.metrics
        self.last_metrics = None
        self.auto_reset = True
        self._ctrl_trg_queue = []
        self.last_loss = None
        self.infer = getattr(cfg, "infer", False)
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        if self.controllable:
            shp = env.observation_space.shape
            self.n_new_obs = 2 * len(want)
            self.metrics_shape = (*shp[:-1], self.n_new_obs)
            self.observation_space = spaces.Box(low=0, high=1, shape=(*shp[:-1], shp[-1] + self.n_new_obs),
                                                dtype=np.float64)
        self.max_loss = self.get_max_loss(ctrl_metrics=want)

    # -- targets
    def set_trgs(self, trgs):
        self._ctrl_trg_queue = [trgs]

    def do_set_trgs(self, trgs):
        self.unwrapped._b.set_trgs(trgs)

    def get_control_bounds(self):
        return {k: self.cond_bounds[k] for k in self.ctrl_metrics}

    def get_control_vals(self):
        return {k: self.metrics[k] for k in self.ctrl_metrics}

    def get_metric_vals(self):
        return self.metrics

    def get_cond_trgs(self):
        return self.metric_trgs

    def get_cond_bounds(self):
        return self.cond_bounds

    # -- losses (host-side evaluation of the same formula the kernel uses, for API parity)
    def _loss_over(self, names):
        if self.metrics is None:
            return 0
        loss = 0
        for m in names:
            trg = self.metric_trgs[m]
            val = self.metrics[m]
            if isinstance(trg, tuple):
                lm = -abs(np.arange(*trg) - val).min()
            else:
                lm = -abs(trg - val)
            loss += lm * self.metric_weights[m]
        return loss

    def get_loss(self):
        return self._loss_over(self.unwrapped._b.all_metrics)

    def get_ctrl_loss(self):
        return self._loss_over(self.ctrl_loss_metrics)

    def get_max_loss(self, ctrl_metrics=()):
        tot = 0
        for k, v in self.static_trgs.items():
            if k in ctrl_metrics:
                continue
            lo, hi = self.cond_bounds[k]
            if isinstance(v, tuple):
                tot += max(abs(v[0] - lo), abs(v[1] - hi)) * self.metric_weights[k]
            else:
                tot += max(abs(v - lo), abs(v - hi)) * self.metric_weights[k]
        return tot

    def _full_obs(self, ob):
        if not self.controllable:
            return ob
        b = self.unwrapped._b
        if b.n_agents > 1:
            rp = self.unwrapped._rep
            return {k: b.observe(dtype=torch.float64, agent=rp._agent_index(k))[0].cpu().numpy() for k in ob}
        return b.observe(dtype=torch.float64)[0].cpu().numpy()

    def reset(self, *, seed=None, options=None):
        if self._ctrl_trg_queue:
            self.do_set_trgs(self._ctrl_trg_queue.pop(0))
        ob, info = self.env.reset()
        self.metrics = self.unwrapped._rep_stats
        self.last_metrics = dict(self.metrics) if self.metrics else None
        self.last_loss = self.get_loss()
        self.n_step = 0
        return self._full_obs(ob), info

    def step(self, action, **kw):
        ob, _, done, trunc, info = self.env.step(action, **kw)
        u = self.unwrapped
        self.metrics = u._rep_stats
        reward = u._last_reward                      # computed in fp64 inside the step kernel
        self.last_loss = self.get_loss()
        self.last_metrics = dict(self.metrics) if self.metrics else None
        self.n_step += 1
        if not self.auto_reset:
            done = trunc = False
        return self._full_obs(ob), reward, done, trunc, info


class UniformNoiseyTargets(spaces.GymWrapper):
    """control_wrappers.py:442-471, intended behaviour (SURVEY A-25): new uniform targets every reset."""

    def __init__(self, env, cfg=None):
        super().__init__(env)
        self.cond_bounds = env.unwrapped.cond_bounds
        self._rng = np.random.default_rng()

    def set_rand_trgs(self):
        trgs = {}
        for k in self.env.ctrl_metrics:
            lb, ub = self.cond_bounds[k]
            trgs[k] = float(self._rng.random() * (ub - lb) + lb)
        self.env.set_trgs(trgs)

    def reset(self, *, seed=None, options=None):
        self.set_rand_trgs()
        return self.env.reset()

    def step(self, action, **kw):
        return self.env.step(action, **kw)


class MultiAgentWrapper(spaces.GymWrapper):
    """wrappers.py:697-736: Dict observation / action spaces keyed 'agent_i'; a step takes {agent: action} and steps
    the env once per agent in dict order (each agent's observation is the one right after its own sub-step);
    done['__all__'] / truncated['__all__'] = every agent of the round saw done."""

    def __init__(self, game, cfg=None):
        super().__init__(game)
        self.n_agents = game.unwrapped._b.n_agents
        self.observation_space = spaces.Dict({f"agent_{i}": game.observation_space for i in range(self.n_agents)})
        self.action_space = spaces.Dict({f"agent_{i}": game.action_space for i in range(self.n_agents)})

    def reset(self, *, seed=None, options=None):
        return self.env.reset()

    def step(self, action):
        obs, rew, done, truncated, info = {}, {}, {}, {}, {}
        for k, v in action.items():
            self.unwrapped._rep.set_active_agent(k)
            obs_k, rew[k], done[k], truncated[k], info[k] = self.env.step({k: v})
            obs.update(obs_k)
        truncated["__all__"] = bool(np.all(list(truncated.values())))
        done["__all__"] = bool(np.all(list(done.values())))
        return obs, rew, done, truncated, info


def make_env(cfg, device="cuda:0"):
    """rl/envs.py:28-81 make_env: pick the wrapper stack by representation, then ControlWrapper."""
    if isinstance(cfg, dict) and "task" in cfg and not hasattr(cfg, "task"):
        from types import SimpleNamespace
        cfg = SimpleNamespace(**cfg)
    c = normalise(cfg)
    rep = REPRESENTATION_ALIASES[c.representation]
    name = c.env_name
    if rep == "wide":
        env = ActionMapImagePCGRLWrapper(name, cfg=cfg, device=device)
    elif rep == "cellular":
        env = CAactionWrapper(name, cfg=cfg, device=device)
    else:
        env = CroppedImagePCGRLWrapper(name, cfg=cfg, device=device)
    env = ControlWrapper(env, ctrl_metrics=c.controls, cfg=cfg)
    if c.controls and not getattr(cfg, "evaluate", False):
        env = UniformNoiseyTargets(env, cfg)
    if c.n_agents != 0:                                                          # rl/envs.py:74-75
        env = MultiAgentWrapper(env, cfg)
    return env
