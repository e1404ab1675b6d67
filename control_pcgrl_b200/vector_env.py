"""PcgrlVectorEnv: the batched seam for RL trainers.

A gymnasium.vector-style API over BatchedPcgrlEnv: one object owns N GPU envs, `step(actions)` returns
batched tensors that stay on the device (the policy runs there too), with auto-reset.  It replaces the
reference's N Python envs per Ray rollout worker (rl/utils.py:402-415) with one launch per step.
"""
from __future__ import annotations

import numpy as np
import torch

from . import spaces
from .batched_env import BatchedPcgrlEnv


class PcgrlVectorEnv:
    def __init__(self, cfg, num_envs: int, device="cuda:0", obs_dtype=torch.float32, env_offset=0, seed=0,
                 uniform_targets: bool | None = None):
        self.env = BatchedPcgrlEnv(cfg, num_envs, device=device, env_offset=env_offset, seed=seed, auto_reset=False)
        self.num_envs = num_envs
        self.obs_dtype = obs_dtype
        b = self.env
        if b.ctrl_metrics and obs_dtype == torch.uint8:
            raise ValueError("target channels are fractional: use a float obs_dtype with controls")
        self.uniform_targets = bool(b.ctrl_metrics) if uniform_targets is None else uniform_targets
        shp = b.obs_shape()
        self.single_observation_space = spaces.Box(0, 1, shape=shp, dtype=np.float32)
        rep = b.representation
        if rep == "narrow":
            self.single_action_space = spaces.Discrete(b.n_tiles)
        elif rep == "turtle":
            self.single_action_space = spaces.Discrete(4 + b.n_tiles)
        elif rep == "wide":
            self.single_action_space = spaces.Discrete(b.obs_window[0] * b.obs_window[1] * b.n_tiles)
        else:
            self.single_action_space = spaces.Box(0, 1, shape=(b.n_tiles * b.cells,), dtype=np.float32)
        self._obs = torch.empty((num_envs, *shp), dtype=obs_dtype, device=b.device)
        self.episode_return = torch.zeros(num_envs, dtype=torch.float64, device=b.device)
        self.episode_length = torch.zeros(num_envs, dtype=torch.int32, device=b.device)

    def reset(self, grids=None, pos=None):
        if self.uniform_targets:
            self.env.sample_uniform_targets()
        self.env.reset(grids=grids, pos=pos)
        self.episode_return.zero_()
        self.episode_length.zero_()
        return self.env.observe(out=self._obs), {}

    def step(self, actions: torch.Tensor):
        """-> obs, reward, terminated(False), truncated(done), info; finished envs are reset in place and
        their `obs` row is the first observation of the new episode (gymnasium autoreset semantics);
        info carries the final stats of finished episodes.  reward / done are the env's own output tensors: a
        reset leaves them alone (include/pcgrl_b200.h, pcgrl_reset), so they need no copies."""
        b = self.env
        reward, done = b.step(actions)
        self.episode_return += reward.double()
        self.episode_length += 1
        info = {}
        # lock-step episodes (no change budget, every env reset together): the episode end is known on the host
        # without a device-to-host sync, and STAYS known afterwards because the reset below is a full one
        lock_step = b.max_changes is None and b._synced_steps is not None
        any_done = (b._synced_steps > b.max_iterations) if lock_step else bool(done.any())
        if any_done:
            d = done.bool()
            info = {"final_stats": b.stats.clone(), "final_return": self.episode_return.clone(),
                    "final_length": self.episode_length.clone(), "_final": d}
            if lock_step:
                if self.uniform_targets:
                    b.sample_uniform_targets()
                b.reset()
                self.episode_return.zero_()
                self.episode_length.zero_()
            else:
                if self.uniform_targets:
                    keep = b.targets.clone()
                    b.sample_uniform_targets()
                    b.targets[~d] = keep[~d]
                b.reset(mask=done)
                self.episode_return[d] = 0
                self.episode_length[d] = 0
        obs = b.observe(out=self._obs)
        return obs, reward, torch.zeros_like(done), done, info
