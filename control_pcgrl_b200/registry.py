"""Env-id registry: "{prob}-{rep}-v0" for every (problem x representation) pair.

Mirrors control_pcgrl/__init__.py:8-37 (ids, kwargs prob/rep, entry point by problem class) without
depending on gymnasium's global registry; when gymnasium is importable the ids are also registered there
under the namespace "b200/" so `gymnasium.make("b200/binary-narrow-v0", cfg=cfg)` works.
"""
from __future__ import annotations

from .problems import PROBLEM_NAMES, REPRESENTATION_ALIASES
from . import spaces

REGISTRY = {}
for _p in PROBLEM_NAMES:
    for _r in REPRESENTATION_ALIASES:
        REGISTRY[f"{_p}-{_r}-v0"] = {"prob": _p, "rep": REPRESENTATION_ALIASES[_r]}


def parse_id(env_id: str):
    if env_id.startswith("b200/"):
        env_id = env_id[5:]
    if env_id not in REGISTRY:
        raise KeyError(f"unknown env id {env_id!r}; ids look like 'binary-narrow-v0'")
    return REGISTRY[env_id]


def make(env_id: str, cfg=None, device="cuda:0", **kwargs):
    """gym.make("<prob>-<rep>-v0", cfg=cfg) -> single-env PcgrlEnv façade (N = 1 on the GPU)."""
    from .envs import PcgrlEnv
    kw = parse_id(env_id)
    if cfg is None:
        from .config import make_config
        cfg = make_config(kw["prob"], kw["rep"])
    return PcgrlEnv(cfg, prob=kw["prob"], rep=kw["rep"], device=device)


def make_batched(env_id: str, n_envs: int, cfg=None, **kwargs):
    """The batched env behind an id: N grids on one GPU."""
    from .batched_env import BatchedPcgrlEnv
    from .config import make_config
    kw = parse_id(env_id)
    if cfg is None:
        cfg = make_config(kw["prob"], kw["rep"])
    return BatchedPcgrlEnv(cfg, n_envs, **kwargs)


if spaces.HAVE_GYMNASIUM:  # pragma: no cover
    try:
        from gymnasium.envs.registration import register
        for _id, _kw in REGISTRY.items():
            register(id=f"b200/{_id}", entry_point="control_pcgrl_b200.envs:PcgrlEnv", kwargs=_kw)
    except Exception:  # noqa: BLE001
        pass
