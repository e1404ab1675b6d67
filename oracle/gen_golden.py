"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REAL reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The fixtures are committed; the GPU box only ever reads the .npz files.

Two kinds of fixtures:
  stats_<problem>.npz    grids + the reference's own get_stats() output for each
  trace_<name>.npz       full episodes through the reference's real env stack
                         (PcgrlEnv + obs wrappers + ControlWrapper): per step action, reward,
                         done, stats, pos, grid (+ a few observations)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim as R  # noqa: E402
from oracle.pcgrl_oracle import STAT_NAMES, TILES  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

ZELDA_W = dict(player=3, key=3, door=3, regions=5, enemies=1, **{"nearest-enemy": 2, "path-length": 1})
BINARY_W = {"regions": 1, "path-length": 1}   # configs/task/binary.yaml


def ref_problem(problem, map_shape):
    R.install()
    from control_pcgrl.envs.probs import PROBLEMS
    cfg = R.make_cfg(problem, "narrow", map_shape, weights={})
    p = PROBLEMS[problem](cfg=cfg)
    p.adjust_param(cfg=cfg)
    return p


def ref_stats(problem_obj, problem, grid):
    H = R.load_helpers()
    helper = H.h3 if grid.ndim == 3 else H.h2
    smap = helper.get_string_map(grid, TILES[problem])
    st = problem_obj.get_stats(smap)
    return [int(st[k]) for k in STAT_NAMES[problem]]


# ------------------------------------------------------------------------------------------ grids
def binary_grids():
    rng = np.random.default_rng(20261017)
    out = []
    for shape in [(16, 16)]:
        for p in (0.1, 0.2, 0.35, 0.5, 0.65, 0.8, 0.9):
            for _ in range(60):
                out.append((rng.random(shape) < p).astype(np.uint8))
    # hand-built known answers
    out.append(np.zeros((16, 16), np.uint8))
    out.append(np.ones((16, 16), np.uint8))
    out.append((np.add.outer(np.arange(16), np.arange(16)) % 2).astype(np.uint8))
    serp = np.ones((16, 16), np.uint8)
    serp[0::2, :] = 0
    for i, y in enumerate(range(1, 16, 2)):
        serp[y, 15 if i % 2 == 0 else 0] = 0
    out.append(serp)
    rng2 = np.random.default_rng(12345)
    for _ in range(3):
        out.append((rng2.random((16, 16)) < 0.5).astype(np.uint8))
    return out


def binary_grids_other_shapes():
    rng = np.random.default_rng(7)
    out = []
    for shape in [(7, 11), (5, 5), (3, 20), (16, 18), (10, 14), (32, 32), (1, 1), (1, 9), (12, 1), (20, 27)]:
        for p in (0.2, 0.5, 0.7):
            for _ in range(8):
                out.append((rng.random(shape) < p).astype(np.uint8))
    R.install()
    from control_pcgrl.envs.probs.binary.eval_maps import binary_eval_maps
    out.append(np.array(binary_eval_maps[0]["map"], dtype=np.uint8))
    return out


def zelda_grids():
    rng = np.random.default_rng(99)
    out = []
    base = np.array([0.58, 0.3, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02])
    for shape in [(7, 11), (7, 11), (7, 11), (16, 16), (5, 9)]:
        for _ in range(60):
            out.append(rng.choice(8, size=shape, p=base).astype(np.uint8))
        # playable-ish: exactly one player/key/door planted on a sparse map
        for _ in range(60):
            g = rng.choice(8, size=shape, p=[0.7, 0.2, 0, 0, 0, 0.04, 0.03, 0.03]).astype(np.uint8)
            cells = rng.choice(g.size, size=3, replace=False)
            for c, t in zip(cells, (2, 3, 4)):
                g.flat[c] = t
            out.append(g)
    # reference level fixtures il/playable_maps/zelda_lvl*.txt (13x9 text incl. border; gen_trajectories.py:18-27)
    chars = {".": 0, "w": 1, "A": 2, "+": 3, "g": 4, "1": 5, "3": 6, "2": 7}
    d = os.path.join(R.REF_ROOT, "control_pcgrl", "il", "playable_maps")
    for i in range(50):
        with open(os.path.join(d, f"zelda_lvl{i}.txt")) as f:
            rows = [ln.rstrip("\n") for ln in f if ln.strip()]
        g = np.array([[chars[c] for c in r] for r in rows], dtype=np.uint8)[1:-1, 1:-1]
        out.append(g)
    return out


def save_stats_fixture(problem, grids, name=None):
    by_shape = {}
    for g in grids:
        by_shape.setdefault(g.shape, []).append(g)
    arrays = {}
    for i, (shape, gs) in enumerate(sorted(by_shape.items())):
        pobj = ref_problem(problem, shape)
        gs = np.stack(gs)
        st = np.array([ref_stats(pobj, problem, g) for g in gs], dtype=np.int64)
        arrays[f"grids_{i}"] = gs
        arrays[f"stats_{i}"] = st
    arrays["stat_names"] = np.array(STAT_NAMES[problem])
    path = os.path.join(OUT, f"stats_{name or problem}.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: v.shape for k, v in arrays.items() if k.startswith("stats_")})


# ------------------------------------------------------------------------------------------ traces
def run_trace(name, problem, rep, map_shape, obs_window, weights, controls=None, n_envs=4, seed=0,
              max_board_scans=3, change_percentage=None, raw_only=False, n_steps=None, init_p=None,
              targets=None, obs_every=97):
    rng = np.random.default_rng(seed)
    cfg = R.make_cfg(problem, rep, map_shape, obs_window=obs_window, weights=weights, controls=controls,
                     max_board_scans=max_board_scans, change_percentage=change_percentage)
    n_tiles = len(TILES[problem])
    rec = dict(grid0=[], pos0=[], actions=[], rewards=[], dones=[], stats=[], pos=[], grids=[], stats0=[],
               obs=[], obs_step=[], obs0=[], changes=[], trg=[])
    for e in range(n_envs):
        env = R.make_wrapped_env(cfg, raw_only=raw_only)
        if init_p is None:
            g0 = rng.integers(0, n_tiles, size=map_shape).astype(np.uint8)
        else:
            g0 = rng.choice(n_tiles, size=map_shape, p=init_p).astype(np.uint8)
        R.inject_map(env, g0)
        trg_e = []
        if targets is not None:
            t = {k: float(rng.uniform(*env.cond_bounds[k])) for k in targets}   # intended UniformNoiseyTargets
            env.set_trgs(t)
            trg_e = [t[k] for k in targets]
        ob, _ = env.reset()
        u = env.unwrapped
        rep_obj = u._rep.unwrapped
        pos0 = [int(v) for v in rep_obj._pos] if rep in ("narrow", "turtle") else [0] * len(map_shape)
        stats0 = [int(u._rep_stats[k]) for k in STAT_NAMES[problem]]
        acts, rews, dones, stats, poss, grids, obs_l, obs_s, chg = [], [], [], [], [], [], [], [], []
        done = False
        t = 0
        while not done and (n_steps is None or t < n_steps):
            if rep == "cellular":
                a = rng.random((n_tiles, *map_shape)).astype(np.float32)
                if t % 3 == 2:      # sometimes keep most of the map: near-no-op logits
                    a = np.eye(n_tiles, dtype=np.float32)[u._rep.unwrapped._map.astype(int)].transpose(2, 0, 1).copy()
                    if t % 6 == 2:
                        a[:, rng.integers(map_shape[0]), rng.integers(map_shape[1])] = rng.random(n_tiles)
                act_store = a
            elif rep == "wide" and raw_only:
                a = [int(rng.integers(s)) for s in map_shape] + [int(rng.integers(n_tiles))]
                act_store = np.array(a)
            else:
                a = int(rng.integers(env.action_space.n))
                act_store = a
            ob, r, done, trunc, info = env.step(a)
            acts.append(act_store)
            rews.append(float(r))
            dones.append(bool(done))
            stats.append([int(u._rep_stats[k]) for k in STAT_NAMES[problem]])
            p = getattr(rep_obj, "_pos", None)
            poss.append([int(v) for v in p] if p is not None and rep in ("narrow", "turtle") else [0] * len(map_shape))
            grids.append(np.array(rep_obj._map, dtype=np.uint8))
            chg.append(int(info["changes"]))
            if (t % obs_every == 0 or done) and not raw_only:
                obs_l.append(np.asarray(ob, dtype=np.float64))
                obs_s.append(t)
            t += 1
        rec["grid0"].append(g0); rec["pos0"].append(pos0); rec["stats0"].append(stats0)
        rec["actions"].append(np.array(acts)); rec["rewards"].append(rews); rec["dones"].append(dones)
        rec["stats"].append(stats); rec["pos"].append(poss); rec["grids"].append(np.stack(grids))
        rec["obs"].append(np.stack(obs_l) if obs_l else np.zeros((0,))); rec["obs_step"].append(obs_s)
        rec["changes"].append(chg); rec["trg"].append(trg_e)
    arrays = {"n_envs": np.array(n_envs)}
    for k, v in rec.items():
        if k == "obs0":
            continue
        for e in range(n_envs):          # episodes can be ragged (change_percentage) -> one array per env
            arrays[f"{k}_{e}"] = np.array(v[e])
    arrays["meta_problem"] = np.array(problem); arrays["meta_rep"] = np.array(rep)
    arrays["meta_map_shape"] = np.array(map_shape); arrays["meta_obs_window"] = np.array(obs_window)
    arrays["meta_max_board_scans"] = np.array(max_board_scans)
    arrays["meta_change_percentage"] = np.array(-1.0 if change_percentage is None else change_percentage)
    arrays["meta_controls"] = np.array(controls or [], dtype=str)
    arrays["meta_weight_keys"] = np.array(list(weights.keys()), dtype=str)
    arrays["meta_weight_vals"] = np.array(list(weights.values()), dtype=np.float64)
    arrays["meta_raw_only"] = np.array(raw_only)
    arrays["meta_targets"] = np.array(targets or [], dtype=str)
    path = os.path.join(OUT, f"trace_{name}.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, "steps", [arrays[f"rewards_{e}"].shape[0] for e in range(n_envs)], "bytes", os.path.getsize(path))


def main(which=None):
    os.makedirs(OUT, exist_ok=True)
    jobs = {
        "stats_binary": lambda: save_stats_fixture("binary", binary_grids()),
        "stats_binary_shapes": lambda: save_stats_fixture("binary", binary_grids_other_shapes(), "binary_shapes"),
        "stats_zelda": lambda: save_stats_fixture("zelda", zelda_grids()),
        "trace_binary_narrow": lambda: run_trace("binary_narrow", "binary", "narrow", (16, 16), (32, 32), BINARY_W),
        "trace_binary_narrow_chg": lambda: run_trace("binary_narrow_chg", "binary", "narrow", (16, 16), (32, 32),
                                                     BINARY_W, change_percentage=0.2, seed=1),
        "trace_binary_turtle": lambda: run_trace("binary_turtle", "binary", "turtle", (16, 16), (32, 32), BINARY_W, seed=2),
        "trace_binary_wide_ctrl": lambda: run_trace("binary_wide_ctrl", "binary", "wide", (16, 16), (16, 16), BINARY_W,
                                                    controls=["regions", "path-length"], seed=3,
                                                    targets=["regions", "path-length"]),
        "trace_binary_cellular": lambda: run_trace("binary_cellular", "binary", "cellular", (16, 16), (16, 16), BINARY_W,
                                                   seed=4, raw_only=True, n_steps=40, n_envs=2),
        "trace_zelda_turtle": lambda: run_trace("zelda_turtle", "zelda", "turtle", (7, 11), (22, 22), ZELDA_W, seed=5,
                                                init_p=[0.58, 0.3, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02], obs_every=61),
        "trace_zelda_narrow": lambda: run_trace("zelda_narrow", "zelda", "narrow", (7, 11), (22, 22), ZELDA_W, seed=6,
                                                init_p=[0.58, 0.3, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02], obs_every=61),
        "trace_zelda_wide_raw": lambda: run_trace("zelda_wide_raw", "zelda", "wide", (7, 11), (7, 11), ZELDA_W, seed=7,
                                                  raw_only=True, n_envs=2,
                                                  init_p=[0.58, 0.3, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02]),
    }
    for k, fn in jobs.items():
        if which and k not in which:
            continue
        fn()


if __name__ == "__main__":
    main(sys.argv[1:] or None)
