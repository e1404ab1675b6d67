"""Loaders for the committed fixtures under tests/golden/ (made by oracle/gen_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_stats(name):
    """-> (stat_names, [(grids[n,*shape], stats[n,K]), ...])"""
    z = np.load(os.path.join(GOLDEN, f"stats_{name}.npz"))
    groups = []
    i = 0
    while f"grids_{i}" in z:
        groups.append((z[f"grids_{i}"], z[f"stats_{i}"]))
        i += 1
    return [str(s) for s in z["stat_names"]], groups


class Trace:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, f"trace_{name}.npz"))
        self.name = name
        self.problem = str(z["meta_problem"])
        self.rep = str(z["meta_rep"])
        self.map_shape = tuple(int(v) for v in z["meta_map_shape"])
        self.obs_window = tuple(int(v) for v in z["meta_obs_window"])
        self.max_board_scans = int(z["meta_max_board_scans"])
        cp = float(z["meta_change_percentage"])
        self.change_percentage = None if cp < 0 else cp
        self.controls = [str(s) for s in z["meta_controls"]] or None
        self.weights = {str(k): float(v) for k, v in zip(z["meta_weight_keys"], z["meta_weight_vals"])}
        self.raw_only = bool(z["meta_raw_only"])
        self.target_names = [str(s) for s in z["meta_targets"]]
        # representation wrappers (newer traces): MultiActionRepresentation patch size, StaticTileRepresentation
        aw = z["meta_act_window"] if "meta_act_window" in z else np.zeros((0,))
        self.act_window = tuple(int(v) for v in aw) or None
        self.static = bool(z["meta_static"]) if "meta_static" in z else False
        self.n_envs = int(z["n_envs"])
        self.envs = []
        for e in range(self.n_envs):
            d = {k: z[f"{k}_{e}"] for k in
                 ("grid0", "pos0", "stats0", "actions", "rewards", "dones", "stats", "pos", "grids",
                  "obs", "obs_step", "changes", "trg")}
            # newer traces keep the grid only every few steps: grid_at[t] -> index into d["grids"]
            steps = z[f"grids_step_{e}"] if f"grids_step_{e}" in z else np.arange(len(d["rewards"]))
            d["grid_at"] = {int(t): i for i, t in enumerate(steps)}
            d["static"] = z[f"static_{e}"] if self.static else None
            self.envs.append(d)

    def targets(self, e):
        return {k: float(v) for k, v in zip(self.target_names, self.envs[e]["trg"])}


TRACES = ["binary_narrow", "binary_narrow_chg", "binary_turtle", "binary_wide_ctrl", "binary_cellular",
          "zelda_turtle", "zelda_narrow", "zelda_wide_raw"]
TRACES_SEARCH = ["sokoban_narrow", "sokoban_turtle", "sokoban_cellular", "smb_narrow", "smb_narrow_small",
                 "smb_turtle_small", "maze3d_narrow", "maze3d_turtle", "maze3d_wide_raw", "maze3d_cellular"]
# representation wrappers of envs/reps/wrappers.py: action patches, frozen tiles
TRACES_WRAPPED = ["binary_patch33", "binary_patch42_chg", "zelda_squeegee", "maze3d_patch", "binary_static_narrow",
                  "zelda_static_turtle", "binary_static_patch", "sokoban_static_narrow"]
# traces that stop before the episode ends (n_steps cap in oracle/gen_golden.py)
TRACES_OPEN_ENDED = ("binary_cellular", "smb_narrow", "maze3d_narrow", "maze3d_cellular")


def load_multiagent():
    """tests/golden/multiagent_turtle.npz (oracle/gen_golden.py multiagent_fixture) -> list of cases, each a dict with
    problem, map_shape, obs_window, n_agents, change_percentage, weights and `envs` (per-env arrays: grid0, pos0
    [A, nd], obs0 [A, ...], stats0, actions [T, A], rewards [T, A], dones [T, A], stats [T, A, K], pos [T, A, A, nd]
    (all agents' positions after each agent's sub-step), grids [T, ...] (after the whole multi-agent step), obs,
    obs_step, iterations, changes)."""
    z = np.load(os.path.join(GOLDEN, "multiagent_turtle.npz"))
    cases = []
    for ci in range(int(z["n_cases"])):
        cp = float(z[f"c{ci}_change_percentage"])
        c = dict(problem=str(z[f"c{ci}_problem"]), map_shape=tuple(int(v) for v in z[f"c{ci}_map_shape"]),
                 obs_window=tuple(int(v) for v in z[f"c{ci}_obs_window"]), n_agents=int(z[f"c{ci}_n_agents"]),
                 change_percentage=None if cp < 0 else cp,
                 show_agents=bool(z[f"c{ci}_show_agents"]) if f"c{ci}_show_agents" in z else False,
                 weights={str(k): float(v) for k, v in zip(z[f"c{ci}_weight_keys"], z[f"c{ci}_weight_vals"])}, envs=[])
        for e in range(int(z[f"c{ci}_n_envs"])):
            pre = f"c{ci}_e{e}_"
            c["envs"].append({k: z[pre + k] for k in ("grid0", "pos0", "obs0", "stats0", "actions", "rewards", "dones",
                                                      "stats", "pos", "grids", "obs", "obs_step", "iterations",
                                                      "changes")})
        cases.append(c)
    return cases
