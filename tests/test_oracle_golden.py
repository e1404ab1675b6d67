"""CPU: pin oracle/pcgrl_oracle.py against fixtures produced by the real reference."""
import numpy as np
import pytest

from oracle import pcgrl_oracle as O
from tests.golden_util import TRACES, TRACES_OPEN_ENDED, TRACES_SEARCH, TRACES_WRAPPED, Trace, load_stats


@pytest.mark.parametrize("name,problem", [("binary", "binary"), ("binary_shapes", "binary"), ("zelda", "zelda"),
                                          ("sokoban", "sokoban"), ("smb", "smb"), ("maze3d", "minecraft_3D_maze"),
                                          ("minecraft_2D_maze", "minecraft_2D_maze"),
                                          ("binary_big", "binary"), ("zelda_big", "zelda")])
def test_stats_match_reference(name, problem):
    names, groups = load_stats(name)
    assert names == O.STAT_NAMES[problem]
    n = 0
    for grids, stats in groups:
        for g, want in zip(grids, stats):
            got = O.stats_vector(problem, O.get_stats(problem, g))
            assert got == [int(v) for v in want], (g.shape, got, want)
            n += 1
    assert n > (15 if name.endswith("_big") else 100)


def test_known_answers():
    # SURVEY.md section C, computed with the reference's helper.py
    assert O.binary_stats(np.zeros((16, 16), int)) == {"regions": 1, "path-length": 30}
    assert O.binary_stats(np.ones((16, 16), int)) == {"regions": 0, "path-length": 0}
    rng = np.random.default_rng(12345)
    want = [(30, 13), (34, 12), (22, 17)]
    for r, p in want:
        g = (rng.random((16, 16)) < 0.5).astype(np.uint8)
        assert O.binary_stats(g) == {"regions": r, "path-length": p}
    assert O.range_reward(3, 1, 1, 1) == -2
    assert O.range_reward(0, 2, 1, 1) == -2
    assert O.range_reward(130, 120, 125, 125) == -10


def replay_oracle(tr: Trace, e: int):
    d = tr.envs[e]
    env = O.OracleEnv(tr.problem, tr.rep, tr.map_shape, weights=tr.weights, controls=tr.controls,
                      max_board_scans=tr.max_board_scans, change_percentage=tr.change_percentage,
                      act_window=tr.act_window)
    st0 = env.reset(d["grid0"], pos=d["pos0"], targets=tr.targets(e), static=d["static"])
    assert O.stats_vector(tr.problem, st0) == [int(v) for v in d["stats0"]]
    n_tiles = len(O.TILES[tr.problem])
    h, w = tr.obs_window[:2]
    for t in range(len(d["rewards"])):
        a = d["actions"][t]
        if tr.rep == "wide" and not tr.raw_only:
            a = O.actionmap_unravel(a, h, w, n_tiles)
        elif tr.rep in ("narrow", "turtle") and tr.act_window is None:
            a = int(a)
        r, done, _ = env.step(a)
        assert done == bool(d["dones"][t]), (tr.name, e, t)
        assert O.stats_vector(tr.problem, env.stats) == [int(v) for v in d["stats"][t]], (tr.name, e, t)
        if t in d["grid_at"]:
            assert np.array_equal(env.grid, d["grids"][d["grid_at"][t]]), (tr.name, e, t)
        assert env.changes == int(d["changes"][t])
        assert r == pytest.approx(float(d["rewards"][t]), rel=1e-6, abs=1e-9), (tr.name, e, t)
        if tr.rep in ("narrow", "turtle"):
            assert env.pos == [int(v) for v in d["pos"][t]], (tr.name, e, t)
        if t in d["obs_step"]:
            want = d["obs"][list(d["obs_step"]).index(t)]
            if tr.rep in ("narrow", "turtle"):
                got = O.cropped_onehot(env.grid, env.pos, tr.obs_window, n_tiles)
            else:
                got = O.full_onehot(env.grid, n_tiles)
            if tr.static:
                sb = O.static_builds_crop(d["static"], env.pos, tr.obs_window)
                got = np.concatenate([got, sb[..., None]], axis=-1)
            if tr.controls:
                ch = O.target_channels(got.shape[:-1], tr.controls, env.targets, env.stats, env.cond_bounds)
                got = np.concatenate([ch, got], axis=-1)
            assert got.shape == want.shape
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    assert bool(d["dones"][-1]) or tr.name in TRACES_OPEN_ENDED


@pytest.mark.parametrize("name", TRACES + TRACES_SEARCH + TRACES_WRAPPED)
def test_trace_matches_reference(name):
    tr = Trace(name)
    for e in range(tr.n_envs):
        replay_oracle(tr, e)


def test_multiagent_turtle_matches_reference():
    """Multi-agent turtle: the reference's CroppedImagePCGRLWrapper + ControlWrapper +
    MultiAgentWrapper stack, replayed by the restatement sub-step by sub-step."""
    from tests.golden_util import load_multiagent
    for c in load_multiagent():
        n_tiles = len(O.TILES[c["problem"]])
        for d in c["envs"]:
            env = O.OracleEnv(c["problem"], "turtle", c["map_shape"], weights=c["weights"],
                              change_percentage=c["change_percentage"])
            st0 = env.reset(d["grid0"], agent_pos=d["pos0"])
            assert O.stats_vector(c["problem"], st0) == [int(v) for v in d["stats0"]]
            def view(i):
                got = O.cropped_onehot(env.grid, env.agent_pos[i], c["obs_window"], n_tiles)
                if c["show_agents"]:      # ToImage appends the 'agent_occupancy' plane (wrappers.py:451-453)
                    occ = O.agent_occupancy_crop(env.agent_pos, env.agent_pos[i], c["obs_window"], c["map_shape"])
                    got = np.concatenate([got, occ[..., None]], axis=-1)
                return got

            for i in range(c["n_agents"]):
                np.testing.assert_allclose(view(i), d["obs0"][i], rtol=1e-12, atol=0)
            for t in range(len(d["actions"])):
                for i in range(c["n_agents"]):
                    r, done, _ = env.step(int(d["actions"][t][i]), agent=i)
                    assert done == bool(d["dones"][t][i]), (t, i)
                    assert r == pytest.approx(float(d["rewards"][t][i]), rel=1e-6, abs=1e-9), (t, i)
                    assert O.stats_vector(c["problem"], env.stats) == [int(v) for v in d["stats"][t][i]], (t, i)
                    assert env.agent_pos == d["pos"][t][i].tolist(), (t, i)
                    if t in d["obs_step"]:
                        want = d["obs"][list(d["obs_step"]).index(t)][i]
                        np.testing.assert_allclose(view(i), want, rtol=1e-12, atol=0)
                assert np.array_equal(env.grid, d["grids"][t]), t
                assert env.iteration == int(d["iterations"][t]) and env.changes == int(d["changes"][t])
            assert bool(d["dones"][-1].all())


def test_legacy_range_reward_matches_reference():
    """Problem.get_reward of the reference's non-ctrl classes (fixture: oracle/gen_golden.py legacy_reward)."""
    import os
    from tests.golden_util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "legacy_reward.npz"))
    for problem in ("binary", "zelda", "sokoban", "smb"):
        names = O.STAT_NAMES[problem]
        for a, b, want in zip(z[f"{problem}_new"], z[f"{problem}_old"], z[f"{problem}_reward"]):
            got = O.legacy_reward(problem, dict(zip(names, a.tolist())), dict(zip(names, b.tolist())))
            assert got == want, (problem, a, b, got, want)
    # minecraft_2D_maze: Minecraft2DmazeProblem.get_reward, minecraft_2D_maze_prob.py:106-115
    z = np.load(os.path.join(GOLDEN, "legacy_reward_minecraft_2D_maze.npz"))
    names = O.STAT_NAMES["minecraft_2D_maze"]
    for a, b, want in zip(z["new"], z["old"], z["reward"]):
        got = O.legacy_reward("minecraft_2D_maze", dict(zip(names, a.tolist())), dict(zip(names, b.tolist())))
        assert got == want, (a, b, got, want)
