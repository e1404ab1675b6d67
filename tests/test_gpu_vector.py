"""GPU: the caller-side seams: the RLlib VectorEnv surface and the single-env facade fixes."""
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def fake_ray(monkeypatch):
    """A stand-in for ray.rllib.env.vector_env (ray is not in this image), like oracle/refshim.py does for the
    reference: only the base class RLlib's samplers type-check against."""
    class VectorEnv:
        def __init__(self, observation_space, action_space, num_envs):
            self.observation_space, self.action_space, self.num_envs = observation_space, action_space, num_envs
            self.base_init_ran = True

    mods = {}
    for name in ("ray", "ray.rllib", "ray.rllib.env", "ray.rllib.env.vector_env"):
        mods[name] = types.ModuleType(name)
    mods["ray.rllib.env.vector_env"].VectorEnv = VectorEnv
    for name, m in mods.items():
        monkeypatch.setitem(sys.modules, name, m)
    return VectorEnv


def test_rllib_vector_env_surface(fake_ray):
    import control_pcgrl_b200 as P
    n = 48
    cfg = P.make_config("binary", "narrow", max_board_scans=0.05, controls=["regions"])
    venv = P.make_rllib_vector_env(cfg, n, seed=4, obs_dtype=np.float32)
    assert isinstance(venv, fake_ray) and venv.base_init_ran and venv.num_envs == n
    twin = P.BatchedPcgrlEnv(cfg, n, seed=4)            # the same shard stepped by hand, same reset call pattern
    obs, infos = venv.vector_reset()
    twin.targets.copy_(venv.env.targets)
    twin.reset()
    assert len(obs) == n and len(infos) == n and obs[0].shape == venv.observation_space.shape == (32, 32, 5)
    np.testing.assert_array_equal(np.stack(obs), twin.observe().cpu().numpy())
    rng = np.random.default_rng(0)
    finished = 0
    for t in range(40):
        acts = [int(a) for a in rng.integers(0, 2, size=n)]
        obs, rew, term, trunc, infos = venv.vector_step(acts)
        r2, d2 = twin.step(torch.tensor(acts, dtype=torch.int32, device=twin.device))
        assert all(isinstance(x, float) for x in rew) and term == [False] * n
        np.testing.assert_array_equal(np.array(rew, dtype=np.float32), r2.cpu().numpy())
        assert trunc == d2.bool().cpu().tolist()
        np.testing.assert_array_equal(np.stack(obs), twin.observe().cpu().numpy())      # terminal obs included
        st = twin.stats.cpu().numpy()
        for i in (0, n // 2, n - 1):
            assert [infos[i][k] for k in twin.stat_names] == st[i].tolist()
        if any(trunc):
            twin.reset(mask=twin.done)     # vector_step already reset its finished envs (one launch for all)
        for i in (0, n // 2, n - 1):
            assert venv.get_sub_environments()[i].metrics == twin.stats_dict(i)
        if any(trunc):
            first = twin.observe().cpu().numpy()
            for i in np.flatnonzero(trunc):
                o, info = venv.reset_at(int(i))
                np.testing.assert_array_equal(o, first[i])
                finished += 1
            assert not venv._fresh.any()
    assert finished >= n                                                                  # every env ended once
    # an explicit reset of a running env is a real (single-env) reset
    o, _ = venv.reset_at(3)
    assert int(venv.env.iteration[3]) == 0 and o.shape == (32, 32, 5)
    assert isinstance(venv.get_sub_environments()[3].metric_trgs["regions"], float)


def test_facade_checks_ids_patch_actions_and_holey_spaces():
    import control_pcgrl_b200 as P
    from control_pcgrl_b200 import envs as E
    with pytest.raises(ValueError, match="problem"):
        P.make("zelda-narrow-v0", cfg=P.make_config("binary", "narrow"))
    with pytest.raises(ValueError, match="representation"):
        P.make("binary-turtle-v0", cfg=P.make_config("binary", "narrow"))
    # MultiActionRepresentation through the id (wrappers.py:438-443: MultiDiscrete([C] * prod(act_window)))
    env = P.make("binary-narrow-v0", cfg=P.make_config("binary", "narrow", act_window=(3, 3)))
    assert list(env.action_space.nvec) == [2] * 9
    env.reset()
    before = env._b.maps[0].clone()
    _, _, _, _, info = env.step(np.ones(9, dtype=np.int64))
    after = env._b.maps[0]
    assert bool((after[:3, :3] == 1).all()) and torch.equal(after[3:], before[3:])
    # declared == returned observation shapes, also for the bordered (holey) and frozen-tile stacks
    for cfg in (P.make_config("binary_holey", "narrow"), P.make_config("binary", "turtle", static_tile_wrapper=True,
                                                                      static_prob=0.3)):
        w = E.CroppedImagePCGRLWrapper(f"{cfg.task.problem}-{cfg.representation}-v0", cfg=cfg)
        o, _ = w.reset()
        assert tuple(o.shape) == tuple(w.observation_space.shape), (o.shape, w.observation_space.shape)


def test_minecraft_2d_maze_is_the_binary_machine_under_other_tile_names():
    """minecraft_2D_maze_prob.py:87-93: regions / longest path over "AIR" (tile 0): same kernels, cache and split
    path as binary, its own ids, tiles and legacy reward weights."""
    import control_pcgrl_b200 as P
    n = 4096
    a = P.BatchedPcgrlEnv(P.make_config("minecraft_2D_maze", "narrow"), n, seed=1, reward_mode="range")
    b = P.BatchedPcgrlEnv(P.make_config("binary", "narrow", map_shape=(14, 14)), n, seed=1)
    assert a.spec.tiles == ["AIR", "DIRT"] and a.map_shape == (14, 14) and a.cache is not None
    a.reset()
    b.reset(grids=a.maps)
    gen = torch.Generator(device=a.device).manual_seed(0)
    for t in range(60):
        act = torch.randint(0, 2, (n,), generator=gen, device=a.device, dtype=torch.int32)
        a.step(act)
        b.step(act)
        assert torch.equal(a.stats, b.stats) and torch.equal(a.grids, b.grids)
    assert torch.equal(a.compute_stats(a.maps), a.stats)
    env = P.make("minecraft_2D_maze-narrow-v0")
    env.reset()
    assert env.get_num_tiles() == 2 and env.action_space.n == 2


def test_policy_input_from_tile_codes_equals_reference_permute():
    """rl/models.py:60-66 feeds `obs.permute(0, 3, 1, 2).float()`; the same tensor from 1-byte tile codes."""
    import control_pcgrl_b200 as P
    from control_pcgrl_b200.policy_input import conv_input_from_codes
    for problem, rep, shape, window in (("binary", "narrow", (16, 16), (32, 32)), ("zelda", "turtle", (7, 11), (22, 22)),
                                        ("binary", "wide", (16, 16), (16, 16))):
        env = P.BatchedPcgrlEnv(P.make_config(problem, rep, map_shape=shape, obs_window=window), 300, seed=4)
        env.reset()
        for i, d in enumerate(shape):
            env.pos[:, i] = torch.randint(0, d, (300,), device=env.device, dtype=torch.int32)
        ref = env.observe(dtype=torch.float64).permute(0, 3, 1, 2).float()
        got = conv_input_from_codes(env.observe(onehot=False), ref.shape[1])
        assert got.shape == ref.shape and torch.equal(got, ref)
        assert got.is_contiguous(memory_format=torch.channels_last)
        conv = torch.nn.Conv2d(ref.shape[1], 8, kernel_size=7, stride=2, padding=3).to(env.device)
        assert torch.allclose(conv(got), conv(ref.contiguous()), atol=1e-5)


@pytest.mark.parametrize("problem,rep,change_pct", [("binary", "narrow", None), ("zelda", "turtle", 0.3)])
def test_sharded_vector_env_equals_one_shard(problem, rep, change_pct):
    """PcgrlVectorEnv(shards=k) steps and observes consecutive env ranges on their own streams (the search of one
    range beside the observation writes of another); maps, observations, rewards, dones and the episode
    bookkeeping must be those of the one-shard env -- resets are counter-based on the global env index, with and
    without lock-step episodes (a change budget makes envs finish at different steps)."""
    import control_pcgrl_b200 as P
    from control_pcgrl_b200.vector_env import PcgrlVectorEnv
    n = 5_003
    kw = dict(change_percentage=change_pct) if change_pct else {}
    cfg = P.make_config(problem, rep, max_board_scans=0.08, **kw)
    one = PcgrlVectorEnv(cfg, n, seed=5, obs_dtype=torch.uint8)
    many = PcgrlVectorEnv(cfg, n, seed=5, obs_dtype=torch.uint8, shards=3)
    assert len(many.shards) == 3 and sum(b.n_envs for b in many.shards) == n
    o1, _ = one.reset()
    o2, _ = many.reset()
    assert torch.equal(o1, o2)
    n_act = one.single_action_space.n
    gen = torch.Generator(device=o1.device).manual_seed(2)
    ends = 0
    for t in range(int(one.env.max_iterations * 2.3) + 5):
        act = torch.randint(0, n_act, (n,), generator=gen, device=o1.device, dtype=torch.int32)
        a1, r1, _, d1, i1 = one.step(act)
        a2, r2, _, d2, i2 = many.step(act)
        assert torch.equal(r1, r2) and torch.equal(d1, d2), t
        assert torch.equal(a1, a2), t
        assert torch.equal(one.episode_return, many.episode_return) and torch.equal(one.episode_length, many.episode_length)
        assert bool(i1) == bool(i2), t
        if i1:
            ends += 1
            f = i1["_final"]
            assert torch.equal(f, i2["_final"])
            for k in ("final_stats", "final_return", "final_length"):
                assert torch.equal(i1[k][f], i2[k][f]), (t, k)
    assert ends >= 2
    assert torch.equal(torch.cat([b.maps for b in many.shards]), one.env.maps)


def test_vector_env_tile_code_observations():
    """PcgrlVectorEnv(onehot=False) hands out the crop's tile codes (1 byte per pixel); expanded by
    policy_input.conv_input_from_codes they are the one-hot observation of the same env, step by step."""
    import control_pcgrl_b200 as P
    from control_pcgrl_b200.policy_input import conv_input_from_codes
    from control_pcgrl_b200.vector_env import PcgrlVectorEnv
    n = 777
    cfg = P.make_config("zelda", "turtle", max_board_scans=0.1)
    hot = PcgrlVectorEnv(cfg, n, seed=3, obs_dtype=torch.float32)
    codes = PcgrlVectorEnv(cfg, n, seed=3, onehot=False, shards=2)
    o1, _ = hot.reset()
    o2, _ = codes.reset()
    assert o2.dtype == torch.uint8 and tuple(o2.shape) == (n, *hot.env.obs_window, 1)
    gen = torch.Generator(device=o1.device).manual_seed(0)
    for t in range(40):
        assert torch.equal(conv_input_from_codes(o2, hot.env.n_tiles + 1), o1.permute(0, 3, 1, 2)), t
        act = torch.randint(0, hot.single_action_space.n, (n,), generator=gen, device=o1.device, dtype=torch.int32)
        o1, r1, _, d1, _ = hot.step(act)
        o2, r2, _, d2, _ = codes.step(act)
        assert torch.equal(r1, r2) and torch.equal(d1, d2), t
