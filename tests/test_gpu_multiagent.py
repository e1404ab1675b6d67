"""Multi-agent turtle (envs/reps/wrappers.py:612-651, wrappers.py:697-736) on the GPU against the
fixture recorded from the reference's real CroppedImagePCGRLWrapper + ControlWrapper + MultiAgentWrapper stack
(tests/golden/multiagent_turtle.npz, oracle/gen_golden.py multiagent_fixture): the batched env (all envs of a case
in one batch, one kernel step per agent) and the single-env façade built by make_env."""
import numpy as np
import pytest
import torch

from tests.golden_util import load_multiagent

pytestmark = pytest.mark.gpu


def _cfg(c):
    import control_pcgrl_b200 as P
    cfg = P.make_config(c["problem"], "turtle", map_shape=c["map_shape"], obs_window=c["obs_window"],
                        weights=c["weights"], change_percentage=c["change_percentage"])
    cfg.multiagent.n_agents = c["n_agents"]
    cfg.show_agents = c["show_agents"]      # ShowAgentRepresentation's 'agent_occupancy' plane (fixture case 3)
    return cfg


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_batched_multiagent_matches_reference(ci):
    import control_pcgrl_b200 as P
    c = load_multiagent()[ci]
    envs, A = c["envs"], c["n_agents"]
    n, nd = len(envs), len(c["map_shape"])
    env = P.BatchedPcgrlEnv(_cfg(c), n, random_init_probs=False)
    assert env.n_agents == A
    env.reset(grids=np.stack([d["grid0"] for d in envs]), pos=np.stack([d["pos0"] for d in envs]))
    assert env.stats.cpu().numpy().tolist() == [d["stats0"].tolist() for d in envs]
    for a in range(A):
        ob = env.observe(dtype=torch.float64, agent=a).cpu().numpy()
        for e, d in enumerate(envs):
            np.testing.assert_allclose(ob[e], d["obs0"][a], rtol=1e-12, atol=0)
    T = max(len(d["actions"]) for d in envs)
    for t in range(T):
        live = [e for e, d in enumerate(envs) if t < len(d["actions"])]
        for a in range(A):
            act = torch.tensor([int(d["actions"][t][a]) if t < len(d["actions"]) else 0 for d in envs],
                               dtype=torch.int32, device=env.device)
            r, dn = env.step(act, agent=a)
            r, dn, st = r.cpu().numpy(), dn.cpu().numpy(), env.stats.cpu().numpy()
            ap = env.agent_pos.cpu().numpy()
            for e in live:
                d = envs[e]
                assert bool(dn[e]) == bool(d["dones"][t][a]), (e, t, a)
                assert r[e] == pytest.approx(float(d["rewards"][t][a]), rel=1e-6, abs=1e-6), (e, t, a)
                assert st[e].tolist() == d["stats"][t][a].tolist(), (e, t, a)
                assert ap[:, e, :nd].tolist() == d["pos"][t][a].tolist(), (e, t, a)
            if any(t in envs[e]["obs_step"] for e in live):
                ob = env.observe(dtype=torch.float64, agent=a).cpu().numpy()
                for e in live:
                    d = envs[e]
                    if t in d["obs_step"]:
                        np.testing.assert_allclose(ob[e], d["obs"][list(d["obs_step"]).index(t)][a], rtol=1e-12, atol=0)
        maps, it, ch = env.maps.cpu().numpy(), env.iteration.cpu().numpy(), env.changes.cpu().numpy()
        for e in live:
            d = envs[e]
            assert np.array_equal(maps[e], d["grids"][t]), (e, t)
            assert int(it[e]) == int(d["iterations"][t]) and int(ch[e]) == int(d["changes"][t]), (e, t)
    env.check_status()


def test_make_env_multiagent_matches_reference():
    """The drop-in stack: make_env(cfg) with cfg.multiagent.n_agents -> MultiAgentWrapper, dict in / dict out."""
    import control_pcgrl_b200 as P
    c = load_multiagent()[3]          # zelda, three agents, show_agents
    d = c["envs"][0]
    A = c["n_agents"]
    names = [f"agent_{i}" for i in range(A)]
    env = P.make_env(_cfg(c))
    assert isinstance(env, P.MultiAgentWrapper) and sorted(env.action_space.spaces) == names
    u = env.unwrapped
    u._rep._old_map, u._rep._random_start = d["grid0"].copy(), False
    obs, _ = env.reset()
    u._b.agent_pos[:, 0, :2] = torch.as_tensor(d["pos0"], dtype=torch.int32)   # the spawn draw is the env's own
    for t in range(len(d["actions"])):
        ob, rew, done, trunc, info = env.step({k: int(d["actions"][t][i]) for i, k in enumerate(names)})
        for i, k in enumerate(names):
            assert rew[k] == pytest.approx(float(d["rewards"][t][i]), rel=1e-6, abs=1e-6), (t, k)
            assert bool(done[k]) == bool(d["dones"][t][i]), (t, k)
            if t in d["obs_step"]:
                np.testing.assert_allclose(ob[k], d["obs"][list(d["obs_step"]).index(t)][i], rtol=1e-12, atol=0)
        assert done["__all__"] == bool(d["dones"][t].all())
        assert info[names[-1]]["iterations"] == int(d["iterations"][t])
        assert u._rep.agent_positions.tolist() == d["pos"][t][-1].tolist()
    assert done["__all__"]


def test_step_agents_round_and_spawn():
    """step_agents == the per-agent steps; random spawns are distinct cells; auto-reset after a whole round."""
    import control_pcgrl_b200 as P
    cfg = P.make_config("binary", "turtle", map_shape=(6, 6), obs_window=(12, 12))
    cfg.multiagent.n_agents = 3
    a_env = P.BatchedPcgrlEnv(cfg, 512, seed=3, auto_reset=True)
    b_env = P.BatchedPcgrlEnv(cfg, 512, seed=3)
    a_env.reset()
    b_env.reset()
    assert torch.equal(a_env.agent_pos, b_env.agent_pos) and torch.equal(a_env.grids, b_env.grids)
    cellidx = a_env.agent_pos[:, :, 0] * 6 + a_env.agent_pos[:, :, 1]
    assert int(a_env.agent_pos[:, :, :2].max()) < 6 and int(a_env.agent_pos.min()) >= 0
    assert bool((cellidx[0] != cellidx[1]).all() and (cellidx[0] != cellidx[2]).all() and (cellidx[1] != cellidx[2]).all())
    g = torch.Generator(device=a_env.device).manual_seed(1)
    rounds = (int(a_env.max_iterations) - 1) // 3 + 2   # the first round whose first agent is past max_iterations
    for t in range(rounds):
        acts = torch.randint(0, 6, (512, 3), generator=g, device=a_env.device, dtype=torch.int32)
        r, d = a_env.step_agents(acts)
        for a in range(3):
            r1, d1 = b_env.step(acts[:, a].contiguous(), agent=a)
            assert torch.equal(r[a], r1) and torch.equal(d[a], d1)
        if t < rounds - 1:
            assert torch.equal(a_env.grids, b_env.grids) and torch.equal(a_env.agent_pos, b_env.agent_pos)
            assert not bool(d.min(dim=0).values.any())
    assert bool(d.min(dim=0).values.all())
    assert int(a_env.iteration.max()) == 0 and int(b_env.iteration.min()) == rounds * 3    # a_env restarted
    sd = a_env.state_dict()
    assert "agent_pos" in sd and tuple(sd["agent_pos"].shape) == (3, 512, 3)
