"""GPU parity tests for the search-based problems (minecraft_3D_maze, sokoban, smb): CUDA path through the
C ABI vs the CPU oracle on fresh random grids, plus size-independent properties at BASELINE sizes.

Bit-exact: grids, stats, done flags, counters.  Rewards: rel 1e-6 (BASELINE.json).
(The reference-generated fixtures and traces for these problems run in tests/test_gpu_parity.py.)
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(problem, rep, map_shape, n, **kw):
    import control_pcgrl_b200 as P
    cfg = P.make_config(problem, rep, map_shape=map_shape, **{k: v for k, v in kw.items()
                                                              if k in ("obs_window", "weights", "controls",
                                                                       "max_board_scans", "change_percentage")})
    extra = {k: v for k, v in kw.items() if k in ("action_kind", "auto_reset", "seed", "env_offset",
                                                  "random_init_probs")}
    return P.BatchedPcgrlEnv(cfg, n, **extra)


def _floors(rng, size, wall_p, hole_p):
    g = (rng.random((size,) * 3) < wall_p).astype(np.int8)
    for z in range(3, size, 3):
        g[z] = (rng.random((size, size)) >= hole_p).astype(np.int8)
    return g


def test_maze3d_random_grids_vs_oracle():
    from oracle import pcgrl_oracle as O
    rng = np.random.default_rng(77)
    for size, n in [(14, 120), (9, 120), (5, 80), (16, 40), (3, 40)]:
        grids = np.stack([(rng.random((size,) * 3) < rng.choice([0.1, 0.3, 0.5, 0.7])).astype(np.int8)
                          if i % 3 else _floors(rng, size, rng.choice([0.05, 0.2, 0.35]), rng.choice([0.05, 0.2]))
                          for i in range(n)])
        env = _mk("minecraft_3D_maze", "narrow", (size,) * 3, 1)
        got = env.compute_stats(grids).cpu().numpy()
        env.check_status()
        for i in range(n):
            want = O.stats_vector("minecraft_3D_maze", O.get_stats("minecraft_3D_maze", grids[i]))
            assert got[i].tolist() == want, (size, i, got[i].tolist(), want)


def test_maze3d_full_size_properties():
    """BASELINE config #5 size (14^3, 8192 envs on this GPU): size-independent properties."""
    n = 8192
    env = _mk("minecraft_3D_maze", "narrow", (14, 14, 14), n, seed=5, random_init_probs=True)
    env.reset()
    assert torch.equal(env.compute_stats(env.maps), env.stats)
    g = torch.Generator(device=env.device).manual_seed(0)
    prev_stats, prev_maps = env.stats.clone(), env.maps.clone()
    for t in range(12):
        a = torch.randint(0, 2, (n,), generator=g, device=env.device, dtype=torch.int32)
        reward, done = env.step(a)
        ch = env.changed.bool()
        assert torch.equal(env.stats[~ch], prev_stats[~ch])
        assert float(reward[~ch].abs().max()) == 0.0
        diff = (env.maps != prev_maps).flatten(1).sum(1)
        assert torch.equal(diff, ch.long())
        assert not bool(done.any())
        prev_stats, prev_maps = env.stats.clone(), env.maps.clone()
    assert torch.equal(env.compute_stats(env.maps), env.stats)
    env.check_status()
    # a sample of the batch against the CPU oracle
    from oracle import pcgrl_oracle as O
    maps = env.maps[:64].cpu().numpy()
    st = env.stats[:64].cpu().numpy()
    for i in range(64):
        assert st[i].tolist() == O.stats_vector("minecraft_3D_maze", O.get_stats("minecraft_3D_maze", maps[i]))
