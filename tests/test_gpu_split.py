"""GPU: the three equivalent step paths of the bit-board problems must agree bit for bit.

  fused  k_step_bitboard (one launch: update, searches, reward) -- pinned on the reference fixtures elsewhere
  split  k_split_act / k_split_stats / k_split_out (global work list, from-scratch searches with full warps)
  inc    the same with the incremental binary search from the per-env cache (BinaryIncMachine)
  incfused  k_step_inc: update, incremental search and reward in one launch, every warp on its own env slice

PCGRL_STEP_PATH selects the path per pcgrl_step call.  Besides path-vs-path equality, every few steps the stats of
EVERY env are recomputed from scratch from the current grids (pcgrl_stats), which is what pins the incremental
machine: any drift of its cache would show up as a stat that a fresh computation does not give.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _envs(problem, rep, shape, n, paths, **kw):
    import control_pcgrl_b200 as P
    cfg_kw = {k: v for k, v in kw.items() if k in ("controls", "obs_window", "max_board_scans", "change_percentage",
                                                   "act_window", "static_tile_wrapper", "static_prob", "n_static_walls")}
    cfg = P.make_config(problem, rep, map_shape=shape, **cfg_kw)
    return [P.BatchedPcgrlEnv(cfg, n, seed=11, auto_reset=True, split_step=(p != "fused")) for p in paths]


def _n_act(env, rep, shape):
    return {"narrow": env.n_tiles, "turtle": 4 + env.n_tiles, "wide": shape[0] * shape[1] * env.n_tiles}[rep]


CASES = [
    ("binary", "narrow", (16, 16), None),
    ("binary", "turtle", (16, 16), None),
    ("binary", "wide", (16, 16), ["regions", "path-length"]),
    ("binary", "narrow", (10, 10), None),      # generic staging path (rows narrower than 16)
    ("binary", "narrow", (7, 5), None),        # 4-word boards
    ("binary", "turtle", (3, 4), None),        # 2-word boards
    ("binary", "narrow", (2, 16), None),       # 1-word boards
    ("binary", "narrow", (20, 24), None),      # one row per word: split path, no cache
    ("zelda", "turtle", (7, 11), None),
    ("zelda", "narrow", (7, 11), None),
    ("binary_holey", "narrow", (16, 16), None),
]


@pytest.mark.parametrize("problem,rep,shape,controls", CASES)
def test_step_paths_agree_and_match_fresh_stats(problem, rep, shape, controls, monkeypatch):
    n = 12_345
    kw = dict(controls=controls, max_board_scans=0.6)
    if rep == "wide":
        kw["obs_window"] = shape
    paths = ["fused", "split", "inc", "incfused"]
    envs = _envs(problem, rep, shape, n, paths, **kw)
    has_cache = envs[2].cache is not None
    assert has_cache == (problem == "binary" and max(shape) <= 16)
    assert envs[0].worklist is None and envs[1].worklist is not None
    if controls:
        g = torch.Generator(device=envs[0].device).manual_seed(1)
        envs[0].sample_uniform_targets(generator=g)
        for e in envs[1:]:
            e.targets.copy_(envs[0].targets)
    for e in envs:
        e.reset()
    ref = envs[0]
    for e in envs[1:]:
        assert torch.equal(e.grids, ref.grids) and torch.equal(e.stats, ref.stats)
    n_act = _n_act(ref, rep, shape)
    gen = torch.Generator(device=ref.device).manual_seed(5)
    steps = int(ref.max_iterations * 2.5) + 7
    for t in range(steps):
        act = torch.randint(0, n_act, (n,), generator=gen, device=ref.device, dtype=torch.int32)
        outs = []
        for path, e in zip(paths, envs):
            monkeypatch.setenv("PCGRL_STEP_PATH", path)
            r, d = e.step(act)
            outs.append((r.clone(), d.clone()))
        for (r, d), e, path in zip(outs[1:], envs[1:], paths[1:]):
            assert torch.equal(r, outs[0][0]), (path, t, int((r != outs[0][0]).sum()))
            assert torch.equal(d, outs[0][1]), (path, t)
            assert torch.equal(e.stats, ref.stats), (path, t, int((e.stats != ref.stats).any(dim=1).sum()))
            assert torch.equal(e.grids, ref.grids) and torch.equal(e.pos, ref.pos), (path, t)
            assert torch.equal(e.changes, ref.changes) and torch.equal(e.iteration, ref.iteration), (path, t)
            assert torch.equal(e.changed, ref.changed), (path, t)
        if t % 16 == 5 or t == steps - 1:
            for e in envs[2:]:
                fresh = e.compute_stats(e.maps, holes=e.holes) if e.holey else e.compute_stats(e.maps)
                assert torch.equal(fresh, e.stats), (t, int((fresh != e.stats).any(dim=1).sum()))
    for e in envs:
        e.check_status()


def test_incremental_cache_is_consistent_with_the_grids(monkeypatch):
    """The cache row of every env holds the passable board of its CURRENT grid after resets and after steps, and a
    shard restored from state_dict continues bit-identically (the cache is part of the state)."""
    import control_pcgrl_b200 as P
    n = 4096
    cfg = P.make_config("binary", "narrow", map_shape=(16, 16), max_board_scans=0.3)
    env = P.BatchedPcgrlEnv(cfg, n, seed=2, auto_reset=True)
    assert env.cache is not None and env.cache_stride == 80
    env.reset()

    def board_of(maps):
        m = (maps.reshape(n, 256) == 0).to(torch.int64)              # tile 0 = empty = passable
        w = (m.reshape(n, 8, 32) << torch.arange(32, device=m.device)).sum(dim=2)   # bit y*16+x of word y//2
        return w.to(torch.int64)

    def cached_board():
        c = env.cache.view(torch.int32).reshape(n, 20)[:, :8].to(torch.int64) & 0xFFFFFFFF
        return c

    assert torch.equal(cached_board(), board_of(env.maps))
    gen = torch.Generator(device=env.device).manual_seed(0)
    for t in range(200):
        env.step(torch.randint(0, 2, (n,), generator=gen, device=env.device, dtype=torch.int32))
        if t % 25 == 0:
            assert torch.equal(cached_board(), board_of(env.maps)), t
    sd = env.state_dict()
    twin = P.BatchedPcgrlEnv(cfg, n, seed=99, auto_reset=True)
    twin.load_state_dict(sd)
    for t in range(120):
        a = torch.randint(0, 2, (n,), generator=gen, device=env.device, dtype=torch.int32)
        r0, _ = env.step(a)
        r1, _ = twin.step(a)
        assert torch.equal(r0, r1) and torch.equal(env.stats, twin.stats) and torch.equal(env.grids, twin.grids), t


def test_incremental_hand_built_maps():
    """Edits that merge, split, create and delete components on hand-built maps, including the component that holds
    the longest path (forces the re-sweep of the untouched components) -- each step checked against a fresh
    pcgrl_stats of the edited map."""
    import control_pcgrl_b200 as P
    cfg = P.make_config("binary", "wide", map_shape=(16, 16), obs_window=(16, 16))
    env = P.BatchedPcgrlEnv(cfg, 1, action_kind="wide_coords")
    g = np.ones((16, 16), dtype=np.int8)
    g[0, :] = 0           # a long corridor (path 15)
    g[2, 0:6] = 0         # a shorter one (path 5)
    g[4, 3] = 0           # an isolated cell
    g[6:9, 6:9] = 0       # a 3x3 room
    env.reset(grids=g[None])
    edits = [(1, 0, 0),    # join corridor 1 and corridor 2 through (1, 0)
             (0, 7, 1),    # cut the long corridor in two
             (0, 7, 0),    # and restore it
             (1, 0, 1),    # separate the two corridors again
             (4, 3, 1),    # delete the isolated cell
             (4, 3, 0),    # and bring it back
             (7, 7, 1),    # punch the room's centre (a ring)
             (0, 0, 1), (0, 1, 1), (0, 2, 1),   # eat the longest corridor from its first tile
             (15, 15, 0), (15, 14, 0), (14, 15, 0),  # grow a new component in the corner
             (5, 3, 0), (3, 3, 0),      # connect the isolated cell to corridor 2 and beyond
             (2, 3, 1)]                 # split corridor 2 (now a T) at the junction
    for y, x, v in edits:
        a = torch.tensor([[y, x, v]], dtype=torch.int32, device=env.device)
        env.step(a)
        fresh = env.compute_stats(env.maps)
        assert torch.equal(fresh, env.stats), ((y, x, v), fresh.tolist(), env.stats.tolist())
    env.check_status()
