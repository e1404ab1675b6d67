#!/usr/bin/env python
"""Device time of one pcgrl_step per step path for any problem / representation / map shape:
    python scripts/bench_paths.py PROBLEM REP HxW ENVS [PATH ...]        (PATH: fused split inc incfused lg)
Random actions, auto-reset, CUDA events around 200 steps after 20 warm ones, 256 MB L2 flush between steps
(excluded).  One line per path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import control_pcgrl_b200 as P  # noqa: E402

problem, rep, shape, n = sys.argv[1], sys.argv[2], tuple(int(x) for x in sys.argv[3].split("x")), int(sys.argv[4])
paths = sys.argv[5:] or ["fused", "split"]
for path in paths:
    os.environ["PCGRL_STEP_PATH"] = path
    kw = dict(obs_window=shape) if rep == "wide" else {}
    env = P.BatchedPcgrlEnv(P.make_config(problem, rep, map_shape=shape, **kw), n, seed=1, auto_reset=True)
    env.reset()
    n_act = {"narrow": env.n_tiles, "turtle": 4 + env.n_tiles, "wide": shape[0] * shape[1] * env.n_tiles}[rep]
    g = torch.Generator(device=env.device).manual_seed(0)
    acts = torch.randint(0, n_act, (220, n), generator=g, device=env.device, dtype=torch.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)
    for t in range(20):
        env.step(acts[t])
    torch.cuda.synchronize()
    tot = 0.0
    evs = []
    for t in range(200):
        flush.fill_(t & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.step(acts[20 + t])
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    env.check_status()
    print(f"{problem}-{rep} {shape} envs={n} path={path}: {ms:.4f} ms per step, {n / ms * 1e3:.4g} env-steps/s", flush=True)
    del env, acts, flush
    torch.cuda.empty_cache()
