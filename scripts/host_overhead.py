#!/usr/bin/env python
"""Where a small shard's host step goes: the whole BatchedPcgrlEnv.step_host call against the bare C call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import control_pcgrl_b200 as P  # noqa: E402

for problem, rep, n in (("zelda", "turtle", 1 << 16), ("binary", "narrow", 1 << 16), ("binary", "narrow", 1 << 12)):
    env = P.BatchedPcgrlEnv(P.make_config(problem, rep), n, compact_host_io=True, auto_reset=True)
    env.reset()
    # a pool of pinned action batches (a constant batch stops changing the maps after one board scan, and the two
    # timed loops would see different device work); the two call styles alternate step by step for the same reason
    rng = np.random.default_rng(0)
    n_act = env.n_tiles + (4 if rep == "turtle" else 0)
    bufs = []
    for _ in range(16):
        b = env.host_action_buffer(None)
        b.numpy()[...] = rng.integers(0, n_act, size=b.numpy().shape).astype(b.numpy().dtype)
        bufs.append(b)
    for i in range(30):
        env.step_host(bufs[i % 16])
    h = env._host_io()
    st, cc, lib = env._st, env._cc, env.lib
    dev_ptr, rec_ptr, stream = h.act_dev.data_ptr(), h.rec.data_ptr(), env._stream()
    ptrs = [b.data_ptr() for b in bufs]
    K = 300
    t_api = t_c = 0.0
    for i in range(K):
        t0 = time.perf_counter()
        env.step_host(bufs[i % 16])
        t1 = time.perf_counter()
        lib.pcgrl_step_host_packed(cc, st, ptrs[(i + 8) % 16], dev_ptr, h.nbytes, rec_ptr, stream)
        t2 = time.perf_counter()
        env._after_step()          # (keeps the episode clock / auto-reset in step; not timed)
        t_api += t1 - t0
        t_c += t2 - t1
    t_api, t_c = t_api / K * 1e6, t_c / K * 1e6
    a_dev = torch.from_numpy(bufs[0].numpy().copy()).to(env.device)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        env.step(a_dev)
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / K * 1e3
    print("%s-%s %d envs: step_host %.1f us per call, bare C call %.1f us, device step %.1f us" % (problem, rep, n, t_api, t_c, t_dev))
