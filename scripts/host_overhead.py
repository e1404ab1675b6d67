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
    buf = env.host_action_buffer(None)
    buf.numpy()[...] = np.random.default_rng(0).integers(0, 2, size=buf.numpy().shape).astype(buf.numpy().dtype)
    for _ in range(30):
        env.step_host(buf)
    K = 300
    t0 = time.perf_counter()
    for _ in range(K):
        env.step_host(buf)
    t_api = (time.perf_counter() - t0) / K * 1e6
    h = env._host_io()
    st, cc, lib = env._st, env._cc, env.lib
    a_ptr, dev_ptr, rec_ptr, stream = buf.data_ptr(), h.act_dev.data_ptr(), h.rec.data_ptr(), env._stream()
    t0 = time.perf_counter()
    for _ in range(K):
        lib.pcgrl_step_host_packed(cc, st, a_ptr, dev_ptr, h.nbytes, rec_ptr, stream)
    t_c = (time.perf_counter() - t0) / K * 1e6
    a_dev = torch.zeros(n, dtype=torch.uint8, device=env.device)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        env.step(a_dev)
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / K * 1e3
    print("%s-%s %d envs: step_host %.1f us per call, bare C call %.1f us, device step %.1f us" % (problem, rep, n, t_api, t_c, t_dev))
