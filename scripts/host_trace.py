#!/usr/bin/env python
"""Device timeline of the pipelined host step (the stand-in for an nsys trace of the host leg):
    PCGRL_HOST_TRACE=3 [PCGRL_HOST_CHUNKS=c] python scripts/host_trace.py [--envs N]
runs a few warm-up calls of BatchedPcgrlEnv(compact_host_io=True).step_host on binary-narrow 16x16 and lets the
library print, for the traced calls, when every chunk's upload / kernels / download finished (microseconds from the
fork on the caller's stream), next to the wall time of the call."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import control_pcgrl_b200 as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
a = ap.parse_args()
n_trace = int(os.environ.get("PCGRL_HOST_TRACE", "0"))   # the library traces calls 41 .. 40 + n
env = P.BatchedPcgrlEnv(P.make_config("binary", "narrow"), a.envs, compact_host_io=True, auto_reset=True)
env.reset()
rng = np.random.default_rng(0)
bufs = []
for _ in range(4):
    b = env.host_action_buffer(None)
    b.numpy()[...] = rng.integers(0, 2, size=a.envs).astype(np.uint8)
    bufs.append(b)
for i in range(40 + n_trace + 10):
    t0 = time.perf_counter()
    env.step_host(bufs[i % 4])
    dt = (time.perf_counter() - t0) * 1e6
    if i >= 30:
        print("call %d: %.1f us wall (python side)" % (i + 1, dt), file=sys.stderr)
