"""Time single sokoban get_stats calls on the GPU for the hardest fixture grids (solver latency)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import control_pcgrl_b200 as P
from tests.golden_util import load_stats
_, groups = load_stats("sokoban")
for grids, stats in groups:
    shape = grids.shape[1:]
    env = P.BatchedPcgrlEnv(P.make_config("sokoban", "narrow", map_shape=shape), 1)
    ts = []
    for g in grids:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        env.compute_stats(g[None]); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    order = np.argsort(-ts)[:6]
    print(shape, "n", len(ts), "median ms %.3f" % np.median(ts), "top:", [(round(float(ts[i]), 2), stats[i][4:6].tolist()) for i in order])
