#!/usr/bin/env python
"""Device-time the observation writer (pcgrl_observe: crop + one-hot + target planes) against the measured HBM peak.

    python scripts/bench_observe.py [--envs N] [--reps R]
Prints one JSON line per case: algorithmic bytes = output bytes + grid bytes read + pos bytes; GB/s; fraction of
MEASURED_PEAKS.json's HBM bandwidth.  The output of every case is larger than L2 (126 MB), so no flush is needed.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import control_pcgrl_b200 as P  # noqa: E402

CASES = [  # (problem, rep, map_shape, obs_window, controls, out dtype; "codes" = uint8 tile codes, not one-hot)
    ("binary", "narrow", (16, 16), (32, 32), None, torch.uint8),
    ("binary", "narrow", (16, 16), (32, 32), None, "codes"),
    ("binary", "narrow", (16, 16), (32, 32), None, torch.float32),
    ("binary", "wide", (16, 16), (16, 16), ["regions", "path-length"], torch.float32),
    ("zelda", "turtle", (7, 11), (22, 22), None, torch.uint8),
    ("zelda", "turtle", (7, 11), (22, 22), None, torch.float32),
    ("minecraft_3D_maze", "narrow", (14, 14, 14), (14, 14, 14), None, torch.uint8),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 18)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default="", help="also append the JSON lines to this file")
    a = ap.parse_args()
    peak = 6544.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    for problem, rep, shape, obs, controls, dt in CASES:
        cfg = P.make_config(problem, rep, map_shape=shape, obs_window=obs, controls=controls)
        env = P.BatchedPcgrlEnv(cfg, a.envs, action_kind="wide_flat" if rep == "wide" else None)
        if controls:
            env.sample_uniform_targets()
        env.reset()
        if rep in ("narrow", "turtle"):   # spread the crop centres
            for i, d in enumerate(shape):
                env.pos[:, i] = torch.randint(0, d, (a.envs,), device=env.device, dtype=torch.int32)
        onehot = dt != "codes"
        dt = dt if onehot else torch.uint8
        out = torch.empty((a.envs, *env.obs_shape(onehot)), dtype=dt, device=env.device)
        obs_shape = list(env.obs_shape(onehot))
        _observe = env.observe
        env.observe = lambda out: _observe(out=out, onehot=onehot)
        for _ in range(3):
            env.observe(out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            env.observe(out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        nbytes = out.numel() * out.element_size() + a.envs * (env.row_stride + 12)
        gbs = nbytes / ms / 1e6
        line = json.dumps({"case": f"{problem}-{rep}-{'x'.join(map(str, shape))}", "dtype": str(dt).split('.')[-1] + ("" if onehot else " tile codes"),
                          "controls": bool(controls), "envs": a.envs, "obs_shape": obs_shape,
                          "ms": round(ms, 4), "bytes": nbytes, "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3),
                          "env_obs_per_s": round(a.envs / ms * 1e3)})
        print(line, flush=True)
        if a.out:
            with open(a.out, "a") as f:
                f.write(line + "\n")
        del env, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
