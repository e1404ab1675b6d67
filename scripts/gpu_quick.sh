#!/bin/bash
# quick GPU check: the whole parity suite (+ optional extra command)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_quick.txt
for wl in "binary narrow 64 64" "zelda turtle 64 64"; do
python - $wl <<'PY'
import sys, time, torch
sys.path.insert(0, ".")
import control_pcgrl_b200 as P
prob, rep, h, w = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
n = 65536
env = P.BatchedPcgrlEnv(P.make_config(prob, rep, map_shape=(h, w)), n, auto_reset=True)
env.reset()
n_act = env.n_tiles + (4 if rep == "turtle" else 0)
acts = torch.randint(0, n_act, (60, n), device=env.device, dtype=torch.int32)
for t in range(10): env.step(acts[t])
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(10, 60): env.step(acts[t])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"{prob}-{rep} {h}x{w}: {n} envs, {ms:.3f} ms/step, {n / ms * 1e3:.4g} env-steps/s")
PY
done | tee gpurun_out/r02_bigboard_rates.txt
