#!/bin/bash
# TEST / BENCH INFRASTRUCTURE -- installs the UNMODIFIED reference (smearle/control-pcgrl) into baseline/_ref so the
# CPU arm of bench.py (--impl reference, cpu_baseline) can run the reference's own code on the GPU box, where
# /root/reference does not exist.  baseline/_ref is git-ignored (no reference source enters the history) but not
# gpurun-ignored, so it travels with the repo snapshot.
#
# The reference's setup.py lists only the top-level package (find_namespace_packages(include=["hydra_plugins.*",
# "control_pcgrl"])): it is meant for `pip install -e .`, and a --target install of it ships no
# control_pcgrl/envs at all.  So the install runs from a copy under /tmp (the reference tree is read-only) whose
# setup.py package list -- packaging metadata, not code -- is widened to "control_pcgrl.*"; every .py file is
# byte-identical to the reference's.  Dependencies (gymnasium, ray, hydra) are not in the offline wheelhouse:
# --no-deps, and oracle/refshim.py supplies the fake third-party modules at import time.
set -eu
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d /tmp/refcopy.XXXX)
cp -r "$REF/setup.py" "$REF/README.md" "$REF/bin" "$REF/control_pcgrl" "$REF/hydra_plugins" "$TMP/"
sed -i 's/include=\["hydra_plugins\.\*", "control_pcgrl"\]/include=["hydra_plugins.*", "control_pcgrl", "control_pcgrl.*"]/' "$TMP/setup.py"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP"
rm -rf "$TMP"
# the install must carry the files of the step path, byte-identical
for f in envs/pcgrl_env.py envs/helper.py envs/helper_3D.py envs/reps/narrow_rep.py envs/probs/binary/binary_prob.py \
         control_wrappers.py wrappers.py envs/probs/sokoban/sokoban/engine.py envs/probs/smb/smb/engine.py; do
  cmp -s "$REF/control_pcgrl/$f" "$ROOT/baseline/_ref/control_pcgrl/$f" || { echo "MISSING/DIFFERENT: $f"; exit 1; }
done
echo "reference installed into baseline/_ref ($(find "$ROOT/baseline/_ref" -name '*.py' | wc -l) python files)"
