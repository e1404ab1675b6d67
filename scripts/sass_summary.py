#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libpcgrl_sm100.so (cuobjdump -sass), for profiles/: which kernels use 128-bit
global accesses, bulk asynchronous copies (UBLKCP = cp.async.bulk, the TMA engine), warp votes / matches / reductions,
atomics; no tensor-core instruction is expected anywhere on this path."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "control_pcgrl_b200/libpcgrl_sm100.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["LDG.E.128", "STG.E.128", "LDG.E.64", "STG.E.64", "LDS.128", "STS.128", "UBLKCP", "UTMASTG", "UTMALDG", "SYNCS",
        "MATCH", "VOTE", "REDUX", "POPC", "FLO", "SHFL", "ATOMG", "ATOMS", "RED", "BAR.SYNC", "DFMA", "DMUL", "DADD",
        "HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA"]
cur, counts, total, arch = None, collections.OrderedDict(), collections.Counter(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_all"] += 1
    for w in WANT:
        if op.startswith(w):
            counts[cur][w] += 1
            total[w] += 1
print("library:", lib, " cubin archs:", sorted(arch))
print("totals:", {k: v for k, v in total.items()})
print("tensor-core mnemonics (HMMA/IMMA/UTC*MMA):", sum(total[k] for k in ("HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA")))
for fn, c in counts.items():
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
    items = " ".join(f"{k}={v}" for k, v in c.items() if k != "_all")
    print(f"{c['_all']:6d} instr  {name}\n        {items}")
