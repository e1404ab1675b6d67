#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('headline value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
for k,v in d['configs'].items(): print('   ', k, 'value %.4g e2e %.4g' % (v['value'], v['e2e']['value']))"
for n in 131072 262144 524288; do for c in 1 2 3 4; do
PCGRL_HOST_CHUNKS=$c timeout 200 python bench.py --envs $n --steps 200 --warmup 10 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('envs $n chunks $c: e2e %.4g value %.4g' % (d['e2e']['value'], d['value']))"
done; done | tee gpurun_out/r02_e2e_chunks_by_size.txt
tail -3 gpurun_out/ab.err
