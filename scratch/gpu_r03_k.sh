#!/bin/bash
# GPU call K: generation 7 (k_chain_direct6 + a memory warp per stream) against generation 4; whole GPU suite with 7 as the default
O=gpurun_out/r03k
mkdir -p $O
run() { n=$1; shift; env "$@" timeout 120 python bench.py --steps 20 --no-cpu-baseline --no-other-configs > $O/$n.json 2> $O/$n.err; }
run hl_v7 B200S_CHAIN_V=7
run hl_v4 B200S_CHAIN_V=4
run hl_v7_b B200S_CHAIN_V=7
B200S_DUAL=2 timeout 120 python bench.py --config 5 --steps 10 --sweep-filter "5/4" > $O/sw_dual2.jsonl 2> $O/sw_dual2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r03k/hl_*.json')) + sorted(glob.glob('gpurun_out/r03k/sw_*.jsonl')):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split('/')[-1], d['config'].get('preset',''), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],4))
PY
B200S_CHAIN_V=7 timeout 200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_gen7.log 2>&1
tail -4 $O/pytest_gpu_gen7.log
