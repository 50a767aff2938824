#!/bin/bash
# GPU call H: k_chain_direct6 with channel-interleaved Band::output rows (128-byte write-back pieces) against planar rows and generation 4;
# the whole GPU test suite with generation 6 as the default
O=gpurun_out/r03h
mkdir -p $O
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs > $O/$n.json 2> $O/$n.err; }
run hl_v4 B200S_CHAIN_V=4
run hl_v6_yil1 B200S_CHAIN_V=6 B200S_YIL=1
run hl_v6_yil0 B200S_CHAIN_V=6 B200S_YIL=0
run hl_v6_yil1_b B200S_CHAIN_V=6 B200S_YIL=1
run hl_v4_b B200S_CHAIN_V=4
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r03h/hl_*.json')):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, 'e2e', round(d['e2e']['ms_per_step'],2))
PY
B200S_CHAIN_V=6 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_gen6.log 2>&1
tail -4 $O/pytest_gpu_gen6.log
du -sh gpurun_out
