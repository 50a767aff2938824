#!/bin/bash
# GPU call G: k_chain_direct6 with 4-step chunks and fills three chunks ahead; A/B against generation 4, memory probes, mono pairs
O=gpurun_out/r03g
mkdir -p $O
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs > $O/$n.json 2> $O/$n.err; }
run hl_v4 B200S_CHAIN_V=4
run hl_v6 B200S_CHAIN_V=6
for p in 6 7; do run hl_v6_probe$p B200S_CHAIN_V=6 B200S_CHAIN_PROBE=$p; done
run hl_v6_b B200S_CHAIN_V=6
for d in 0 1; do
  B200S_DUAL=$d timeout 300 python bench.py --config 5 --steps 10 --sweep-filter "5/4" > $O/sw_dual$d.jsonl 2> $O/sw_dual$d.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r03g/hl_*.json')) + sorted(glob.glob('gpurun_out/r03g/sw_*.jsonl')):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split('/')[-1], d['config'].get('preset',''), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
PY
timeout 600 python -m pytest tests -m gpu -x -q -k "generations or mono_stream or benchmark_shape or golden" > $O/pytest_subset.log 2>&1
tail -4 $O/pytest_subset.log
du -sh gpurun_out
