#!/bin/bash
# GPU call F: memory-side ablation probes of k_chain_direct6 (profiling build): no prev-input fetch / no write-back / no fetch at all
O=gpurun_out/r03f
mkdir -p $O
run() { n=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs > $O/$n.json 2> $O/$n.err; }
run hl_v6 B200S_CHAIN_V=6
for p in 5 6 7; do run hl_v6_probe$p B200S_CHAIN_V=6 B200S_CHAIN_PROBE=$p; done
for b in 512; do
  for p in 0 7; do B200S_CHAIN_V=6 B200S_CHAIN_PROBE=$p timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-other-configs --batch $b > $O/hl_v6_batch${b}_probe$p.json 2> $O/hl_v6_batch${b}_probe$p.err; done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r03f/hl_*.json')):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split('/')[-1], round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
PY
du -sh gpurun_out
