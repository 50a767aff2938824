#!/bin/bash
# GPU call A (round 3 session): A/B of chain generation 6 vs 4, mono pairs on/off, GPU parity subset, ncu of the new kernels
mkdir -p gpurun_out/r03a
O=gpurun_out/r03a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
for v in 4 6 4 6; do
  B200S_CHAIN_V=$v timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs >> $O/ab_chain_v$v.jsonl 2>> $O/ab_chain_v$v.err
done
timeout 600 python bench.py --config 5 --steps 10 > $O/sweep_dual1.jsonl 2> $O/sweep_dual1.err
B200S_DUAL=0 timeout 300 python bench.py --config 5 --steps 10 --sweep-filter "5/4" > $O/sweep_dual0.jsonl 2> $O/sweep_dual0.err
timeout 900 python -m pytest tests -m gpu -x -q -k "generations or mono_stream or free_run or golden or teacher or benchmark_shape or cheaper" > $O/pytest_subset.log 2>&1
echo "pytest rc=$?" >> $O/pytest_subset.log
B200S_CHAIN_V=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_direct6 -c 1 -o $O/k6_stereo python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > $O/ncu_k6_stereo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_direct6 -c 1 -o $O/k6_dual python bench.py --config 5 --steps 1 --warmup 1 --sweep-filter "presetDefault:5/4" > $O/ncu_k6_dual.log 2>&1
ls -la $O
tail -3 $O/pytest_subset.log
for f in $O/ab_chain_v4.jsonl $O/ab_chain_v6.jsonl; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: continue
    print(sys.argv[1], round(d['ms_per_step'],3), d['roofline']['kernel_ms_per_step'], 'e2e', round(d['e2e']['ms_per_step'],2))
PY
done
python - <<'PY'
import json
for f in ('gpurun_out/r03a/sweep_dual1.jsonl','gpurun_out/r03a/sweep_dual0.jsonl'):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f[-17:], d['config']['preset'], d['config']['in_over_out'], round(d['ms_per_step'],2), round(d['roofline']['frac'],3), {k: round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})
PY
