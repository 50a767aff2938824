#!/bin/bash
# GPU call I (final artefacts of the session): GPU test suite, smoke, default bench + reference arm, sweep, ncu launch list
O=gpurun_out/r03i
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 700 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 600 python bench.py --config 5 --steps 10 > $O/sweep.jsonl 2> $O/sweep.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > $O/ncu_launches.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03i/bench_default.json').read().strip().splitlines()[-1])
print('default', round(d['ms_per_step'],3), d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], 'e2e', d['e2e']['value'], {k:(round(v['ms_per_step'],2), round(v['roofline_frac'],3)) for k,v in d.get('other_configs',{}).items()})
PY
du -sh gpurun_out
