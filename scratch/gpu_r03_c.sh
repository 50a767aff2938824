#!/bin/bash
# GPU call C: whole GPU test suite, steady-state ncu of the stereo chain kernels (generations 4 and 6; text exports only),
# final default bench + reference arm + sweep
O=gpurun_out/r03c
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
for v in 4 6; do
  B200S_CHAIN_V=$v timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_chain_direct -s 3 -c 1 -f -o /tmp/chain_v$v python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > $O/ncu_chain_v$v.log 2>&1
  ncu -i /tmp/chain_v$v.ncu-rep --page raw --csv > $O/ncu_chain_v${v}_raw.csv 2>/dev/null
  ncu -i /tmp/chain_v$v.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_v$v.csv 2>/dev/null
  python profiles/ncu_source_hot.py /tmp/src_v$v.csv > $O/ncu_chain_v${v}_hot.txt 2>&1 || cp /tmp/src_v$v.csv $O/ncu_chain_v${v}_source.csv
  ls -la /tmp/chain_v$v.ncu-rep /tmp/src_v$v.csv
done
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 600 python bench.py --config 5 --steps 10 > $O/sweep.jsonl 2> $O/sweep.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03c/bench_default.json').read().strip().splitlines()[-1])
print('default', round(d['ms_per_step'],3), d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], 'e2e', d['e2e']['value'], d.get('other_configs',{}).keys())
PY
du -sh gpurun_out
