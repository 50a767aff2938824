#!/bin/bash
# round-2 GPU run #6: full suite, headline with the forked commit, configs 3/4, the config-5 sweep, the live loop
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run6_pytest.log
timeout 300 python bench.py --steps 20 > gpurun_out/r02_run6_bench.log 2> gpurun_out/r02_run6_bench.err
for c in 3 4; do
  timeout 300 python bench.py --config $c --steps 5 > gpurun_out/r02_run6_bench_c$c.log 2> gpurun_out/r02_run6_bench_c$c.err
done
timeout 900 python bench.py --config 5 --steps 5 > gpurun_out/r02_run6_sweep.jsonl 2> gpurun_out/r02_run6_sweep.err
timeout 300 python bench.py --live --steps 10 > gpurun_out/r02_run6_live.log 2> gpurun_out/r02_run6_live.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02_run6_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/r02_run6_launches.log 2>&1
du -sh gpurun_out
