#!/bin/bash
# round-2 GPU run #13: final validation of the tree (tests, smoke, default bench, reference arm) + the sweep with the fast mono chain
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run13_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_run13_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_run13_bench_ref.log 2> gpurun_out/r02_run13_bench_ref.err
timeout 600 python bench.py > gpurun_out/r02_run13_bench.log 2> gpurun_out/r02_run13_bench.err
timeout 900 python bench.py --config 5 --steps 5 > gpurun_out/r02_run13_sweep.jsonl 2> gpurun_out/r02_run13_sweep.err
timeout 300 python bench.py --live --steps 20 > gpurun_out/r02_run13_live.log 2> gpurun_out/r02_run13_live.err
du -sh gpurun_out
