#!/bin/bash
# round-2 GPU run #5: coalesced serial passes, random-time-factor path, full test-suite, benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run5_pytest.log
for c in 3 4; do
  timeout 300 python bench.py --config $c --steps 5 > gpurun_out/r02_run5_bench_c$c.log 2> gpurun_out/r02_run5_bench_c$c.err
  timeout 900 ncu --set full --clock-control none -k regex:'k_prep|k_chain|k_passes|k_energy' -s 12 -c 4 -o gpurun_out/r02_run5_cfg$c python bench.py --config $c --steps 1 --warmup 3 --no-e2e > gpurun_out/r02_run5_ncu_c$c.log 2>&1
  python profiles/summarize_ncu.py "r02 run5 config $c (k_energy + k_passes (staged) + k_prep)" "" gpurun_out/r02_run5_cfg$c.ncu-rep > gpurun_out/r02_run5_cfg${c}_summary.md 2>&1
  rm -f gpurun_out/r02_run5_cfg$c.ncu-rep
done
timeout 300 python bench.py --steps 20 > gpurun_out/r02_run5_bench.log 2> gpurun_out/r02_run5_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_run5_bench_ref.log 2> gpurun_out/r02_run5_bench_ref.err
du -sh gpurun_out
