#!/bin/bash
# round-2 GPU run #10: step-major path, second iteration (hoisted row pointers, two rows ahead, branch-free exact div / sqrt)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "config3 or config4 or formant or freq_map or golden or teacher or benchmark_shape or random" > gpurun_out/r02_run10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run10_pytest.log
for c in 3 4; do
  timeout 300 python bench.py --config $c --steps 5 > gpurun_out/r02_run10_bench_c$c.log 2> gpurun_out/r02_run10_bench_c$c.err
  timeout 900 ncu --set full --clock-control none -k regex:'k_prep|k_chain|k_products' -s 9 -c 3 -o gpurun_out/r02_run10_cfg$c python bench.py --config $c --steps 1 --warmup 3 --no-e2e > gpurun_out/r02_run10_ncu_c$c.log 2>&1
  python profiles/summarize_ncu.py "r02 run9 config $c (step-major path v2)" "" gpurun_out/r02_run10_cfg$c.ncu-rep > gpurun_out/r02_run10_cfg${c}_summary.md 2>&1
  rm -f gpurun_out/r02_run10_cfg$c.ncu-rep
done
du -sh gpurun_out
