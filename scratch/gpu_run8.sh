#!/bin/bash
# round-2 GPU run #8: step-major path for the mapped / formant configurations
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run8_pytest.log
for c in 3 4; do
  timeout 300 python bench.py --config $c --steps 5 > gpurun_out/r02_run8_bench_c$c.log 2> gpurun_out/r02_run8_bench_c$c.err
  timeout 900 ncu --set full --clock-control none -k regex:'k_prep|k_chain|k_passes|k_energy|k_products' -s 18 -c 6 -o gpurun_out/r02_run8_cfg$c python bench.py --config $c --steps 1 --warmup 3 --no-e2e > gpurun_out/r02_run8_ncu_c$c.log 2>&1
  python profiles/summarize_ncu.py "r02 run8 config $c (step-major path: k_energy, k_passes, k_prep map-only, k_products, k_chain_t)" "" gpurun_out/r02_run8_cfg$c.ncu-rep > gpurun_out/r02_run8_cfg${c}_summary.md 2>&1
  rm -f gpurun_out/r02_run8_cfg$c.ncu-rep
done
timeout 400 python bench.py --steps 20 > gpurun_out/r02_run8_bench.log 2> gpurun_out/r02_run8_bench.err
du -sh gpurun_out
