#!/bin/bash
# round-2 GPU run #11: k_chain_t with the 12-step statically renamed body; full GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run11_pytest.log
for c in 3 4; do
  timeout 300 python bench.py --config $c --steps 5 > gpurun_out/r02_run11_bench_c$c.log 2> gpurun_out/r02_run11_bench_c$c.err
  timeout 900 ncu --set full --clock-control none -k regex:'k_prep|k_chain|k_products|k_passes|k_energy' -s 15 -c 5 -o gpurun_out/r02_run11_cfg$c python bench.py --config $c --steps 1 --warmup 3 --no-e2e > gpurun_out/r02_run11_ncu_c$c.log 2>&1
  python profiles/summarize_ncu.py "r02 run11 config $c (step-major path, final)" "" gpurun_out/r02_run11_cfg$c.ncu-rep > gpurun_out/r02_run11_cfg${c}_summary.md 2>&1
  rm -f gpurun_out/r02_run11_cfg$c.ncu-rep
done
du -sh gpurun_out
