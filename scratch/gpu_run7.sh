#!/bin/bash
# round-2 GPU run #7: live loop with the device bank, launch list, source-level profiles of the two FFT kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "live or api_sequence or golden" > gpurun_out/r02_run7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run7_pytest.log
timeout 300 python bench.py --live --steps 20 > gpurun_out/r02_run7_live.log 2> gpurun_out/r02_run7_live.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 40 --csv --log-file gpurun_out/r02_run7_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e > gpurun_out/r02_run7_launches.log 2>&1
for k in k_analyse2 k_synth2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/r02_run7_$k python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/r02_run7_ncu_$k.log 2>&1
  python profiles/summarize_ncu.py "r02 run7 $k" "" gpurun_out/r02_run7_$k.ncu-rep > gpurun_out/r02_run7_${k}_summary.md 2>&1
  ncu -i gpurun_out/r02_run7_$k.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_run7_${k}_source.csv 2>/dev/null
  python profiles/ncu_source_hot.py gpurun_out/r02_run7_${k}_source.csv 70 > gpurun_out/r02_run7_${k}_hot.txt 2>&1
  rm -f gpurun_out/r02_run7_${k}_source.csv gpurun_out/r02_run7_$k.ncu-rep
done
du -sh gpurun_out
