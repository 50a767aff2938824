#!/bin/bash
# round-2 GPU run #3: ws chain with the producer's interior path; new parity tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run3_pytest.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r02_run3_bench_ws.log 2> gpurun_out/r02_run3_bench_ws.err
B200S_CHAIN_V=4 timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r02_run3_bench_v4.log 2> gpurun_out/r02_run3_bench_v4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_chain_ws' -s 3 -c 1 -o gpurun_out/r02_run3_ws python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/r02_run3_ncu.log 2>&1
python profiles/summarize_ncu.py "r02 run3 k_chain_ws (interior producer path)" "" gpurun_out/r02_run3_ws.ncu-rep > gpurun_out/r02_run3_ws_summary.md 2>&1
ncu -i gpurun_out/r02_run3_ws.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_run3_ws_source.csv 2>/dev/null
python profiles/ncu_source_hot.py gpurun_out/r02_run3_ws_source.csv 60 > gpurun_out/r02_run3_ws_hot.txt 2>&1
rm -f gpurun_out/r02_run3_ws_source.csv
du -sh gpurun_out
