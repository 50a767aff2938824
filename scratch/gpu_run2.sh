#!/bin/bash
# round-2 GPU run #2: the warp-specialised chain kernel -- tests, A/B bench, sanitizer, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run2_pytest.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r02_run2_bench_ws.log 2> gpurun_out/r02_run2_bench_ws.err
B200S_CHAIN_V=4 timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r02_run2_bench_v4.log 2> gpurun_out/r02_run2_bench_v4.err
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_run2_racecheck.log python __graft_entry__.py smoke > gpurun_out/r02_run2_racecheck.out 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_run2_racecheck.out
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_run2_memcheck.log python __graft_entry__.py smoke > gpurun_out/r02_run2_memcheck.out 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_run2_memcheck.out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_chain_ws' -s 3 -c 1 -o gpurun_out/r02_run2_ws python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/r02_run2_ncu.log 2>&1
python profiles/summarize_ncu.py "r02 run2 k_chain_ws" "" gpurun_out/r02_run2_ws.ncu-rep > gpurun_out/r02_run2_ws_summary.md 2>&1
ncu -i gpurun_out/r02_run2_ws.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_run2_ws_source.csv 2>/dev/null
python profiles/ncu_source_hot.py gpurun_out/r02_run2_ws_source.csv 60 > gpurun_out/r02_run2_ws_hot.txt 2>&1
ls -la gpurun_out
du -sh gpurun_out
