#!/bin/bash
# round-2 GPU run #1 (baseline of the round-1 code): tests, sanitizers, batch-size experiment, config 3/4 captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_run1_smi.log 2>&1
lscpu | head -30 > gpurun_out/r02_run1_lscpu.log 2>&1
nproc >> gpurun_out/r02_run1_lscpu.log; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02_run1_lscpu.log 2>&1
python -c "import os; print('affinity', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())" >> gpurun_out/r02_run1_lscpu.log 2>&1
nvidia-smi topo -m >> gpurun_out/r02_run1_lscpu.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
for b in 256 512 1024; do
  timeout 300 python bench.py --batch $b --steps 10 --no-cpu-baseline > gpurun_out/r02_run1_bench_b$b.log 2> gpurun_out/r02_run1_bench_b$b.err
done
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_run1_memcheck.log python __graft_entry__.py smoke > gpurun_out/r02_run1_memcheck.out 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_run1_memcheck.out
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_run1_racecheck.log python __graft_entry__.py smoke > gpurun_out/r02_run1_racecheck.out 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_run1_racecheck.out
for c in 3 4; do
  timeout 300 python bench.py --config $c --steps 5 > gpurun_out/r02_run1_bench_c$c.log 2> gpurun_out/r02_run1_bench_c$c.err
  timeout 900 ncu --set full --clock-control none -k regex:'k_prep|k_chain|k_analyse2|k_synth2|k_pitch' -s 12 -c 4 -o gpurun_out/r02_run1_cfg$c python bench.py --config $c --steps 1 --warmup 3 --no-e2e > gpurun_out/r02_run1_ncu_c$c.log 2>&1
  python profiles/summarize_ncu.py "r02 run1 config $c (round-1 kernels)" "" gpurun_out/r02_run1_cfg$c.ncu-rep > gpurun_out/r02_run1_cfg${c}_summary.md 2>&1
  rm -f gpurun_out/r02_run1_cfg$c.ncu-rep
done
ls -la gpurun_out | tail -30
