#!/bin/bash
O=gpurun_out/r03j
mkdir -p $O
timeout 240 compute-sanitizer --tool racecheck python scratch/smoke_gen6.py > $O/sanitizer_racecheck_gen6.log 2>&1; tail -5 $O/sanitizer_racecheck_gen6.log
timeout 200 compute-sanitizer --tool memcheck python scratch/smoke_gen6.py > $O/sanitizer_memcheck_gen6.log 2>&1; tail -5 $O/sanitizer_memcheck_gen6.log
