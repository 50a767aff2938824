#!/bin/bash
# GPU call B: bulk L2 prefetch in k_chain_direct6 (A/B over the piece size), stereo headline + one mono sweep point,
# then the whole GPU test suite, then raw ncu metrics (CSV, no report files: gpurun_out must stay small)
O=gpurun_out/r03b
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
for pf in 0 16 32 64 128; do
  B200S_CHAIN_V=6 B200S_L2PF=$pf timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs > $O/hl_v6_pf$pf.json 2> $O/hl_v6_pf$pf.err
done
B200S_CHAIN_V=4 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs > $O/hl_v4.json 2> $O/hl_v4.err
for pf in 0 32 64 128; do
  B200S_L2PF=$pf timeout 300 python bench.py --config 5 --steps 10 --sweep-filter "5/4" > $O/sw_dual_pf$pf.jsonl 2> $O/sw_dual_pf$pf.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r03b/hl_*.json')) + sorted(glob.glob('gpurun_out/r03b/sw_*.jsonl')):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split('/')[-1], d['config'].get('preset',''), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
PY
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,dram__sectors_read.sum,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__average_warp_latency_issue_stalled_lg_throttle.ratio,smsp__average_warp_latency_issue_stalled_mio_throttle.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio,smsp__average_warp_latency_issue_stalled_drain.ratio,smsp__average_warp_latency_issue_stalled_membar.ratio,smsp__average_warp_latency_issue_stalled_sleeping.ratio
for pf in 0 32; do
  B200S_CHAIN_V=6 B200S_L2PF=$pf timeout 300 ncu --metrics $M --clock-control none -k regex:'k_chain_direct6|k_analyse2|k_synth2' -c 3 --csv --log-file $O/ncu_hl_pf$pf.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > $O/ncu_hl_pf$pf.log 2>&1
  B200S_L2PF=$pf timeout 300 ncu --metrics $M --clock-control none -k regex:'k_chain_direct6' -c 1 --csv --log-file $O/ncu_dual_pf$pf.csv python bench.py --config 5 --steps 1 --warmup 1 --sweep-filter "presetDefault:5/4" > $O/ncu_dual_pf$pf.log 2>&1
done
du -sh gpurun_out
