#!/bin/bash
# round-2 GPU run #12 (2 GPUs): the driver's launch shape -- torchrun, NCCL, NUMA binding per rank
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_run12_topo.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_run12_bench_2gpu.log 2> gpurun_out/r02_run12_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_run12_ref_2gpu.log 2> gpurun_out/r02_run12_ref_2gpu.err
tail -c 600 gpurun_out/r02_run12_bench_2gpu.err
