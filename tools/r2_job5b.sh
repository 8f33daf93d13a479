#!/bin/bash
# 2 GPUs: BASELINE config 5 through NCCL scatter -> solve -> gather (strong scaling), exactly as the driver launches it
O=gpurun_out/r2; mkdir -p $O
nvidia-smi topo -m > $O/topo_2gpu.txt 2>&1
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 1500 $O/bench_n2.err; tail -c 3000 $O/bench_n2.json
