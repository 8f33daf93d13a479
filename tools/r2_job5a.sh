#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
(time timeout 1700 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -12 $O/pytest_gpu.log
timeout 400 bash tools/prof_cfg2_source.sh
(time timeout 500 python bench.py --impl reference --gpus 2 --steps 1 --warmup 1) > $O/bench_ref_cfg5.json 2> $O/bench_ref_cfg5.err; tail -c 600 $O/bench_ref_cfg5.json
