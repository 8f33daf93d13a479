#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for v in "" _genws; do SCB_LIB=$PWD/safe_control_b200/libscb$v.so timeout 300 python tools/mpc_variants.py cfg3 du5 kb5 q5 >> $O/mpc_variants_lds.txt 2>&1; done; cat $O/mpc_variants_lds.txt
(time timeout 1700 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.err
(time timeout 600 python bench.py --impl reference --gpus 2 --steps 3 --warmup 1) > $O/bench_ref_cfg5.json 2> $O/bench_ref_cfg5.err; tail -c 900 $O/bench_ref_cfg5.json; tail -4 $O/bench_ref_cfg5.err
