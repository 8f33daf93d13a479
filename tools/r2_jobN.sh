#!/bin/bash
# N GPUs: BASELINE config 5 through NCCL scatter -> solve -> gather (strong scaling), exactly as the driver launches it
N=${1:-4}; O=gpurun_out/r2; mkdir -p $O
nvidia-smi topo -m > $O/topo_${N}gpu.txt 2>&1
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3) > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -c 600 $O/bench_n$N.err; python - <<PY
import json
d = json.loads(open("$O/bench_n$N.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value %.4g" % d["value"], "ms %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["config"]["phases_ms"], "eff", d.get("strong_scaling_efficiency_vs_base"))
PY
