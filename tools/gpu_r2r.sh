#!/bin/bash
# 8-GPU box: the driver's scaling run, N = 8 (and N = 4) under torchrun
mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 100 --cpu-seconds 2 2> gpurun_out/r2r_b$n.err | tail -1 > gpurun_out/r2r_bench_${n}gpu.json; tail -2 gpurun_out/r2r_b$n.err
done
python - <<PY
import json
for n in (8,4):
    d=json.loads(open("gpurun_out/r2r_bench_%dgpu.json"%n).read().strip().splitlines()[-1])
    print(n, "value %.1f e2e %.1f ms/step %.4f host_us %.1f train %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["host_us_per_scene"], json.dumps(d["train_C3"])[:160]))
PY
