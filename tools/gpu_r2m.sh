#!/bin/bash
# 2-GPU box: torchrun path of bench.py (scenes sharded, DDP training) and the C5 eval sweep at 1 and 2 ranks
mkdir -p gpurun_out
python tools/eval_sweep.py --workload C5 --scenes 8 2> gpurun_out/r2m_sweep1.err | tail -1 > gpurun_out/r2m_eval_sweep_C5_1gpu.json; tail -2 gpurun_out/r2m_sweep1.err; cut -c1-600 gpurun_out/r2m_eval_sweep_C5_1gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/eval_sweep.py --workload C5 --scenes 8 2> gpurun_out/r2m_sweep2.err | tail -1 > gpurun_out/r2m_eval_sweep_C5_2gpu.json; tail -2 gpurun_out/r2m_sweep2.err; cut -c1-600 gpurun_out/r2m_eval_sweep_C5_2gpu.json
python bench.py --gpus 1 --steps 100 --cpu-seconds 2 2> gpurun_out/r2m_b1.err > gpurun_out/r2m_bench_1gpu.json; tail -2 gpurun_out/r2m_b1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --cpu-seconds 2 2> gpurun_out/r2m_b2.err > gpurun_out/r2m_bench_2gpu.json; tail -3 gpurun_out/r2m_b2.err
python - <<PY
import json
for n in (1,2):
    d=json.loads(open("gpurun_out/r2m_bench_%dgpu.json"%n).read().strip().splitlines()[-1])
    print(n, "value %.1f e2e %.1f host_us %.1f train %s" % (d["value"], d["e2e"]["value"], d["host_us_per_scene"], json.dumps(d["train_C3"])[:200]))
PY
