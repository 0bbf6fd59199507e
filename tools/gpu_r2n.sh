#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --cpu-seconds 2 --train-steps 0 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -2 gpurun_out/r2n_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_bench.json").read())
print("value %.1f ms/step %.4f passes %s conv kernel_ms %.4f flushed step %.4f e2e %.1f launches %d" % (d["value"], d["ms_per_step"], d["passes_ms_per_step"], d["roofline"]["kernel_ms"], d["step_ms_flushed"]["median"], d["e2e"]["value"], d["gpu_launches"]))
PY
