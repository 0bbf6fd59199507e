#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --cpu-seconds 1 --steps 60 2>gpurun_out/r2q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('module_api', d['e2e']['module_api_ms_per_scene'], 'value', d['value'], 'train', d['train_C3'])"
tail -2 gpurun_out/r2q.err
