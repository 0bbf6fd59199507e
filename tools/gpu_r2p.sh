#!/bin/bash
mkdir -p gpurun_out
python tools/profile_train.py > gpurun_out/r2p_profile_train.txt 2>&1; tail -30 gpurun_out/r2p_profile_train.txt
python bench.py --cpu-seconds 1 --train-steps 0 --steps 60 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('module_api', d['e2e']['module_api_ms_per_scene'], 'value', d['value'])"
