#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_vote_gpu.py -x -q 2>&1 | tail -15 ) > gpurun_out/r2d_vote_tests.log 2>&1
cat gpurun_out/r2d_vote_tests.log
python tools/time_vote.py all > gpurun_out/r2d_time_vote.txt 2>&1; cat gpurun_out/r2d_time_vote.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_' -c 40 --csv --log-file gpurun_out/r2d_vote_launches.csv python tools/time_vote.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2d_vote_launches.csv")) if len(r)>10]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
for r in rows[1:14]: print(r[k][:50], r[v])
PY
python -m pytest tests/test_parity_gpu.py -q -s > gpurun_out/r2d_parity_tests.log 2>&1
grep -E "engine C2|class_pred|boxes|training step|passed|failed|Error" gpurun_out/r2d_parity_tests.log
python -m pytest tests/test_sparse_gpu.py -x -q 2>&1 | tail -5
